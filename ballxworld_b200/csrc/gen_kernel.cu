// gen_kernel.cu — bxw_world::stdgen::StdGenerator as sm_100a kernels.
//
// The reference evaluates calc_voxel_params for the 1024 columns of EVERY chunk (stdgen.rs:335-347), so a vertical
// stack of chunks recomputes the same columns. Here one CTA owns a vertical stack: it computes the parameters of its 32x32
// columns once (fill_kernel<64 / 32>; they also go to the region's height map), derives the chunk summaries and the page
// table from them -- a chunk that is all void or all stone is not written at all, its page-table entry points at a shared
// page -- and writes the block ids of the chunks in between with 16-byte coalesced stores (stdgen.rs:349-373).
// heightmap_kernel computes the same parameters alone (raw getters, FP32 report).
//
// The f64 path performs exactly the IEEE operations of the reference, in its order, with explicit
// round-to-nearest intrinsics so that nvcc never contracts a*b+c into an FMA (Rust/LLVM does not contract).
#include "device_common.cuh"

namespace bxw {
namespace {

// ---- noise 0.8.2 super_simplex_2d constants (see oracle/noise.c for provenance: crate source not vendored) ----
constexpr double kToReal = -0.211324865405187;
constexpr double kToSimplex = 0.366025403784439;
constexpr double kNorm = 1.0 / 0.05428295288661623;
constexpr double kDiag = 0.70710678118654752440;

#define LA 0.788675134594813
#define LB 0.211324865405187
#define LC 0.577350269189626
#define LD 1.366025403784439
#define LE 0.36602540378443904
const double h_lat_off[32][2] = {
    {0.0, 0.0}, {-LC, -LC}, {LA, -LB},  {-LB, LA}, {0.0, 0.0}, {-LC, -LC}, {LB, -LA},  {-LA, LB},
    {0.0, 0.0}, {-LC, -LC}, {-LA, LB},  {-LB, LA}, {0.0, 0.0}, {-LC, -LC}, {-LD, -LE}, {-LA, LB},
    {0.0, 0.0}, {-LC, -LC}, {LA, -LB},  {LB, -LA}, {0.0, 0.0}, {-LC, -LC}, {LB, -LA},  {-LE, -LD},
    {0.0, 0.0}, {-LC, -LC}, {-LA, LB},  {LB, -LA}, {0.0, 0.0}, {-LC, -LC}, {-LD, -LE}, {-LE, -LD},
};
const signed char h_lat_pt[32][2] = {
    {0, 0}, {1, 1}, {-1, 0}, {0, -1}, {0, 0}, {1, 1}, {0, 1}, {1, 0}, {0, 0}, {1, 1}, {1, 0},  {0, -1}, {0, 0}, {1, 1}, {2, 1}, {1, 0},
    {0, 0}, {1, 1}, {-1, 0}, {0, 1},  {0, 0}, {1, 1}, {0, 1}, {1, 2}, {0, 0}, {1, 1}, {1, 0},  {0, 1},  {0, 0}, {1, 1}, {2, 1}, {1, 2},
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// Where the generator's tables are read from. Global: a pointer (height map / sample kernels). Shared: fill_kernel stages the
// tables per CTA, and the noise subroutine -- which is not inlined, so a pointer would arrive as a GENERIC address: 64-bit
// address arithmetic and generic loads for its 16 lookups -- gets the 32-bit shared-memory address and uses ld.shared.
constexpr uint32_t kTabGrad = (uint32_t)offsetof(GenTables, grad), kTabLat = (uint32_t)offsetof(GenTables, lat);
struct TabsGlobal {
    const GenTables *T;
    __device__ __forceinline__ unsigned u8(uint32_t off) const { return reinterpret_cast<const uint8_t *>(T)[off]; }
    __device__ __forceinline__ double2 f64x2(uint32_t off) const {
        return *reinterpret_cast<const double2 *>(reinterpret_cast<const uint8_t *>(T) + off);
    }
    __device__ __forceinline__ int2 s32x2(uint32_t off) const {
        return *reinterpret_cast<const int2 *>(reinterpret_cast<const uint8_t *>(T) + off);
    }
};
struct TabsShared {
    uint32_t ts; // shared-window address of the CTA's GenTables copy (written before a __syncthreads, read-only afterwards)
    __device__ __forceinline__ unsigned u8(uint32_t off) const {
        unsigned v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(ts + off));
        return v;
    }
    __device__ __forceinline__ double2 f64x2(uint32_t off) const {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ts + off));
        return v;
    }
    __device__ __forceinline__ int2 s32x2(uint32_t off) const {
        int2 v;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(ts + off));
        return v;
    }
};

// noise::SuperSimplex::get([x, y]) (f64). One copy per kernel (not inlined): the height map kernel calls it from 11 sites and
// fully inlined it was 141 KB of SASS, more than half of its warp stalls were instruction-cache misses.
template <class TB>
__device__ __noinline__ double super_simplex_2d(const TB T, int which, double px, double py) {
    const uint32_t perm = (uint32_t)which * 256u;
    const double to_simplex_offset = dmul(dadd(px, py), kToSimplex);
    const double sx = dadd(px, to_simplex_offset), sy = dadd(py, to_simplex_offset);
    const double bxf = floor(sx), byf = floor(sy);
    // only the low 8 bits of the lattice coordinates reach the permutation table
    const unsigned bxi = (unsigned)__double2ll_rz(bxf), byi = (unsigned)__double2ll_rz(byf);
    const double rx = dsub(sx, bxf), ry = dsub(sy, byf);
    const double region_sum = floor(dadd(rx, ry));
    const double half_rs = dmul(region_sum, 0.5);
    unsigned group = region_sum >= 1.0 ? 1u : 0u;
    group |= (dsub(dadd(dsub(rx, dmul(ry, 0.5)), 1.0), half_rs) >= 1.0 ? 1u : 0u) << 1;
    group |= (dsub(dadd(dsub(ry, dmul(rx, 0.5)), 1.0), half_rs) >= 1.0 ? 1u : 0u) << 2;
    const double to_real_offset = dmul(dadd(rx, ry), kToReal);
    const double qx = dadd(rx, to_real_offset), qy = dadd(ry, to_real_offset);
    double value = 0.0;
    auto point = [&](double ox, double oy, unsigned lpx, unsigned lpy) {
        const double dx = dadd(qx, ox), dy = dadd(qy, oy);
        const double attn = dsub(2.0 / 3.0, dadd(dmul(dx, dx), dmul(dy, dy)));
        // The reference adds the term only if attn > 0. Here the term is always computed and +0.0 is added otherwise: the sum
        // starts at +0.0 and can never become -0.0 under round-to-nearest, so adding +0.0 leaves it bit for bit unchanged -- and
        // without the four branches the lookup chains of the four lattice points overlap (a warp nearly always had a lane
        // inside every point's radius anyway).
        const unsigned h = T.u8(perm + (T.u8(perm + ((bxi + lpx) & 0xffu)) ^ ((byi + lpy) & 0xffu)));
        // gradient::grad2(h % 8) = (1,0) (-1,0) (0,1) (0,-1) (d,d) (-d,d) (d,-d) (-d,-d): one 16-byte table load
        const double2 gr = T.f64x2(kTabGrad + (h & 7u) * 16u);
        const double a2 = dmul(attn, attn);
        const double term = dmul(dmul(a2, a2), dadd(dmul(gr.x, dx), dmul(gr.y, dy)));
        value = dadd(value, attn > 0.0 ? term : 0.0);
    };
    // lattice points 0 and 1 are (0, 0) and (1, 1) in every group; 2 and 3 come from the group's table rows
    point(0.0, 0.0, 0u, 0u);
    point(-LC, -LC, 1u, 1u);
#pragma unroll
    for (unsigned k = 0; k < 2; k++) {
        const uint32_t e = kTabLat + (group * 2u + k) * (uint32_t)sizeof(LatticeEntry);
        const double2 o = T.f64x2(e);
        const int2 lp = T.s32x2(e + 16u);
        point(o.x, o.y, (unsigned)lp.x, (unsigned)lp.y);
    }
    return dmul(value, kNorm);
}

// FP32 evaluation of the same noise (the north_star-literal variant). The lattice cell (floor of the skewed coordinates) and the
// three bits that pick the lattice-point group are taken in f64 exactly as above -- a dozen operations; at world coordinates of
// 10^4 an f32 position has an error of 10^-3, which would move lattice decisions and broke the 1e-5 bound in round 1 -- so the
// same lattice points and gradients are used as in the f64 path. Everything after that works on the relative coordinates in
// [0, 1): offsets, attenuation, gradient dot products and the sum are f32 (about 80 % of the arithmetic).
template <class TB>
__device__ __noinline__ float super_simplex_2d_f32(const TB T, int which, double px, double py) {
    const uint32_t perm = (uint32_t)which * 256u;
    const double to_simplex_offset = dmul(dadd(px, py), kToSimplex);
    const double sx = dadd(px, to_simplex_offset), sy = dadd(py, to_simplex_offset);
    const double bxf = floor(sx), byf = floor(sy);
    const unsigned bxi = (unsigned)__double2ll_rz(bxf), byi = (unsigned)__double2ll_rz(byf);
    const double rxd = dsub(sx, bxf), ryd = dsub(sy, byf);
    const double region_sum = floor(dadd(rxd, ryd));
    const double half_rs = dmul(region_sum, 0.5);
    unsigned group = region_sum >= 1.0 ? 1u : 0u;
    group |= (dsub(dadd(dsub(rxd, dmul(ryd, 0.5)), 1.0), half_rs) >= 1.0 ? 1u : 0u) << 1;
    group |= (dsub(dadd(dsub(ryd, dmul(rxd, 0.5)), 1.0), half_rs) >= 1.0 ? 1u : 0u) << 2;
    const float rx = (float)rxd, ry = (float)ryd;
    const float to_real_offset = __fmul_rn(__fadd_rn(rx, ry), (float)kToReal);
    const float qx = __fadd_rn(rx, to_real_offset), qy = __fadd_rn(ry, to_real_offset);
    float value = 0.0f;
    auto point = [&](float ox, float oy, unsigned lpx, unsigned lpy) {
        const float dx = __fadd_rn(qx, ox), dy = __fadd_rn(qy, oy);
        const float attn = __fsub_rn(2.0f / 3.0f, __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        const unsigned h = T.u8(perm + (T.u8(perm + ((bxi + lpx) & 0xffu)) ^ ((byi + lpy) & 0xffu)));
        const double2 gr = T.f64x2(kTabGrad + (h & 7u) * 16u);
        const float a2 = __fmul_rn(attn, attn);
        const float term = __fmul_rn(__fmul_rn(a2, a2), __fadd_rn(__fmul_rn((float)gr.x, dx), __fmul_rn((float)gr.y, dy)));
        value = __fadd_rn(value, attn > 0.0f ? term : 0.0f); // (as in the f64 path: adding +0.0 is exact)
    };
    point(0.0f, 0.0f, 0u, 0u);
    point((float)-LC, (float)-LC, 1u, 1u);
#pragma unroll
    for (unsigned k = 0; k < 2; k++) {
        const uint32_t e = kTabLat + (group * 2u + k) * (uint32_t)sizeof(LatticeEntry);
        const double2 o = T.f64x2(e);
        const int2 lp = T.s32x2(e + 16u);
        point((float)o.x, (float)o.y, (unsigned)lp.x, (unsigned)lp.y);
    }
    return __fmul_rn(value, (float)kNorm);
}

// precision-generic noise returning double
template <int PREC, class TB>
__device__ __forceinline__ double noise_at(const TB T, int which, double px, double py) {
    if (PREC == 32) return (double)super_simplex_2d_f32(T, which, px, py);
    return super_simplex_2d(T, which, px, py);
}

// The scaled column coordinates x / sf, z / sf of the four noise scales (elevation 2560, plains 1200, hills 1600, mountains
// 2000). A true division each (sf has no exact reciprocal), and a software routine of ~25 FP64 operations on this chip: the
// fill kernel computes them once per CTA row / column (64 x 4 divisions) instead of eight per column.
struct ColCoords {
    double ex, ez, px, pz, hx, hz, mx, mz;
};
__device__ __forceinline__ ColCoords col_coords(int32_t x, int32_t z) {
    ColCoords c;
    c.ex = ddiv((double)x, 2560.0);
    c.ez = ddiv((double)z, 2560.0);
    c.px = ddiv((double)x, 1200.0);
    c.pz = ddiv((double)z, 1200.0);
    c.hx = ddiv((double)x, 1600.0);
    c.hz = ddiv((double)z, 1600.0);
    c.mx = ddiv((double)x, 2000.0);
    c.mz = ddiv((double)z, 2000.0);
    return c;
}

// (n + 1) / 2 of the reference is computed as (n + 1) * 0.5: halving is exact in binary floating point either way
template <int PREC, class TB>
__device__ __forceinline__ double elevation_noise_at(const TB T, double ex, double ez) {
    return dmul(dadd(noise_at<PREC>(T, 5, ex, ez), 1.0), 0.5);
}

template <int PREC, class TB>
__device__ __forceinline__ double elevation_noise(const TB T, int32_t x, int32_t z) {
    // stdgen.rs:171-175: GLOBAL_BIOME_SCALE * GLOBAL_SCALE_MOD = 2560
    const double sf = 256.0 * 10.0;
    return elevation_noise_at<PREC>(T, ddiv((double)x, sf), ddiv((double)z, sf));
}

template <int PREC, class TB>
__device__ __forceinline__ double h_n(const TB T, int i, double px, double pz) {
    return dmul(dadd(noise_at<PREC>(T, i, px, pz), 1.0), 0.5); // stdgen.rs:178-180 ((n + 1) / 2, exact as * 0.5)
}
template <int PREC, class TB>
__device__ __forceinline__ double h_rn(const TB T, int i, double px, double pz) {
    return dmul(2.0, dsub(0.5, fabs(dsub(0.5, h_n<PREC>(T, i, px, pz))))); // stdgen.rs:183-185
}

template <int PREC, class TB>
__device__ double plains_height(const TB T, double px, double pz) {
    // stdgen.rs:187-192; px, pz = pos / (10 * 120)
    const double a = dmul(0.75, h_n<PREC>(T, 0, px, pz));
    const double b = dmul(0.25, h_n<PREC>(T, 1, dmul(px, 2.0), dmul(pz, 2.0)));
    return dadd(dmul(dadd(a, b), 5.0), 10.0);
}

template <int PREC, class TB>
__device__ double hills_height(const TB T, double px, double pz) {
    // stdgen.rs:194-200; px, pz = pos / (10 * 160)
    const double a = dmul(0.60, h_n<PREC>(T, 0, px, pz));
    const double b = dmul(0.25, h_n<PREC>(T, 1, dmul(px, 1.5), dmul(pz, 1.5)));
    const double c = dmul(0.15, h_n<PREC>(T, 2, dmul(px, 3.0), dmul(pz, 3.0)));
    return dadd(dmul(dadd(dadd(a, b), c), 30.0), 15.0);
}

template <int PREC, class TB>
__device__ double mountains_height(const TB T, double px, double pz) {
    // stdgen.rs:202-213; px, pz = pos / (10 * 200)
    const double h0 = dmul(0.50, h_rn<PREC>(T, 0, px, pz));
    const double h01 = dadd(dmul(0.25, h_rn<PREC>(T, 1, dmul(px, 2.0), dmul(pz, 2.0))), h0);
    const double q = ddiv(h01, 0.75);
    const double t1 = dmul(dmul(q, 0.15), h_n<PREC>(T, 2, dmul(px, 5.0), dmul(pz, 5.0)));
    const double t2 = dmul(dmul(q, 0.05), h_rn<PREC>(T, 3, dmul(px, 9.0), dmul(pz, 9.0)));
    return dadd(dmul(dadd(dadd(h01, t1), t2), 750.0), 40.0);
}

__device__ __forceinline__ uint64_t sm64_next(uint64_t &x) {
    x += 0x9e3779b97f4a7c15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// CellGen::get_cell_points (stdgen.rs:116-147): one thread per super-cell
template <int PREC>
__global__ void cellpoints_kernel(uint64_t seed, int32_t cell_x0, int32_t cell_z0, int32_t cells_w, int32_t cells_d,
                                  const GenTables *T, CellPointDev *cells) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells_w * cells_d) return;
    const int32_t cx = cell_x0 + i % cells_w, cz = cell_z0 + i / cells_w;
    uint64_t s = seed ^ ((((uint64_t)(int64_t)cx) << 32) | (((uint64_t)(int64_t)cz) & 0xFFFFFFFFULL));
    uint64_t st[4]; // Xoshiro256StarStar::seed_from_u64 (rand_xoshiro 0.6.0)
    for (int k = 0; k < 4; k++) st[k] = sm64_next(s);
    const int bx[4] = {32, 96, 0, 64}, bz[4] = {32, 32, 96, 96};
    for (int k = 0; k < 4; k++) {
        uint32_t r[2];
        for (int q = 0; q < 2; q++) {
            const uint64_t res = rotl64(st[1] * 5, 7) * 9;
            const uint64_t t = st[1] << 17;
            st[2] ^= st[0];
            st[3] ^= st[1];
            st[1] ^= st[2];
            st[0] ^= st[3];
            st[2] ^= t;
            st[3] = rotl64(st[3], 45);
            r[q] = (uint32_t)(res >> 32);
        }
        CellPointDev cp;
        cp.px = cx * 128 + bx[k] + ((int32_t)(r[0] % 32u) - 16);
        cp.pz = cz * 128 + bz[k] + ((int32_t)(r[1] % 32u) - 16);
        cp.elevation_class = elevation_noise<PREC>(TabsGlobal{T}, cp.px, cp.pz);
        cells[i * 4 + k] = cp;
    }
}

// CellGen::calc_voxel_params (stdgen.rs:215-276) of column (ix, iz) of the height map: returns height*4 + elevation class
template <int PREC, class TB>
// cands != nullptr: the 36 cell points around the column's super-cell in the reference's visiting order (cdx outer, cdy inner,
// point index innermost), gathered once per CTA (every column of a chunk lies in one super-cell); T: the permutation tables
__device__ __forceinline__ int32_t column_params(const GenParams &p, const TB T, const CellPointDev *cands, int n_cands, int ix,
                                                 int iz, double *raw_out, const ColCoords &cc) {
    const int32_t x = p.x0 + ix, z = p.z0 + iz;

    // find_nearest_cell_points(pos, 1) (stdgen.rs:149-169): truncating division, first minimum wins
    int32_t best = 0x7FFFFFFF;
    double cec = 0.0;
    bool have = false;
    if (cands) {
#pragma unroll 4
        for (int k = 0; k < n_cands; k++) {
            const int32_t dx = cands[k].px - x, dz = cands[k].pz - z;
            const int32_t dist = dx * dx + dz * dz;
            if (!have || dist < best) {
                have = true;
                best = dist;
                cec = cands[k].elevation_class;
            }
        }
    } else {
        const int32_t cellx = x / 128, cellz = z / 128;
        for (int cdx = -1; cdx <= 1; cdx++)
            for (int cdy = -1; cdy <= 1; cdy++) {
                const CellPointDev *c = p.cells + ((size_t)(cellz + cdy - p.cell_z0) * p.cells_w + (cellx + cdx - p.cell_x0)) * 4;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int32_t dx = c[k].px - x, dz = c[k].pz - z;
                    const int32_t dist = dx * dx + dz * dz;
                    if (!have || dist < best) {
                        have = true;
                        best = dist;
                        cec = c[k].elevation_class;
                    }
                }
            }
    }
    const double ec = elevation_noise_at<PREC>(T, cc.ex, cc.ez);
    // stdgen.rs:222-243: <0.4 plains / <0.5 lerp plains-hills / <0.7 hills / <0.8 lerp hills-mountains / else mountains.
    // Each biome height is evaluated at most once per column and at one call site.
    const bool need_p = ec < 0.5, need_h = !(ec < 0.4) && ec < 0.8, need_m = !(ec < 0.7);
    double ph = 0.0, hh = 0.0, mh = 0.0;
    if (need_p) ph = plains_height<PREC>(T, cc.px, cc.pz);
    if (need_h) hh = hills_height<PREC>(T, cc.hx, cc.hz);
    if (need_m) mh = mountains_height<PREC>(T, cc.mx, cc.mz);
    double height;
    if (ec < 0.4) {
        height = ph;
    } else if (ec < 0.5) {
        const double nlin = ddiv(dsub(ec, 0.4), 0.1);
        const double olin = dsub(1.0, nlin);
        height = dadd(dmul(olin, ph), dmul(nlin, hh));
    } else if (ec < 0.7) {
        height = hh;
    } else if (ec < 0.8) {
        const double nlin = ddiv(dsub(ec, 0.7), 0.1);
        const double olin = dsub(1.0, nlin);
        height = dadd(dmul(olin, hh), dmul(nlin, mh));
    } else {
        height = mh;
    }
    int elevation; // stdgen.rs:245-267
    if (cec < 0.4) {
        elevation = 0;
    } else if (cec < 0.5 || (!(cec < 0.7) && cec < 0.8)) {
        const int lo = cec < 0.5 ? 0 : 1;
        elevation = noise_at<PREC>(T, 4, (double)x, (double)z) < 0.0 ? lo : lo + 1;
    } else if (cec < 0.7) {
        elevation = 1;
    } else {
        elevation = 2;
    }
    // height.round() as i32 (stdgen.rs:272): half away from zero, saturating
    double r = round(height);
    int32_t h;
    if (!(r == r)) h = 0;
    else if (r >= 536870911.0) h = 536870911; // packed into 30 bits below; the terrain never approaches this
    else if (r <= -536870912.0) h = -536870912;
    else h = (int32_t)r;
    if (raw_out) *raw_out = height;
    return h * 4 + elevation;
}

// one thread per column
template <int PREC>
__global__ void __launch_bounds__(128) heightmap_kernel(GenParams p) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iz = blockIdx.y;
    if (ix >= p.W) return;
    double raw;
    p.hmap[(size_t)iz * p.W + ix] = column_params<PREC>(p, TabsGlobal{p.tables}, nullptr, 0, ix, iz, &raw, col_coords(p.x0 + ix, p.z0 + iz));
    if (p.raw) p.raw[(size_t)iz * p.W + ix] = raw;
}


// StdGenerator::generate_chunk block selection (stdgen.rs:349-373) for a vertical stack of chunks.
// CTA = one (sx,sz) chunk column; thread = 4 consecutive x of one z row; loops over all layers.
// GEN = 0: column parameters come from the height map. GEN = 32 / 64: the CTA first computes the parameters of its 32x32
// columns itself (4 per thread, FP32 / f64 noise) and stores them to the height map, then writes the stack: CTAs in their
// FP64 phase share an SM with CTAs in their store phase, which hides the height map behind the HBM writes.
// chunk column of CTA b along x (FillParams::col_mode)
__device__ __forceinline__ int fill_column(const FillParams &p, int b) {
    const int nx = p.SX - 2;
    if (p.col_mode == 1) return b == 0 ? 1 : nx;
    if (p.col_mode == 2) return b == 0 ? 0 : (nx == 1 ? 2 : (b < nx - 1 ? b + 1 : b + 2));
    return b;
}

template <int GEN>
__global__ void __launch_bounds__(256, 4) fill_kernel(FillParams p, GenParams g) {
    int sx, sz, sy_lo, sy_hi;
    if (p.list) { // load deltas: one listed chunk per CTA (the reference's granularity, generation.rs:96-148)
        const uint32_t c = p.list[blockIdx.x];
        sx = (int)(c % (uint32_t)p.SX);
        sz = (int)((c / (uint32_t)p.SX) % (uint32_t)p.SZ);
        sy_lo = (int)(c / (uint32_t)(p.SX * p.SZ));
        sy_hi = sy_lo + 1;
    } else {
        sx = fill_column(p, (int)blockIdx.x);
        sz = blockIdx.y;
        sy_lo = 0;
        sy_hi = p.SY;
    }
    const int q = threadIdx.x & 7, zr = threadIdx.x >> 3;
    int4 hm;
    if (GEN == 0) {
        hm = __ldg(reinterpret_cast<const int4 *>(p.hmap + (size_t)(sz * 32 + zr) * p.W + sx * 32 + q * 4));
    } else {
        // scaled coordinates of the CTA's 32 x values and 32 z values at the four noise scales: one division per thread
        __shared__ double s_coord[4][2][32];
        __shared__ GenTables s_tab;         // the seven permutation tables (1.8 KB): two dependent lookups per lattice point
        __shared__ CellPointDev s_cand[36]; // the cell points around this chunk column's super-cell, in visiting order ...
        __shared__ CellPointDev s_near[36]; // ... and those of them that can be the nearest one of some column of this chunk
        __shared__ int s_keep[36], s_bound, s_ncand;
        for (int i = threadIdx.x; i < (int)(sizeof(GenTables) / 4); i += 256)
            reinterpret_cast<uint32_t *>(&s_tab)[i] = reinterpret_cast<const uint32_t *>(g.tables)[i];
        // A chunk is 32 columns wide and a super-cell 128, so its columns share pos / 128 -- unless the chunk starts at a
        // negative multiple of 128: the division truncates (stdgen.rs:150), -128 / 128 = -1 but -127 / 128 = 0. Those chunks
        // search per column.
        const int32_t cx_lo = g.x0 + sx * 32, cz_lo = g.z0 + sz * 32;
        const bool one_cell = (cx_lo / 128 == (cx_lo + 31) / 128) && (cz_lo / 128 == (cz_lo + 31) / 128);
        // Pruning: a point whose smallest distance to the chunk's 32 x 32 columns exceeds the LARGEST distance of some other
        // point is farther than that one from every column, so it can neither win nor tie (find_nearest_cell_points keeps the
        // first minimum, stdgen.rs:149-169; the survivors keep their visiting order). 36 -> typically 5-10 candidates.
        int32_t c_min = 0, c_max = 0;
        if (threadIdx.x == 0) s_bound = 0x7FFFFFFF;
        if (threadIdx.x < 36 && one_cell) {
            const int32_t cellx = cx_lo / 128, cellz = cz_lo / 128;
            const int k = threadIdx.x, cdx = k / 12 - 1, cdy = (k / 4) % 3 - 1;
            const CellPointDev c = g.cells[((size_t)(cellz + cdy - g.cell_z0) * g.cells_w + (cellx + cdx - g.cell_x0)) * 4 + (k & 3)];
            s_cand[k] = c;
            const int32_t ax = cx_lo - c.px, bx = c.px - (cx_lo + 31), az = cz_lo - c.pz, bz = c.pz - (cz_lo + 31);
            const int32_t nx = max(max(ax, bx), 0), nz = max(max(az, bz), 0);   // nearest column
            const int32_t fx = max(abs(ax), abs(bx)), fz = max(abs(az), abs(bz)); // farthest column
            c_min = nx * nx + nz * nz;
            c_max = fx * fx + fz * fz;
        }
        {
            const int t = threadIdx.x, sc = t >> 6, axis = (t >> 5) & 1, i = t & 31;
            const double sf = sc == 0 ? 2560.0 : (sc == 1 ? 1200.0 : (sc == 2 ? 1600.0 : 2000.0));
            const int32_t coord = axis ? g.z0 + sz * 32 + i : g.x0 + sx * 32 + i;
            s_coord[sc][axis][i] = ddiv((double)coord, sf);
        }
        __syncthreads();
        if (threadIdx.x < 36 && one_cell) atomicMin(&s_bound, c_max);
        __syncthreads();
        if (threadIdx.x < 36 && one_cell) s_keep[threadIdx.x] = c_min <= s_bound;
        __syncthreads();
        if (threadIdx.x < 36 && one_cell) {
            int pos = 0;
            for (int k = 0; k < (int)threadIdx.x; k++) pos += s_keep[k];
            if (s_keep[threadIdx.x]) s_near[pos] = s_cand[threadIdx.x];
            if (threadIdx.x == 35) s_ncand = pos + s_keep[35];
        }
        __syncthreads();
        const int n_cands = one_cell ? s_ncand : 0;
        int hv[4];
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            const int xi = q * 4 + k;
            ColCoords cc;
            cc.ex = s_coord[0][0][xi];
            cc.ez = s_coord[0][1][zr];
            cc.px = s_coord[1][0][xi];
            cc.pz = s_coord[1][1][zr];
            cc.hx = s_coord[2][0][xi];
            cc.hz = s_coord[2][1][zr];
            cc.mx = s_coord[3][0][xi];
            cc.mz = s_coord[3][1][zr];
            hv[k] = column_params<(GEN == 0 ? 64 : GEN)>(g, TabsShared{(uint32_t)__cvta_generic_to_shared(&s_tab)}, one_cell ? s_near : nullptr, n_cands, sx * 32 + q * 4 + k, sz * 32 + zr, nullptr, cc);
        }
        hm = make_int4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<int4 *>(g.hmap + (size_t)(sz * 32 + zr) * p.W + sx * 32 + q * 4) = hm;
    }
    const int hh[4] = {hm.x >> 2, hm.y >> 2, hm.z >> 2, hm.w >> 2};
    const int el[4] = {hm.x & 3, hm.y & 3, hm.z & 3, hm.w & 3};
    // the block at y == h: snow grass on mountains above 80, else grass (stdgen.rs:361-366) -- a per-column constant
    uint32_t top[4];
#pragma unroll
    for (int k = 0; k < 4; k++) top[k] = (el[k] == 2 && hh[k] > 80) ? p.id_snow : p.id_grass;
    // chunk summaries of the stack: above every column -> all void; more than 5 below every column -> all stone; at or below
    // every column (of the chunk / of one of its four side planes) -> solid terrain (device_common.cuh: kSolidShift)
    __shared__ int red[6][8];
    const int big = 0x7FFFFFFF;
    int lo = min(min(hh[0], hh[1]), min(hh[2], hh[3])), hi = max(max(hh[0], hh[1]), max(hh[2], hh[3]));
    const int t_lo = lo, t_hi = hi; // lowest / highest surface among this thread's four columns
    int lx0 = q == 0 ? hh[0] : big, lx1 = q == 7 ? hh[3] : big;
    int lz0 = zr == 0 ? lo : big, lz1 = zr == 31 ? lo : big;
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    // range of the per-voxel selection in the store loop: per thread when only the surface chunks are written (fewer selected
    // layers), per warp when every chunk is (most chunks are all stone / all void: no divergence)
    const int b_lo = p.elide ? t_lo : lo, b_hi = p.elide ? t_hi : hi;
    lx0 = __reduce_min_sync(0xFFFFFFFFu, lx0);
    lx1 = __reduce_min_sync(0xFFFFFFFFu, lx1);
    lz0 = __reduce_min_sync(0xFFFFFFFFu, lz0);
    lz1 = __reduce_min_sync(0xFFFFFFFFu, lz1);
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        red[0][w] = lo;
        red[1][w] = hi;
        red[2][w] = lx0;
        red[3][w] = lx1;
        red[4][w] = lz0;
        red[5][w] = lz1;
    }
    __syncthreads();
    for (int k = 0; k < 8; k++) {
        lo = min(lo, red[0][k]);
        hi = max(hi, red[1][k]);
        lx0 = min(lx0, red[2][k]);
        lx1 = min(lx1, red[3][k]);
        lz0 = min(lz0, red[4][k]);
        lz1 = min(lz1, red[5][k]);
    }
    if (p.summary) {
        for (int sy = sy_lo + (int)threadIdx.x; sy < sy_hi; sy += 256) {
            const int ybase = (p.cy_min + sy) * 32, ytop = ybase + 31;
            unsigned long long sum = 0;
            if (ybase > hi) sum = kUniformBit | p.id_void;
            else if (ytop < lo - 5) sum = kUniformBit | p.id_stone;
            // a voxel at height y of a column with surface h is grass / snow grass / dirt / stone iff y <= h
            const unsigned solid = (ytop <= lo ? 1u : 0u) | (ytop <= lx0 ? 2u : 0u) | (ytop <= lx1 ? 4u : 0u) | (ybase <= lo ? 8u : 0u) |
                                   (ytop <= lo ? 16u : 0u) | (ytop <= lz0 ? 32u : 0u) | (ytop <= lz1 ? 64u : 0u);
            p.summary[(size_t)(sx + p.SX * (sz + p.SZ * sy))] = sum | ((unsigned long long)solid << kSolidShift);
        }
    }
    // Uniform chunks are not materialised: their page entry points at the shared all-void / all-stone page (every reader goes
    // through the page table; the first writer materialises the chunk). Same test as the summary above; thread per chunk.
    if (p.page) {
        for (int sy = sy_lo + (int)threadIdx.x; sy < sy_hi; sy += 256) {
            const uint32_t c = (uint32_t)(sx + p.SX * (sz + p.SZ * sy));
            const int ybase = (p.cy_min + sy) * 32;
            uint32_t pg = c;
            if (p.elide) {
                if (ybase > hi) pg = p.page_void;
                else if (ybase + 31 < lo - 5) pg = p.page_stone;
            }
            p.page[c] = pg;
        }
    }
    // the chunks in between are written: ybase <= hi and ybase + 31 >= lo - 5 (floor divisions: heights may be negative)
    int sy_m0 = sy_lo, sy_m1 = sy_hi;
    if (p.elide) {
        sy_m0 = max(sy_lo, ((lo - 5) >> 5) - p.cy_min);
        sy_m1 = min(sy_hi, (hi >> 5) - p.cy_min + 1);
    }
    const uint4 stone4 = make_uint4(p.id_stone, p.id_stone, p.id_stone, p.id_stone), void4 = make_uint4(p.id_void, p.id_void, p.id_void, p.id_void);
    for (int sy = sy_m0; sy < sy_m1; sy++) {
        const uint32_t c = (uint32_t)(sx + p.SX * (sz + p.SZ * sy));
        const int ybase = (p.cy_min + sy) * 32;
        uint4 *dst = reinterpret_cast<uint4 *>(p.blocks + (size_t)c * kChunkVoxels + zr * 32 + q * 4);
        // layers [0, y_mixed) are stone and layers [y_void, 32) are void for the four columns of this thread (terrain is smooth:
        // the per-voxel selection runs for six or seven layers); 1024 voxels per y layer = 256 uint4
        const int y_mixed = min(max(b_lo - 5 - ybase, 0), 32), y_void = min(max(b_hi + 1 - ybase, y_mixed), 32);
#pragma unroll 4
        for (int yc = 0; yc < y_mixed; yc++) __stcs(&dst[yc * 256], stone4); // streaming stores
#pragma unroll 2
        for (int yc = y_mixed; yc < y_void; yc++) {
            const int y = ybase + yc;
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int t = hh[k] - y; // depth below the surface (stdgen.rs:359-372)
                o[k] = t < 0 ? p.id_void : (t == 0 ? top[k] : (t <= 5 ? p.id_dirt : p.id_stone));
            }
            __stcs(&dst[yc * 256], make_uint4(o[0], o[1], o[2], o[3]));
        }
#pragma unroll 4
        for (int yc = y_void; yc < 32; yc++) __stcs(&dst[yc * 256], void4);
    }
}

// The x=0 / x=31 voxel planes of every chunk, straight from the column parameters (same selection rule as fill_kernel).
// One CTA per (sx, sz) chunk column, thread = (side, z); loops over all layers; coalesced 128-byte stores.
__global__ void __launch_bounds__(64) fill_xface_kernel(FillParams p) {
    int sx, sz, sy_lo, sy_hi;
    if (p.list) {
        const uint32_t c = p.list[blockIdx.x];
        sx = (int)(c % (uint32_t)p.SX);
        sz = (int)((c / (uint32_t)p.SX) % (uint32_t)p.SZ);
        sy_lo = (int)(c / (uint32_t)(p.SX * p.SZ));
        sy_hi = sy_lo + 1;
    } else {
        sx = fill_column(p, (int)blockIdx.x);
        sz = blockIdx.y;
        sy_lo = 0;
        sy_hi = p.SY;
    }
    const int side = threadIdx.x >> 5, z = threadIdx.x & 31;
    const int hm = __ldg(p.hmap + (size_t)(sz * 32 + z) * p.W + sx * 32 + (side ? 31 : 0));
    const int h = hm >> 2, el = hm & 3;
    // which chunks of the stack are materialised: one page-table load per thread up front (inside the loop below it was a
    // dependent global load per chunk, 18 memory latencies in a row)
    __shared__ uint32_t s_mat[32];
    if (threadIdx.x < 32) s_mat[threadIdx.x] = 0u;
    __syncthreads();
    for (int sy = sy_lo + (int)threadIdx.x; sy < sy_hi; sy += 64) {
        const uint32_t c = (uint32_t)(sx + p.SX * (sz + p.SZ * sy));
        if (!p.page || p.page[c] == c) atomicOr(&s_mat[((sy - sy_lo) >> 5) & 31], 1u << ((sy - sy_lo) & 31));
    }
    __syncthreads();
    for (int sy = sy_lo; sy < sy_hi; sy++) {
        const uint32_t c = (uint32_t)(sx + p.SX * (sz + p.SZ * sy));
        if (sy - sy_lo < 1024 ? !((s_mat[(sy - sy_lo) >> 5] >> ((sy - sy_lo) & 31)) & 1u) : (p.page && p.page[c] != c))
            continue; // not materialised: the shared page carries its planes
        uint32_t *dst = p.xface + (size_t)c * 2048 + side * 1024 + z;
        const int ybase = (p.cy_min + sy) * 32;
        for (int yc = 0; yc < 32; yc++) {
            const int y = ybase + yc;
            uint32_t v;
            if (y == h) v = (el == 2 && y > 80) ? p.id_snow : p.id_grass;
            else if (y < h - 5) v = p.id_stone;
            else if (y < h) v = p.id_dirt;
            else v = p.id_void;
            __stcs(&dst[yc * 32], v);
        }
    }
}

// raw SuperSimplex::get samples of one of the generator's seven noise functions (parity tests / the FP32 report)
template <int PREC>
__global__ void noise_samples_kernel(const GenTables *T, int which, const double *xs, const double *ys, size_t n, double *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = noise_at<PREC>(TabsGlobal{T}, which, xs[i], ys[i]);
}

} // namespace

void fill_noise_constants(GenTables &t) {
    const double d = kDiag;
    const double g[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {d, d}, {-d, d}, {d, -d}, {-d, -d}};
    for (int i = 0; i < 8; i++) t.grad[i][0] = g[i][0], t.grad[i][1] = g[i][1];
    for (int grp = 0; grp < 8; grp++)
        for (int k = 0; k < 2; k++) {
            LatticeEntry &e = t.lat[grp][k];
            e.ox = h_lat_off[grp * 4 + 2 + k][0];
            e.oy = h_lat_off[grp * 4 + 2 + k][1];
            e.px = h_lat_pt[grp * 4 + 2 + k][0];
            e.py = h_lat_pt[grp * 4 + 2 + k][1];
            e.pad[0] = e.pad[1] = 0;
        }
}

void launch_cellpoints_kernel(uint64_t seed, int32_t cell_x0, int32_t cell_z0, int32_t cells_w, int32_t cells_d,
                              const GenTables *tables, CellPointDev *cells, int precision, cudaStream_t stream) {
    const int n = cells_w * cells_d;
    if (precision == 32)
        cellpoints_kernel<32><<<(n + 127) / 128, 128, 0, stream>>>(seed, cell_x0, cell_z0, cells_w, cells_d, tables, cells);
    else
        cellpoints_kernel<64><<<(n + 127) / 128, 128, 0, stream>>>(seed, cell_x0, cell_z0, cells_w, cells_d, tables, cells);
}

void launch_noise_samples(const GenTables *tables, int which, const double *xs, const double *ys, size_t n, double *out, int precision,
                          cudaStream_t stream) {
    if (!n) return;
    const unsigned nb = (unsigned)((n + 127) / 128);
    if (precision == 32) noise_samples_kernel<32><<<nb, 128, 0, stream>>>(tables, which, xs, ys, n, out);
    else noise_samples_kernel<64><<<nb, 128, 0, stream>>>(tables, which, xs, ys, n, out);
}

void launch_heightmap_kernel(const GenParams &p, int precision, cudaStream_t stream) {
    dim3 grid((p.W + 127) / 128, p.D);
    if (precision == 32) heightmap_kernel<32><<<grid, 128, 0, stream>>>(p);
    else heightmap_kernel<64><<<grid, 128, 0, stream>>>(p);
}

static dim3 fill_grid(const FillParams &p) {
    if (p.list) return dim3(p.n_list);
    const int nx = p.SX - 2;
    const int cols = p.col_mode == 1 ? (nx == 1 ? 1 : 2) : (p.col_mode == 2 ? p.SX - (nx == 1 ? 1 : 2) : p.SX);
    return dim3(cols, p.SZ);
}

void launch_fill_kernel(const FillParams &p, cudaStream_t stream) {
    if (p.list && !p.n_list) return;
    fill_kernel<0><<<fill_grid(p), 256, 0, stream>>>(p, GenParams{});
}

void launch_gen_fill_kernel(const FillParams &p, const GenParams &g, int precision, cudaStream_t stream) {
    if (p.list && !p.n_list) return;
    if (precision == 32) fill_kernel<32><<<fill_grid(p), 256, 0, stream>>>(p, g);
    else fill_kernel<64><<<fill_grid(p), 256, 0, stream>>>(p, g);
}

void launch_fill_xface_kernel(const FillParams &p, cudaStream_t stream) {
    if (p.list && !p.n_list) return;
    if (p.xface) fill_xface_kernel<<<fill_grid(p), 64, 0, stream>>>(p);
}

// the shared single-datum pages behind the storage (blocks + x-face planes)
__global__ void uniform_page_kernel(uint32_t *blocks, uint32_t *xface, uint32_t slot, uint32_t datum) {
    const uint4 v = make_uint4(datum, datum, datum, datum);
    uint4 *b = reinterpret_cast<uint4 *>(blocks + (size_t)slot * kChunkVoxels);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kChunkVoxels / 4; i += gridDim.x * blockDim.x) b[i] = v;
    uint4 *x = reinterpret_cast<uint4 *>(xface + (size_t)slot * 2048);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 512; i += gridDim.x * blockDim.x) x[i] = v;
}

void launch_uniform_page(uint32_t *blocks, uint32_t *xface, uint32_t slot, uint32_t datum, cudaStream_t stream) {
    uniform_page_kernel<<<8, 256, 0, stream>>>(blocks, xface, slot, datum);
}

} // namespace bxw
