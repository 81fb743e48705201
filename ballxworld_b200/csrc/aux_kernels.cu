// aux_kernels.cu — the small kernels around the two hot ones: slab-boundary halo pack/unpack (multi-GPU), voxel
// edits + dirty tracking (World::apply_voxel_changes), the VChunk RLE codec, and bxw_terragen's SuperSimplex.
#include "aux_kernels.cuh"

namespace bxw {
namespace {

// ------------------------------------------------------------------------------------------------------------
// Halo planes. One x = const voxel plane over the whole storage extent in y and z (shell rows included so that
// edges and corners travel too: AO samples diagonal neighbours, stdshapes.rs:131-152). Layout [gy][gz].
// ------------------------------------------------------------------------------------------------------------
__global__ void halo_plane_kernel(uint32_t *blocks, uint32_t *xface, unsigned long long *summary, const uint32_t *page, int SX, int SY,
                                  int SZ, int sx, int lx, uint32_t *plane, int import) {
    const int gz = blockIdx.x * blockDim.x + threadIdx.x;
    const int gy = blockIdx.y;
    if (gz >= SZ * 32) return;
    const int sy = gy >> 5, ly = gy & 31, sz = gz >> 5, lz = gz & 31;
    const size_t c = (size_t)(sx + SX * (sz + SZ * sy));
    uint32_t *p = plane + (size_t)gy * (SZ * 32) + gz;
    if (import) {
        // a chunk that is still not materialised was checked by halo_check_kernel: the plane equals its uniform datum
        if (page && page[c] != (uint32_t)c) return;
        blocks[c * kChunkVoxels + lx + 32 * lz + 1024 * ly] = *p;
        xface[c * 2048 + (lx ? 1024 : 0) + ly * 32 + lz] = *p; // lx is 0 or 31 here
        if (summary) { // the chunk stays known-uniform only if every imported voxel equals its uniform datum
            const unsigned long long sm = summary[c];
            if (sm && (!(sm & kUniformBit) || (uint32_t)sm != *p)) summary[c] = 0; // (also drops the solid-terrain bits)
        }
    } else {
        // the boundary plane is one of the chunk's x-face planes: read it from there (128-byte rows) instead of one voxel per
        // 128-byte row of the block storage
        const size_t g = page ? page[c] : c;
        *p = xface[g * 2048 + (lx ? 1024 : 0) + ly * 32 + lz];
    }
}

// import, step 1: which not-materialised chunks of the shell column receive a voxel that differs from their uniform datum?
__global__ void halo_check_kernel(const unsigned long long *summary, const uint32_t *page, int SX, int SY, int SZ, int sx,
                                  const uint32_t *plane, uint8_t *need) {
    const int gz = blockIdx.x * blockDim.x + threadIdx.x;
    const int gy = blockIdx.y;
    if (gz >= SZ * 32) return;
    const size_t c = (size_t)(sx + SX * ((gz >> 5) + SZ * (gy >> 5)));
    if (page[c] != (uint32_t)c && plane[(size_t)gy * (SZ * 32) + gz] != (uint32_t)summary[c]) need[c] = 1;
}

// Copies the shared uniform page of every listed / flagged chunk into the chunk's own storage and points its page entry back
// at itself. chunks == nullptr: the (sy, sz) chunks of storage column `sx` whose need[] flag is set (halo import).
__global__ void __launch_bounds__(256) materialise_kernel(uint32_t *blocks, uint32_t *xface, uint32_t *page, const uint32_t *chunks,
                                                          uint8_t *need, int SX, int sx) {
    const size_t c = chunks ? (size_t)chunks[blockIdx.x] : (size_t)sx + (size_t)SX * blockIdx.x;
    if (need) {
        if (!need[c]) return;
    }
    const uint32_t g = page[c];
    if (g != (uint32_t)c) {
        const uint4 *s = reinterpret_cast<const uint4 *>(blocks + (size_t)g * kChunkVoxels);
        uint4 *d = reinterpret_cast<uint4 *>(blocks + c * kChunkVoxels);
        for (int i = threadIdx.x; i < kChunkVoxels / 4; i += 256) d[i] = s[i];
        const uint4 *sxf = reinterpret_cast<const uint4 *>(xface + (size_t)g * 2048);
        uint4 *dxf = reinterpret_cast<uint4 *>(xface + c * 2048);
        for (int i = threadIdx.x; i < 512; i += 256) dxf[i] = sxf[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        page[c] = (uint32_t)c;
        if (need) need[c] = 0;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Edits. The host groups the changes by voxel (stable), one thread replays one voxel's changes in order:
// `if chunk.blocks_yzx[bidx] == change.from { = change.to }` (generation.rs:52-57). Every change marks
// bpos.touching_chunks() dirty whether or not it applied (worldmgr.rs:329-334, lib.rs:110-135).
// ------------------------------------------------------------------------------------------------------------
__global__ void apply_changes_kernel(EditParams p) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_runs) return;
    const uint32_t b = p.run_start[r], e = p.run_start[r + 1];
    const bxw_voxel_change c0 = p.changes[b];
    // storage-local voxel coordinates (the host dropped changes outside the storage box)
    const int gx = c0.x - p.x_min, gy = c0.y - p.y_min, gz = c0.z - p.z_min;
    const int sx = gx >> 5, sy = gy >> 5, sz = gz >> 5, lx = gx & 31, ly = gy & 31, lz = gz & 31;
    // load-anchor mode: a chunk without block data is skipped (worldmgr.rs:321-323 `get_chunk_index(cpos).is_none() -> continue`);
    // the dirty marks below are set all the same, as in the reference
    const bool present = !p.blk_loaded || p.blk_loaded[(size_t)(sx + p.SX * (sz + p.SZ * sy))];
    if (present) {
        uint32_t *v = p.blocks + ((size_t)(sx + p.SX * (sz + p.SZ * sy))) * kChunkVoxels + lx + 32 * lz + 1024 * ly;
        const uint32_t before = *v;
        uint32_t cur = before;
        for (uint32_t k = b; k < e; k++) {
            const bxw_voxel_change c = p.changes[k];
            if (cur == c.from) cur = c.to;
        }
        *v = cur;
        if (cur != before && p.summary) p.summary[(size_t)(sx + p.SX * (sz + p.SZ * sy))] = 0; // no longer known uniform
        if (lx == 0 || lx == 31) p.xface[((size_t)(sx + p.SX * (sz + p.SZ * sy))) * 2048 + (lx ? 1024 : 0) + ly * 32 + lz] = cur;
    }
    auto mark = [&](int cx, int cy, int cz) { // interior chunks only carry meshes
        if (cx >= 1 && cx <= p.SX - 2 && cy >= 1 && cy <= p.SY - 2 && cz >= 1 && cz <= p.SZ - 2)
            p.dirty[(cx - 1) + (p.SX - 2) * ((cz - 1) + (p.SZ - 2) * (cy - 1))] = 1;
    };
    mark(sx, sy, sz);
    if (lx == 0) mark(sx - 1, sy, sz);
    if (lx == 31) mark(sx + 1, sy, sz);
    if (ly == 0) mark(sx, sy - 1, sz);
    if (ly == 31) mark(sx, sy + 1, sz);
    if (lz == 0) mark(sx, sy, sz - 1);
    if (lz == 31) mark(sx, sy, sz + 1);
}

__global__ void dirty_list_kernel(uint8_t *dirty, const uint8_t *mesh_loaded, uint32_t n, int nx, int nz, int SX, int SZ, uint32_t *work,
                                  uint32_t *work_span, uint32_t *count, int clear) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !dirty[i]) return;
    if (clear) dirty[i] = 0;
    // a chunk whose mesh is not loaded gets no update task (worldmgr.rs:346-348: status Unloaded -> continue); it is meshed
    // from the then-current blocks when a load anchor asks for it
    if (mesh_loaded && !mesh_loaded[i]) return;
    const uint32_t k = atomicAdd(count, 1u);
    const uint32_t ix = i % nx, iz = (i / nx) % nz, iy = i / (nx * nz);
    work[k] = (ix + 1) + SX * ((iz + 1) + SZ * (iy + 1));
    work_span[k] = i;
}

// ------------------------------------------------------------------------------------------------------------
// VChunk RLE (bxw_world/src/lib.rs:263-345). compress_rle emits, for every maximal run of a value v of length L,
// [v] if L == 1 and [v, v, L-2] otherwise (a count word follows every equal pair; the state machine at
// lib.rs:268-291 never pairs across a count). One CTA per chunk.
// ------------------------------------------------------------------------------------------------------------
constexpr int RT = 256;            // threads per CTA
constexpr int RSEG = 32768 / RT;   // 128 contiguous voxels per thread

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_tot, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t a = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += a;
    }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (int k = 0; k < RT / 32; k++) {
        uint32_t t = warp_tot[k];
        if (k < warp) base += t;
        tot += t;
    }
    total = tot;
    return base + incl - v;
}

// Each thread owns the run HEADS inside its 128-voxel segment. A run's length is the distance to the next head, which may
// lie in a later segment: the first head of every segment goes to shared memory and a suffix-minimum gives each thread
// the first head after its segment, so no thread ever walks a long run voxel by voxel.
template <bool WRITE>
__global__ void __launch_bounds__(RT) rle_compress_kernel(const uint32_t *dense, const uint32_t *ids, const uint32_t *page, size_t n_chunks,
                                                          const uint64_t *offsets, uint32_t *out_words, uint32_t *out_len) {
    __shared__ uint32_t warp_tot[RT / 32];
    __shared__ uint32_t first_head[RT + 1];
    const size_t c = blockIdx.x;
    if (c >= n_chunks) return;
    const size_t cc = ids ? (size_t)ids[c] : c;
    const uint32_t *d = dense + (page ? (size_t)page[cc] : cc) * kChunkVoxels; // a chunk that is not materialised reads its uniform page
    const int t = threadIdx.x;
    const int b = t * RSEG;
    // pass 1: heads of this segment as a 128-bit mask; words they contribute
    uint32_t hm[RSEG / 32];
    uint32_t prev = b ? d[b - 1] : 0;
    uint32_t words = 0;
#pragma unroll
    for (int w = 0; w < RSEG / 32; w++) {
        uint32_t m = 0;
        for (int k = 0; k < 32; k++) {
            const int i = b + w * 32 + k;
            const uint32_t v = d[i];
            const bool head = (i == 0) || (v != prev);
            m |= (uint32_t)head << k;
            prev = v;
        }
        hm[w] = m;
    }
    int fh = kChunkVoxels;
#pragma unroll
    for (int w = RSEG / 32 - 1; w >= 0; w--)
        if (hm[w]) fh = b + w * 32 + __ffs(hm[w]) - 1;
    first_head[t] = (uint32_t)fh;
    if (t == 0) first_head[RT] = kChunkVoxels;
    __syncthreads();
    // first head strictly after this segment (suffix minimum; 256 entries, a short serial loop per thread is cheap
    // because it stops at the first segment that has a head, which for run-heavy data is far but for those chunks
    // there are few heads overall)
    uint32_t next_after = kChunkVoxels;
    for (int k = t + 1; k < RT; k++) {
        if (first_head[k] < (uint32_t)kChunkVoxels) {
            next_after = first_head[k];
            break;
        }
    }
    // run length of every head: distance to the next head (in this segment, else next_after)
    auto for_each_head = [&](auto &&fn) {
#pragma unroll
        for (int w = 0; w < RSEG / 32; w++) {
            uint32_t m = hm[w];
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const int i = b + w * 32 + k;
                // next head after i
                uint32_t nx = next_after;
                uint32_t rest = m;
                if (rest) nx = b + w * 32 + __ffs(rest) - 1;
                else
                    for (int w2 = w + 1; w2 < RSEG / 32; w2++)
                        if (hm[w2]) {
                            nx = b + w2 * 32 + __ffs(hm[w2]) - 1;
                            break;
                        }
                fn(i, nx - (uint32_t)i);
            }
        }
    };
    for_each_head([&](int, uint32_t len) { words += len >= 2 ? 3u : 1u; });
    uint32_t total;
    uint32_t off = block_exclusive_scan(words, warp_tot, total);
    if (!WRITE) {
        if (t == 0) out_len[c] = total;
        return;
    }
    uint32_t *o = out_words + offsets[c];
    for_each_head([&](int i, uint32_t len) {
        const uint32_t v = d[i];
        o[off++] = v;
        if (len >= 2) { // [v, v, L-2] (lib.rs:263-294)
            o[off++] = v;
            o[off++] = len - 2;
        }
    });
}

__global__ void offsets_scan_kernel(const uint32_t *len, size_t n, uint64_t *offsets) {
    // single CTA: exclusive scan of per-chunk word counts into 64-bit offsets (n+1 entries)
    __shared__ uint64_t carry;
    __shared__ uint32_t warp_tot[RT / 32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < n; base += RT) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < n ? len[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, warp_tot, total);
        if (i < n) offsets[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = carry;
}

// decompress_rle (lib.rs:304-345). The parse state before word i is one of N (no previous value), V (previous word
// is a value that may pair) or C (a count is expected); the transition depends only on whether w[i] == w[i-1]:
//   C -> N,  N -> V,  V -> (w[i]==w[i-1] ? C : V).
// Each thread folds the transitions of its contiguous word segment for all three start states, the CTA scans the
// composed maps, and each thread then replays its segment from its true start state.
__device__ __forceinline__ int rle_step(int s, bool eq) { return s == 2 ? 0 : (s == 0 ? 1 : (eq ? 2 : 1)); }

__global__ void __launch_bounds__(RT) rle_decompress_kernel(const uint32_t *words, const uint64_t *offsets, size_t n_chunks,
                                                            uint32_t *out_dense, uint32_t *status) {
    __shared__ uint32_t warp_tot[RT / 32];
    __shared__ uint8_t seg_map[RT];   // 2 bits per start state -> end state
    __shared__ uint8_t seg_start[RT];
    __shared__ uint32_t n_long;
    __shared__ uint32_t long_runs[512][3]; // value, start, length of runs too long for one thread (32768 / 64 = 512 at most)
    __shared__ int bad;
    const size_t c = blockIdx.x;
    if (c >= n_chunks) return;
    const uint32_t *w = words + offsets[c];
    const uint64_t len64 = offsets[c + 1] - offsets[c];
    uint32_t *out = out_dense + c * kChunkVoxels;
    const int t = threadIdx.x;
    if (t == 0) {
        bad = 0;
        n_long = 0;
    }
    __syncthreads();
    if (len64 < 3 || len64 > 49152) { // lib.rs:305-307; an encoder never exceeds 1.5 words/voxel
        if (t == 0) status[c] = 4;
        return;
    }
    const uint32_t len = (uint32_t)len64;
    const uint32_t seg = (len + RT - 1) / RT;
    const uint32_t b = min(len, t * seg), e = min(len, b + seg);
    // fold
    int m[3] = {0, 1, 2};
    for (uint32_t i = b; i < e; i++) {
        const bool eq = i > 0 && w[i] == w[i - 1];
#pragma unroll
        for (int k = 0; k < 3; k++) m[k] = rle_step(m[k], eq);
    }
    seg_map[t] = (uint8_t)(m[0] | (m[1] << 2) | (m[2] << 4));
    __syncthreads();
    if (t == 0) {
        int s = 0; // decompress_rle starts with no previous value
        for (int k = 0; k < RT; k++) {
            seg_start[k] = (uint8_t)s;
            s = (seg_map[k] >> (2 * s)) & 3;
        }
        if (s == 2) bad = 3; // stream ends where a count is expected: MissingRleRepeatN
    }
    __syncthreads();
    // count voxels produced by this segment
    int s = seg_start[t];
    uint64_t vox = 0;
    for (uint32_t i = b; i < e; i++) {
        const bool eq = i > 0 && w[i] == w[i - 1];
        vox += (s == 2) ? (uint64_t)w[i] : 1ull;
        s = rle_step(s, eq);
    }
    if (vox > 65536) {
        vox = 65536; // clamp so the 32-bit scan cannot wrap; flagged below through the total
        bad = 2;
    }
    uint32_t total;
    uint32_t pos = block_exclusive_scan((uint32_t)vox, warp_tot, total);
    __syncthreads();
    if (bad || total != (uint32_t)kChunkVoxels) {
        if (t == 0) status[c] = bad ? (uint32_t)bad : (total > (uint32_t)kChunkVoxels ? 2u : 4u);
        return;
    }
    // write
    s = seg_start[t];
    for (uint32_t i = b; i < e; i++) {
        const bool eq = i > 0 && w[i] == w[i - 1];
        if (s == 2) {
            const uint32_t n = w[i], v = w[i - 1];
            if (n >= 64) {
                const uint32_t q = atomicAdd(&n_long, 1u);
                if (q < 512) {
                    long_runs[q][0] = v;
                    long_runs[q][1] = pos;
                    long_runs[q][2] = n;
                } else {
                    for (uint32_t k = 0; k < n; k++) out[pos + k] = v; // unreachable: a valid stream holds at most 32768 / 66 long runs
                }
            } else {
                for (uint32_t k = 0; k < n; k++) out[pos + k] = v;
            }
            pos += n;
        } else {
            out[pos++] = w[i];
        }
        s = rle_step(s, eq);
    }
    __syncthreads();
    const uint32_t nl = min(n_long, 512u);
    for (uint32_t q = 0; q < nl; q++) {
        const uint32_t v = long_runs[q][0], st = long_runs[q][1], n = long_runs[q][2];
        for (uint32_t k = t; k < n; k += RT) out[st + k] = v;
    }
    if (t == 0) status[c] = 0;
}

// ------------------------------------------------------------------------------------------------------------
// bxw_terragen::SuperSimplex::noise2 (bxw_terragen/src/supersimplex.rs:53-107), one point per thread.
// ------------------------------------------------------------------------------------------------------------
__global__ void terragen_noise2_kernel(const TerragenTables *T, const double *xs, const double *ys, size_t n, double *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = xs[i], y = ys[i];
    const double s = __dmul_rn(__dadd_rn(x, y), 0.366025403784439);
    const double xs_ = __dadd_rn(x, s), ys_ = __dadd_rn(y, s);
    // widefloor_f64_i32 (math.rs:68-78): truncating cast, minus one where x < trunc(x)
    int32_t xsb = __double2int_rz(xs_), ysb = __double2int_rz(ys_);
    if (xs_ < (double)xsb) xsb -= 1;
    if (ys_ < (double)ysb) ysb -= 1;
    const double xsi = __dsub_rn(xs_, (double)xsb), ysi = __dsub_rn(ys_, (double)ysb);
    const int32_t a = __double2int_rz(__dadd_rn(xsi, ysi));
    const double af = (double)a;
    const double half_a = __ddiv_rn(af, 2.0);
    const int32_t index = (a << 2) |
                          (__double2int_rz(__dsub_rn(__dadd_rn(__dsub_rn(xsi, __ddiv_rn(ysi, 2.0)), 1.0), half_a)) << 3) |
                          (__double2int_rz(__dsub_rn(__dadd_rn(__dsub_rn(ysi, __ddiv_rn(xsi, 2.0)), 1.0), half_a)) << 4);
    const double ssi = __dmul_rn(__dadd_rn(xsi, ysi), -0.211324865405187);
    const double xi = __dadd_rn(xsi, ssi), yi = __dadd_rn(ysi, ssi);
    double value = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int li = (index + k) & 31;
        const double dx = __dadd_rn(xi, T->lat_dx[li]), dy = __dadd_rn(yi, T->lat_dy[li]);
        const double attn = __dsub_rn(__dsub_rn(2.0 / 3.0, __dmul_rn(dx, dx)), __dmul_rn(dy, dy));
        if (attn > 0.0) {
            const unsigned pxm = (unsigned)((xsb + T->lat_x[li]) & 2047), pym = (unsigned)((ysb + T->lat_y[li]) & 2047);
            const unsigned gi = (unsigned)T->perm[pxm] ^ pym;
            const double ex = __dadd_rn(__dmul_rn(T->gx[gi], dx), __dmul_rn(T->gy[gi], dy));
            const double a2 = __dmul_rn(attn, attn);
            value = __dadd_rn(value, __dmul_rn(__dmul_rn(a2, a2), ex));
        }
    }
    out[i] = value;
}

// x=0 / x=31 planes of a list of chunks (after uploads / RLE decodes): thread per (side, y, z)
__global__ void xface_extract_kernel(const uint32_t *blocks, uint32_t *xface, const uint32_t *chunks, uint32_t first, uint32_t n) {
    const uint32_t k = blockIdx.x;
    if (k >= n) return;
    const size_t c = chunks ? chunks[k] : (size_t)first + k;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
        const int side = i >> 10, y = (i >> 5) & 31, z = i & 31;
        xface[c * 2048 + i] = blocks[c * kChunkVoxels + (side ? 31 : 0) + 32 * z + 1024 * y];
    }
}

// Synthetic input of BASELINE config 2: every voxel is a pure function of its global position.
__global__ void fill_synthetic_kernel(uint32_t *blocks, int SX, int SY, int SZ, int32_t cx_min, int32_t cy_min, int32_t cz_min,
                                      uint64_t seed, uint32_t void_threshold) {
    const size_t chunk = blockIdx.x;
    const int sx = (int)(chunk % SX), sz = (int)((chunk / SX) % SZ), sy = (int)(chunk / ((size_t)SX * SZ));
    for (int i = threadIdx.x; i < kChunkVoxels; i += blockDim.x) {
        const long long gx = (long long)(cx_min + sx) * 32 + (i & 31), gz = (long long)(cz_min + sz) * 32 + ((i >> 5) & 31),
                        gy = (long long)(cy_min + sy) * 32 + (i >> 10);
        const uint64_t k = ((uint64_t)gx & 0x1FFFFF) | (((uint64_t)gy & 0x1FFFFF) << 21) | (((uint64_t)gz & 0x1FFFFF) << 42);
        uint64_t z = (seed ^ k) + 0x9e3779b97f4a7c15ULL; // splitmix64 (bxw_terragen/src/math.rs:24-30)
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        const uint64_t r = z ^ (z >> 31);
        uint32_t datum = 0;
        if ((r & 0xFF) >= void_threshold) {
            const uint32_t id = 1u + (uint32_t)((r >> 8) % 7);
            const uint32_t s = (uint32_t)(r >> 16) & 15u;
            const uint32_t shape = s < 10 ? 0u : (s < 12 ? 1u : (s < 14 ? 2u : 3u));
            const uint32_t orient = (uint32_t)((r >> 20) % 24);
            datum = (id << 16) | (orient * 64u + shape);
        }
        blocks[chunk * kChunkVoxels + i] = datum;
    }
}

} // namespace

void launch_xface_extract(const uint32_t *blocks, uint32_t *xface, const uint32_t *chunks, uint32_t first, uint32_t n, cudaStream_t stream) {
    if (n) xface_extract_kernel<<<n, 256, 0, stream>>>(blocks, xface, chunks, first, n);
}

void launch_fill_synthetic(uint32_t *blocks, int SX, int SY, int SZ, int32_t cx_min, int32_t cy_min, int32_t cz_min, uint64_t seed,
                           uint32_t void_threshold, cudaStream_t stream) {
    fill_synthetic_kernel<<<(unsigned)((size_t)SX * SY * SZ), 256, 0, stream>>>(blocks, SX, SY, SZ, cx_min, cy_min, cz_min, seed, void_threshold);
}

void launch_halo_plane(uint32_t *blocks, uint32_t *xface, unsigned long long *summary, uint32_t *page, uint8_t *need, int SX, int SY, int SZ,
                       int sx, int lx, uint32_t *plane, int import, cudaStream_t stream) {
    dim3 grid((SZ * 32 + 127) / 128, SY * 32);
    if (import && page) {
        // chunks of the shell column that are not materialised stay that way unless the plane brings a different voxel
        halo_check_kernel<<<grid, 128, 0, stream>>>(summary, page, SX, SY, SZ, sx, plane, need);
        materialise_kernel<<<(unsigned)(SY * SZ), 256, 0, stream>>>(blocks, xface, page, nullptr, need, SX, sx);
    }
    halo_plane_kernel<<<grid, 128, 0, stream>>>(blocks, xface, summary, page, SX, SY, SZ, sx, lx, plane, import);
}

void launch_materialise(uint32_t *blocks, uint32_t *xface, uint32_t *page, const uint32_t *chunks, uint32_t n, cudaStream_t stream) {
    if (n) materialise_kernel<<<n, 256, 0, stream>>>(blocks, xface, page, chunks, nullptr, 0, 0);
}

void launch_apply_changes(const EditParams &p, cudaStream_t stream) {
    if (p.n_runs == 0) return;
    apply_changes_kernel<<<(p.n_runs + 127) / 128, 128, 0, stream>>>(p);
}

void launch_dirty_list(uint8_t *dirty, const uint8_t *mesh_loaded, uint32_t n, int nx, int nz, int SX, int SZ, uint32_t *work,
                       uint32_t *work_span, uint32_t *count, int clear, cudaStream_t stream) {
    dirty_list_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dirty, mesh_loaded, n, nx, nz, SX, SZ, work, work_span, count, clear);
}

void launch_rle_compress(const uint32_t *dense, const uint32_t *ids, const uint32_t *page, size_t n_chunks, uint32_t *len, uint64_t *offsets,
                         uint32_t *out_words, int phase, cudaStream_t stream) {
    if (phase == 0) {
        rle_compress_kernel<false><<<(unsigned)n_chunks, RT, 0, stream>>>(dense, ids, page, n_chunks, nullptr, nullptr, len);
        offsets_scan_kernel<<<1, RT, 0, stream>>>(len, n_chunks, offsets);
    } else {
        rle_compress_kernel<true><<<(unsigned)n_chunks, RT, 0, stream>>>(dense, ids, page, n_chunks, offsets, out_words, nullptr);
    }
}

// Copy n decoded chunks into region storage at chunk indices ids[], refresh their x-face planes and mark every interior
// chunk within one chunk of them dirty (their meshes read the new voxels).
__global__ void chunk_scatter_kernel(const uint32_t *src, const uint32_t *ids, uint32_t *blocks, uint32_t *xface,
                                     unsigned long long *summary, uint32_t *page, uint8_t *blk_loaded, int SX, int SY, int SZ, uint8_t *dirty) {
    const size_t k = blockIdx.x;
    const size_t c = ids[k];
    if (blk_loaded && threadIdx.x == 0) blk_loaded[c] = 1;
    if (summary && threadIdx.x == 0) summary[c] = 0; // no longer known uniform
    if (page && threadIdx.x == 0) page[c] = (uint32_t)c; // every voxel is overwritten below: materialised
    const uint4 *s = reinterpret_cast<const uint4 *>(src + k * kChunkVoxels);
    uint4 *d = reinterpret_cast<uint4 *>(blocks + c * kChunkVoxels);
    for (int i = threadIdx.x; i < kChunkVoxels / 4; i += blockDim.x) d[i] = s[i];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
        const int side = i >> 10, y = (i >> 5) & 31, z = i & 31;
        xface[c * 2048 + i] = src[k * kChunkVoxels + (side ? 31 : 0) + 32 * z + 1024 * y];
    }
    if (dirty && threadIdx.x < 27) {
        const int sx = (int)(c % SX), sz = (int)((c / SX) % SZ), sy = (int)(c / ((size_t)SX * SZ));
        const int t = threadIdx.x;
        const int cx = sx + t % 3 - 1, cz = sz + (t / 3) % 3 - 1, cy = sy + t / 9 - 1;
        if (cx >= 1 && cx <= SX - 2 && cy >= 1 && cy <= SY - 2 && cz >= 1 && cz <= SZ - 2)
            dirty[(cx - 1) + (SX - 2) * ((cz - 1) + (SZ - 2) * (cy - 1))] = 1;
    }
}

void launch_chunk_scatter(const uint32_t *src, const uint32_t *ids, size_t n, uint32_t *blocks, uint32_t *xface,
                          unsigned long long *summary, uint32_t *page, uint8_t *blk_loaded, int SX, int SY, int SZ, uint8_t *dirty,
                          cudaStream_t stream) {
    if (n) chunk_scatter_kernel<<<(unsigned)n, 256, 0, stream>>>(src, ids, blocks, xface, summary, page, blk_loaded, SX, SY, SZ, dirty);
}

// page[c] = c for c in [0, n): every chunk lives in its own storage
__global__ void page_identity_kernel(uint32_t *page, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) page[i] = i;
}
void launch_page_identity(uint32_t *page, uint32_t n, cudaStream_t stream) {
    if (n) page_identity_kernel<<<(n + 255) / 256, 256, 0, stream>>>(page, n);
}

void launch_rle_decompress(const uint32_t *words, const uint64_t *offsets, size_t n_chunks, uint32_t *out_dense, uint32_t *status,
                           cudaStream_t stream) {
    rle_decompress_kernel<<<(unsigned)n_chunks, RT, 0, stream>>>(words, offsets, n_chunks, out_dense, status);
}

// ------------------------------------------------------------------------------------------------------------
// Candidate list of a whole-interior mesh pass over storage (see StorageCandParams). Three small launches: keep flags +
// per-block counts, a one-CTA scan of the counts, ordered compaction. Work index w enumerates the interior in the mesher's
// tiled order (x-tiles of 8 columns; z-rows, columns, vertical stack fastest) so neighbouring chunks stay neighbours in time.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tiled_coords(uint32_t w, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t &ix, uint32_t &iy, uint32_t &iz) {
    const uint32_t per_tile = 8u * ny * nz;
    const uint32_t t = w / per_tile, rem = w - t * per_tile;
    const uint32_t tw = min(8u, nx - 8u * t);
    iz = rem / (tw * ny);
    const uint32_t r2 = rem - iz * tw * ny;
    ix = 8u * t + r2 / ny;
    iy = r2 % ny;
}

__global__ void __launch_bounds__(256) cand_flags_kernel(StorageCandParams p, uint32_t n) {
    const uint32_t w = blockIdx.x * 256u + threadIdx.x;
    int keep = 0;
    if (w < n) {
        const uint32_t nx = (uint32_t)p.SX - 2u, ny = (uint32_t)p.SY - 2u, nz = (uint32_t)p.SZ - 2u;
        uint32_t ix, iy, iz;
        tiled_coords(w, nx, ny, nz, ix, iy, iz);
        const size_t c0 = (ix + 1u) + (size_t)p.SX * ((iz + 1u) + (size_t)p.SZ * (iy + 1u));
        const unsigned long long sc = p.summary[c0];
        keep = 1;
        if (sc & kUniformBit) {
            const uint32_t d = (uint32_t)sc, id = d >> 16;
            const uint32_t kind = id < p.reg.n ? p.reg.kind[id] : 0u; // 0 absent (the mesher reports it), 1 no mesh, 2 meshed
            const uint32_t shape = d & 63u;
            if (kind == 1u) {
                keep = 0; // is_chunk_trivial (voxmesh.rs:16-25): nothing in this chunk emits a side
            } else if (kind == 2u && (shape == 0u || shape > 3u)) {
                // a chunk of cubes: empty mesh iff every neighbour is the same uniform datum (voxmesh.rs:119)
                int same = 1;
                for (int t = 0; t < 27; t++) {
                    const int rx = t % 3 - 1, rz = (t / 3) % 3 - 1, ry = t / 9 - 1;
                    const unsigned long long sn = p.summary[(size_t)((long long)c0 + rx + (long long)p.SX * (rz + (long long)p.SZ * ry))];
                    same &= (sn == sc);
                }
                keep = !same;
            }
        }
        if (keep && p.solids_are_cubes && ((sc >> kSolidShift) & 1ull)) {
            // solid terrain throughout: no side is visible iff each of the six neighbours shows a solid plane
            auto plane = [&](long long off, int d) { return (int)((p.summary[(size_t)((long long)c0 + off)] >> (kSolidShift + 1 + d)) & 1ull); };
            const long long sxs = 1, szs = p.SX, sys = (long long)p.SX * p.SZ;
            keep = !(plane(-sxs, 1) & plane(sxs, 0) & plane(-sys, 3) & plane(sys, 2) & plane(-szs, 5) & plane(szs, 4));
        }
        p.keep[w] = (uint8_t)keep;
    }
    const int total = __syncthreads_count(keep);
    if (threadIdx.x == 0) p.blk[blockIdx.x] = (uint32_t)total;
}

__global__ void __launch_bounds__(1024) cand_scan_kernel(uint32_t *blk, uint32_t n_blocks, uint32_t *count) {
    // exclusive scan of the per-block counts in place (one CTA; n_blocks is n_work / 256)
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry, pass_total;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? blk[i] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31u) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t wt = warp_tot[lane];
            uint32_t wi = wt;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (lane >= (uint32_t)o) wi += t;
            }
            warp_tot[lane] = wi - wt; // exclusive warp offsets
            if (lane == 31u) pass_total = wi;
        }
        __syncthreads();
        if (i < n_blocks) blk[i] = carry + warp_tot[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += pass_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

__global__ void __launch_bounds__(256) cand_write_kernel(StorageCandParams p, uint32_t n) {
    __shared__ uint32_t warp_tot[8];
    const uint32_t w = blockIdx.x * 256u + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int keep = w < n ? p.keep[w] : 0;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    if (w >= n) return;
    uint32_t off = p.blk[blockIdx.x] + __popc(bal & ((1u << lane) - 1u));
    for (uint32_t k = 0; k < warp; k++) off += warp_tot[k];
    const uint32_t nx = (uint32_t)p.SX - 2u, ny = (uint32_t)p.SY - 2u, nz = (uint32_t)p.SZ - 2u;
    uint32_t ix, iy, iz;
    tiled_coords(w, nx, ny, nz, ix, iy, iz);
    const uint32_t slot = ix + nx * (iz + nz * iy);
    if (keep) {
        p.work[off] = (ix + 1u) + (uint32_t)p.SX * ((iz + 1u) + (uint32_t)p.SZ * (iy + 1u));
        p.work_span[off] = slot;
    } else {
        bxw_mesh_span sp;
        sp.v_off = sp.i_off = 0;
        sp.v_count = sp.i_count = 0;
        p.spans[slot] = sp;
        if (p.span_cap) p.span_cap[slot] = make_uint2(0u, 0u);
    }
}

void launch_storage_candidates(const StorageCandParams &p, cudaStream_t stream) {
    const uint32_t n = (uint32_t)((p.SX - 2) * (p.SY - 2) * (p.SZ - 2));
    const uint32_t nb = (n + 255u) / 256u;
    cand_flags_kernel<<<nb, 256, 0, stream>>>(p, n);
    cand_scan_kernel<<<1, 1024, 0, stream>>>(p.blk, nb, p.count);
    cand_write_kernel<<<nb, 256, 0, stream>>>(p, n);
}

// ------------------------------------------------------------------------------------------------------------
// Arena compaction: every live span moves to the prefix sum of the reservations before it (slot order), into a second arena.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) compact_offsets_kernel(const bxw_mesh_span *spans, uint2 *span_cap, uint32_t n,
                                                               unsigned long long *new_off, MeshCounters *counters) {
    // one CTA: exclusive scan of the (tightened) reservations; new_off[2 s] / [2 s + 1] = new vertex / index offset of slot s
    __shared__ unsigned long long wv[32], wi[32];
    __shared__ unsigned long long carry_v, carry_i;
    if (threadIdx.x == 0) carry_v = carry_i = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += 1024u) {
        const uint32_t s = base + threadIdx.x;
        unsigned long long cv = 0, ci = 0;
        if (s < n && (spans[s].v_count | spans[s].i_count)) {
            // keep the slack of re-meshed chunks, drop reservations of empty meshes
            cv = span_cap[s].x;
            ci = span_cap[s].y;
        }
        unsigned long long iv = cv, ii = ci;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long a = __shfl_up_sync(0xFFFFFFFFu, iv, o), b = __shfl_up_sync(0xFFFFFFFFu, ii, o);
            if (lane >= (uint32_t)o) {
                iv += a;
                ii += b;
            }
        }
        if (lane == 31u) {
            wv[warp] = iv;
            wi[warp] = ii;
        }
        __syncthreads();
        unsigned long long bv = carry_v, bi = carry_i;
        for (uint32_t k = 0; k < warp; k++) {
            bv += wv[k];
            bi += wi[k];
        }
        if (s < n) {
            new_off[2 * (size_t)s] = bv + iv - cv;
            new_off[2 * (size_t)s + 1] = bi + ii - ci;
            if (!(cv | ci)) span_cap[s] = make_uint2(0u, 0u);
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry_v = bv + iv;
            carry_i = bi + ii;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counters->v_alloc = carry_v;
        counters->i_alloc = carry_i;
        counters->v_garbage = 0;
        counters->i_garbage = 0;
    }
}

__global__ void __launch_bounds__(256) compact_copy_kernel(bxw_mesh_span *spans, const unsigned long long *new_off, const bxw_vertex *v_old,
                                                           const uint32_t *i_old, bxw_vertex *v_new, uint32_t *i_new) {
    const uint32_t s = blockIdx.x;
    const bxw_mesh_span sp = spans[s];
    if (!(sp.v_count | sp.i_count)) return;
    const unsigned long long nv = new_off[2 * (size_t)s], ni = new_off[2 * (size_t)s + 1];
    // spans start on multiples of 4 vertices (160 B) / 8 indices (32 B): 16-byte units on both sides
    const uint4 *sv = reinterpret_cast<const uint4 *>(v_old + sp.v_off);
    uint4 *dv = reinterpret_cast<uint4 *>(v_new + nv);
    const uint32_t n16v = (uint32_t)(((size_t)sp.v_count * 40 + 15) / 16);
    for (uint32_t k = threadIdx.x; k < n16v; k += 256) dv[k] = sv[k];
    const uint4 *si = reinterpret_cast<const uint4 *>(i_old + sp.i_off);
    uint4 *di = reinterpret_cast<uint4 *>(i_new + ni);
    const uint32_t n16i = (sp.i_count + 3u) / 4u;
    for (uint32_t k = threadIdx.x; k < n16i; k += 256) di[k] = si[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        spans[s].v_off = nv;
        spans[s].i_off = ni;
    }
}

void launch_compact_offsets(const bxw_mesh_span *spans, uint2 *span_cap, uint32_t n, unsigned long long *new_off, MeshCounters *counters,
                            cudaStream_t stream) {
    compact_offsets_kernel<<<1, 1024, 0, stream>>>(spans, span_cap, n, new_off, counters);
}
void launch_compact_copy(bxw_mesh_span *spans, const unsigned long long *new_off, uint32_t n, const bxw_vertex *v_old, const uint32_t *i_old,
                         bxw_vertex *v_new, uint32_t *i_new, cudaStream_t stream) {
    if (n) compact_copy_kernel<<<n, 256, 0, stream>>>(spans, new_off, v_old, i_old, v_new, i_new);
}

void launch_terragen_noise2(const TerragenTables *T, const double *xs, const double *ys, size_t n, double *out, cudaStream_t stream) {
    terragen_noise2_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(T, xs, ys, n, out);
}

void build_terragen_tables(uint64_t seed, TerragenTables &t) {
    // SuperSimplex::new (supersimplex.rs:20-43) and lookup_2d (:121-274)
    static const double grad2[24][2] = {
        {0.130526192220052, 0.99144486137381},    {0.38268343236509, 0.923879532511287},   {0.608761429008721, 0.793353340291235},
        {0.793353340291235, 0.608761429008721},   {0.923879532511287, 0.38268343236509},   {0.99144486137381, 0.130526192220051},
        {0.99144486137381, -0.130526192220051},   {0.923879532511287, -0.38268343236509},  {0.793353340291235, -0.60876142900872},
        {0.608761429008721, -0.793353340291235},  {0.38268343236509, -0.923879532511287},  {0.130526192220052, -0.99144486137381},
        {-0.130526192220052, -0.99144486137381},  {-0.38268343236509, -0.923879532511287}, {-0.608761429008721, -0.793353340291235},
        {-0.793353340291235, -0.608761429008721}, {-0.923879532511287, -0.38268343236509}, {-0.99144486137381, -0.130526192220052},
        {-0.99144486137381, 0.130526192220051},   {-0.923879532511287, 0.38268343236509},  {-0.793353340291235, 0.608761429008721},
        {-0.608761429008721, 0.793353340291235},  {-0.38268343236509, 0.923879532511287},  {-0.130526192220052, 0.99144486137381}};
    const double N2 = 0.05481866495625118;
    uint16_t source[2048];
    for (int i = 0; i < 2048; i++) source[i] = (uint16_t)i;
    for (int i = 2047; i >= 0; i--) {
        seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint64_t r = (seed + 31ULL) % ((uint64_t)i + 1ULL);
        t.perm[i] = source[r];
        t.gx[i] = grad2[t.perm[i] % 24][0] / N2;
        t.gy[i] = grad2[t.perm[i] % 24][1] / N2;
        source[r] = source[i];
    }
    for (int i = 0; i < 8; i++) {
        int i1, j1, i2, j2;
        if ((i & 1) == 0) {
            i1 = (i & 2) == 0 ? -1 : 1;
            j1 = 0;
            i2 = 0;
            j2 = (i & 4) == 0 ? -1 : 1;
        } else {
            i1 = (i & 2) != 0 ? 2 : 0;
            j1 = 1;
            i2 = 1;
            j2 = (i & 4) != 0 ? 2 : 0;
        }
        const int px[4] = {0, 1, i1, i2}, py[4] = {0, 1, j1, j2};
        for (int k = 0; k < 4; k++) {
            const double ssv = (double)(px[k] + py[k]) * -0.211324865405187;
            t.lat_x[i * 4 + k] = px[k];
            t.lat_y[i * 4 + k] = py[k];
            t.lat_dx[i * 4 + k] = (double)(-px[k]) - ssv;
            t.lat_dy[i * 4 + k] = (double)(-py[k]) - ssv;
        }
    }
}

} // namespace bxw
