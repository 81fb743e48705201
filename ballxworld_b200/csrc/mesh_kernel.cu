// mesh_kernel.cu — first half of the chunk mesher: everything that needs the 34^3 neighbourhood on chip.
//
// Semantics: src/client/render/voxmesh.rs:45-121 (mesh_from_chunk up to the culling decision, plus the emission ORDER) for
// every chunk of a work list, reading the GPU-resident dense block storage directly (the reference decodes 27 RLE streams
// into a 34^3 table first, voxmesh.rs:53-80; RleVoxelIterator::skip_until_index returns exactly blocks_yzx[idx]).
// Output: per chunk its arena span (vertex / index counts and offsets) and one face record per visible side in the
// reference's emission order; emit_kernel.cu turns the records into vertices and indices (voxmesh.rs:122-177).
//
// One persistent CTA (256 threads) takes one chunk at a time. Two instantiations (SmemT below): the general kernel with the
// class table (two CTAs per SM) and the cubes-only kernel without it (four per SM; generated terrain).
//   A  stage the chunk plus its 1-voxel halo (34^3) in shared memory as one 34-bit SOLID word per (y,z) row and, in the
//      general kernel, one CLASS BYTE per voxel (class = shape x orientation of the datum, 0 = no mesh). The rows stream
//      through warp-private cp.async rings (16-byte copies, a few units ahead); the x-halo comes from the neighbour chunks'
//      x-face planes. A unit (384 voxels) that is one single datum costs one match + one vote (terrain is mostly
//      uniform); a 34^3 neighbourhood that is one void/cube datum ends the chunk after one barrier. In height-map mode
//      (bxw_region_generate_and_mesh) the table is derived from the 34x34 column parameters instead of reading voxels.
//   B  per-row vertex/index/face counts. Cubes-and-void chunks: one thread per row, visibility by bit operations
//      on the solid words (voxmesh.rs:119 reduces to solid & ~neighbour-solid). Chunks containing slopes/corners:
//      one warp per row, lane = x, the general orientation-aware test (voxmesh.rs:105-121); the side masks are kept for C.
//   S  block-wide exclusive scan of the 1024 row counts = the reference's emission order (cells y,z,x-major,
//      directions 0..5 minor, voxmesh.rs:94,104); one atomicAdd triple reserves the chunk's vertex span, index span and
//      face records (or the chunk's old reservation is reused, in re-mesh passes).
//   C  face records: one uint4 per visible side (device_common.cuh) with its first vertex / index, class, block id and the
//      27-bit occupancy of the cell's 3x3x3 neighbourhood (all the AO test needs, voxmesh.rs:128-137). Cubes-and-void
//      chunks: tiles of <=1024 sides, (C1) thread per row orders the sides, (C2) thread per side completes the record.
//      Chunks with other shapes: warp per row, every lane writes the records of its cell directly.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "device_common.cuh"

namespace bxw {
namespace {

constexpr int NW = 8;        // warps per CTA
constexpr int NT = NW * 32;  // threads per CTA
// class table: row (y',z') = 40 bytes: [3] = x=-1, [4..35] = x=0..31, [36] = x=32 (so x +- 1 is address +- 1)
constexpr int ROWB = kClsRowBytes;
constexpr int PLANE = kClsPlaneBytes;
constexpr int XOFF = 4;
constexpr int CLS_BYTES = 34 * PLANE;
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int FC = 1024;     // sides per record tile
constexpr int VC = 4096;     // vertices per record tile (the tile-relative first vertex has 14 bits)

// phase-A staging: every warp streams its share of the 34^3 neighbourhood through a private cp.async ring.
// One unit = a third of a y' layer: 3 warp copies of 4 rows (12 rows x 128 B) + (first third only) the layer's x-halo.
constexpr int RSLOT_BYTES = 3 * 512 + 128 + 128 + 16;

struct Scratch { // phases B, S, C
    uint32_t rowV[1028];   // exclusive prefix over rows: first vertex of the row (chunk-relative); [1024] = total
    uint32_t rowI[1028];
    uint32_t rowF[1028];
    // record tile of a cubes-and-void chunk: written by C1
    uint32_t frecA[FC];   // row | x<<10 | dir<<15 | vstart<<18 (tile-relative first vertex)
    // chunks with other shapes: the visible-side mask of every cell, kept from phase B for phase C (the general test costs
    // six class lookups + six table lookups per cell)
    uint8_t vis[32 * 32 * 32];
    // ... and the 34-bit solid words (row + x-halo bits) combined once, for the 27-cell occupancy of every visible cell
    unsigned long long solid64[34 * 34];
};
struct ScratchLean { // the cubes-only kernel: every side is a quad, C2 recomputes the rest
    uint32_t rowV[1028];
    uint32_t rowI[1028];
    uint32_t rowF[1028];
    uint32_t frecA[FC];
};

// Two instantiations of the kernel share this file's code:
//   LEAN = false: the general kernel. 34^3 class table (46 KB) + four-slot rings: 110 KB, two CTAs per SM.
//   LEAN = true:  cubes and void only (all of generated terrain). No class table -- visibility needs only the solid words --
//                 two-slot rings, a quarter of the record scratch: 36 KB, so four to five CTAs per SM hide the staging latency
//                 that bounds the general kernel on surface chunks. A chunk in which it meets another shape is handed to the
//                 general kernel through a fallback list.
template <bool LEAN>
struct SmemT {
    static constexpr int RSLOTS = LEAN ? 2 : 4;
    uint8_t cls[LEAN ? 16 : CLS_BYTES];
    union {
        std::conditional_t<LEAN, ScratchLean, Scratch> sc;   // phases B, S, C
        alignas(16) uint8_t ring[NW][RSLOTS][RSLOT_BYTES];   // phase A
        struct {                                             // phase A of the fused generate+mesh path (general kernel)
            int32_t hm[34 * 36];                             // packed column parameters (height*4 + elevation), [zz][xx]
            int32_t red[2][NW];
        } gen;
    };
    uint32_t solid[34 * 34];         // per (y',z') row: bit x set when voxel x (0..31) has a mesh (class != 0)
    uint8_t solid_halo[34 * 34 + 4]; // bit0: x=-1 solid, bit1: x=32 solid
    uint32_t ctab[kNumClasses + 3];
    uint16_t spread6[64];            // bit d of the index -> bits 2d, 2d+1 (mask over the 2-bit per-direction codes)
    uint8_t reg_kind[256];
    uint32_t kpack;                  // reg_kind[0..15], two bits each
    uint32_t klut[2];                // byte i of the pair: bit 0 = block id i (0..7) has a mesh, bit 7 = id i is not registered
    uint32_t nbr[27 + 5];
    uint32_t warp_tot[3][NW];
    unsigned long long vbase, ibase, fbase;
    uint32_t work_item[2]; // work index of this / the next iteration (double buffered)
    uint32_t work_c0[2], work_sp[2]; // its list entries, when there is a list
    uint32_t wref[NW];     // per-warp reference datum (trivial-chunk detection)
    uint32_t wflags[NW];   // per-warp shapes | mixed<<1
    uint32_t gen_cls[5];   // class of the five generator datums (void, grass, dirt, stone, snow_grass)
    int emit;
};
using Smem = SmemT<false>;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <class S>
__device__ __forceinline__ uint32_t classify(uint32_t d, const S &s, const DeviceRegistry &reg, uint32_t &err) {
    // voxmesh.rs:76-78 -> stdshapes.rs:51-72
    const uint32_t id = d >> 16;
    const uint32_t k = id < 256u ? (uint32_t)s.reg_kind[id] : (id < reg.n ? (uint32_t)__ldg(reg.kind + id) : 0u);
    if (k == 0u) err = 1u;          // voxregistry.rs:141-143 would panic
    if (k != 2u) return 0u;         // VoxelMesh::None
    const uint32_t meta = d & 0xFFFFu;
    uint32_t shape = meta & 63u;    // StdMeta::from_meta: meta % 64
    if (shape > 3u) shape = 0u;     // unknown shape -> cube (stdshapes.rs:59)
    uint32_t o = meta >> 6;         // meta / 64
    if (o >= 24u) o = 0u;           // from_index(..).unwrap_or_default()
    return 1u + shape * 24u + o;
}

// The cubes-only kernel needs less of a voxel in phase A: 0 = no mesh, 1 = a cube (any orientation), 25 = another shape (which
// sends the chunk to the general kernel).
// kpack: the registry kinds of block ids 0..15, two bits each (SmemT::kpack) -- a register lookup for the ids of generated
// terrain instead of a shared-memory load per voxel; other ids take the table (a real branch, so that the common path holds no
// load).
template <class S>
__device__ __noinline__ uint32_t kind_from_table(const S &s, const DeviceRegistry &reg, uint32_t id) {
    return id < 256u ? (uint32_t)s.reg_kind[id] : (id < reg.n ? (uint32_t)__ldg(reg.kind + id) : 0u);
}
template <class S>
__device__ __forceinline__ uint32_t classify_lean(uint32_t d, const S &s, const DeviceRegistry &reg, uint32_t &err, uint32_t kpack) {
    const uint32_t id = d >> 16;
    uint32_t k = (kpack >> (2u * (id & 15u))) & 3u;
    if (id >= 16u) k = kind_from_table(s, reg, id);
    if (k == 0u) err = 1u;
    if (k != 2u) return 0u;
    const uint32_t shape = d & 63u;
    return (shape - 1u < 3u) ? 25u : 1u;
}

// visibility of the 6 sides of the cell at byte `at` -- voxmesh.rs:104-121
template <class S>
__device__ __forceinline__ uint32_t cell_visibility(const S &s, int at, uint32_t t) {
    const uint32_t ec = (t >> 6) & 63u, eu = (t >> 12) & 63u;
    uint32_t vis = eu;
    if (ec) {
        const uint32_t n0 = s.cls[at - 1], n1 = s.cls[at + 1];
        const uint32_t n2 = s.cls[at - PLANE], n3 = s.cls[at + PLANE];
        const uint32_t n4 = s.cls[at - ROWB], n5 = s.cls[at + ROWB];
        // neighbour in world dir d must can_clip on its side facing us (dir d^1)
        const uint32_t blocked = ((s.ctab[n0] >> 1) & 1u) | (((s.ctab[n1] >> 0) & 1u) << 1) | (((s.ctab[n2] >> 3) & 1u) << 2) |
                                 (((s.ctab[n3] >> 2) & 1u) << 3) | (((s.ctab[n4] >> 5) & 1u) << 4) |
                                 (((s.ctab[n5] >> 4) & 1u) << 5);
        vis |= ec & ~blocked;
    }
    return vis;
}

// vertices / indices / faces of a cell, packed V | I<<10 | F<<21. The class word holds a 2-bit vertex-count code per direction
// (0 none, 1 three, 2 four, 3 six vertices; indices 3 / 6 / 6): mask the codes of the visible directions and count the codes of
// each kind with three popcounts.
template <class S>
__device__ __forceinline__ uint32_t cell_counts(const S &s, uint32_t vis, uint32_t t) {
    const uint32_t c = (t >> 18) & s.spread6[vis];
    const uint32_t lo = c & 0x555u, hi = (c >> 1) & 0x555u;
    const uint32_t n1 = __popc(lo & ~hi), n2 = __popc(hi & ~lo), n3 = __popc(lo & hi);
    const uint32_t v = 3u * n1 + 4u * n2 + 6u * n3, i = 3u * n1 + 6u * (n2 + n3);
    return v | (i << 10) | ((uint32_t)__popc(vis) << 21);
}

// the 34-bit solid word of row rr (bit 0: x=-1, bits 1..32: x=0..31, bit 33: x=32)
template <class S>
__device__ __forceinline__ unsigned long long solid_word(const S &s, int rr) {
    const unsigned long long h = s.solid_halo[rr];
    return ((unsigned long long)s.solid[rr] << 1) | (h & 1ull) | ((h >> 1) << 33);
}

// occupancy (class != 0) of the 3x3x3 neighbourhood of cell (x, y, z), bit (dy+1)*9 + (dz+1)*3 + (dx+1)
template <class S>
__device__ __forceinline__ uint32_t neighbourhood27(const S &s, int y, int z, int x) {
    uint32_t nb = 0;
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dz = 0; dz < 3; dz++) {
            const unsigned long long w = solid_word(s, (y + dy) * 34 + (z + dz)); // rows y-1..y+1, z-1..z+1 (+1 halo offset)
            nb |= ((uint32_t)(w >> x) & 7u) << (3 * (dy * 3 + dz)); // bit x of w is cell x-1
        }
    return nb;
}

// the same from the combined words (low / high halves; bit x of a word is cell x-1)
__device__ __forceinline__ uint32_t neighbourhood27_words(const unsigned long long *w64, int y, int z, int x) {
    uint32_t nb = 0;
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dz = 0; dz < 3; dz++) {
            const uint2 w = *reinterpret_cast<const uint2 *>(&w64[(y + dy) * 34 + (z + dz)]);
            nb |= (__funnelshift_r(w.x, w.y, x) & 7u) << (3 * (dy * 3 + dz));
        }
    return nb;
}

// visible-side words of a cubes-and-void row: bit x+1 of out[d] = side d of cell x is visible
template <class S>
__device__ __forceinline__ void row_visibility(const S &s, int y, int z, unsigned long long out[6]) {
    const int r = (y + 1) * 34 + (z + 1);
    const unsigned long long C = solid_word(s, r), M = 0x1FFFFFFFEull;
    out[0] = C & ~(C << 1) & M;
    out[1] = C & ~(C >> 1) & M;
    out[2] = C & ~solid_word(s, r - 34) & M;
    out[3] = C & ~solid_word(s, r + 34) & M;
    out[4] = C & ~solid_word(s, r - 1) & M;
    out[5] = C & ~solid_word(s, r + 1) & M;
}

// the slot's arena reservation is given up (its mesh moved or became empty): count it as garbage
__device__ __forceinline__ void release_span(const MeshParams &p, uint32_t span_slot) {
    if (!p.span_cap) return;
    if (p.reuse) {
        const uint2 oc = p.span_cap[span_slot];
        if (oc.x) atomicAdd(&p.counters->v_garbage, (unsigned long long)oc.x);
        if (oc.y) atomicAdd(&p.counters->i_garbage, (unsigned long long)oc.y);
    }
    p.span_cap[span_slot] = make_uint2(0u, 0u);
}

// x-halo class bytes / solid bits of layer yy: lane z owns row zz = z+1, lanes 0 and 2 also the corner rows 0 / 33
template <bool LEAN>
__device__ __forceinline__ void halo_store(SmemT<LEAN> &s, int lane, int yy, uint32_t cm, uint32_t cp, uint32_t ce, uint32_t co) {
    if (!LEAN) {
        uint8_t *row = &s.cls[yy * PLANE + (lane + 1) * ROWB + XOFF];
        row[-1] = (uint8_t)cm;
        row[32] = (uint8_t)cp;
    }
    s.solid_halo[yy * 34 + lane + 1] = (uint8_t)((cm != 0u) | ((cp != 0u) << 1));
    if (lane == 0 || lane == 2) {
        const int zz = lane ? 33 : 0;
        if (!LEAN) {
            s.cls[yy * PLANE + zz * ROWB + XOFF - 1] = (uint8_t)ce;
            s.cls[yy * PLANE + zz * ROWB + XOFF + 32] = (uint8_t)co;
        }
        s.solid_halo[yy * 34 + zz] = (uint8_t)((ce != 0u) | ((co != 0u) << 1));
    }
}

template <bool LEAN>
__device__ __forceinline__ uint32_t classify_any(uint32_t d, const SmemT<LEAN> &s, const DeviceRegistry &reg, uint32_t &err, uint32_t kpack) {
    return LEAN ? classify_lean(d, s, reg, err, kpack) : classify(d, s, reg, err);
}

// A staged unit (a third of a layer: 12 rows + in the first third the layer's x-halo) that is NOT one single datum: per-voxel
// classification into class bytes and solid words. Kept out of line on purpose: the unit loop of phase A is unrolled three
// ways and inlining this body into each copy made the kernel outgrow the instruction cache (a third of all stall cycles on
// surface chunks were instruction fetches). Returns shapes | err << 1.
template <bool LEAN>
__device__ __forceinline__ uint32_t classify_unit_body(SmemT<LEAN> &s, const uint8_t *reg_kind, uint32_t reg_n, int yy, int PART, uint4 v0,
                                                       uint4 v1, uint4 v2, uint32_t hm, uint32_t hp, uint32_t he) {
    const int lane = threadIdx.x & 31, q = lane & 7, zsub = lane >> 3;
    DeviceRegistry reg;
    reg.kind = reg_kind;
    reg.texf = nullptr;
    reg.color = nullptr;
    reg.n = reg_n;
    const uint32_t kpack = LEAN ? s.kpack : 0u;
    const uint32_t lut0 = LEAN ? s.klut[0] : 0u, lut1 = LEAN ? s.klut[1] : 0u;
    uint32_t err = 0, shapes = 0, tiny_seen = 0;
    uint32_t last_d = 0xFFFFFFFFu, last_c = 0; // (datum 0xFFFFFFFF: block id 65535 with meta 65535 -- classified like any other on a miss)
    bool have = false;
    if (PART == 0) {
        const uint32_t cm = classify_any<LEAN>(hm, s, reg, err, kpack), cp = classify_any<LEAN>(hp, s, reg, err, kpack);
        const uint32_t ce = lane < 4 ? classify_any<LEAN>(he, s, reg, err, kpack) : 0u;
        const uint32_t co = __shfl_xor_sync(FULL, ce, 1); // the other side of the same corner row
        shapes |= (cm > 24u) | (cp > 24u) | (ce > 24u);
        halo_store<LEAN>(s, lane, yy, cm, cp, ce, co);
    }
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
        const uint4 v = jj == 0 ? v0 : (jj == 1 ? v1 : v2);
        const int zz = 4 * (3 * PART + jj) + zsub;
        const uint32_t first = __shfl_sync(FULL, v.x, 0);
        const uint32_t diff = ((v.x ^ first) | (v.y ^ first)) | ((v.z ^ first) | (v.w ^ first));
        uint32_t b, rowmask;
        if (__all_sync(FULL, diff == 0u || zz >= 34)) { // all 128 voxels equal: classify once (cached)
            if (!have || first != last_d) {
                have = true;
                last_d = first;
                last_c = classify_any<LEAN>(first, s, reg, err, kpack);
            }
            b = last_c * 0x01010101u;
            rowmask = last_c ? 0xFFFFFFFFu : 0u;
            shapes |= (last_c > 24u);
        } else if (LEAN && __all_sync(FULL, (((v.x | v.y) | (v.z | v.w)) & 0xFFF8FFFFu) == 0u)) {
            // Cubes-only kernel, all 128 datums are block ids 0..7 with meta 0 (all of generated terrain): the four lookups of a
            // lane are ONE byte permute -- the selector's nibbles are the four ids, the two lookup words hold a byte per id -- and
            // the four "has a mesh" bits are gathered with one multiply.
            const uint32_t sel = ((v.x >> 16) | (v.y >> 12)) | ((v.z >> 8) | (v.w >> 4));
            const uint32_t r = zz < 34 ? __byte_perm(lut0, lut1, sel) : 0u; // (rows 34, 35 of the last third hold stale ring data)
            tiny_seen |= r;
            uint32_t m = (((r & 0x01010101u) * 0x01020408u) >> 24) << (q * 4);
            m |= __shfl_xor_sync(FULL, m, 1);
            m |= __shfl_xor_sync(FULL, m, 2);
            m |= __shfl_xor_sync(FULL, m, 4);
            b = 0;
            rowmask = m; // (a row beyond the 34 is not stored)
        } else {
            const bool valid = zz < 34;
            const uint32_t k0 = valid ? classify_any<LEAN>(v.x, s, reg, err, kpack) : 0u, k1 = valid ? classify_any<LEAN>(v.y, s, reg, err, kpack) : 0u,
                           k2 = valid ? classify_any<LEAN>(v.z, s, reg, err, kpack) : 0u, k3 = valid ? classify_any<LEAN>(v.w, s, reg, err, kpack) : 0u;
            b = k0 | (k1 << 8) | (k2 << 16) | (k3 << 24);
            shapes |= (k0 > 24u) | (k1 > 24u) | (k2 > 24u) | (k3 > 24u);
            uint32_t m = ((k0 != 0u) | ((k1 != 0u) << 1) | ((k2 != 0u) << 2) | ((k3 != 0u) << 3)) << (q * 4);
            m |= __shfl_xor_sync(FULL, m, 1);
            m |= __shfl_xor_sync(FULL, m, 2);
            m |= __shfl_xor_sync(FULL, m, 4);
            rowmask = m;
        }
        if (zz < 34) {
            if (!LEAN) *reinterpret_cast<uint32_t *>(&s.cls[yy * PLANE + zz * ROWB + XOFF + q * 4]) = b;
            if (q == 0) s.solid[yy * 34 + zz] = rowmask;
        }
    }
    if (tiny_seen & 0x80808080u) err = 1u; // an unregistered id among them (voxregistry.rs:141-143 would panic)
    return shapes | (err << 1);
}
template <bool LEAN>
__device__ __noinline__ uint32_t classify_unit_call(SmemT<LEAN> &s, const uint8_t *reg_kind, uint32_t reg_n, int yy, int PART, uint4 v0, uint4 v1,
                                                    uint4 v2, uint32_t hm, uint32_t hp, uint32_t he) {
    return classify_unit_body<LEAN>(s, reg_kind, reg_n, yy, PART, v0, v1, v2, hm, hp, he);
}

#ifdef BXW_PHASE_TIMING
#define PHASE_MARK(var) const long long var = clock64()
#define PHASE_ADD(slot, a, b) do { if (tid == 0) atomicAdd(&p.counters->phase[slot], (unsigned long long)((b) - (a))); } while (0)
#else
#define PHASE_MARK(var)
#define PHASE_ADD(slot, a, b)
#endif

// CALL_SLOW: the per-voxel branch of phase A is a function call instead of inline code (smaller kernel: better where most
// chunks take that branch at high occupancy -- the candidate list of a terrain pass; worse where nearly every unit is uniform).
template <bool LEAN, bool CALL_SLOW>
__global__ void __launch_bounds__(NT, LEAN ? 4 : 2) mesh_kernel(MeshParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SM = SmemT<LEAN>;
    SM &s = *reinterpret_cast<SM *>(smem_raw);
    constexpr int RSLOTS = SM::RSLOTS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const MeshTables *T = p.tables;

    for (int i = tid; i < kNumClasses; i += NT) s.ctab[i] = T->ctab[i];
    for (int i = tid; i < 256; i += NT) s.reg_kind[i] = (uint32_t)i < p.reg.n ? p.reg.kind[i] : (uint8_t)0;
    if (tid == 0) {
        uint32_t kp = 0;
        for (uint32_t i = 0; i < 16u && i < p.reg.n; i++) kp |= ((uint32_t)p.reg.kind[i] & 3u) << (2u * i);
        s.kpack = kp;
        uint32_t lut[2] = {0u, 0u};
        for (uint32_t i = 0; i < 8u; i++) {
            const uint32_t k = i < p.reg.n ? (uint32_t)p.reg.kind[i] : 0u;
            lut[i >> 2] |= ((k == 2u ? 1u : 0u) | (k == 0u ? 0x80u : 0u)) << (8u * (i & 3u));
        }
        s.klut[0] = lut[0];
        s.klut[1] = lut[1];
    }
    if (tid < 64) {
        uint32_t m = 0;
        for (int d = 0; d < 6; d++) m |= ((tid >> d) & 1) ? (3u << (2 * d)) : 0u;
        s.spread6[tid] = (uint16_t)m;
    }
    if (tid < 5) { // class of the generator's datums (meta 0): cube (class 1) when the block has a mesh, else 0
        const uint32_t id = p.gen_datum[tid] >> 16;
        s.gen_cls[tid] = (p.hmap && id < p.reg.n && p.reg.kind[id] == 2) ? 1u : 0u;
    }
    const uint32_t n_work = p.n_work_dev ? __ldg(p.n_work_dev) : p.n_work;
    // Work items come from an atomic cursor. The last warp (it has no outer-layer unit in phase A, so it reaches the phase's
    // barrier first) fetches the next item while the current chunk is staged: the atomic is issued at the top of the iteration
    // and its result -- plus the list entries it selects -- is published just before that barrier, off the critical path.
    constexpr int PF_TID = NT - 32;
    if (tid == PF_TID) {
        const uint32_t w0 = (uint32_t)atomicAdd(p.cursor, 1ull);
        s.work_item[0] = w0;
        if (p.work && w0 < n_work) {
            s.work_c0[0] = p.work[w0];
            s.work_sp[0] = p.work_span[w0];
        }
    }

    for (uint32_t iter = 0;; iter++) {
        __syncthreads();
        const uint32_t w = s.work_item[iter & 1];
        if (w >= n_work) break;
        PHASE_MARK(t_fetch);
        uint32_t next_w = 0;
        if (tid == PF_TID) next_w = (uint32_t)atomicAdd(p.cursor, 1ull); // consumed before the phase-A barrier
        uint32_t c0, span_slot;
        if (p.work) {
            c0 = s.work_c0[iter & 1];
            span_slot = s.work_sp[iter & 1];
        } else {
            // The whole interior without a work list (no dependent global load). Order: x-tiles of 8 columns; inside a
            // tile z-rows, then the 8 columns, then the vertical stack fastest. Chunks whose halos overlap are then
            // in flight on neighbouring CTAs at about the same time, so halo rows/planes hit L2 instead of HBM.
            const uint32_t nx = (uint32_t)p.SX - 2u, ny = (uint32_t)p.SY - 2u, nz = (uint32_t)p.SZ - 2u;
            const uint32_t per_tile = 8u * ny * nz;
            const uint32_t t = w / per_tile, rem = w - t * per_tile;
            const uint32_t tw = min(8u, nx - 8u * t);
            const uint32_t iz = rem / (tw * ny), r2 = rem - iz * tw * ny;
            const uint32_t ix = 8u * t + r2 / ny, iy = r2 % ny;
            c0 = (ix + 1u) + (uint32_t)p.SX * ((iz + 1u) + (uint32_t)p.SZ * (iy + 1u));
            span_slot = ix + nx * (iz + nz * iy);
        }
        if (tid < 27) {
            const int sx = (int)(c0 % (uint32_t)p.SX), sz = (int)((c0 / (uint32_t)p.SX) % (uint32_t)p.SZ),
                      sy = (int)(c0 / (uint32_t)(p.SX * p.SZ));
            const int rx = tid % 3, rz = (tid / 3) % 3, ry = tid / 9;
            const uint32_t nc = (uint32_t)((sx + rx - 1) + p.SX * ((sz + rz - 1) + p.SZ * (sy + ry - 1)));
            s.nbr[tid] = p.page ? __ldg(p.page + nc) : nc; // where the neighbour's voxels live (itself or a shared uniform page)
        }
        __syncthreads();

        PHASE_MARK(t_a0);
        PHASE_ADD(0, t_fetch, t_a0);
        PHASE_ADD(5, 0, 1);
        // ---------------- phase A: stage 34^3 class bytes + per-row solid words ----------------
        uint32_t err = 0;
        uint32_t shapes = 0;  // saw a non-cube shape (class > 24)
        uint32_t mixed = 0;   // saw a voxel different from this warp's reference voxel
        if (!LEAN && p.hmap) {
            // Fused generate+mesh: the blocks of pristine terrain are a pure function of the column parameters
            // (stdgen.rs:349-373), so the 34^3 table is derived from the 34x34 columns under the chunk instead of reading
            // 148 KB of voxels back from HBM.
            const int sx = (int)(c0 % (uint32_t)p.SX), sz = (int)((c0 / (uint32_t)p.SX) % (uint32_t)p.SZ), sy = (int)(c0 / (uint32_t)(p.SX * p.SZ));
            const int y_lo = (p.cy_min + sy) * 32 - 1; // world y of layer yy = 0
            int hmin = 0x7FFFFFFF, hmax = -0x7FFFFFFF;
            for (int i = tid; i < 34 * 34; i += NT) {
                const int zz = i / 34, xx = i - zz * 34;
                const int v = __ldg(p.hmap + (size_t)(sz * 32 - 1 + zz) * p.hmap_W + sx * 32 - 1 + xx);
                s.gen.hm[zz * 36 + xx] = v;
                hmin = min(hmin, v >> 2);
                hmax = max(hmax, v >> 2);
            }
            hmin = __reduce_min_sync(FULL, hmin);
            hmax = __reduce_max_sync(FULL, hmax);
            if (lane == 0) {
                s.gen.red[0][warp] = hmin;
                s.gen.red[1][warp] = hmax;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < NW; k++) {
                hmin = min(hmin, s.gen.red[0][k]);
                hmax = max(hmax, s.gen.red[1][k]);
            }
            const uint32_t c_void = s.gen_cls[0];
            // all 34 layers above every column -> one void datum; all at or below every column's dirt/stone -> solid cubes
            const bool all_void = y_lo > hmax;
            const bool all_solid = (y_lo + 33 <= hmin) && p.gen_all_cubes; // every voxel at or below its column's surface: terrain cubes only
            if ((all_void && c_void == 0u) || all_solid) {
                mixed = 0;
                shapes = 0;
            } else {
                mixed = 1;
                const uint32_t g0 = s.gen_cls[0], g1 = s.gen_cls[1], g2 = s.gen_cls[2], g3 = s.gen_cls[3], g4 = s.gen_cls[4];
                // Layers above every column are all void, layers more than 5 below every column all stone: those are
                // filled with one class word per 4 cells. Only the band in between looks at the columns.
                const int yy_lo = max(0, min(34, (hmin - 5) - y_lo)); // first layer that can hold something else than stone
                const int yy_hi = max(0, min(34, hmax - y_lo + 1));   // first layer that is void everywhere
                for (int i = tid; i < 34 * 34; i += NT) {
                    const int yy = i / 34, zz = i - yy * 34;
                    if (yy >= yy_lo && yy < yy_hi) continue;
                    const uint32_t c = yy < yy_lo ? g3 : g0;
                    uint32_t *row = reinterpret_cast<uint32_t *>(&s.cls[yy * PLANE + zz * ROWB]);
#pragma unroll
                    for (int k = 0; k < ROWB / 4; k++) row[k] = c * 0x01010101u; // x = -1 .. 32 live at bytes XOFF-1 .. XOFF+32
                    s.solid[yy * 34 + zz] = c ? 0xFFFFFFFFu : 0u;
                    s.solid_halo[yy * 34 + zz] = c ? 3u : 0u;
                    shapes |= (c > 24u);
                }
                // warp per z-row of columns, lane = x (lanes 0/1 also own the x = -1 / 32 halo columns); the column
                // parameters stay in registers while the band's layers are emitted
                auto cls_of = [&](int hv, int y) {
                    const int h = hv >> 2, el = hv & 3;
                    return (y == h) ? ((el == 2 && y > 80) ? g4 : g1) : (y < h - 5 ? g3 : (y < h ? g2 : g0));
                };
                for (int zz = warp; zz < 34; zz += NW) {
                    const int hv = s.gen.hm[zz * 36 + lane + 1];
                    const int hh = s.gen.hm[zz * 36 + (lane & 1 ? 33 : 0)]; // used by lanes 0 and 1
#pragma unroll 2
                    for (int yy = yy_lo; yy < yy_hi; yy++) {
                        const int y = y_lo + yy;
                        const uint32_t c = cls_of(hv, y);
                        const uint32_t rowmask = __ballot_sync(FULL, c != 0u);
                        shapes |= (c > 24u);
                        s.cls[yy * PLANE + zz * ROWB + XOFF + lane] = (uint8_t)c;
                        if (lane < 2) {
                            const uint32_t ch = cls_of(hh, y);
                            shapes |= (ch > 24u);
                            s.cls[yy * PLANE + zz * ROWB + XOFF + (lane ? 32 : -1)] = (uint8_t)ch;
                            const uint32_t other = __shfl_xor_sync(0x3u, ch, 1);
                            if (lane == 0) {
                                s.solid[yy * 34 + zz] = rowmask;
                                s.solid_halo[yy * 34 + zz] = (uint8_t)((ch != 0u) | ((other != 0u) << 1));
                            }
                        }
                    }
                }
            }
            if (lane == 0) s.wref[warp] = 0;
        } else {
            const int q = lane & 7, zsub = lane >> 3;
            auto cls_of = [&](uint32_t d, const SM &sm, const DeviceRegistry &rg, uint32_t &e) {
                return LEAN ? classify_lean(d, sm, rg, e, sm.kpack) : classify(d, sm, rg, e);
            };
            uint32_t last_d = 0, last_c = 0, wref = 0;
            bool have_ref = false;
            auto note_ref = [&](uint32_t x) { // first datum this warp sees: reference for trivial-chunk detection
                if (!have_ref) {
                    have_ref = true;
                    wref = __shfl_sync(FULL, x, 0);
                    last_d = wref;
                    last_c = cls_of(wref, s, p.reg, err);
                }
            };
            // source pointers of a layer's rows: zz = 0 / 1..32 / 33 live in the three z-column chunks at this ry
            auto row_src = [&](int ry, int by, int zz) {
                const int rz = zz == 0 ? 0 : (zz == 33 ? 2 : 1), bz = (zz - 1) & 31;
                return p.blocks + (size_t)s.nbr[1 + 3 * rz + 9 * ry] * kChunkVoxels + by * 1024 + bz * 32 + q * 4;
            };
            // The 34 layers are cut into 102 units (layer, third); warp w takes units w, w+NW, ... and streams them through
            // its private ring, two units (6 warp copies + halo) ahead of their use.
            constexpr int N_UNITS = 34 * 3;
            uint8_t *const ring = &s.ring[warp][0][0];
            // rows of the chunk itself (yy 1..32, zz 1..32) come straight off this pointer; only the boundary rows look up
            // a neighbour chunk
            const uint32_t *const center_q = p.blocks + (size_t)s.nbr[13] * kChunkVoxels + q * 4;
            auto issue = [&](int g, int slot) {
                if (g < N_UNITS) {
                    const int yy = g / 3, part = g - yy * 3;
                    uint8_t *dst = ring + slot * RSLOT_BYTES + lane * 16;
                    const bool inner = (yy >= 1) & (yy <= 32);
                    const int ry = yy == 0 ? 0 : (yy <= 32 ? 1 : 2), by = (yy - 1) & 31;
                    const int zb = 12 * part + zsub; // zz of copy jj is zb + 4*jj
                    const uint32_t *mid = (inner ? center_q : p.blocks + (size_t)s.nbr[1 + 3 + 9 * ry] * kChunkVoxels + q * 4) + by * 1024 + (zb - 1) * 32;
                    if (part == 1) {
                        cp_async16(dst, mid);
                        cp_async16(dst + 512, mid + 128);
                        cp_async16(dst + 1024, mid + 256);
                    } else if (part == 0) {
                        cp_async16(dst, zb == 0 ? row_src(ry, by, 0) : mid);
                        cp_async16(dst + 512, mid + 128);
                        cp_async16(dst + 1024, mid + 256);
                        // x-halo of the layer from the x-neighbours' face planes (coalesced 128 B each) + 4 corner voxels
                        uint8_t *hd = ring + slot * RSLOT_BYTES + 1536 + lane * 4;
                        cp_async4(hd, p.xface + (size_t)s.nbr[0 + 3 + 9 * ry] * 2048 + 1024 + by * 32 + lane);
                        cp_async4(hd + 128, p.xface + (size_t)s.nbr[2 + 3 + 9 * ry] * 2048 + by * 32 + lane);
                        if (lane < 4) {
                            const int erx = (lane & 1) ? 2 : 0, erz = (lane & 2) ? 2 : 0;
                            cp_async4(hd + 256, p.xface + (size_t)s.nbr[erx + 3 * erz + 9 * ry] * 2048 + ((lane & 1) ? 0 : 1024) + by * 32 + ((lane & 2) ? 0 : 31));
                        }
                    } else {
                        cp_async16(dst, mid);
                        cp_async16(dst + 512, mid + 128);
                        if (zb + 8 == 33) cp_async16(dst + 1024, row_src(ry, by, 33));
                        else if (zb + 8 < 33) cp_async16(dst + 1024, mid + 256);
                    }
                }
                cp_async_commit();
            };
            // one unit (PART = which third of the layer). One copy of this code per kernel: the unit loop below is not unrolled --
            // unrolled six ways the kernel outgrew the instruction cache (a third of all stall cycles were instruction fetches).
            auto process = [&](int yy, const uint8_t *src, const int PART) {
                uint4 v[3];
#pragma unroll
                for (int jj = 0; jj < 3; jj++) v[jj] = *reinterpret_cast<const uint4 *>(src + jj * 512 + lane * 16);
                const bool last_valid = PART < 2 || zsub < 2; // rows 34, 35 of the last copy do not exist
                if (PART == 2 && !last_valid) v[2] = make_uint4(0, 0, 0, 0);
                uint32_t hm = 0, hp = 0, he = 0;
                if (PART == 0) {
                    hm = *reinterpret_cast<const uint32_t *>(src + 1536 + lane * 4);
                    hp = *reinterpret_cast<const uint32_t *>(src + 1664 + lane * 4);
                    if (lane < 4) he = *reinterpret_cast<const uint32_t *>(src + 1792 + lane * 4);
                }
                // whole-unit test first: 384 (+68) voxels equal to one datum (the common case in terrain) costs one match +
                // one vote instead of a shuffle/vote chain per 128-voxel group
                const uint32_t x0 = v[0].x; // rows of v[0] always exist
                note_ref(x0);
                uint32_t d = ((v[0].x ^ x0) | (v[0].y ^ x0)) | ((v[0].z ^ x0) | (v[0].w ^ x0));
                d |= ((v[1].x ^ x0) | (v[1].y ^ x0)) | ((v[1].z ^ x0) | (v[1].w ^ x0));
                const uint32_t d2 = ((v[2].x ^ x0) | (v[2].y ^ x0)) | ((v[2].z ^ x0) | (v[2].w ^ x0));
                if (last_valid) d |= d2;
                if (PART == 0) d |= (hm ^ x0) | (hp ^ x0) | (lane < 4 ? (he ^ x0) : 0u);
                int same_x0;
                __match_all_sync(FULL, x0, &same_x0);
                if (__all_sync(FULL, d == 0u) && same_x0) {
                    if (x0 != last_d) {
                        last_d = x0;
                        last_c = cls_of(x0, s, p.reg, err);
                    }
                    mixed |= (x0 != wref);
                    shapes |= (last_c > 24u);
                    const uint32_t b = last_c * 0x01010101u, rowmask = last_c ? 0xFFFFFFFFu : 0u;
                    uint32_t *so = &s.solid[yy * 34 + 12 * PART + zsub];
                    if (!LEAN) {
                        uint8_t *cl = &s.cls[yy * PLANE + (12 * PART + zsub) * ROWB + XOFF + q * 4];
                        *reinterpret_cast<uint32_t *>(cl) = b;
                        *reinterpret_cast<uint32_t *>(cl + 4 * ROWB) = b;
                        if (last_valid) *reinterpret_cast<uint32_t *>(cl + 8 * ROWB) = b;
                    }
                    if (q == 0) {
                        so[0] = rowmask;
                        so[4] = rowmask;
                        if (last_valid) so[8] = rowmask;
                    }
                    if (PART == 0) halo_store<LEAN>(s, lane, yy, last_c, last_c, last_c, last_c);
                    return;
                }
                mixed = 1;
                const uint32_t fl = CALL_SLOW ? classify_unit_call<LEAN>(s, p.reg.kind, p.reg.n, yy, PART, v[0], v[1], v[2], hm, hp, he)
                                              : classify_unit_body<LEAN>(s, p.reg.kind, p.reg.n, yy, PART, v[0], v[1], v[2], hm, hp, he);
                shapes |= fl & 1u;
                err |= fl >> 1;
            };
            // Unit schedule. The 32 layers that belong to the chunk's own y range (yy = 1..32, neighbours all at ry = 1) are
            // dealt 4 per warp: their addresses are per-chunk constant pointers + by*stride. The two outer layers (yy = 0, 33:
            // the y-neighbour chunks) make 6 more units, one each for warps 0..5, issued through the generic path.
            const uint32_t *const Zm = p.blocks + (size_t)s.nbr[1 + 0 + 9] * kChunkVoxels + 31 * 32 + q * 4; // z- neighbour, row z=31
            const uint32_t *const Zp = p.blocks + (size_t)s.nbr[1 + 6 + 9] * kChunkVoxels + q * 4;           // z+ neighbour, row z=0
            const uint32_t *const Xm = p.xface + (size_t)s.nbr[0 + 3 + 9] * 2048 + 1024 + lane;              // x- neighbour, its x=31 plane
            const uint32_t *const Xp = p.xface + (size_t)s.nbr[2 + 3 + 9] * 2048 + lane;                     // x+ neighbour, its x=0 plane
            const uint32_t *const Xc = p.xface + (size_t)s.nbr[((lane & 1) ? 2 : 0) + 3 * ((lane & 2) ? 2 : 0) + 9] * 2048 +
                                       ((lane & 1) ? 0 : 1024) + ((lane & 2) ? 0 : 31); // corner rows (lanes 0..3)
            auto issue_inner = [&](int by, int slot_, const int PART) {
                uint8_t *dst = ring + slot_ * RSLOT_BYTES + lane * 16;
                const uint32_t *rowbase = center_q + by * 1024 + (12 * PART + zsub - 1) * 32; // row of copy 0 (zz - 1)
                if (PART == 0) {
                    cp_async16(dst, zsub == 0 ? Zm + by * 1024 : rowbase);
                    cp_async16(dst + 512, rowbase + 128);
                    cp_async16(dst + 1024, rowbase + 256);
                    uint8_t *hd = ring + slot_ * RSLOT_BYTES + 1536 + lane * 4;
                    cp_async4(hd, Xm + by * 32);
                    cp_async4(hd + 128, Xp + by * 32);
                    if (lane < 4) cp_async4(hd + 256, Xc + by * 32);
                } else if (PART == 1) {
                    cp_async16(dst, rowbase);
                    cp_async16(dst + 512, rowbase + 128);
                    cp_async16(dst + 1024, rowbase + 256);
                } else {
                    cp_async16(dst, rowbase);
                    cp_async16(dst + 512, rowbase + 128);
                    if (zsub == 0) cp_async16(dst + 1024, rowbase + 256);
                    else if (zsub == 1) cp_async16(dst + 1024, Zp + by * 1024);
                }
                cp_async_commit();
            };
            const int tail_g = warp < 6 ? (warp < 3 ? 0 : 33) * 3 + warp % 3 : N_UNITS; // N_UNITS = none (empty commit)
            // The warp's 12 inner units u = 3 li + part (and then its outer unit, u = 12) go through slot u % RSLOTS, RSLOTS - 1
            // units ahead of their use.
            constexpr int DIST = RSLOTS - 1;
            auto issue_unit = [&](int v, const int part) { // v: unit index, part = v % 3 (known at compile time at every call)
                const int li2 = v / 3;
                if (li2 < 4) issue_inner(warp + NW * li2, v % RSLOTS, part);
                else if (v == 12) issue(tail_g, v % RSLOTS); // this warp's outer-layer unit
                else cp_async_commit();
            };
#pragma unroll
            for (int v = 0; v < DIST; v++) issue_unit(v, v % 3);
#pragma unroll 1
            for (int li = 0; li < 4; li++) {
                // the three thirds of a layer are unrolled (their row-validity tests fold away), the layers are not
#pragma unroll
                for (int part = 0; part < 3; part++) {
                    const int u = 3 * li + part;
                    __syncwarp();
                    issue_unit(u + DIST, (part + DIST) % 3);
                    cp_async_wait<DIST>();
                    __syncwarp();
                    process(warp + NW * li + 1, ring + (u % RSLOTS) * RSLOT_BYTES, part);
                }
            }
            if (tail_g < N_UNITS) {
                cp_async_wait<0>();
                __syncwarp();
                const int yy = tail_g / 3;
                process(yy, ring, tail_g - yy * 3);
            }
            cp_async_wait<0>();
            if (lane == 0) s.wref[warp] = wref;
        }
        if (err) atomicOr(&p.counters->flags, kFlagUnknownId);
        if (tid == PF_TID) { // publish the next work item (see above)
            s.work_item[(iter + 1) & 1] = next_w;
            if (p.work && next_w < n_work) {
                s.work_c0[(iter + 1) & 1] = p.work[next_w];
                s.work_sp[(iter + 1) & 1] = p.work_span[next_w];
            }
        }
        // one barrier for both flags: each warp publishes shapes | mixed<<1 (warp-reduced) next to its reference datum
        {
            const uint32_t wf = __reduce_or_sync(FULL, shapes | (mixed << 1));
            if (lane == 0) s.wflags[warp] = wf;
        }
        __syncthreads();
        PHASE_MARK(t_a1);
        PHASE_ADD(1, t_a0, t_a1);
        int has_shapes = 0, any_mixed = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            has_shapes |= (int)(s.wflags[k] & 1u);
            any_mixed |= (int)(s.wflags[k] >> 1) | (int)(s.wref[k] != s.wref[0]); // did all warps see the same single datum?
        }
        if (LEAN && has_shapes) {
            // not a cubes-and-void neighbourhood: the general kernel meshes this chunk (second launch over the fallback list)
            if (tid == 0) {
                const uint32_t k = atomicAdd(p.fallback_count, 1u);
                p.fallback_work[k] = c0;
                p.fallback_span[k] = span_slot;
            }
            continue;
        }
        if (!has_shapes && !any_mixed) {
            // The whole 34^3 neighbourhood is one datum and it is void or a cube: every side is either absent or
            // clipped by an identical neighbour (voxmesh.rs:100,119) -> empty mesh, nothing to count or emit.
            if (tid == 0) {
                bxw_mesh_span sp;
                sp.v_off = sp.i_off = 0;
                sp.v_count = sp.i_count = 0;
                p.spans[span_slot] = sp;
                release_span(p, span_slot);
            }
            continue;
        }

        // ---------------- phase B: per-row counts (packed V | I<<10 | F<<21 into rowV) ----------------
        if (LEAN || !has_shapes) {
            for (int row = tid; row < 1024; row += NT) {
                unsigned long long vw[6];
                row_visibility(s, row >> 5, row & 31, vw);
                const uint32_t nf = __popcll(vw[0]) + __popcll(vw[1]) + __popcll(vw[2]) + __popcll(vw[3]) + __popcll(vw[4]) + __popcll(vw[5]);
                s.sc.rowV[row] = (nf * 4u) | ((nf * 6u) << 10) | (nf << 21);
            }
        } else if constexpr (!LEAN) {
            for (int row = warp; row < 1024; row += NW) {
                const int y = row >> 5, z = row & 31;
                const int at = (y + 1) * PLANE + (z + 1) * ROWB + XOFF + lane;
                const uint32_t c = s.cls[at];
                uint32_t cnt = 0;
                if (__any_sync(FULL, c != 0)) {
                    const uint32_t t = s.ctab[c];
                    const uint32_t vis = cell_visibility(s, at, t);
                    s.sc.vis[row * 32 + lane] = (uint8_t)vis; // (rows without a side are skipped by phase C)
                    cnt = __reduce_add_sync(FULL, cell_counts(s, vis, t));
                }
                if (lane == 0) s.sc.rowV[row] = cnt;
            }
        }
        __syncthreads();

        // ---------------- scan of the 1024 row counts ----------------
        {
            uint32_t v[4], i[4], f[4], tv = 0, ti = 0, tf = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t c = s.sc.rowV[tid * 4 + k];
                v[k] = tv;
                i[k] = ti;
                f[k] = tf;
                tv += c & 1023u;
                ti += (c >> 10) & 2047u;
                tf += c >> 21;
            }
            uint32_t sv = tv, si = ti, sf = tf;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t a = __shfl_up_sync(FULL, sv, o), b = __shfl_up_sync(FULL, si, o), c = __shfl_up_sync(FULL, sf, o);
                if (lane >= o) {
                    sv += a;
                    si += b;
                    sf += c;
                }
            }
            if (lane == 31) {
                s.warp_tot[0][warp] = sv;
                s.warp_tot[1][warp] = si;
                s.warp_tot[2][warp] = sf;
            }
            __syncthreads();
            uint32_t bv = 0, bi = 0, bf = 0;
            for (int k = 0; k < warp; k++) {
                bv += s.warp_tot[0][k];
                bi += s.warp_tot[1][k];
                bf += s.warp_tot[2][k];
            }
            bv += sv - tv;
            bi += si - ti;
            bf += sf - tf;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                s.sc.rowV[tid * 4 + k] = bv + v[k];
                s.sc.rowI[tid * 4 + k] = bi + i[k];
                s.sc.rowF[tid * 4 + k] = bf + f[k];
            }
            if (tid == NT - 1) {
                const unsigned long long TV = bv + tv, TI = bi + ti;
                s.sc.rowV[1024] = bv + tv;
                s.sc.rowI[1024] = bi + ti;
                s.sc.rowF[1024] = bf + tf;
                // spans start on multiples of 16 vertices (640 B: emit_kernel's 32-vertex pieces are then whole 128-byte lines) and
                // 8 indices
                const unsigned long long padV = (TV + 15ull) & ~15ull, padI = (TI + 7ull) & ~7ull;
                const unsigned long long padF = ((unsigned long long)(bf + tf) + 31ull) & ~31ull; // emit_kernel: a warp never straddles chunks
                unsigned long long vb = 0, ib = 0, fb = 0;
                int emit = 0;
                if (TV | TI) {
                    const uint2 oc = p.reuse ? p.span_cap[span_slot] : make_uint2(0u, 0u);
                    if (p.reuse && padV <= oc.x && padI <= oc.y) {
                        // the new mesh fits the chunk's old reservation: rewrite it in place
                        vb = p.spans[span_slot].v_off;
                        ib = p.spans[span_slot].i_off;
                        fb = atomicAdd(&p.counters->f_alloc, padF);
                        if (!p.count_only) {
                            if (fb + padF <= p.f_cap) emit = 1;
                            else atomicOr(&p.counters->flags, kFlagArenaOverflow);
                        }
                    } else {
                        // re-meshes reserve some slack so that the next edit of the chunk can stay in place
                        const unsigned long long capV = p.reuse ? ((TV + TV / 8ull + 16ull + 15ull) & ~15ull) : padV;
                        const unsigned long long capI = p.reuse ? ((TI + TI / 8ull + 24ull + 7ull) & ~7ull) : padI;
                        release_span(p, span_slot);
                        vb = atomicAdd(&p.counters->v_alloc, capV);
                        ib = atomicAdd(&p.counters->i_alloc, capI);
                        fb = atomicAdd(&p.counters->f_alloc, padF);
                        if (p.span_cap) p.span_cap[span_slot] = make_uint2((uint32_t)capV, (uint32_t)capI);
                        if (!p.count_only) {
                            if (vb + capV <= p.v_cap && ib + capI <= p.i_cap && fb + padF <= p.f_cap) emit = 1;
                            else atomicOr(&p.counters->flags, kFlagArenaOverflow);
                        }
                    }
                } else {
                    release_span(p, span_slot);
                }
                bxw_mesh_span sp;
                sp.v_off = vb;
                sp.i_off = ib;
                sp.v_count = (uint32_t)TV;
                sp.i_count = (uint32_t)TI;
                p.spans[span_slot] = sp;
                s.vbase = vb;
                s.ibase = ib;
                s.fbase = fb;
                s.emit = emit;
            }
        }
        __syncthreads();
        PHASE_MARK(t_b1);
        PHASE_ADD(2, t_a1, t_b1);
        if (!s.emit) continue;
        PHASE_ADD(4, 0, 1);

        // ---------------- phase C: the chunk's face records, in emission order, tile by tile ----------------
        const uint32_t *center = p.blocks + (size_t)s.nbr[13] * kChunkVoxels;
        const unsigned long long fbase = s.fbase;
        const bool quads_only = LEAN || !has_shapes; // every side has 4 vertices / 6 indices
        if constexpr (!LEAN) {
            if (!quads_only) {
                // Chunks with slopes / corners: warp per row, lane per cell; every lane writes the finished records of its cell's
                // visible sides straight to the record buffer (a row's records are one contiguous run, so the scattered 16-byte
                // stores of a warp fall into a few lines). No staging tile, no second pass, no block barriers.
                PHASE_MARK(t_t1);
                for (int i = tid; i < 34 * 34; i += NT) s.sc.solid64[i] = solid_word(s, i);
                __syncthreads();
                // block ids of the row come from global memory (an L2 hit: phase A just streamed the chunk); the load for the
                // next row is issued a whole iteration ahead
                uint32_t datum_next = __ldg(center + warp * 32 + lane);
                for (uint32_t row = warp; row < 1024; row += NW) {
                    const uint32_t datum = datum_next;
                    if (row + NW < 1024u) datum_next = __ldg(center + (row + NW) * 32 + lane);
                    const uint32_t rf = s.sc.rowF[row];
                    if (s.sc.rowF[row + 1] - rf == 0) continue;
                    const int at = ((row >> 5) + 1) * PLANE + ((row & 31) + 1) * ROWB + XOFF + lane;
                    const uint32_t c = s.cls[at];
                    const uint32_t t = s.ctab[c];
                    const uint32_t vis = s.sc.vis[row * 32 + lane];
                    const uint32_t pk = cell_counts(s, vis, t);
                    uint32_t incl = pk;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t a = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += a;
                    }
                    const uint32_t excl = incl - pk;
                    if (vis) {
                        const uint32_t nb = neighbourhood27_words(s.sc.solid64, (int)(row >> 5), (int)(row & 31), lane); // warp-uniform loads
                        const uint32_t id = datum >> 16;
                        uint32_t vs = s.sc.rowV[row] + (excl & 1023u), is = s.sc.rowI[row] + ((excl >> 10) & 2047u);
                        uint4 *out = p.frec + fbase + rf + (excl >> 21);
                        const uint32_t codes = t >> 18;
                        const uint32_t a0 = row | ((uint32_t)lane << 10) | (c << 18), idlo = (id & 0xFFFu) << 20, idhi = (id >> 12) << 21;
#pragma unroll
                        for (int d = 0; d < 6; d++) {
                            if ((vis >> d) & 1u) {
                                const uint32_t code = (codes >> (2 * d)) & 3u;
                                *out++ = make_uint4(a0 | ((uint32_t)d << 15), vs | idlo, is | idhi, nb);
                                vs += (0x6430u >> (4u * code)) & 15u; // 3 / 4 / 6 vertices
                                is += (0x6630u >> (4u * code)) & 15u; // 3 / 6 / 6 indices
                            }
                        }
                    }
                }
                PHASE_MARK(t_t2);
                PHASE_ADD(7, t_t1, t_t2);
            }
        }
        for (uint32_t r0 = quads_only ? 0u : 1024u; r0 < 1024;) {
            PHASE_MARK(t_t0);
            // largest r1 with faces(r0..r1) <= FC and vertices(r0..r1) <= VC (a single row never exceeds 192 / 768). The test is
            // monotone in r1: every warp finds it with two 32-way votes (rows r0+1, r0+33, ... then the 32 rows behind the hit).
            const uint32_t f0 = s.sc.rowF[r0], v0 = s.sc.rowV[r0], i0 = s.sc.rowI[r0];
            uint32_t lo = 1024;
            if (!(s.sc.rowF[1024] - f0 <= (uint32_t)FC && s.sc.rowV[1024] - v0 <= (uint32_t)VC)) { // else: the rest fits one tile (terrain)
                auto fits = [&](uint32_t r) {
                    r = min(r, 1024u);
                    return s.sc.rowF[r] - f0 <= (uint32_t)FC && s.sc.rowV[r] - v0 <= (uint32_t)VC;
                };
                const uint32_t coarse = __ballot_sync(FULL, fits(r0 + 1u + 32u * (uint32_t)lane)); // lane 0 always fits
                const uint32_t base = r0 + 1u + 32u * (31u - (uint32_t)__clz(coarse));
                const uint32_t fine = __ballot_sync(FULL, fits(base + (uint32_t)lane));
                lo = min(base + (31u - (uint32_t)__clz(fine)), 1024u);
            }
            const uint32_t r1 = lo;
            const uint32_t tileF = s.sc.rowF[r1] - f0;
            if (tileF == 0) {
                r0 = r1;
                continue;
            }
            PHASE_MARK(t_t1);
            PHASE_ADD(6, t_t0, t_t1);
            // ---- C1: face records in emission order ----
            if (quads_only) {
                // thread per row (a surface chunk has a few hundred rows with ~3 faces each: a warp per row would idle 29 lanes)
                for (uint32_t row = r0 + tid; row < r1; row += NT) {
                    uint32_t f = s.sc.rowF[row] - f0;
                    if (s.sc.rowF[row + 1] - s.sc.rowF[row] == 0) continue;
                    unsigned long long vw[6];
                    row_visibility(s, row >> 5, row & 31, vw);
                    unsigned long long any = vw[0] | vw[1] | vw[2] | vw[3] | vw[4] | vw[5];
                    while (any) {
                        const int b = __ffsll((long long)any) - 1; // bit x+1
                        any &= any - 1;
#pragma unroll
                        for (int d = 0; d < 6; d++) {
                            if ((vw[d] >> b) & 1ull) {
                                s.sc.frecA[f] = row | ((uint32_t)(b - 1) << 10) | ((uint32_t)d << 15) | ((f * 4u) << 18);
                                f++;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            PHASE_MARK(t_t2);
            PHASE_ADD(7, t_t1, t_t2);
            // ---- C2: complete the records and hand them to emit_kernel (coalesced: thread = face) ----
            // (the side's block datum is the only global load -- an L2 hit, phase A just streamed the chunk --: the one of the
            // thread's next side is issued before this side's record is built)
            uint32_t a_next = 0, datum_next = 0;
            if (tid < tileF) {
                a_next = s.sc.frecA[tid];
                datum_next = __ldg(center + (a_next & 1023u) * 32 + ((a_next >> 10) & 31u));
            }
            for (uint32_t f = tid; f < tileF; f += NT) {
                const uint32_t a = a_next, datum = datum_next;
                if (f + NT < tileF) {
                    a_next = s.sc.frecA[f + NT];
                    datum_next = __ldg(center + (a_next & 1023u) * 32 + ((a_next >> 10) & 31u));
                }
                const uint32_t row = a & 1023u, x = (a >> 10) & 31u;
                const uint32_t y = row >> 5, z = row & 31u;
                const uint32_t id = datum >> 16;
                uint32_t cx;
                if (LEAN) { // no class table: a cube's class is 1 + its orientation
                    uint32_t e2 = 0;
                    cx = classify(datum, s, p.reg, e2);
                } else {
                    cx = s.cls[(y + 1) * PLANE + (z + 1) * ROWB + XOFF + (int)x];
                }
                const uint32_t nb = neighbourhood27(s, (int)y, (int)z, (int)x);
                const uint32_t is = f * 6u;
                p.frec[fbase + f0 + f] = make_uint4((a & 0x3FFFFu) | (cx << 18), (v0 + (a >> 18)) | ((id & 0xFFFu) << 20),
                                                    (i0 + is) | ((id >> 12) << 21), nb);
            }
            PHASE_MARK(t_t3);
            PHASE_ADD(8, t_t2, t_t3);
            __syncthreads(); // the face records are rewritten by the next tile
            PHASE_MARK(t_t4);
            PHASE_ADD(9, t_t3, t_t4);
            r0 = r1;
        }
        {
            const uint32_t TF = s.sc.rowF[1024];
            const uint32_t padF = (TF + 31u) & ~31u;
            if (tid < padF - TF) p.frec[fbase + TF + tid] = make_uint4(kFacePad, 0u, 0u, 0u);
            const unsigned long long vb = s.vbase, ib = s.ibase;
            // v_off, i_off (40 bits) and the pass's tag: in a pipelined pass emit_kernel polls the tag, so every record of the
            // chunk has to be visible before it
            const uint4 where = make_uint4((uint32_t)vb, (uint32_t)(vb >> 32), (uint32_t)ib, ((uint32_t)(ib >> 32) & 0xFFu) | (p.epoch << 8));
            if (p.pipelined) {
                __threadfence();
                __syncthreads();
            }
            for (uint32_t t = tid; t < (padF >> 5); t += NT) p.fslot[(fbase >> 5) + t] = where;
        }
        PHASE_MARK(t_c1);
        PHASE_ADD(3, t_b1, t_c1);
    }
    if (p.pipelined) { // this CTA's records are all published: tell emit_kernel (it stops when every mesh CTA has retired)
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(&p.counters->mesh_done, 1ull);
        }
    }
}

} // namespace

// One warp per interior (ix, iz) column of chunks: min / max surface height over the 34x34 columns under it, then the same
// trivial test as the mesher's fused phase A for every chunk of the vertical stack.
__global__ void __launch_bounds__(256) hmap_candidates_kernel(CandidateParams p) {
    const int nx = p.SX - 2, ny = p.SY - 2, nz = p.SZ - 2;
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (col >= nx * nz) return;
    const int ix = col % nx, iz = col / nx;
    int hmin = 0x7FFFFFFF, hmax = -0x7FFFFFFF;
    for (int i = lane; i < 34 * 34; i += 32) {
        const int zz = i / 34, xx = i - zz * 34;
        const int v = __ldg(p.hmap + (size_t)((iz + 1) * 32 - 1 + zz) * p.hmap_W + (ix + 1) * 32 - 1 + xx) >> 2;
        hmin = min(hmin, v);
        hmax = max(hmax, v);
    }
    hmin = __reduce_min_sync(0xFFFFFFFFu, hmin);
    hmax = __reduce_max_sync(0xFFFFFFFFu, hmax);
    for (int iy = lane; iy < ny; iy += 32) {
        const int y_lo = (p.cy_min + iy + 1) * 32 - 1;
        const bool all_void = y_lo > hmax && p.void_is_empty;
        const bool all_solid = (y_lo + 33 <= hmin) && p.gen_all_cubes;
        const uint32_t slot = (uint32_t)ix + (uint32_t)nx * ((uint32_t)iz + (uint32_t)nz * (uint32_t)iy);
        if (all_void || all_solid) {
            bxw_mesh_span sp;
            sp.v_off = sp.i_off = 0;
            sp.v_count = sp.i_count = 0;
            p.spans[slot] = sp;
            if (p.span_cap) p.span_cap[slot] = make_uint2(0u, 0u);
        } else {
            const uint32_t k = atomicAdd(p.count, 1u);
            p.work[k] = (uint32_t)(ix + 1) + (uint32_t)p.SX * ((uint32_t)(iz + 1) + (uint32_t)p.SZ * (uint32_t)(iy + 1));
            p.work_span[k] = slot;
        }
    }
}

void launch_hmap_candidates(const CandidateParams &p, cudaStream_t stream) {
    const int cols = (p.SX - 2) * (p.SZ - 2);
    hmap_candidates_kernel<<<(cols + 7) / 8, 256, 0, stream>>>(p);
}

// two CTAs per SM: 228 KB of shared memory per SM, 1 KB reserved per CTA
static_assert(sizeof(SmemT<false>) <= 113 * 1024 - 512, "mesh_kernel's shared memory no longer fits twice on an SM");
static_assert(sizeof(SmemT<true>) <= 44 * 1024, "the cubes-only mesh_kernel's shared memory no longer fits five times on an SM");

int mesh_kernel_max_ctas_per_sm() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mesh_kernel<false, false>, NT, sizeof(SmemT<false>));
    return n;
}

template <bool LEAN, bool CALL_SLOW>
static unsigned mesh_grid_t(const MeshParams &p, int sm_count, int max_per_sm) {
    // per device (a process may own contexts on several GPUs); cheap enough to repeat on every launch
    cudaFuncSetAttribute(mesh_kernel<LEAN, CALL_SLOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<LEAN>));
    // the same shared-memory carve-out for every kernel of the mesher: CTAs of mesh_kernel and emit_kernel share SMs in a
    // pipelined pass, and an SM cannot change its carve-out while it holds CTAs
    cudaFuncSetAttribute(mesh_kernel<LEAN, CALL_SLOW>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         p.pipelined ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mesh_kernel<LEAN, CALL_SLOW>, NT, sizeof(SmemT<LEAN>));
    if (per_sm < 1) per_sm = 1;
    if (max_per_sm > 0) per_sm = std::min(per_sm, max_per_sm);
    if (const char *e = getenv("BXW_EXP_MESH_CTAS")) per_sm = std::min(per_sm, std::max(1, atoi(e))); // experiments only
    unsigned grid = (unsigned)(sm_count * per_sm);
    if (grid > p.n_work) grid = p.n_work;
    return grid;
}

template <bool LEAN, bool CALL_SLOW>
static void launch_mesh_kernel_t(const MeshParams &p, int sm_count, cudaStream_t stream, int max_per_sm) {
    const unsigned grid = mesh_grid_t<LEAN, CALL_SLOW>(p, sm_count, max_per_sm);
    if (grid == 0) return;
    mesh_kernel<LEAN, CALL_SLOW><<<grid, NT, sizeof(SmemT<LEAN>), stream>>>(p);
}

unsigned mesh_kernel_grid(const MeshParams &p, int sm_count, bool lean, int max_per_sm) {
    return lean ? mesh_grid_t<true, true>(p, sm_count, max_per_sm) : mesh_grid_t<false, false>(p, sm_count, max_per_sm);
}

void launch_mesh_kernel(const MeshParams &p, int sm_count, cudaStream_t stream, bool lean, bool mostly_uniform, int max_per_sm) {
    if (const char *e = getenv("BXW_EXP_MESH_INLINE")) mostly_uniform = atoi(e) != 0; // experiments only
    if (!lean) launch_mesh_kernel_t<false, false>(p, sm_count, stream, max_per_sm);
    else if (mostly_uniform) launch_mesh_kernel_t<true, false>(p, sm_count, stream, max_per_sm);
    else launch_mesh_kernel_t<true, true>(p, sm_count, stream, max_per_sm);
}

} // namespace bxw
