// emit_kernel.cu — second half of mesh_from_chunk (src/client/render/voxmesh.rs:122-177): vertices and indices of every
// side that mesh_kernel found visible.
//
// The work here is table-driven vertex construction with no large shared state, split from mesh_kernel (whose shared memory
// holds the 34^3 neighbourhood) so that each half gets the kernel shape it needs. Input: the face records mesh_kernel wrote in
// the reference's emission order (device_common.cuh: kFacePad); every chunk's records start on a multiple of 32, so the 32
// sides of a batch belong to one chunk and their vertices / indices form one contiguous run of its arena span. A warp takes a
// batch: a side pass (lane = side) leaves a 32-byte side record per lane in shared memory, a vertex pass (lane = vertex, up
// to five vertices per lane) builds the 40-byte vertices in a warp-private staging buffer, from which the run -- normally
// the whole batch, ~5 KB -- leaves with one bulk copy; the index run is staged and copied out with 16-byte stores.
//
// Bit-exactness: colours are as_rgba8(debug_color * ao) with ao = the sequential f32 product of 0.88 factors
// (voxmesh.rs:27-34,128-152; only the *number* of occluding samples matters), the barycentric colour offset is the in-order
// f32 sum over the side's vertices (voxmesh.rs:124,153,172-174), positions are exact integers (every vertex is a voxel
// corner: field = 16 * corner, voxmesh.rs:36-41).
#include "device_common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace bxw {
namespace {

// Kernel shape. Measured on B200 (tools/exp_emit_variants.sh; hashed shapes / bench terrain, ms): the kernel is issue bound
// per scheduler and its stores like few, large pieces: 12 warps per SM in ONE CTA with passes of up to 160 vertices (a batch
// of 32 sides is ~128 vertices, so a batch leaves as one bulk copy of ~5 KB) 9.04 / 0.744; 16 warps x 64 vertices 9.0-9.1 /
// 0.85-0.86; 24 warps x 32 vertices (round 2's first shape) 9.73 / 0.90; 13, 14, 20, 24 warps are 10-20 % slower (the warp
// count has to divide evenly over the four schedulers, and more writers slow the write stream down).
#ifndef BXW_EMIT_THREADS
#define BXW_EMIT_THREADS 384
#endif
#ifndef BXW_EMIT_PV
#define BXW_EMIT_PV 160
#endif
constexpr int EREGC = 32;     // registry entries cached in shared memory (ids beyond take the per-side fetch path)
constexpr uint32_t FULL = 0xFFFFFFFFu;

struct alignas(16) SideRec { // what the vertex pass needs of a side (written by its lane in the side pass)
    uint32_t cell_id;  // cell (x + 32 z + 1024 y) | block id << 16
    uint32_t ks_tab;   // occluder count of vertex j at bits 3j (6 x 3 bits) | (class * 36 + world dir * 6) << 18
    uint32_t tex;      // texture id of the local side as f32 bits
    uint32_t pos0;     // as_wxyz10 of the cell's minimum corner: x << 24 | y << 14 | z << 4
    uint32_t fb[4];    // barycentric_color_offset
};

// 32-bit vertex descriptor [class][dir][vertex], derived from MeshTables::vdesc when a CTA starts: everything the vertex pass
// adds to the side's words sits at its final bit position.
//   bit 0 texcoord u, bit 1 texcoord v, bits 4 / 14 / 24 corner z / y / x (position delta), bits 17-19 barycentric flags (index
//   word), bits 5-12 index into aolut (corner * 27 + AO normal code)
constexpr uint32_t kVdPos = (1u << 24) | (1u << 14) | (1u << 4);
constexpr uint32_t kVdFlags = 7u << 17;

template <int ENT, int PV>
struct ESmemT {
    static constexpr int ENW = ENT / 32;                 // warps per CTA
    static constexpr int STGW = (PV * 10 + 2 + 3) / 4 * 4; // staging words per warp: PV vertices x 10 words + 2 of misalignment, rounded to 16 bytes
    static_assert(PV % 16 == 0 && STGW >= 196, "a pass ends on a multiple of 16 vertices; the index run of a batch is staged in the same buffer");
    alignas(16) uint32_t wstage[ENW][2][STGW]; // two staging buffers per warp: one is filled while the bulk copy drains the other
    SideRec rec[ENW][32];
    float4 bigreg[ENW][32];           // debug_color + local-side texture id of sides whose block id is not in the cached registry
    uint8_t vmap[ENW][192];           // vertex of the batch -> side (lane) << 3 | vertex of the side
    uint32_t stab[kNumClasses * 6 + 2];
    uint32_t aolut[8 * 27];           // AO sample mask per (vertex corner, AO normal code)
    uint32_t col_lut[EREGC * 8];      // as_rgba8(debug_color * ao_k) per cached block id and occluder count k
    float reg_color[EREGC * 3];
    float reg_texf[EREGC * 6];
    float ao_lut[8];
    uint32_t vd32[kNumClasses * 36];  // vertex descriptors [class][dir][vertex] (above)
};

// Warp copy of n 32-bit words (n even) from the warp's staging slice to global memory with 16-byte stores. The data was staged
// at stage16 + misal, misal = (dst / 4) & 3 (0 or 2 here), so source and destination share their position inside a 16-byte unit.
__device__ __forceinline__ void warp_copy_out(uint32_t *dst, const uint32_t *stage16, uint32_t misal, uint32_t n, uint32_t lane) {
    uint32_t *d0 = dst - misal; // 16-byte aligned
    const uint32_t total = misal + n, nfull = total >> 2;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(stage16);
    uint4 *d4 = reinterpret_cast<uint4 *>(d0);
    // unit 0 may start with `misal` words that belong to the previous side: its own words go out one by one
    if (misal) {
        if (lane >= misal && lane < 4u && lane < total) d0[lane] = stage16[lane];
    } else if (lane == 0 && nfull) {
        __stcs(&d4[0], s4[0]);
    }
#pragma unroll 3
    for (uint32_t k = lane + 1u; k < nfull; k += 32) __stcs(&d4[k], s4[k]); // the meshes are not read again on the device
    const uint32_t rem = total & 3u; // tail words of the last, partial unit
    if (lane < rem && (nfull || !misal)) d0[nfull * 4u + lane] = stage16[nfull * 4u + lane]; // (a short misaligned run was all head)
}

// Registry entry of a block id beyond the entries cached in shared memory (rare). Not inlined on purpose: inlined, the loads
// become predicated instructions on the common path, and the first use of their (unused) result waits on the same scoreboard
// as the record prefetch of the next batch.
__device__ __noinline__ void fetch_big_registry_entry(const DeviceRegistry reg, uint32_t id, uint32_t ls, float4 *out) {
    float4 bc;
    bc.x = __ldg(reg.color + id * 3);
    bc.y = __ldg(reg.color + id * 3 + 1);
    bc.z = __ldg(reg.color + id * 3 + 2);
    bc.w = __ldg(reg.texf + id * 6 + ls);
    *out = bc;
}

// One vertex pass leaves the warp's staging buffer: n words (10 per vertex) staged at stage16 + misal (misal 0 or 2, the run
// starts 8 bytes into a 16-byte unit when the first vertex index is odd). The whole 16-byte units go out as ONE bulk copy
// shared -> global issued by lane 0 (cp.async.bulk, the TMA engine: no LDS / STG wavefronts on the L1 path, which is what
// bounded this kernel); the 8-byte halves at either end are ordinary stores. The caller has fenced the staging writes towards
// the async proxy and synchronised the warp.
__device__ __forceinline__ void warp_bulk_out_vertices(uint32_t *dst, const uint32_t *stage16, uint32_t misal, uint32_t n, uint32_t lane) {
    const uint32_t total = misal + n, nfull = total >> 2;
    const uint32_t first = misal ? 1u : 0u;
    if (lane == 0u) {
#ifndef BXW_EXP_NO_STORE
        if (nfull > first) {
#else
        if (nfull > first && dst == nullptr) {
#endif
            const uint32_t bytes = (nfull - first) * 16u;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst - misal + first * 4u),
                         "r"((uint32_t)__cvta_generic_to_shared(stage16 + first * 4u)), "r"(bytes)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // lane 1: head half (run starts odd), lane 2: tail half (run ends odd). Pieces end on multiples of 16 vertices, so only a
    // batch's first piece can start odd and only its last can end odd: the (warp-uniform) test skips this for the others.
    if (misal | (total & 2u)) {
        const uint32_t w = lane == 1u ? 2u : nfull * 4u;
        if ((lane == 1u && misal) || (lane == 2u && (total & 2u)))
            __stcs(reinterpret_cast<uint2 *>(dst - misal + w), *reinterpret_cast<const uint2 *>(stage16 + w));
    }
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ENT threads per CTA; PV vertices per pass (PV / 32 per lane).
// PIPE = false: launched after mesh_kernel has finished, over the f_alloc / 32 batches it left.
// PIPE = true: launched NEXT TO mesh_kernel (another stream; mesh_kernel is compute / latency bound, this kernel memory bound):
// a batch is consumed as soon as its chunk's record block is published -- mesh_kernel writes the batch's fslot entry last,
// tagged with the pass's epoch, after a fence -- and the loop ends when every mesh CTA has retired (counters->mesh_done) and
// the batch index has passed the final f_alloc.
template <bool PIPE, int ENT, int PV>
__global__ void __launch_bounds__(ENT, PIPE ? 4 : 1) emit_kernel(MeshParams p) {
    using ESmem = ESmemT<ENT, PV>;
    constexpr int ENW = ESmem::ENW, STGW = ESmem::STGW;
    extern __shared__ __align__(16) unsigned char esmem_raw[];
    ESmem &s = *reinterpret_cast<ESmem *>(esmem_raw);
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const MeshTables *T = p.tables;

    for (int i = tid; i < kNumClasses * 6; i += ENT) s.stab[i] = (&T->stab[0][0])[i];
    for (int i = tid; i < kNumClasses * 36; i += ENT) {
        const uint32_t e = (&T->vdesc[0][0][0])[i];
        s.vd32[i] = ((e >> 3) & 3u) | ((e & 4u) << 2) | ((e & 2u) << 13) | ((e & 1u) << 24) | (((e >> 5) & 7u) << 17) |
                    (((e & 7u) * 27u + (e >> 8)) << 5);
    }
    for (int i = tid; i < 8 * 27; i += ENT) s.aolut[i] = T->aolut[i];
    if (tid < 8) s.ao_lut[tid] = T->ao_lut[tid];
    for (int i = tid; i < EREGC * 3; i += ENT) s.reg_color[i] = (uint32_t)(i / 3) < p.reg.n ? p.reg.color[i] : 0.0f;
    for (int i = tid; i < EREGC * 6; i += ENT) s.reg_texf[i] = (uint32_t)(i / 6) < p.reg.n ? p.reg.texf[i] : 0.0f;
    for (int i = tid; i < EREGC * 8; i += ENT) {
        const uint32_t id = (uint32_t)i >> 3;
        uint32_t col = 0;
        if (id < p.reg.n) { // as_rgba8 (voxmesh.rs:29-34,147-152); Rust's `as u32` saturates like cvt.rzi.u32.f32
            const float ao = T->ao_lut[i & 7];
            const float c0f = __fmul_rn(p.reg.color[id * 3], ao), c1f = __fmul_rn(p.reg.color[id * 3 + 1], ao), c2f = __fmul_rn(p.reg.color[id * 3 + 2], ao);
            col = ((__float2uint_rz(__fmul_rn(c0f, 255.0f)) & 0xFFu) << 24) | ((__float2uint_rz(__fmul_rn(c1f, 255.0f)) & 0xFFu) << 16) |
                  ((__float2uint_rz(__fmul_rn(c2f, 255.0f)) & 0xFFu) << 8) | 0xFFu;
        }
        s.col_lut[i] = col;
    }
    __syncthreads();

    // a chunk that did not fit left its records unwritten: the host grows the arenas and runs both kernels again
    if (!PIPE && (p.counters->flags & kFlagArenaOverflow)) return;
    // (pipelined: the final count is not known yet; no batch can lie beyond the record buffer)
    const unsigned long long n_batches = PIPE ? (p.f_cap >> 5) : (p.counters->f_alloc >> 5);
    uint32_t *const stw = &s.wstage[warp][0][0];
    uint32_t npass = 0; // vertex passes of this warp so far (selects the staging buffer)
    SideRec *const rec = &s.rec[warp][0];
    uint8_t *const vmap = &s.vmap[warp][0];

    // the records of the warp's next batch (one 16-byte load per lane + the batch's span slot) are loaded one batch ahead
    const unsigned long long bstep = (unsigned long long)gridDim.x * ENW;
    // NB these two are the ONLY global loads inside the loop: a wait on any other load would also wait for the prefetch behind
    // it (scoreboards count in order), i.e. expose a full memory latency per batch.
    uint4 q = make_uint4(kFacePad, 0u, 0u, 0u);
    uint4 qw = make_uint4(0u, 0u, 0u, 0u);
    bool have = false; // PIPE: q / qw hold the (published) batch of the coming iteration
    auto fetch = [&](unsigned long long bb) { // bb < n_batches
        if (PIPE) {
            // only what is published already may be read ahead (the tag is the last word mesh_kernel writes for a batch)
            have = (ld_acquire_u32(&p.fslot[bb].w) >> 8) == p.epoch;
            if (have) {
                q = __ldcg(p.frec + bb * 32ull + lane);
                qw = __ldcg(p.fslot + bb);
            }
        } else {
            q = __ldg(p.frec + bb * 32ull + lane);
            qw = __ldg(p.fslot + bb);
        }
    };
    if ((unsigned long long)blockIdx.x * ENW + warp < n_batches) fetch((unsigned long long)blockIdx.x * ENW + warp);
    for (unsigned long long b = (unsigned long long)blockIdx.x * ENW + warp; b < n_batches; b += bstep) {
        if (PIPE && !have) {
            // wait for the batch to be published, or for the end of production
            const long long t_start = clock64();
            int stop = 0;
            for (;;) {
                if ((ld_acquire_u32(&p.fslot[b].w) >> 8) == p.epoch) break;
                if (ld_acquire_u64(&p.counters->mesh_done) == p.mesh_ctas) {
                    // every mesh CTA has retired: f_alloc is final. A batch below it that is still untagged belongs to a chunk
                    // that did not fit the arenas (the host grows them and runs the pass again).
                    if ((ld_acquire_u32(&p.fslot[b].w) >> 8) == p.epoch) break;
                    stop = 1;
                    break;
                }
                if (clock64() - t_start > (1ll << 32)) { // ~2 s: something is wrong; fail loudly instead of hanging the device
                    if (lane == 0) atomicOr(&p.counters->flags, kFlagPipelineTimeout);
                    stop = 1;
                    break;
                }
                __nanosleep(256);
            }
            if (stop) break;
            q = __ldcg(p.frec + b * 32ull + lane);
            qw = __ldcg(p.fslot + b);
        }
        const uint32_t r0 = q.x, r1 = q.y, r2 = q.z, nb = q.w;
        // the chunk's arena span (fslot: v_off as 64 bits, i_off as 40 bits, the pass's epoch in the top 24)
        const unsigned long long v_off = (unsigned long long)qw.x | ((unsigned long long)qw.y << 32),
                                 i_off = (unsigned long long)qw.z | ((unsigned long long)(qw.w & 0xFFu) << 32);
        if (b + bstep < n_batches) fetch(b + bstep);
        else have = false;
        const bool act = r0 != kFacePad;          // padding only ever follows the faces of a chunk
        const uint32_t nbf = __popc(__ballot_sync(FULL, act));
        // ---------------- side pass: lane = side (voxmesh.rs:122-153) ----------------
        uint32_t vq = 0, iq = 0, nv = 0, ni = 0;
        if (act) {
            vq = r1 & 0xFFFFFu;
            iq = r2 & 0x1FFFFFu;
            const uint32_t id = (r1 >> 20) | ((r2 >> 21) << 12);
            const uint32_t tabi = (r0 >> 18) * 6u + ((r0 >> 15) & 7u); // class * 6 + world dir
            const uint32_t st = s.stab[tabi];
            nv = (st >> 3) & 7u;
            ni = (st >> 6) & 7u;
            const uint32_t row = r0 & 1023u, x = (r0 >> 10) & 31u;
            const uint32_t ls = st & 7u, signs = st >> 9;
            // block ids beyond the cached registry entries (rare) fetch theirs into the warp's slots first, so that the common
            // path holds no global load
            const bool big = id >= (uint32_t)EREGC;
            if (big) fetch_big_registry_entry(p.reg, id, ls, &s.bigreg[warp][lane]);
            const float *rc = big ? &s.bigreg[warp][lane].x : &s.reg_color[id * 3];
            const float dc0 = rc[0], dc1 = rc[1], dc2 = rc[2];
            const float texid = big ? rc[3] : s.reg_texf[id * 6 + ls]; // texture of the LOCAL side (voxmesh.rs:146)
            // per vertex: number of occluding AO samples (voxmesh.rs:128-137; every meshed shape occludes) and the in-order
            // barycentric colour sum (voxmesh.rs:124,153,172-174): sum += sign as f32 * colour with sign in {-1, 0, 1}. -1 * c is
            // exactly -c and a + (-c) is a - c; a 0 * c term is +-0 and never changes the sum (which starts at +0 and cannot
            // become -0 under round-to-nearest), so the multiplications by the sign fold into add / subtract / skip. The fourth
            // component sums sign * 1.0: small integers, exact in any order.
            const uint32_t *vds = &s.vd32[tabi * 6u];
            float b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;
            uint32_t ks = 0;
#pragma unroll
            for (uint32_t j = 0; j < 6; j++) {
                if (j < nv) {
                    const uint32_t k = __popc(nb & s.aolut[(vds[j] >> 5) & 255u]);
                    const uint32_t code = (signs >> (2 * j)) & 3u; // sign + 1
                    const float aj = s.ao_lut[k];
                    ks |= k << (3 * j);
                    if (code != 1u) {
                        const uint32_t neg = (code == 0u) ? 0x80000000u : 0u;
                        b0 = __fadd_rn(b0, __uint_as_float(__float_as_uint(__fmul_rn(dc0, aj)) ^ neg));
                        b1 = __fadd_rn(b1, __uint_as_float(__float_as_uint(__fmul_rn(dc1, aj)) ^ neg));
                        b2 = __fadd_rn(b2, __uint_as_float(__float_as_uint(__fmul_rn(dc2, aj)) ^ neg));
                    }
                }
            }
            // sum of (code - 1) over the side's vertices (codes past nv are 0 in the table)
            const float b3 = (float)((int)(__popc(signs & 0x555u) + 2 * __popc(signs & 0xAAAu)) - (int)nv);
            SideRec sr;
            sr.cell_id = (x + 32u * (row & 31u) + 1024u * (row >> 5)) | (id << 16);
            sr.ks_tab = ks | ((tabi * 6u) << 18);
            sr.tex = __float_as_uint(texid);
            sr.pos0 = (x << 24) | ((row >> 5) << 14) | ((row & 31u) << 4);
            sr.fb[0] = __float_as_uint(b0);
            sr.fb[1] = __float_as_uint(b1);
            sr.fb[2] = __float_as_uint(b2);
            sr.fb[3] = __float_as_uint(b3);
            rec[lane] = sr;
        }
        const uint32_t vq0 = __shfl_sync(FULL, vq, 0), iq0 = __shfl_sync(FULL, iq, 0);
        const uint32_t vqe = __shfl_sync(FULL, vq + nv, (int)nbf - 1), iqe = __shfl_sync(FULL, iq + ni, (int)nbf - 1);
        if (act) {
            const uint32_t vs = vq - vq0; // the sides of a batch are consecutive in the chunk's vertex buffer
#pragma unroll
            for (uint32_t j = 0; j < 6; j++)
                if (j < nv) vmap[vs + j] = (uint8_t)((lane << 3) | j);
        }
        __syncwarp();
        bxw_vertex *const vout = p.varena + v_off;
        uint32_t *const iout = p.iarena + i_off;
        // ---------------- vertex pass: lane = vertex ----------------
        // record: position, color, texcoord (u, v, texture id), index | barycentric flags, barycentric_color_offset[4]
        // (src/client/render/voxrender.rs:39-49). A vertex is 40 bytes and v_off is a multiple of 4 vertices, so an even vertex
        // index is 16-byte aligned: passes are cut at even indices (the first takes 31 vertices when the batch starts odd) and
        // every pass leaves as whole 16-byte units except for one 8-byte half at the very start / end of the batch's run. A pass
        // takes up to PV vertices, two per lane.
        const uint32_t nvt = vqe - vq0;
        for (uint32_t tb = 0; tb < nvt;) {
            const uint32_t mis = ((vq0 + tb) & 1u) * 2u;
            // passes end on multiples of 16 vertices of the chunk's span (which starts on a 128-byte line): every piece but the
            // first and last of a batch is then ten whole lines
            const uint32_t cnt = min(nvt - tb, (uint32_t)PV - ((vq0 + tb) & 15u));
            uint32_t *const buf = stw + (npass & 1u) * STGW;
            uint2 *const st2 = reinterpret_cast<uint2 *>(buf);
            npass++;
            // the bulk copy that last read this buffer (two passes ago) must be through with it
            if (lane == 0u) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (uint32_t h = 0; h < ((uint32_t)PV + 31u) / 32u; h++) {
                const uint32_t v = lane + 32u * h; // vertex of the pass
#ifdef BXW_EXP_NO_COMPUTE
                if (v < cnt) {
                    uint2 *o = st2 + (mis >> 1) + v * 5u;
                    o[0] = make_uint2(tb, lane);
                    o[1] = make_uint2(tb, lane);
                    o[2] = make_uint2(tb, lane);
                    o[3] = make_uint2(tb, lane);
                    o[4] = make_uint2(tb, lane);
                }
                if (false) {
#else
                if (v < cnt) {
#endif
                    const uint32_t m = vmap[tb + v];
                    const SideRec *sr = &rec[m >> 3];
                    const uint4 a = *reinterpret_cast<const uint4 *>(sr);
                    const uint4 fbv = *reinterpret_cast<const uint4 *>(&sr->fb[0]);
                    const uint32_t j = m & 7u;
                    const uint32_t cell = a.x & 0x7FFFu, id = a.x >> 16;
                    const uint32_t vd = s.vd32[(a.y >> 18) + j];
                    const uint32_t k = (a.y >> (3u * j)) & 7u;
                    uint32_t col;
                    if (id < (uint32_t)EREGC) {
                        col = s.col_lut[id * 8 + k];
                    } else { // as_rgba8(debug_color * ao) (voxmesh.rs:29-34,147-152)
                        const float aj = s.ao_lut[k];
                        const float4 bc = s.bigreg[warp][m >> 3];
                        const float c0f = __fmul_rn(bc.x, aj), c1f = __fmul_rn(bc.y, aj), c2f = __fmul_rn(bc.z, aj);
                        col = ((__float2uint_rz(__fmul_rn(c0f, 255.0f)) & 0xFFu) << 24) | ((__float2uint_rz(__fmul_rn(c1f, 255.0f)) & 0xFFu) << 16) |
                              ((__float2uint_rz(__fmul_rn(c2f, 255.0f)) & 0xFFu) << 8) | 0xFFu;
                    }
                    // as_wxyz10 of (ipos + voffset + 0.5)/32 (voxmesh.rs:36-41,139-145,165): every vertex sits on a voxel corner, so
                    // each 10-bit field is 16 * corner coordinate (<= 512); w = (1/32*2) as u32 = 0
                    uint2 *o = st2 + (mis >> 1) + v * 5u; // lanes 40 bytes apart: conflict-free 8-byte stores
                    o[0] = make_uint2(a.w + (vd & kVdPos), col);
                    o[1] = make_uint2((vd & 1u) * 0x3F800000u, ((vd >> 1) & 1u) * 0x3F800000u); // texcoord u, v: 0.0f or 1.0f
                    o[2] = make_uint2(a.z, cell | (vd & kVdFlags));
                    o[3] = make_uint2(fbv.x, fbv.y);
                    o[4] = make_uint2(fbv.z, fbv.w);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // staging writes -> visible to the bulk copy engine
            __syncwarp();
            warp_bulk_out_vertices(reinterpret_cast<uint32_t *>(vout + vq0 + tb), buf, mis, cnt * 10u, lane);
            tb += cnt;
        }
        // indices: QUAD_INDICES [0,1,2,2,3,0] for quads (stdshapes.rs:198), 0..n-1 for the triangle lists; they refer to the
        // chunk's own vertex buffer (voff = vbuf.len(), voxmesh.rs:123,175)
        {
            const uint32_t imis = iq0 & 3u; // i_off is a multiple of 8
            // staged in the buffer the next vertex pass would take: only the last pass's bulk copy (other buffer) may be in flight
            uint32_t *const ist = stw + (npass & 1u) * STGW;
            if (lane == 0u) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            if (act) {
                uint32_t *o = ist + imis + (iq - iq0);
                const uint32_t pat = nv == 4u ? 0x032210u : 0x543210u;
#pragma unroll
                for (uint32_t k = 0; k < 6; k++)
                    if (k < ni) o[k] = vq + ((pat >> (4 * k)) & 15u);
            }
            __syncwarp();
            warp_copy_out(iout + iq0, ist, imis, iqe - iq0, lane);
            __syncwarp();
        }
    }
    // bulk copies still in flight read this CTA's shared memory
    if (lane == 0u) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

} // namespace

// the sequential pass: one CTA per SM; the pipelined pass has to share the SMs with mesh_kernel: round 2's first shape
using ESmemSeq = ESmemT<BXW_EMIT_THREADS, BXW_EMIT_PV>;
using ESmemPipe = ESmemT<256, 32>;
static_assert(sizeof(ESmemSeq) <= 226 * 1024, "emit_kernel's shared memory no longer fits an SM");
static_assert(sizeof(ESmemPipe) <= 56 * 1024, "the pipelined emit_kernel no longer fits beside mesh_kernel");

void launch_emit_kernel(const MeshParams &p, int sm_count, cudaStream_t stream) {
    auto kern = emit_kernel<false, BXW_EMIT_THREADS, BXW_EMIT_PV>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ESmemSeq));
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BXW_EMIT_THREADS, sizeof(ESmemSeq));
    if (getenv("BXW_DEBUG_OCCUPANCY"))
        fprintf(stderr, "emit_kernel: %d CTAs/SM possible, %zu bytes of shared memory per CTA\n", per_sm, sizeof(ESmemSeq));
    per_sm = std::min(per_sm, 1);
    if (per_sm < 1) per_sm = 1;
    kern<<<(unsigned)(sm_count * per_sm), BXW_EMIT_THREADS, sizeof(ESmemSeq), stream>>>(p);
}

void launch_emit_kernel_pipelined(const MeshParams &p, int sm_count, int ctas_per_sm, cudaStream_t stream) {
    auto kern = emit_kernel<true, 256, 32>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ESmemPipe));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    kern<<<(unsigned)(sm_count * ctas_per_sm), 256, sizeof(ESmemPipe), stream>>>(p);
}

} // namespace bxw
