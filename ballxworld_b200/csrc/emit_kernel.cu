// emit_kernel.cu — second half of mesh_from_chunk (src/client/render/voxmesh.rs:122-177): vertices and indices of every
// side that mesh_kernel found visible.
//
// mesh_kernel needs ~110 KB of shared memory per CTA (the 34^3 class table) and therefore runs 16 warps per SM, too few to
// hide the dependent table lookups of vertex construction. Here the work is one face per lane with no large shared state,
// so many warps are resident. Input: the face records mesh_kernel wrote in the reference's emission order
// (device_common.cuh: kFacePad); every chunk's records start on a multiple of 32, so the 32 faces of a warp belong to one
// chunk and their vertices / indices form one contiguous run of its arena span. The warp transposes that run through its
// private shared-memory slice and writes it with coalesced 16-byte stores.
//
// Bit-exactness: colours are as_rgba8(debug_color * ao) with ao = the sequential f32 product of 0.88 factors
// (voxmesh.rs:27-34,128-152; only the *number* of occluding samples matters), the barycentric colour offset is the in-order
// f32 sum over the side's vertices (voxmesh.rs:124,153,172-174), positions are exact integers (every vertex is a voxel
// corner: field = 16 * corner, voxmesh.rs:36-41).
#include "device_common.cuh"

#include <type_traits>

namespace bxw {
namespace {

constexpr int ENT = 256;      // threads per CTA
constexpr int ENW = ENT / 32; // warps per CTA
constexpr int EREGC = 64;     // registry entries cached in shared memory
constexpr int STG4 = 244;     // staging per warp in 16-byte units: 16 sides x 6 vertices x 40 B + one unit of misalignment
constexpr uint32_t FULL = 0xFFFFFFFFu;

struct ESmem {
    alignas(16) uint4 wstage[ENW][STG4];
    uint32_t stab[kNumClasses * 6 + 2];
    uint32_t aolut[8 * 27];           // AO sample mask per (vertex corner, AO normal code)
    uint32_t col_lut[EREGC * 8];      // as_rgba8(debug_color * ao_k) per cached block id and occluder count k
    float reg_color[EREGC * 3];
    float reg_texf[EREGC * 6];
    float ao_lut[8];
    uint16_t vdesc[kNumClasses * 36]; // vertex descriptors [class][dir][vertex] (tables.hpp)
};

// Warp copy of n 32-bit words from the warp's staging slice to global memory with 16-byte stores. The data was staged at
// stage16 + misal, misal = (dst / 4) & 3, so source and destination share their position inside a 16-byte unit.
__device__ __forceinline__ void warp_copy_out(uint32_t *dst, const uint32_t *stage16, uint32_t misal, uint32_t n, uint32_t lane) {
    uint32_t *d0 = dst - misal; // 16-byte aligned
    const uint32_t total = misal + n, nfull = total >> 2;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(stage16);
    uint4 *d4 = reinterpret_cast<uint4 *>(d0);
    // unit 0 may start with `misal` words that belong to the previous side: lanes 1..3 write its own words one by one
    if (misal) {
        if (lane >= misal && lane < 4u && lane < total) d0[lane] = stage16[lane];
    } else if (lane == 0 && nfull) {
        d4[0] = s4[0];
    }
#pragma unroll 2
    for (uint32_t k = lane + 1u; k < nfull; k += 32) __stcs(&d4[k], s4[k]); // the meshes are not read again on the device
    const uint32_t rem = total & 3u; // tail words of the last, partial unit
    if (lane < rem && (nfull || !misal)) d0[nfull * 4u + lane] = stage16[nfull * 4u + lane]; // (a short misaligned run was all head)
}

__global__ void __launch_bounds__(ENT, 4) emit_kernel(MeshParams p) {
    __shared__ ESmem s;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const MeshTables *T = p.tables;

    for (int i = tid; i < kNumClasses * 6; i += ENT) s.stab[i] = (&T->stab[0][0])[i];
    for (int i = tid; i < kNumClasses * 36; i += ENT) s.vdesc[i] = (&T->vdesc[0][0][0])[i];
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
    if (p.counters->flags & kFlagArenaOverflow) return;
    const unsigned long long n_batches = p.counters->f_alloc >> 5;
    uint4 *const stg = &s.wstage[warp][0];
    uint2 *const st2 = reinterpret_cast<uint2 *>(stg);
    uint32_t *const stw = reinterpret_cast<uint32_t *>(stg);

    // the records of the warp's next batch (one 16-byte load per lane + the batch's span slot) are loaded one batch ahead
    const unsigned long long bstep = (unsigned long long)gridDim.x * ENW;
    uint4 q = make_uint4(kFacePad, 0u, 0u, 0u);
    uint32_t qslot = 0;
    auto fetch = [&](unsigned long long bb) {
        if (bb < n_batches) {
            q = __ldg(p.frec + bb * 32ull + lane);
            qslot = __ldg(p.fslot + bb);
        }
    };
    fetch((unsigned long long)blockIdx.x * ENW + warp);
    for (unsigned long long b = (unsigned long long)blockIdx.x * ENW + warp; b < n_batches; b += bstep) {
        const uint32_t r0 = q.x, r1 = q.y, r2 = q.z, nb = q.w, slot0 = qslot;
        fetch(b + bstep);
        const bool act = r0 != kFacePad;          // padding only ever follows the faces of a chunk
        const uint32_t nbf = __popc(__ballot_sync(FULL, act));
        uint32_t vq = 0, iq = 0, nv = 0, ni = 0, texbits = 0, cell = 0, st = 0, id = 0;
        uint32_t fb0 = 0, fb1 = 0, fb2 = 0, fb3 = 0;
        uint32_t w_pos[6], w_col[6], w_uvi[6]; // w_uvi: u | v<<1 | barycentric flags<<2
#pragma unroll
        for (int j = 0; j < 6; j++) w_pos[j] = w_col[j] = w_uvi[j] = 0;
        if (act) {
            vq = r1 & 0xFFFFFu;
            iq = r2 & 0x1FFFFFu;
            id = (r1 >> 20) | ((r2 >> 21) << 12);
            st = s.stab[(r0 >> 18) * 6 + ((r0 >> 15) & 7u)];
            nv = (st >> 3) & 7u;
            ni = (st >> 6) & 7u;
        }
        // Per-face data. NVMAX = 4 when every side of the batch is a quad (all of a cube chunk; the common case everywhere):
        // the vertex loop is then free of predicates.
        auto face = [&](auto nvmax_tag) {
            constexpr uint32_t NVMAX = decltype(nvmax_tag)::value;
            const uint32_t row = r0 & 1023u, x = (r0 >> 10) & 31u, d = (r0 >> 15) & 7u, cx = r0 >> 18;
            const uint32_t y = row >> 5, z = row & 31u;
            const uint32_t ls = st & 7u, signs = st >> 9;
            float dc0, dc1, dc2, texid;
            const bool small_id = id < (uint32_t)EREGC;
            if (small_id) {
                dc0 = s.reg_color[id * 3];
                dc1 = s.reg_color[id * 3 + 1];
                dc2 = s.reg_color[id * 3 + 2];
                texid = s.reg_texf[id * 6 + ls]; // texture of the LOCAL side (voxmesh.rs:146)
            } else {
                dc0 = __ldg(p.reg.color + id * 3);
                dc1 = __ldg(p.reg.color + id * 3 + 1);
                dc2 = __ldg(p.reg.color + id * 3 + 2);
                texid = __ldg(p.reg.texf + id * 6 + ls);
            }
            texbits = __float_as_uint(texid);
            cell = x + 32u * z + 1024u * y;
            // as_wxyz10 of (ipos + voffset + 0.5)/32 (voxmesh.rs:36-41,139-145,165): every vertex sits on a voxel corner, so
            // each 10-bit field is 16 * corner coordinate (<= 512); w = (1/32*2) as u32 = 0
            const uint32_t pos0 = (x << 24) | (y << 14) | (z << 4);
            // per vertex: descriptor, number of occluding AO samples (voxmesh.rs:128-137; every meshed shape occludes), colour,
            // position, and the in-order barycentric colour sum
            const uint16_t *vds = &s.vdesc[cx * 36u + d * 6u];
            float b0 = 0.0f, b1 = 0.0f, b2 = 0.0f, b3 = 0.0f;
#pragma unroll
            for (uint32_t j = 0; j < NVMAX; j++) {
                if (NVMAX == 4 || j < nv) {
                    const uint32_t e = vds[j];
                    const uint32_t k = __popc(nb & s.aolut[(e & 7u) * 27u + (e >> 8)]);
                    const float sg = (float)((int)((signs >> (2 * j)) & 3u) - 1);
                    const float aj = s.ao_lut[k];
                    b0 = __fadd_rn(b0, __fmul_rn(sg, __fmul_rn(dc0, aj)));
                    b1 = __fadd_rn(b1, __fmul_rn(sg, __fmul_rn(dc1, aj)));
                    b2 = __fadd_rn(b2, __fmul_rn(sg, __fmul_rn(dc2, aj)));
                    b3 = __fadd_rn(b3, __fmul_rn(sg, 1.0f));
                    w_pos[j] = pos0 + (((e & 1u) << 24) | ((e & 2u) << 13) | ((e & 4u) << 2));
                    if (small_id) {
                        w_col[j] = s.col_lut[id * 8 + k];
                    } else {
                        const float c0f = __fmul_rn(dc0, aj), c1f = __fmul_rn(dc1, aj), c2f = __fmul_rn(dc2, aj);
                        w_col[j] = ((__float2uint_rz(__fmul_rn(c0f, 255.0f)) & 0xFFu) << 24) | ((__float2uint_rz(__fmul_rn(c1f, 255.0f)) & 0xFFu) << 16) |
                                   ((__float2uint_rz(__fmul_rn(c2f, 255.0f)) & 0xFFu) << 8) | 0xFFu;
                    }
                    w_uvi[j] = (e >> 3) & 31u;
                }
            }
            fb0 = __float_as_uint(b0);
            fb1 = __float_as_uint(b1);
            fb2 = __float_as_uint(b2);
            fb3 = __float_as_uint(b3);
        };
        const bool all_quads = __all_sync(FULL, !act || nv == 4u);
        if (all_quads) {
            if (act) face(std::integral_constant<uint32_t, 4>());
        } else {
            if (act) face(std::integral_constant<uint32_t, 6>());
        }
        // the chunk's arena span
        const unsigned long long v_off = p.spans[slot0].v_off, i_off = p.spans[slot0].i_off;
        bxw_vertex *const vout = p.varena + v_off;
        uint32_t *const iout = p.iarena + i_off;
        const uint32_t vq0 = __shfl_sync(FULL, vq, 0), iq0 = __shfl_sync(FULL, iq, 0);
        const uint32_t vqe = __shfl_sync(FULL, vq + nv, (int)nbf - 1), iqe = __shfl_sync(FULL, iq + ni, (int)nbf - 1);
        // record: position, color, texcoord (u, v, texture id), index | barycentric flags, barycentric_color_offset[4]
        // (src/client/render/voxrender.rs:39-49)
        auto tu = [&](int j) { return (w_uvi[j] & 1u) ? 0x3F800000u : 0u; };
        auto tv = [&](int j) { return (w_uvi[j] & 2u) ? 0x3F800000u : 0u; };
        auto ix = [&](int j) { return cell | ((w_uvi[j] >> 2) << 17); };
        if (all_quads && !(vq0 & 1u)) {
            // 32 quads = 128 vertices from a 16-byte aligned address: two passes of 16 quads (2560 B)
#pragma unroll
            for (uint32_t h = 0; h < 2; h++) {
                if (act && (lane >> 4) == h) {
                    uint4 *o = stg + (lane & 15u) * 10u;
                    o[0] = make_uint4(w_pos[0], w_col[0], tu(0), tv(0));
                    o[1] = make_uint4(texbits, ix(0), fb0, fb1);
                    o[2] = make_uint4(fb2, fb3, w_pos[1], w_col[1]);
                    o[3] = make_uint4(tu(1), tv(1), texbits, ix(1));
                    o[4] = make_uint4(fb0, fb1, fb2, fb3);
                    o[5] = make_uint4(w_pos[2], w_col[2], tu(2), tv(2));
                    o[6] = make_uint4(texbits, ix(2), fb0, fb1);
                    o[7] = make_uint4(fb2, fb3, w_pos[3], w_col[3]);
                    o[8] = make_uint4(tu(3), tv(3), texbits, ix(3));
                    o[9] = make_uint4(fb0, fb1, fb2, fb3);
                }
                __syncwarp();
                const uint32_t nfh = nbf > 16u * h ? min(16u, nbf - 16u * h) : 0u;
                uint4 *gq = reinterpret_cast<uint4 *>(vout + vq0 + h * 64u);
                for (uint32_t k = lane; k < nfh * 10u; k += 32) __stcs(&gq[k], stg[k]);
                __syncwarp();
            }
        } else {
            // mixed sides (3, 4 or 6 vertices): 16 sides (<= 96 vertices) per pass; a vertex is 40 bytes, so the run starts 0 or
            // 8 bytes into a 16-byte unit
#pragma unroll 1
            for (uint32_t h = 0; h < 2; h++) {
                if (h * 16u >= nbf) break;
                const int first = (int)(h * 16u), last = (int)min(h * 16u + 15u, nbf - 1u);
                const uint32_t vb = __shfl_sync(FULL, vq, first), ve = __shfl_sync(FULL, vq + nv, last);
                const uint32_t mis = (vb & 1u) * 2u; // v_off is a multiple of 4
                if (act && (lane >> 4) == h) {
                    uint2 *o = st2 + (mis >> 1) + (vq - vb) * 5u;
#pragma unroll
                    for (uint32_t j = 0; j < 6; j++) {
                        if (j < nv) {
                            o[j * 5 + 0] = make_uint2(w_pos[j], w_col[j]);
                            o[j * 5 + 1] = make_uint2(tu(j), tv(j));
                            o[j * 5 + 2] = make_uint2(texbits, ix(j));
                            o[j * 5 + 3] = make_uint2(fb0, fb1);
                            o[j * 5 + 4] = make_uint2(fb2, fb3);
                        }
                    }
                }
                __syncwarp();
                warp_copy_out(reinterpret_cast<uint32_t *>(vout + vb), stw, mis, (ve - vb) * 10u, lane);
                __syncwarp();
            }
        }
        (void)vqe;
        // indices: QUAD_INDICES [0,1,2,2,3,0] for quads (stdshapes.rs:198), 0..n-1 for the triangle lists; they refer to the
        // chunk's own vertex buffer (voff = vbuf.len(), voxmesh.rs:123,175)
        {
            const uint32_t imis = iq0 & 3u; // i_off is a multiple of 8
            if (act) {
                uint32_t *o = stw + imis + (iq - iq0);
                const uint32_t pat = nv == 4u ? 0x032210u : 0x543210u;
#pragma unroll
                for (uint32_t k = 0; k < 6; k++)
                    if (k < ni) o[k] = vq + ((pat >> (4 * k)) & 15u);
            }
            __syncwarp();
            warp_copy_out(iout + iq0, stw, imis, iqe - iq0, lane);
            __syncwarp();
        }
    }
}

} // namespace

void launch_emit_kernel(const MeshParams &p, int sm_count, cudaStream_t stream) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, emit_kernel, ENT, 0);
    if (per_sm < 1) per_sm = 1;
    emit_kernel<<<(unsigned)(sm_count * per_sm), ENT, 0, stream>>>(p);
}

} // namespace bxw
