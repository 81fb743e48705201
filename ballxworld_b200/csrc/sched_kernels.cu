// sched_kernels.cu — the load-anchor scheduler on the device: World::recalculate_load_deltas
// (bxw_world/src/worldmgr.rs:633-732) for a region box.
//
// The reference walks every allocated chunk (what to unload) and every coordinate inside an anchor sphere
// (iter_coords_to_load, :734-742; what to load), gates mesh requests on the block data of the 27-neighbourhood
// (can_request_load, :761-790), sorts both lists by the squared distance to the nearest anchor and interleaves them
// (unload, load, unload, load ...; the list is reversed and popped from the back, i.e. consumed nearest first, :699-731).
// Here one thread classifies one chunk of the box against all anchors; a counting sort over the squared distance (one bin per
// value, so the order is exact; ties are in arbitrary order, as they are in the reference's hash maps) places every delta.
#include "sched_kernels.cuh"

namespace bxw {
namespace {

__device__ __forceinline__ bool covers(const bxw_load_anchor &a, int x, int y, int z, int &d2) {
    const int dx = x - a.cx, dy = y - a.cy, dz = z - a.cz;
    d2 = dx * dx + dy * dy + dz * dz;
    return (unsigned)d2 <= a.radius * a.radius; // does_anchor_load_coords (worldmgr.rs:744-746)
}

// per chunk: what (if anything) has to happen, and its sort key. code: bits 0-1 kinds (1 block data, 2 mesh), bit 2 unload
__global__ void __launch_bounds__(256) delta_classify_kernel(SchedParams p) {
    const uint32_t c = blockIdx.x * 256u + threadIdx.x;
    if (c >= p.n_storage) return;
    const int sx = (int)(c % (uint32_t)p.SX), sz = (int)((c / (uint32_t)p.SX) % (uint32_t)p.SZ), sy = (int)(c / (uint32_t)(p.SX * p.SZ));
    const int x = p.cx_min + sx, y = p.cy_min + sy, z = p.cz_min + sz;
    const bool interior = sx >= 1 && sx <= p.SX - 2 && sy >= 1 && sy <= p.SY - 2 && sz >= 1 && sz <= p.SZ - 2;
    const uint32_t slot = interior ? (uint32_t)((sx - 1) + (p.SX - 2) * ((sz - 1) + (p.SZ - 2) * (sy - 1))) : 0u;
    const bool blk = p.blk_loaded[c] != 0, msh = interior && p.mesh_loaded[slot] != 0;
    bool want_blk = false, want_msh = false;
    int dmin_all = 0x7FFFFFFF, dmin_blk = 0x7FFFFFFF, dmin_msh = 0x7FFFFFFF;
    for (uint32_t k = 0; k < p.n_anchors; k++) {
        const bxw_load_anchor a = p.anchors[k];
        int d2;
        const bool in = covers(a, x, y, z, d2);
        dmin_all = min(dmin_all, d2);
        if (in) { // the block handler loads for every anchor, the mesh handler for anchors with load_mesh
            want_blk = true;
            dmin_blk = min(dmin_blk, d2);
            if (a.load_mesh) {
                want_msh = true;
                dmin_msh = min(dmin_msh, d2);
            }
        }
    }
    uint32_t code = 0;
    int dist = 0;
    const uint32_t un = ((blk && !want_blk) ? 1u : 0u) | ((msh && !want_msh) ? 2u : 0u);
    if (un) { // chunks_to_unload: loaded kinds no anchor asks for, keyed by the distance to the nearest anchor (:636-661)
        code = un | 4u;
        dist = p.n_anchors ? dmin_all : 0x7FFFFFFF;
    } else if (want_blk && !blk) {
        code = 1u; // block data has no dependency (can_request_load: true)
        dist = dmin_blk;
    } else if (want_msh && !msh && interior && blk) {
        // the mesh needs the block data of the chunk and its 26 neighbours loaded NOW (can_request_load, :761-790)
        bool ok = true;
        for (int t = 0; t < 27; t++) {
            const int rx = t % 3 - 1, rz = (t / 3) % 3 - 1, ry = t / 9 - 1;
            ok &= p.blk_loaded[(size_t)((sx + rx) + p.SX * ((sz + rz) + p.SZ * (sy + ry)))] != 0;
        }
        if (ok) {
            code = 2u;
            dist = dmin_msh;
        }
    }
    p.code[c] = (uint8_t)code;
    if (code) {
        const uint32_t bin = min((uint32_t)dist, p.n_bins - 1u);
        p.key[c] = (uint32_t)dist;
        atomicAdd(&p.hist[(code & 4u) ? bin : p.n_bins + bin], 1u); // two histograms: unloads, loads
    }
}

// exclusive scan of both histograms in place (one CTA; n = 2 * n_bins); totals[0] = unloads, totals[1] = loads
__global__ void __launch_bounds__(1024) delta_scan_kernel(uint32_t *hist, uint32_t n_bins, uint32_t *totals) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry, pass_total;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int h = 0; h < 2; h++) {
        uint32_t *a = hist + (size_t)h * n_bins;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < n_bins; base += 1024u) {
            const uint32_t i = base + threadIdx.x;
            const uint32_t v = i < n_bins ? a[i] : 0u;
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
                warp_tot[lane] = wi - wt;
                if (lane == 31u) pass_total = wi;
            }
            __syncthreads();
            if (i < n_bins) a[i] = carry + warp_tot[warp] + incl - v;
            __syncthreads();
            if (threadIdx.x == 0) carry += pass_total;
            __syncthreads();
        }
        if (threadIdx.x == 0) totals[h] = carry;
        __syncthreads();
    }
}

// every delta to its place: unload j at 2j (j < n_load) or n_load + j, load j at 2j + 1 (j < n_unload) or n_unload + j --
// the reference's interleaving (:699-731). Also the per-kind work lists, each nearest first.
__global__ void __launch_bounds__(256) delta_scatter_kernel(SchedParams p) {
    const uint32_t c = blockIdx.x * 256u + threadIdx.x;
    if (c >= p.n_storage) return;
    const uint32_t code = p.code[c];
    if (!code) return;
    const uint32_t n_un = p.totals[0], n_ld = p.totals[1];
    const uint32_t bin = min(p.key[c], p.n_bins - 1u);
    const bool un = (code & 4u) != 0;
    const uint32_t j = atomicAdd(&p.hist[un ? bin : p.n_bins + bin], 1u); // (the scanned histogram doubles as the fill cursor)
    const uint32_t pos = un ? (j < n_ld ? 2u * j : n_ld + j) : (j < n_un ? 2u * j + 1u : n_un + j);
    const int sx = (int)(c % (uint32_t)p.SX), sz = (int)((c / (uint32_t)p.SX) % (uint32_t)p.SZ), sy = (int)(c / (uint32_t)(p.SX * p.SZ));
    bxw_chunk_delta d;
    d.cx = p.cx_min + sx;
    d.cy = p.cy_min + sy;
    d.cz = p.cz_min + sz;
    d.kinds = code & 3u;
    d.unload = un ? 1u : 0u;
    d.min_anchor_distance = (int32_t)p.key[c];
    p.deltas[pos] = d;
    p.order[pos] = c;
}

// Walk the ordered list (at most `limit` deltas) and split it into the executable lists, keeping the order: chunks to
// generate, chunks to mesh, and the state changes of the unloads. One thread per delta; the three output lists are compacted
// with a block scan over the order, so each stays nearest first.
__global__ void __launch_bounds__(1024) delta_lists_kernel(SchedParams p, uint32_t limit) {
    __shared__ uint32_t warp_tot[3][32];
    __shared__ uint32_t carry[3];
    const uint32_t n = min(limit, p.totals[0] + p.totals[1]);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3) carry[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        uint32_t c = 0, code = 0;
        if (i < n) {
            c = p.order[i];
            code = p.code[c];
        }
        const bool un = (code & 4u) != 0;
        const uint32_t f[3] = {(!un && (code & 1u)) ? 1u : 0u, (!un && (code & 2u)) ? 1u : 0u, (un && (code & 2u)) ? 1u : 0u};
        uint32_t off[3];
        for (int k = 0; k < 3; k++) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, f[k]);
            off[k] = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) warp_tot[k][warp] = __popc(bal);
        }
        __syncthreads();
        for (int k = 0; k < 3; k++) {
            uint32_t b = carry[k];
            for (uint32_t w = 0; w < warp; w++) b += warp_tot[k][w];
            off[k] += b;
        }
        if (i < n) {
            const int sx = (int)(c % (uint32_t)p.SX), sz = (int)((c / (uint32_t)p.SX) % (uint32_t)p.SZ), sy = (int)(c / (uint32_t)(p.SX * p.SZ));
            const uint32_t slot = (uint32_t)((sx - 1) + (p.SX - 2) * ((sz - 1) + (p.SZ - 2) * (sy - 1))); // meaningful for interior chunks
            if (f[0]) {
                p.gen_list[off[0]] = c;
                p.blk_loaded[c] = 1;
            }
            if (f[1]) {
                p.mesh_list[off[1]] = c;
                p.mesh_span[off[1]] = slot;
                p.mesh_loaded[slot] = 1;
            }
            if (un) { // swap_data(None) + status Unloaded (worldmgr.rs:478-492)
                if (code & 1u) p.blk_loaded[c] = 0;
                if (code & 2u) {
                    p.unmesh_span[off[2]] = slot;
                    p.mesh_loaded[slot] = 0;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 0; k < 3; k++) {
                uint32_t t = 0;
                for (int w = 0; w < 32; w++) t += warp_tot[k][w];
                carry[k] += t;
            }
        __syncthreads();
    }
    if (threadIdx.x < 3) p.list_counts[threadIdx.x] = carry[threadIdx.x];
    if (threadIdx.x == 3) p.list_counts[3] = n;
}

// unloaded meshes: the span is dropped, its arena reservation becomes garbage
__global__ void unmesh_kernel(const uint32_t *slots, const uint32_t *count, bxw_mesh_span *spans, uint2 *span_cap, MeshCounters *counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const uint32_t s = slots[i];
    const uint2 oc = span_cap[s];
    if (oc.x) atomicAdd(&counters->v_garbage, (unsigned long long)oc.x);
    if (oc.y) atomicAdd(&counters->i_garbage, (unsigned long long)oc.y);
    span_cap[s] = make_uint2(0u, 0u);
    bxw_mesh_span sp;
    sp.v_off = sp.i_off = 0;
    sp.v_count = sp.i_count = 0;
    spans[s] = sp;
}

} // namespace

void launch_delta_build(const SchedParams &p, cudaStream_t stream) {
    cudaMemsetAsync(p.hist, 0, 2ull * p.n_bins * sizeof(uint32_t), stream);
    const unsigned nb = (p.n_storage + 255u) / 256u;
    delta_classify_kernel<<<nb, 256, 0, stream>>>(p);
    delta_scan_kernel<<<1, 1024, 0, stream>>>(p.hist, p.n_bins, p.totals);
    delta_scatter_kernel<<<nb, 256, 0, stream>>>(p);
}

void launch_delta_lists(const SchedParams &p, uint32_t limit, cudaStream_t stream) { delta_lists_kernel<<<1, 1024, 0, stream>>>(p, limit); }

void launch_unmesh(const uint32_t *slots, const uint32_t *count, uint32_t upper, bxw_mesh_span *spans, uint2 *span_cap, MeshCounters *counters,
                   cudaStream_t stream) {
    if (upper) unmesh_kernel<<<(upper + 255u) / 256u, 256, 0, stream>>>(slots, count, spans, span_cap, counters);
}

} // namespace bxw
