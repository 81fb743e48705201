// sched_kernels.cuh — launchers of the load-anchor scheduler kernels (sched_kernels.cu)
#pragma once
#include "device_common.cuh"

namespace bxw {

struct SchedParams {
    int SX, SY, SZ;                 // storage box (shell included)
    int32_t cx_min, cy_min, cz_min; // world chunk coordinate of storage chunk (0,0,0)
    uint32_t n_storage;
    const bxw_load_anchor *anchors; // device
    uint32_t n_anchors;
    uint8_t *blk_loaded;            // [n_storage] block data present (ChunkDataState::Loaded of CHUNK_BLOCK_DATA)
    uint8_t *mesh_loaded;           // [interior] mesh present (CHUNK_MESH_DATA)
    uint8_t *code;                  // [n_storage] scratch: kinds | unload << 2
    uint32_t *key;                  // [n_storage] scratch: squared distance
    uint32_t *hist;                 // [2 * n_bins] unload / load histograms -> offsets -> fill cursors
    uint32_t n_bins;
    uint32_t *totals;               // [2] unloads, loads
    bxw_chunk_delta *deltas;        // [n_storage] the ordered list
    uint32_t *order;                // [n_storage] storage chunk of each list entry
    uint32_t *gen_list;             // chunks to generate, nearest first
    uint32_t *mesh_list, *mesh_span; // chunks to mesh (+ span slots), nearest first
    uint32_t *unmesh_span;          // span slots of unloaded meshes
    uint32_t *list_counts;          // [4] generate, mesh, unmesh, deltas consumed
};

void launch_delta_build(const SchedParams &p, cudaStream_t stream);
void launch_delta_lists(const SchedParams &p, uint32_t limit, cudaStream_t stream);
void launch_unmesh(const uint32_t *slots, const uint32_t *count, uint32_t upper, bxw_mesh_span *spans, uint2 *span_cap, MeshCounters *counters,
                   cudaStream_t stream);

} // namespace bxw
