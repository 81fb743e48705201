// aux_kernels.cuh — launchers of the kernels in aux_kernels.cu
#pragma once
#include "device_common.cuh"

namespace bxw {

// import also clears the summaries of the chunks it writes; chunks of the shell column that are not materialised (page[c] != c)
// are materialised first when the plane brings a voxel that differs from their uniform datum (need: per-chunk scratch flags)
void launch_halo_plane(uint32_t *blocks, uint32_t *xface, unsigned long long *summary, uint32_t *page, uint8_t *need, int SX, int SY, int SZ,
                       int sx, int lx, uint32_t *plane, int import, cudaStream_t stream);
// give the listed chunks their own storage back (copy of their shared uniform page), before a writer touches single voxels
void launch_materialise(uint32_t *blocks, uint32_t *xface, uint32_t *page, const uint32_t *chunks, uint32_t n, cudaStream_t stream);
void launch_page_identity(uint32_t *page, uint32_t n, cudaStream_t stream);

struct EditParams {
    uint32_t *blocks;
    uint32_t *xface;
    unsigned long long *summary; // cleared for every chunk a change is applied to
    int SX, SY, SZ;
    int32_t x_min, y_min, z_min;     // world voxel coordinate of storage voxel (0,0,0)
    const bxw_voxel_change *changes; // grouped by voxel, original order within a group
    const uint32_t *run_start;       // [n_runs + 1]
    uint32_t n_runs;
    uint8_t *dirty;                  // interior chunk flags
    const uint8_t *blk_loaded;       // load-anchor mode: chunks without block data are skipped (nullptr = every chunk holds data)
};
void launch_apply_changes(const EditParams &p, cudaStream_t stream);
void launch_dirty_list(uint8_t *dirty, const uint8_t *mesh_loaded, uint32_t n, int nx, int nz, int SX, int SZ, uint32_t *work,
                       uint32_t *work_span, uint32_t *count, int clear, cudaStream_t stream);

// phase 0: per-chunk word counts -> offsets (n+1); phase 1: write words at offsets
// ids (optional): chunk k is read from dense + ids[k]*32768 instead of dense + k*32768
void launch_rle_compress(const uint32_t *dense, const uint32_t *ids, const uint32_t *page, size_t n_chunks, uint32_t *len, uint64_t *offsets,
                         uint32_t *out_words, int phase, cudaStream_t stream);
void launch_chunk_scatter(const uint32_t *src, const uint32_t *ids, size_t n, uint32_t *blocks, uint32_t *xface,
                          unsigned long long *summary, uint32_t *page, uint8_t *blk_loaded, int SX, int SY, int SZ, uint8_t *dirty,
                          cudaStream_t stream);
void launch_rle_decompress(const uint32_t *words, const uint64_t *offsets, size_t n_chunks, uint32_t *out_dense, uint32_t *status,
                           cudaStream_t stream);

// arena compaction (bxw_region_compact_arena): new offsets of every live span, then the copy into the second arena
void launch_compact_offsets(const bxw_mesh_span *spans, uint2 *span_cap, uint32_t n, unsigned long long *new_off, MeshCounters *counters,
                            cudaStream_t stream);
void launch_compact_copy(bxw_mesh_span *spans, const unsigned long long *new_off, uint32_t n, const bxw_vertex *v_old, const uint32_t *i_old,
                         bxw_vertex *v_new, uint32_t *i_new, cudaStream_t stream);

struct TerragenTables {
    uint16_t perm[2048];
    double gx[2048], gy[2048];
    int32_t lat_x[32], lat_y[32];
    double lat_dx[32], lat_dy[32];
};
void build_terragen_tables(uint64_t seed, TerragenTables &t);
void launch_terragen_noise2(const TerragenTables *T, const double *xs, const double *ys, size_t n, double *out, cudaStream_t stream);

// refresh the x-face planes of chunks[0..n) (or of chunks first..first+n when chunks == nullptr)
void launch_xface_extract(const uint32_t *blocks, uint32_t *xface, const uint32_t *chunks, uint32_t first, uint32_t n, cudaStream_t stream);

// synthetic benchmark input: hashed random blocks incl. shapes/orientations
void launch_fill_synthetic(uint32_t *blocks, int SX, int SY, int SZ, int32_t cx_min, int32_t cy_min, int32_t cz_min, uint64_t seed,
                           uint32_t void_threshold, cudaStream_t stream);

} // namespace bxw
