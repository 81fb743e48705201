// device_common.cuh — shared device-side declarations of libbxw_cuda (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/bxw_cuda.h"
#include "tables.hpp"

namespace bxw {

constexpr int kChunkDim = 32;
constexpr int kChunkVoxels = 32768;

// Registry as the kernels see it (voxregistry.rs:141-143 lookups become array reads).
struct DeviceRegistry {
    const uint8_t *kind;   // [n] 0 = absent, 1 = VoxelMesh::None, 2 = VoxelMesh::CubeAndSlopes
    const float *texf;     // [n][6] texture id as f32 (voxmesh.rs:167 `texid as f32`)
    const float *color;    // [n][3] debug_color
    uint32_t n;
};

// error / status flags accumulated by kernels (atomicOr into MeshCounters::flags)
constexpr unsigned long long kFlagUnknownId = 1ull;
constexpr unsigned long long kFlagArenaOverflow = 2ull;
constexpr unsigned long long kFlagPipelineTimeout = 4ull; // pipelined emit_kernel gave up waiting for mesh_kernel (a bug, not a state)

struct MeshCounters {
    unsigned long long v_alloc; // next free vertex slot in the arena
    unsigned long long i_alloc; // next free index slot
    unsigned long long flags;
    unsigned long long work_next; // dynamic work-list cursor
    unsigned long long work_next2; // the same for the general kernel's pass over the cubes-only kernel's fallback list
    unsigned long long f_alloc;   // next free face-record slot (a chunk takes a multiple of 32)
    unsigned long long mesh_done; // pipelined passes: mesh_kernel CTAs that have retired
    unsigned long long v_garbage; // arena space that no span refers to any more (re-meshes that moved, unloaded meshes)
    unsigned long long i_garbage;
    // -DBXW_PHASE_TIMING builds only: SM cycles of CTA thread 0 summed over chunks: [0] work fetch + neighbour table,
    // [1] phase A, [2] phases B + scan + allocation, [3] phase C, [4] chunks that reached phase C, [5] chunks
    unsigned long long phase[12]; // [6] tile setup, [7] C1, [8] warp 0's batches, [9] tile-end barrier wait
};

struct MeshParams {
    const uint32_t *blocks; // region storage: chunk page g at blocks + g*32768, voxel x + 32 z + 1024 y
    const uint32_t *xface;  // per chunk page the x=0 and x=31 voxel planes, [page][side][y][z] (2 x 4 KB): the mesher's x-halo source
    // page[c] = where chunk c's voxels live: c itself, or one of the shared pages behind the storage that hold a single
    // datum (chunks the generator proved uniform are not materialised; device analogue of a 3-word VChunk, lib.rs:263-294)
    const uint32_t *page;
    int SX, SY, SZ;         // storage dims in chunks (shell included); chunk index sx + SX*(sz + SZ*sy)
    const uint32_t *work;   // [n_work] storage chunk index of each chunk to mesh
    const uint32_t *work_span; // [n_work] span slot of each work item
    uint32_t n_work;           // upper bound of the list length (grid sizing)
    const uint32_t *n_work_dev; // optional: the actual list length lives on the device (candidate list built by a kernel)
    bxw_mesh_span *spans;
    // span_cap[slot] = arena space (vertices, indices) reserved behind the slot's span. reuse = 1 (re-mesh passes): a chunk
    // whose new mesh fits its old reservation is rewritten in place, otherwise it moves to the arena's end (with slack) and the
    // old reservation is counted as garbage (the reference drops the old ChunkBuffers when the new one lands, voxrender.rs:398-413)
    uint2 *span_cap;
    int reuse;
    bxw_vertex *varena;
    uint32_t *iarena;
    unsigned long long v_cap, i_cap;
    // face records between mesh_kernel (stage, cull, count, order) and emit_kernel (vertex / index construction): f_cap
    // records + f_cap / 32 span slots, see kFacePad below
    uint4 *frec;
    uint4 *fslot;
    unsigned long long f_cap;
    MeshCounters *counters;
    unsigned long long *cursor; // the launch's work-list cursor (inside *counters)
    const MeshTables *tables;
    DeviceRegistry reg;
    uint32_t *fallback_work, *fallback_span, *fallback_count; // lean kernel only: chunks left to the general kernel
    // pipelined pass (mesh_kernel and emit_kernel side by side): tag of this pass in the fslot entries, number of mesh CTAs of
    // all mesh launches of the pass
    int pipelined;
    uint32_t epoch;
    unsigned long long mesh_ctas;
    int count_only; // 1: fill spans' counts only (offsets = running totals), no vertex/index writes
    // fused generate+mesh (pristine StdGenerator terrain): derive the 34^3 table from the column parameters
    const int32_t *hmap;   // nullptr = read the block storage
    int32_t hmap_W, cy_min;
    uint32_t gen_datum[5]; // void, grass, dirt, stone, snow_grass datums
    int gen_all_cubes;     // grass, dirt, stone and snow grass all classify as cubes (1..24)
};

// lean = the cubes-and-void instantiation (no class table, 4-5 CTAs per SM); chunks it cannot handle are appended to the
// fallback list for a second launch of the general kernel
// mostly_uniform: the work list is a whole region with the uniform chunks still in it (picks the lean variant tuned for them)
void launch_mesh_kernel(const MeshParams &p, int sm_count, cudaStream_t stream, bool lean, bool mostly_uniform, int max_per_sm = 0);

// Face record: one uint4 per visible side, frec[k], in the reference's emission order; every chunk's records start on a
// multiple of 32 and are padded to one, and fslot[k / 32] holds the chunk's arena offsets (v_off, i_off as two u64), so that
// emit_kernel's only global loads are the two it prefetches one batch ahead.
//   x: row (y*32+z, 10 bits) | x << 10 | world dir << 15 | class << 18; 0xFFFFFFFF = padding (no face)
//   y: first vertex of the side, chunk-relative (20 bits) | (block id & 0xFFF) << 20
//   z: first index of the side, chunk-relative (21 bits) | (block id >> 12) << 21
//   w: occupancy of the cell's 3x3x3 neighbourhood, bit (dy+1)*9 + (dz+1)*3 + (dx+1)
constexpr uint32_t kFacePad = 0xFFFFFFFFu;
// Second half of mesh_from_chunk (voxmesh.rs:122-177): builds the vertices and indices of every recorded side.
void launch_emit_kernel(const MeshParams &p, int sm_count, cudaStream_t stream);
void launch_emit_kernel_pipelined(const MeshParams &p, int sm_count, int ctas_per_sm, cudaStream_t stream);
// grid of the mesh_kernel launch launch_mesh_kernel would make (max_per_sm: cap on CTAs per SM, 0 = none)
unsigned mesh_kernel_grid(const MeshParams &p, int sm_count, bool lean, int max_per_sm);

// Fused generate+mesh: list the interior chunks of pristine terrain that can have a mesh (the surface passes through
// their 34^3 neighbourhood) and zero the spans of all others, from the height map alone.
struct CandidateParams {
    const int32_t *hmap;
    int32_t hmap_W, cy_min;
    int SX, SY, SZ;
    int void_is_empty, gen_all_cubes;
    uint32_t *work, *work_span, *count;
    bxw_mesh_span *spans;
    uint2 *span_cap;
};
void launch_hmap_candidates(const CandidateParams &p, cudaStream_t stream);

// Chunk summaries: the device-side analogue of a 3-word RLE chunk (lib.rs:263-302, voxmesh.rs:16-25). Every kernel that
// writes voxels either proves the chunk uniform (the generator's fill) or clears the entry.
constexpr unsigned long long kUniformBit = 1ull << 32;
// Bits 33..39 (generator only): every voxel of the chunk (bit 33) / of its boundary plane towards world dir d (bit 34 + d,
// d = -x +x -y +y -z +z) lies at or below the terrain surface, i.e. is grass, snow grass, dirt or stone. A chunk that is
// solid throughout and faces a solid plane on all six sides has no visible side when those four blocks are cubes.
constexpr int kSolidShift = 33;

// Whole-interior meshing from storage: keep the chunks that can have a mesh, in the mesher's tiled processing order, and
// zero the spans of the rest. A chunk is skipped when it is uniform and its block has no mesh (is_chunk_trivial), or when
// it and its 26 neighbours are uniform with one cube datum (every side clipped by an identical neighbour), or when it is
// solid generator terrain facing solid planes on all six sides (solids_are_cubes: the four terrain blocks mesh as cubes).
struct StorageCandParams {
    const unsigned long long *summary;
    int solids_are_cubes;
    int SX, SY, SZ;
    DeviceRegistry reg;
    uint8_t *keep;       // [n_work] scratch
    uint32_t *blk;       // [ceil(n_work / 256) + 1] scratch: per-block keep counts -> exclusive offsets
    uint32_t *work, *work_span, *count;
    bxw_mesh_span *spans;
    uint2 *span_cap;
};
void launch_storage_candidates(const StorageCandParams &p, cudaStream_t stream);
int mesh_kernel_max_ctas_per_sm();

// generation -------------------------------------------------------------------------------------------------
// A lattice point of super_simplex_2d's per-region table: offset in real space and integer lattice step
struct alignas(16) LatticeEntry {
    double ox, oy;
    int32_t px, py;
    int32_t pad[2];
};
struct GenTables {
    uint8_t perm[7][256]; // height_map_gen[0..5], elevation_map_gen, moisture_map_gen (stdgen.rs:82-84,103-114)
    // seed-independent tables of the noise function, next to the permutations so that a CTA stages all of it in shared memory
    // with one copy (indexed per lane: constant memory would serialise the lookups)
    alignas(16) double grad[8][2];  // gradient::grad2
    LatticeEntry lat[8][2];          // lattice points 2 and 3 of each of the 8 groups (points 0 and 1 are the same in every group)
};
void fill_noise_constants(GenTables &t); // gen_kernel.cu

struct CellPointDev {
    int32_t px, pz;
    double elevation_class;
};

struct GenParams {
    uint64_t seed;
    int32_t x0, z0;      // world voxel coordinate of column (0,0) of the height map
    int32_t W, D;        // columns along x, z
    int32_t cell_x0, cell_z0, cells_w, cells_d; // super-cell table origin / extent (stdgen.rs:15,149-161)
    const CellPointDev *cells; // [cells_d][cells_w][4]
    const GenTables *tables;
    int32_t *hmap;       // [D][W] height*4 + elevation
    double *raw;         // optional [D][W] raw f64 height
};

void launch_cellpoints_kernel(uint64_t seed, int32_t cell_x0, int32_t cell_z0, int32_t cells_w, int32_t cells_d,
                              const GenTables *tables, CellPointDev *cells, int precision, cudaStream_t stream);
void launch_heightmap_kernel(const GenParams &p, int precision, cudaStream_t stream);
void launch_noise_samples(const GenTables *tables, int which, const double *xs, const double *ys, size_t n, double *out, int precision,
                          cudaStream_t stream);

struct FillParams {
    const int32_t *hmap; // [D][W]
    int32_t W;
    uint32_t *blocks;
    uint32_t *xface;
    unsigned long long *summary; // per chunk: kUniformBit | datum when every voxel of the chunk holds that datum, else 0
    int SX, SY, SZ;
    int32_t cy_min;      // world chunk y of storage layer 0
    uint32_t id_void, id_grass, id_dirt, id_stone, id_snow; // already shifted datums (id << 16)
    // uniform-chunk elision: a chunk proven all void / all stone is not written; page[c] points at the shared page instead
    uint32_t *page;
    int elide;
    uint32_t page_void, page_stone;
    // list mode (load deltas): entry k = storage chunk index; the CTA handles that one chunk instead of a whole stack
    const uint32_t *list;
    uint32_t n_list;
    // stack mode, which chunk columns along x: 0 = all, 1 = the two interior boundary columns (sx = 1 and SX - 2; what a
    // multi-GPU halo export reads), 2 = all the others
    int col_mode;
};
void launch_fill_kernel(const FillParams &p, cudaStream_t stream);
void launch_fill_xface_kernel(const FillParams &p, cudaStream_t stream);
void launch_uniform_page(uint32_t *blocks, uint32_t *xface, uint32_t slot, uint32_t datum, cudaStream_t stream);
// height map + block selection in one kernel (g.hmap is written, g.raw is ignored)
void launch_gen_fill_kernel(const FillParams &p, const GenParams &g, int precision, cudaStream_t stream);

} // namespace bxw
