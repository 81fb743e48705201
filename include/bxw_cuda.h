/*
 * bxw_cuda.h — C ABI of libbxw_cuda.so: BallX World's chunk pipeline (generate -> GPU-resident store -> mesh)
 * as hand-written sm_100a kernels.
 *
 * The reference (eigenraven/ballxworld, Rust) has no FFI of its own; the seam this ABI replaces is three Rust
 * items (file:line relative to the reference tree):
 *   1. bxw_world::stdgen::StdGenerator::{new, generate_chunk}          bxw_world/src/stdgen.rs:297-374
 *   2. client::render::voxmesh::{mesh_from_chunk, is_chunk_trivial}     src/client/render/voxmesh.rs:16-190
 *   3. the ChunkDataHandler plugin trait both sit behind               bxw_world/src/worldmgr.rs:49-77
 *      (WorldBlocks: bxw_world/src/generation.rs:11-203; MeshDataHandler: src/client/render/voxrender.rs:192-448;
 *      edits: World::apply_voxel_changes bxw_world/src/worldmgr.rs:312-357)
 * INTEGRATION.md shows the Rust-side bindings (extern "C" block + shims) a maintainer would add.
 *
 * Conventions: every function returns 0 (BXW_OK) or a negative BXW_E* code and never unwinds; out-pointers are
 * caller-owned; one context per CUDA device; calls on one context must be serialised by the caller (the Rust
 * side wraps the context in a Mutex). There is NO CPU fallback: without a usable CUDA device bxw_create fails.
 */
#ifndef BXW_CUDA_H
#define BXW_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BXW_CHUNK_DIM 32
#define BXW_CHUNK_VOXELS 32768

enum {
    BXW_OK = 0,
    BXW_E_INVALID = -1,    /* bad argument / call order */
    BXW_E_CUDA = -2,       /* CUDA runtime error; see bxw_last_error */
    BXW_E_NOMEM = -3,      /* host or device allocation failed */
    BXW_E_UNKNOWN_ID = -4, /* a voxel's block id has no registry definition (reference panics: voxregistry.rs:141-143) */
    BXW_E_NOSPACE = -5,    /* caller buffer / mesh arena too small */
    BXW_E_BAD_RLE = -6,    /* RLE stream does not decode to 32768 voxels (lib.rs:296-302) */
    BXW_E_NO_REGISTRY = -7 /* registry / generator ids not set */
};

typedef struct bxw_ctx bxw_ctx;

/* VoxelDefinition fields the path reads (bxw_world/src/lib.rs:629-639). mesh_kind: 0 VoxelMesh::None,
 * 1 VoxelMesh::CubeAndSlopes (lib.rs:651-655). tex[] is TextureMapping by signed axis index (lib.rs:624-626). */
typedef struct {
    uint8_t present;
    uint8_t mesh_kind;
    float debug_color[3];
    uint32_t tex[6];
} bxw_block_def;

/* vox::VoxelVertex, #[repr(C)], 40 bytes (src/client/render/voxrender.rs:39-49) */
typedef struct {
    uint32_t position; /* W2 X10 Y10 Z10 */
    uint32_t color;    /* R8 G8 B8 A8 */
    float texcoord[3];
    int32_t index;
    float barycentric_color_offset[4];
} bxw_vertex;

/* VoxelChange as consumed by World::apply_voxel_changes (bxw_world/src/worldmgr.rs:312-357,
 * bxw_world/src/generation.rs:39-64): compare-and-swap of one voxel at a global block position. */
typedef struct {
    int32_t x, y, z;
    uint32_t from, to;
} bxw_voxel_change;

/* Where one chunk's ChunkBuffers{vertices, indices} (voxrender.rs:34-37) live inside the mesh arena.
 * Indices are chunk-relative exactly as in the reference (voxmesh.rs:123,175). */
typedef struct {
    uint64_t v_off, i_off; /* in vertices / indices from the start of the arena */
    uint32_t v_count, i_count;
} bxw_mesh_span;

/* ---- context ---- */
int bxw_create(int device, bxw_ctx **out);
void bxw_destroy(bxw_ctx *ctx);
const char *bxw_last_error(const bxw_ctx *ctx);
/* Launch all work on an externally owned cudaStream_t (e.g. torch's current stream); NULL = the context's own. */
int bxw_set_stream(bxw_ctx *ctx, void *cuda_stream);
/* Tuning switches. BXW_OPT_CHUNK_SUMMARIES (default 1): whole-region mesh passes consult the per-chunk uniformity
 * summaries and skip chunks proven empty; 0 = read and mesh every chunk (same output; used to measure the streaming rate).
 * BXW_OPT_MESH_FROM_HEIGHT_MAP (default 0): bxw_region_generate_and_mesh derives the mesher's class table from the column
 * parameters instead of reading the freshly written voxels (same output).
 * BXW_OPT_ELIDE_UNIFORM (default 1): chunks the generator proves uniform (all void above the terrain, all stone below it)
 * are not written voxel by voxel: their entry of the region's page table points at a shared single-datum page (the device
 * analogue of the reference's 3-word VChunk, bxw_world/src/lib.rs:263-294); the first writer (edit, upload, import, halo
 * import that differs) gives the chunk its own storage back. 0 = every chunk materialised. Same results either way.
 * BXW_OPT_CUBES_ONLY_MESHER (default 1): regions holding generator output (plus edits) are meshed by the cubes-and-void
 * instantiation of the mesher (no class table, twice the occupancy); chunks with other shapes fall back to the general
 * kernel. 0 = general kernel only. Same results either way.
 * BXW_OPT_PIPELINED_MESHER (default 0): the mesher's two kernels run side by side on two streams, emit_kernel consuming
 * each chunk's face records as soon as mesh_kernel has published them. 1 = when the previous pass emitted at least 400
 * sides per listed chunk, 2 = always, 0 = one after the other. Same results; on B200 the two kernels slow each other down
 * by more than the overlap gains (profiles/r02_summary.md), so it is off unless asked for. */
enum { BXW_OPT_CHUNK_SUMMARIES = 0, BXW_OPT_MESH_FROM_HEIGHT_MAP = 1, BXW_OPT_ELIDE_UNIFORM = 2, BXW_OPT_CUBES_ONLY_MESHER = 3,
       BXW_OPT_PIPELINED_MESHER = 4 };
int bxw_set_option(bxw_ctx *ctx, int option, int value);
int bxw_synchronize(bxw_ctx *ctx);
/* Number of this library's kernels launched since creation (bench.py's gpu_launches evidence). */
uint64_t bxw_kernel_launches(const bxw_ctx *ctx);

/* ---- registry (replaces &VoxelRegistry arguments; voxregistry.rs:94-143, blocks/mod.rs:6-65) ---- */
int bxw_set_registry(bxw_ctx *ctx, const bxw_block_def *defs, uint32_t n); /* ids 0..n-1 */
/* The five names StdGenerator::generate_chunk resolves (stdgen.rs:307-326), as ids. */
int bxw_set_gen_ids(bxw_ctx *ctx, uint16_t void_id, uint16_t grass, uint16_t dirt, uint16_t stone,
                    uint16_t snow_grass);

/* ---- per-chunk drop-ins (host pointers) ---- */
/* StdGenerator::new(seed).generate_chunk(chunk{position=(cx,cy,cz)}, registry)  (stdgen.rs:297-374).
 * Overwrites all 32768 voxels of blocks_yzx (index x + 32 z + 1024 y, lib.rs:104-108). */
int bxw_gen_chunk(bxw_ctx *ctx, uint64_t seed, int32_t cx, int32_t cy, int32_t cz, uint32_t *blocks_yzx);
/* mesh_from_chunk(registry, chunks[27]) (voxmesh.rs:45-190) with the 27 neighbours given dense, in
 * iter_neighbors(cpos, true) order: slot = rx + 3 rz + 9 ry (worldmgr.rs:748-759, voxmesh.rs:72).
 * counts->v_count / i_count receive the buffer sizes; fetch with bxw_mesh_fetch. */
int bxw_mesh_chunk27(bxw_ctx *ctx, const uint32_t *const dense[27], bxw_mesh_span *counts);
/* Same, neighbours given as VChunk RLE streams (lib.rs:237-345); decoded on the device. */
int bxw_mesh_chunk27_rle(bxw_ctx *ctx, const uint32_t *const rle[27], const size_t rle_len[27],
                         bxw_mesh_span *counts);
int bxw_mesh_fetch(bxw_ctx *ctx, bxw_vertex *vertices, uint32_t *indices);
/* is_chunk_trivial(chunk, registry) (voxmesh.rs:16-25); host-side, no device work. */
int bxw_is_chunk_trivial(bxw_ctx *ctx, const uint32_t *rle, size_t rle_len, int *out_trivial);

/* ---- VChunk RLE codec on the device (bxw_world/src/lib.rs:263-345), host buffers ---- */
/* compress_rle of n_chunks dense chunks; out_words must hold n_chunks*32768*... see bxw_rle_bound(). offsets
 * gets n_chunks+1 word offsets. */
size_t bxw_rle_bound(size_t n_chunks);
int bxw_rle_compress(bxw_ctx *ctx, const uint32_t *dense, size_t n_chunks, uint32_t *out_words,
                     uint64_t *out_offsets);
int bxw_rle_decompress(bxw_ctx *ctx, const uint32_t *words, const uint64_t *offsets, size_t n_chunks,
                       uint32_t *out_dense);

/* ---- region form: GPU-resident block storage + batched kernels (the form that reaches the roofline) ---- */
/* A region is the box of nx*ny*nz chunks whose minimum chunk is (cx0,cy0,cz0), stored together with a 1-chunk
 * shell so every interior chunk has its 26 neighbours (the reference's precondition, worldmgr.rs:761-790).
 * Local chunk coordinates l* run over -1..n* (shell included). One region per context. */
int bxw_region_create(bxw_ctx *ctx, int32_t cx0, int32_t cy0, int32_t cz0, uint32_t nx, uint32_t ny, uint32_t nz);
int bxw_region_destroy(bxw_ctx *ctx);
/* Fill every chunk (shell included) with StdGenerator(seed) terrain. precision: 64 = f64 noise exactly as the
 * reference computes it; 32 = FP32 noise evaluation (north_star-literal variant; may flip voxels whose noise value
 * sits within rounding of a selection threshold). */
int bxw_region_generate(bxw_ctx *ctx, uint64_t seed, int precision);
/* The same in two calls, for multi-GPU steps: part 1 fills only the two chunk columns next to the x- and x+ faces of the
 * interior (everything bxw_region_halo_export reads), part 2 all the others. Between the two the caller exports the boundary
 * planes and starts their transfer, which then overlaps part 2; the received planes are imported after part 2. Part 2 is
 * queued on an internal side stream that is ordered after everything preceding part 1 but not after part 1 and what the caller
 * queued behind it; the context's stream waits for it before the call returns control of the stream to the caller (later work
 * on the stream sees all of the generation). */
int bxw_region_generate_part(bxw_ctx *ctx, uint64_t seed, int precision, int part);
/* Column parameters of the last generate: height (round(h) as i32) and elevation class (0 LowLand,1 Hill,
 * 2 Mountain) for the (nx+2)*32 x (nz+2)*32 columns, row-major [z][x]; raw f64 height optional (may be NULL). */
int bxw_region_heightmap(bxw_ctx *ctx, int32_t *height, uint8_t *elevation, double *raw_height);
int bxw_region_upload_chunk(bxw_ctx *ctx, int32_t lx, int32_t ly, int32_t lz, const uint32_t *blocks_yzx);
int bxw_region_download_chunk(bxw_ctx *ctx, int32_t lx, int32_t ly, int32_t lz, uint32_t *blocks_yzx);
/* WorldBlocks::get_data / serialize_data (bxw_world/src/generation.rs:80-84, 158-166) served from device storage:
 * compress_rle (lib.rs:263-302) of the n_chunks region chunks lpos[3i..3i+2] = (lx,ly,lz) without moving their voxels to
 * the host. out_offsets receives n_chunks+1 word offsets (the save blob of chunk i is words[offsets[i]..offsets[i+1]) as
 * little-endian bytes). out_words == NULL is a size query; BXW_E_NOSPACE if cap_words < out_offsets[n_chunks]. */
int bxw_region_export_rle(bxw_ctx *ctx, const int32_t *lpos, size_t n_chunks, uint32_t *out_words, size_t cap_words,
                          uint64_t *out_offsets);
/* WorldBlocks::deserialize_data (generation.rs:168-194): decode n_chunks RLE streams on the device into the region
 * chunks lpos[], refresh derived planes and mark the chunks and their 26 neighbours dirty for bxw_region_remesh_dirty.
 * A stream that does not decode returns BXW_E_BAD_RLE and leaves the region unchanged. */
int bxw_region_import_rle(bxw_ctx *ctx, const int32_t *lpos, size_t n_chunks, const uint32_t *words, const uint64_t *offsets);
/* Synthetic benchmark input (BASELINE.json configs[1], SURVEY.md section 8d cfg 2): every voxel (shell included) is a
 * pure hash of its global position: splitmix64(seed ^ key) -> void if low byte < void_threshold, else a random id in
 * 1..7 with a random shape (cube/slope/corner/inner corner) and one of the 24 orientations. */
int bxw_region_fill_synthetic(bxw_ctx *ctx, uint64_t seed, uint32_t void_threshold);
/* Mesh arena capacity (vertices / indices). Without a reservation bxw_region_mesh sizes it with a count pass. */
int bxw_region_mesh_reserve(bxw_ctx *ctx, uint64_t max_vertices, uint64_t max_indices);
/* Mesh every interior chunk (mesh_from_chunk semantics per chunk). Totals include alignment padding between
 * chunks, i.e. they are the arena extents to fetch. */
int bxw_region_mesh(bxw_ctx *ctx, uint64_t *arena_vertices, uint64_t *arena_indices);
/* Fused StdGenerator + mesher over the pristine region: blocks are written once and meshed from on-chip data. */
int bxw_region_generate_and_mesh(bxw_ctx *ctx, uint64_t seed, int precision, uint64_t *arena_vertices,
                                 uint64_t *arena_indices);
/* View-distance scheduling (World::recalculate_load_deltas, bxw_world/src/worldmgr.rs:633-746): mesh exactly the
 * interior chunks inside the load-anchor sphere |c - anchor|^2 <= radius^2 (iter_coords_to_load, worldmgr.rs:734-742),
 * nearest first (the reference sorts its deltas by distance); every other chunk's span is cleared.
 * anchor is a world chunk position. n_meshed = chunks in the sphere that belong to this region. */
int bxw_region_mesh_sphere(bxw_ctx *ctx, int32_t ax, int32_t ay, int32_t az, uint32_t radius, uint32_t *n_meshed,
                           uint64_t *arena_vertices, uint64_t *arena_indices);
/* ---- load anchors on the device (World::recalculate_load_deltas / main_loop_tick, bxw_world/src/worldmgr.rs:407-567,633-790) ----
 * The region box plays the part of the world: every chunk of the box (shell included) can hold block data, every interior
 * chunk a mesh; coordinates outside the box are ignored. Each chunk carries the two load states of the reference's handlers
 * (CHUNK_BLOCK_DATA, CHUNK_MESH_DATA): bxw_region_generate / _mesh set them for the whole box, bxw_region_unload_all clears
 * them. */
typedef struct {
    int32_t cx, cy, cz;  /* anchor chunk position (ChunkPosition::from(CLocation.position)) */
    uint32_t radius;     /* CLoadAnchor.radius, chunks */
    uint32_t load_mesh;  /* CLoadAnchor.load_mesh: 0 = block data only (a server-side anchor) */
} bxw_load_anchor;
enum { BXW_KIND_BLOCKS = 1, BXW_KIND_MESH = 2 };
typedef struct {        /* ChunkDelta (worldmgr.rs:96-102) */
    int32_t cx, cy, cz;
    uint32_t kinds;      /* BXW_KIND_* bits */
    uint32_t unload;
    int32_t min_anchor_distance; /* squared, chunks */
} bxw_chunk_delta;
int bxw_region_unload_all(bxw_ctx *ctx);
/* recalculate_load_deltas on the device: chunks whose loaded kinds no anchor wants any more (unload deltas, keyed by the
 * distance to the nearest anchor) and chunks inside an anchor sphere that miss a kind they may request now (load deltas: block
 * data at once, a mesh when the block data of the chunk and its 26 neighbours is loaded), each list nearest first, interleaved
 * unload / load like the reference's. The list stays on the device; n_unload / n_load receive the counts. */
int bxw_region_update_anchors(bxw_ctx *ctx, const bxw_load_anchor *anchors, uint32_t n_anchors, uint32_t *n_unload, uint32_t *n_load);
/* Copy of the list of the last bxw_region_update_anchors (at most cap entries), in processing order. */
int bxw_region_load_deltas(bxw_ctx *ctx, bxw_chunk_delta *out, uint32_t cap, uint32_t *n);
/* main_loop_tick over the first max_deltas entries of the list (0 = all), batched: unloads drop their data (a mesh's arena
 * space becomes garbage), block loads run the generator for just those chunks, mesh loads run the mesher over just those
 * chunks, nearest first. stats = {deltas consumed, chunks generated, chunks meshed, meshes unloaded}. Meshes need a second
 * update + apply round after their block data has arrived, exactly as in the reference. */
int bxw_region_apply_deltas(bxw_ctx *ctx, uint64_t seed, int precision, uint32_t max_deltas, uint32_t stats[4]);
/* Load state of every chunk: blocks_loaded[(nx+2)(ny+2)(nz+2)] (storage order), mesh_loaded[nx*ny*nz] (span order); either
 * may be NULL. */
int bxw_region_load_state(bxw_ctx *ctx, uint8_t *blocks_loaded, uint8_t *mesh_loaded);
/* spans: nx*ny*nz entries, chunk order x + nx*(z + nz*y). */
int bxw_region_mesh_spans(bxw_ctx *ctx, bxw_mesh_span *spans);
/* Copy arena extents [0,arena_vertices) / [0,arena_indices) to host (pinned memory recommended). */
int bxw_region_mesh_fetch(bxw_ctx *ctx, bxw_vertex *vertices, uint64_t n_vertices, uint32_t *indices,
                          uint64_t n_indices);
/* One piece of the arena, e.g. a single chunk's ChunkBuffers (its span's offsets and counts). */
int bxw_region_mesh_fetch_range(bxw_ctx *ctx, uint64_t v_off, uint64_t n_vertices, bxw_vertex *vertices, uint64_t i_off,
                                uint64_t n_indices, uint32_t *indices);
/* The same copy on the context's copy stream, returning at once: the next generate / mesh pass overlaps the transfer (only
 * the kernel that rewrites the arena waits for it). The host buffers (pinned) are valid after bxw_region_fetch_wait. */
int bxw_region_mesh_fetch_async(bxw_ctx *ctx, bxw_vertex *vertices, uint64_t n_vertices, uint32_t *indices,
                                uint64_t n_indices);
int bxw_region_fetch_wait(bxw_ctx *ctx);
/* World::apply_voxel_changes (worldmgr.rs:312-357): CAS each change in order, mark touching_chunks
 * (lib.rs:110-135) dirty. n_dirty = interior chunks now dirty. */
int bxw_region_apply_changes(bxw_ctx *ctx, const bxw_voxel_change *changes, uint32_t n, uint32_t *n_dirty);
/* Re-mesh the dirty interior chunks; their spans are updated. A chunk whose new mesh fits the arena space reserved behind
 * its old span is rewritten in place, otherwise it moves to the end of the arena and the old space is garbage (the reference
 * drops the old ChunkBuffers when the new one lands, src/client/render/voxrender.rs:398-413); when a quarter of the arena is
 * garbage the live spans are compacted (offsets change: re-read the spans). The arena extent to fetch afterwards is
 * bxw_region_arena_extent. */
int bxw_region_remesh_dirty(bxw_ctx *ctx, uint32_t *n_remeshed);
/* Current arena extents (what bxw_region_mesh returned, plus whatever re-mesh passes appended since). */
int bxw_region_arena_extent(bxw_ctx *ctx, uint64_t *arena_vertices, uint64_t *arena_indices);
typedef struct {
    uint64_t vertices_used, indices_used;         /* arena extents */
    uint64_t vertices_garbage, indices_garbage;   /* inside the extents, referenced by no span */
    uint64_t vertices_capacity, indices_capacity; /* allocated */
} bxw_arena_stats;
int bxw_region_arena_stats(bxw_ctx *ctx, bxw_arena_stats *out);
/* Move every live span to the front of a fresh arena (slot order); spans' offsets change, garbage becomes 0. */
int bxw_region_compact_arena(bxw_ctx *ctx, uint64_t *arena_vertices, uint64_t *arena_indices);
/* Re-position the region box without reallocating it (slab streaming: generate -> mesh -> retire, box by box; a sliding
 * view window). The stored voxels are stale until the next generate / upload. */
int bxw_region_set_origin(bxw_ctx *ctx, int32_t cx0, int32_t cy0, int32_t cz0);
/* Multi-GPU slab boundary: the 1-voxel plane adjacent to face 0 (x-) or 1 (x+) of the interior, packed for the
 * neighbouring rank ((nz+2)*32 x (ny+2)*32 u32, [y][z]). export packs interior boundary voxels into dev_buf;
 * import writes a received plane into this region's shell. dev_buf is device memory (e.g. a torch tensor). */
size_t bxw_region_halo_bytes(const bxw_ctx *ctx);
int bxw_region_halo_export(bxw_ctx *ctx, int face, void *dev_buf);
int bxw_region_halo_import(bxw_ctx *ctx, int face, const void *dev_buf);
/* Device pointers for zero-copy consumers (e.g. graphics interop); valid until the next mesh/destroy call. */
int bxw_region_device_ptrs(bxw_ctx *ctx, void **blocks, void **vertices, void **indices);
/* The region's page table (device, u32 per storage chunk, index sx + SX*(sz + SZ*sy) over the shell-padded box): chunk c's
 * voxels are at blocks + pages[c] * 32768; pages[c] == c unless the chunk is uniform and not materialised, in which case it
 * is one of the n_shared single-datum pages stored behind the box. */
int bxw_region_page_table(bxw_ctx *ctx, void **pages, uint32_t *n_shared);
/* The same table copied to the host: (nx+2)*(ny+2)*(nz+2) entries. */
int bxw_region_download_pages(bxw_ctx *ctx, uint32_t *pages);

/* ---- bxw_terragen in-tree SuperSimplex (bxw_terragen/src/supersimplex.rs:20-107), batched ---- */
int bxw_terragen_noise2(bxw_ctx *ctx, uint64_t seed, const double *xs, const double *ys, size_t n, double *out);

/* ---- the generator's noise functions, raw (noise::SuperSimplex::get behind bxw_world/src/stdgen.rs:171-213) ----
 * which: 0..4 = height_map_gen[i], 5 = elevation_map_gen, 6 = moisture_map_gen of StdGenerator(seed) (stdgen.rs:103-114).
 * precision 64 = the reference's f64 arithmetic; 32 = the FP32 variant (lattice cell chosen in f64, attenuation / gradient
 * arithmetic in f32): values agree with f64 within 1e-5 of the noise range. */
int bxw_stdgen_noise(bxw_ctx *ctx, uint64_t seed, int which, int precision, const double *xs, const double *ys, size_t n, double *out);

/* ---- device timing helpers (CUDA events on the context's stream) ---- */
int bxw_timer_start(bxw_ctx *ctx);
int bxw_timer_stop_ms(bxw_ctx *ctx, float *ms); /* records, synchronises, returns elapsed */
/* Accumulated device time of the mesher kernel since the last reset (events around each launch). */
int bxw_mesh_kernel_time_ms(bxw_ctx *ctx, float *ms, uint64_t *launches, int reset);
/* The same for the generator's kernels: which = BXW_KERNEL_MESH / _FILL (block selection, stdgen.rs:349-373) /
 * _HEIGHTMAP (calc_voxel_params, stdgen.rs:215-276). */
enum { BXW_KERNEL_MESH = 0, BXW_KERNEL_FILL = 1, BXW_KERNEL_HEIGHTMAP = 2, BXW_KERNEL_EMIT = 3 /* emit_kernel's share of _MESH */ };
int bxw_kernel_time_ms(bxw_ctx *ctx, int which, float *ms, uint64_t *launches, int reset);
/* Statistics of the last whole-region mesh pass: chunks that were actually meshed (the rest were proven empty from
 * their summaries, the device analogue of is_chunk_trivial on a 3-word RLE chunk, voxmesh.rs:16-25) / interior chunks. */
int bxw_region_mesh_candidates(bxw_ctx *ctx, uint32_t *listed, uint32_t *interior);
/* Profiling builds (-DBXW_PHASE_TIMING) only, zeros otherwise: SM cycles per mesher phase of the last region mesh launch,
 * summed over chunks: fetch, phase A, phases B+scan, phase C, chunks that emitted, chunks, then phase C split into tile
 * setup, face ordering, record hand-off, tile-end barrier wait (2 spare). */
int bxw_debug_phase_cycles(bxw_ctx *ctx, uint64_t out[12]);

#ifdef __cplusplus
}
#endif
#endif
