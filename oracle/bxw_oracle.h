/*
 * bxw_oracle.h — CPU restatement of BallX World's chunk pipeline (generate -> RLE store -> mesh).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library. The product path
 * (ballxworld_b200/, libbxw_cuda.so) never links, imports or calls anything under oracle/.
 *
 * Each function cites the reference file:line (relative to /root/reference) it follows.
 *
 * Parity status:
 *   - RLE codec: pinned by the reference's own KATs (bxw_world/src/lib.rs:486-546).
 *   - Orientation tables: pinned by the reference's property tests (bxw_util/src/direction.rs:354-448).
 *   - Mesher: every table is in-tree source; restated loop-for-loop; the reference has no mesher
 *     test, so it is pinned by hand-derived known answers only.
 *   - Generator: PARITY UNPINNED. The noise arithmetic lives in the third-party crate
 *     `noise 0.8.2` (+ `rand 0.7.3`, `rand_xorshift 0.2.0`, `rand_xoshiro 0.6.0`), none of which is
 *     vendored under /root/reference, and the reference has no generator test or golden vector.
 *     The published algorithms are restated here; the Rust reference cannot be built in this image.
 */
#ifndef BXW_ORACLE_H
#define BXW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BXO_CHUNK_DIM 32
#define BXO_CHUNK_DIM2 (32 * 32)
#define BXO_CHUNK_DIM3 (32 * 32 * 32)

/* ---- directions / orientations: bxw_util/src/direction.rs ---- */
/* Direction index: 0 X-, 1 X+, 2 Y-, 3 Y+, 4 Z-, 5 Z+ (direction.rs:9-22, 128-138). */
void bxo_dir_to_vec(int dir, int out[3]);                 /* direction.rs:89-99 */
int bxo_dir_from_vec(const int v[3]);                     /* direction.rs:76-87; -1 = None */
int bxo_dir_opposite(int dir);                            /* direction.rs:43-53 */
int bxo_lh_cross(int a, int b);                           /* direction.rs:157-161, 452-492; -1 = None */
int bxo_orient_from_index(int idx, int *right, int *up);  /* direction.rs:267-284; 0 ok, -1 None */
int bxo_orient_to_index(int right, int up);               /* direction.rs:249-264 */
int bxo_orient_from_dirs(int right, int up, int front);   /* direction.rs:216-222; 1 if Some */
void bxo_orient_matrix(int idx, int m[9]);                /* direction.rs:287-293; row-major m[r*3+c] */
int bxo_orient_apply(int idx, int dir);                   /* direction.rs:300-302 */
int bxo_orient_unapply(int idx, int dir);                 /* direction.rs:312-314 */

/* ---- block registry: bxw_world/src/voxregistry.rs, bxw_world/src/lib.rs:629-655 ---- */
typedef struct {
    uint8_t present;       /* definitions[id].is_some() */
    uint8_t mesh_kind;     /* 0 = VoxelMesh::None, 1 = VoxelMesh::CubeAndSlopes */
    float debug_color[3];
    uint32_t tex[6];       /* TextureMapping, indexed by signed axis index (lib.rs:624-626) */
} bxo_block_def;

typedef struct {
    uint32_t n;
    const bxo_block_def *defs;
} bxo_registry;

/* Standard registry of the client path (voxregistry.rs:94-111 + blocks/mod.rs:6-65), texture ids =
 * rank of the texture name among the sorted keys of res/manifest.toml (resources.rs:149-182,
 * toml 0.5 BTreeMap order). Fills 8 entries. */
void bxo_standard_registry(bxo_block_def out[8]);

/* ---- shapes: bxw_world/src/blocks/stdshapes.rs ---- */
typedef struct {
    float offset[3];
    float texcoord[2];
    float barycentric[3];
    int32_t barycentric_sign;
    int32_t n_ao;
    int32_t ao[5][3];
} bxo_vsvertex;

typedef struct {
    int32_t can_clip, can_be_clipped;
    int32_t n_verts;
    bxo_vsvertex verts[6];
    int32_t n_idx;
    uint32_t idx[6];
} bxo_vsside;

typedef struct {
    int32_t causes_ambient_occlusion;
    bxo_vsside sides[6];
} bxo_shape;

/* shape index: 0 none, 1 cube, 2 slope, 3 corner, 4 inner corner (stdshapes.rs:103-109) */
const bxo_shape *bxo_shape_table(int shape_index);
/* stdshapes.rs:51-72: shape index / orientation index of a datum given its definition's mesh kind */
int bxo_block_shape(uint32_t datum, int mesh_kind);
int bxo_block_orientation(uint32_t datum, int mesh_kind);

/* ---- RLE chunk storage: bxw_world/src/lib.rs:263-479 ---- */
/* compress_rle (lib.rs:263-294). out must hold up to 2*n words... returns word count. */
size_t bxo_compress_rle(const uint32_t *data, size_t n, uint32_t *out);
/* decompress_rle (lib.rs:304-345). 0 ok; 1 TooMuchUncompressedData; 2 TooMuchRleData;
 * 3 MissingRleRepeatN; 4 FinalPosMismatch. */
int bxo_decompress_rle(const uint32_t *data, size_t len, uint32_t out[BXO_CHUNK_DIM3]);

typedef struct {
    const uint32_t *rem;
    size_t rem_len;
    int state; /* 0 First, 1 Middle, 2 InRepetition, 3 Finished (lib.rs:348-353) */
    uint32_t prev, what, count;
    uint32_t pos;
} bxo_rle_iter;
void bxo_rle_iter_init(bxo_rle_iter *it, const uint32_t *data, size_t len); /* lib.rs:364-372 */
int bxo_rle_iter_next(bxo_rle_iter *it, uint32_t *datum, uint32_t *idx);    /* lib.rs:418-469; 1 = Some */
int bxo_rle_iter_skip_until(bxo_rle_iter *it, uint32_t target, uint32_t *datum, uint32_t *idx); /* lib.rs:374-403 */

/* ---- mesher: src/client/render/voxmesh.rs, src/client/render/voxrender.rs:39-49 ---- */
typedef struct {
    uint32_t position;  /* W2 X10 Y10 Z10 */
    uint32_t color;     /* R8 G8 B8 A8 */
    float texcoord[3];
    int32_t index;
    float barycentric_color_offset[4];
} bxo_vertex; /* 40 bytes, repr(C) */

typedef struct {
    bxo_vertex *vertices;
    size_t n_vertices, cap_vertices;
    uint32_t *indices;
    size_t n_indices, cap_indices;
} bxo_mesh;

void bxo_mesh_free(bxo_mesh *m);
/* voxmesh.rs:16-25 */
int bxo_is_chunk_trivial(const bxo_registry *reg, const uint32_t *rle, size_t len);
/* voxmesh.rs:45-190; chunks = 27 RLE streams in iter_neighbors(cpos,true) order (worldmgr.rs:748-759,
 * slot = rx + 3*rz + 9*ry). Returns 0 ok, -1 unknown block id (reference panics), -2 iterator ran dry. */
int bxo_mesh_from_chunk(const bxo_registry *reg, const uint32_t *const rle[27], const size_t rle_len[27],
                        bxo_mesh *out);
/* Same assemble stage, but the 34^3 decode reads 27 dense chunks directly (the iterator returns exactly
 * blocks_yzx[idx]; equality of the two entry points is tested). */
int bxo_mesh_from_dense27(const bxo_registry *reg, const uint32_t *const dense[27], bxo_mesh *out);

/* ---- RNGs: rand_xoshiro 0.6.0, rand_xorshift 0.2.0, rand 0.7.3 (published algorithms) ---- */
typedef struct { uint64_t x; } bxo_splitmix64;
uint32_t bxo_splitmix64_next_u32(bxo_splitmix64 *s);
uint64_t bxo_splitmix64_next_u64(bxo_splitmix64 *s);
typedef struct { uint64_t s[4]; } bxo_xoshiro256ss;
void bxo_xoshiro256ss_seed_from_u64(bxo_xoshiro256ss *r, uint64_t seed);
uint64_t bxo_xoshiro256ss_next_u64(bxo_xoshiro256ss *r);
uint32_t bxo_xoshiro256ss_next_u32(bxo_xoshiro256ss *r);

/* ---- noise 0.8.2 SuperSimplex (2-D) ---- */
void bxo_permutation_table(uint32_t seed, uint8_t out[256]);
double bxo_super_simplex_2d(const uint8_t perm[256], double x, double y);

/* ---- generator: bxw_world/src/stdgen.rs ---- */
typedef struct bxo_stdgen bxo_stdgen;
bxo_stdgen *bxo_stdgen_new(uint64_t seed);          /* stdgen.rs:90-114, 297-302 */
void bxo_stdgen_free(bxo_stdgen *g);
const uint8_t *bxo_stdgen_perm(const bxo_stdgen *g, int which); /* 0..4 height, 5 elevation, 6 moisture */
typedef struct {
    int32_t height;      /* height.round() as i32 (stdgen.rs:272) */
    int32_t elevation;   /* 0 LowLand, 1 Hill, 2 Mountain */
    double height_f64;   /* before rounding */
    double ec;           /* elevation_noise(pos) (stdgen.rs:219) */
    double cec;          /* nearest cell point's elevation_class (stdgen.rs:218) */
} bxo_voxel_params;
void bxo_calc_voxel_params(bxo_stdgen *g, int32_t x, int32_t z, bxo_voxel_params *out); /* stdgen.rs:215-276 */
/* stdgen.rs:304-374. ids = {void, grass, dirt, stone, snow_grass}. */
void bxo_generate_chunk(bxo_stdgen *g, int32_t cx, int32_t cy, int32_t cz, const uint16_t ids[5],
                        uint32_t blocks_yzx[BXO_CHUNK_DIM3]);

/* ---- bxw_terragen in-tree SuperSimplex: bxw_terragen/src/supersimplex.rs ---- */
typedef struct bxo_tg_simplex bxo_tg_simplex;
bxo_tg_simplex *bxo_tg_simplex_new(uint64_t seed);   /* supersimplex.rs:20-43 */
void bxo_tg_simplex_free(bxo_tg_simplex *s);
double bxo_tg_simplex_noise2(const bxo_tg_simplex *s, double x, double y); /* supersimplex.rs:53-107 */
const uint16_t *bxo_tg_simplex_perm(const bxo_tg_simplex *s);
uint64_t bxo_tg_splitmix64(uint64_t seed, uint64_t *newseed); /* bxw_terragen/src/math.rs:24-30 */

/* ---- CPU baseline driver (what the reference's worker pool does: bxw_util/src/taskpool.rs:54-169,
 *      generation.rs:96-148, voxrender.rs:224-279) ---- */
typedef struct {
    double gen_seconds;   /* wall time of the generate+compress_rle phase */
    double mesh_seconds;  /* wall time of the mesh phase (RLE iterator decode + assemble) */
    uint64_t chunks;      /* interior chunks meshed */
    uint64_t gen_chunks;  /* chunks generated (interior + 1-chunk shell) */
    uint64_t total_vertices, total_indices;
    uint64_t nontrivial;  /* chunks not skipped by is_chunk_trivial */
    int threads;
} bxo_baseline_result;
/* Generates the (nx+2)x(ny+2)x(nz+2) shell-padded region with origin chunk (cx0,cy0,cz0) (interior),
 * then meshes the nx*ny*nz interior chunks, one task per chunk, on `threads` pthreads. */
int bxo_baseline_region(uint64_t seed, int32_t cx0, int32_t cy0, int32_t cz0, int nx, int ny, int nz,
                        int threads, bxo_baseline_result *out);

#ifdef __cplusplus
}
#endif
#endif
