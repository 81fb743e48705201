/*
 * oracle/stdgen.c — TEST INFRASTRUCTURE (see bxw_oracle.h). bxw_world::stdgen::StdGenerator restated from
 * bxw_world/src/stdgen.rs. The noise arithmetic it calls is in noise.c (PARITY UNPINNED, see there).
 */
#include "bxw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GLOBAL_SCALE_MOD 10.0    /* stdgen.rs:13 */
#define GLOBAL_BIOME_SCALE 256.0 /* stdgen.rs:14 */
#define SUPERGRID_SIZE 128       /* stdgen.rs:15 */

typedef struct {
    int32_t pos[2];
    double elevation_class;
} cell_point;

/* The reference memoises get_cell_points in a 64-entry LRU (stdgen.rs:85,96,122-124,145). The value is a pure
 * function of (seed, cell), so any memo gives the same results; a small direct-mapped table is used here. */
#define MEMO_SLOTS 256
typedef struct {
    int valid;
    int32_t cx, cz;
    cell_point pts[4];
} memo_entry;

struct bxo_stdgen {
    uint64_t seed;
    uint8_t perm[7][256]; /* height_map_gen[0..5], elevation_map_gen, moisture_map_gen */
    memo_entry memo[MEMO_SLOTS];
};

bxo_stdgen *bxo_stdgen_new(uint64_t seed) {
    /* CellGen::new + set_seed, stdgen.rs:90-114 */
    bxo_stdgen *g = (bxo_stdgen *)calloc(1, sizeof *g);
    if (!g) return 0;
    bxo_splitmix64 sd = {seed}; /* SplitMix64::seed_from_u64(seed) */
    g->seed = seed;
    for (int i = 0; i < 7; i++) bxo_permutation_table(bxo_splitmix64_next_u32(&sd), g->perm[i]);
    return g;
}

void bxo_stdgen_free(bxo_stdgen *g) { free(g); }
const uint8_t *bxo_stdgen_perm(const bxo_stdgen *g, int which) { return g->perm[which]; }

static double elevation_noise(const bxo_stdgen *g, int32_t x, int32_t z) {
    /* stdgen.rs:171-175 */
    double scale_factor = GLOBAL_BIOME_SCALE * GLOBAL_SCALE_MOD;
    double px = (double)x / scale_factor, pz = (double)z / scale_factor;
    return (bxo_super_simplex_2d(g->perm[5], px, pz) + 1.0) / 2.0;
}

static double h_n(const bxo_stdgen *g, int i, double px, double pz) {
    /* stdgen.rs:178-180 */
    return (bxo_super_simplex_2d(g->perm[i], px, pz) + 1.0) / 2.0;
}

static double h_rn(const bxo_stdgen *g, int i, double px, double pz) {
    /* stdgen.rs:183-185 */
    return 2.0 * (0.5 - fabs(0.5 - h_n(g, i, px, pz)));
}

static double plains_height_noise(const bxo_stdgen *g, int32_t x, int32_t z) {
    /* stdgen.rs:187-192 */
    double scale_factor = GLOBAL_SCALE_MOD * 120.0;
    double px = (double)x / scale_factor, pz = (double)z / scale_factor;
    return (0.75 * h_n(g, 0, px, pz) + 0.25 * h_n(g, 1, px * 2.0, pz * 2.0)) * 5.0 + 10.0;
}

static double hills_height_noise(const bxo_stdgen *g, int32_t x, int32_t z) {
    /* stdgen.rs:194-200 */
    double scale_factor = GLOBAL_SCALE_MOD * 160.0;
    double px = (double)x / scale_factor, pz = (double)z / scale_factor;
    return (0.60 * h_n(g, 0, px, pz) + 0.25 * h_n(g, 1, px * 1.5, pz * 1.5) + 0.15 * h_n(g, 2, px * 3.0, pz * 3.0)) *
               30.0 +
           15.0;
}

static double mountains_height_noise(const bxo_stdgen *g, int32_t x, int32_t z) {
    /* stdgen.rs:202-213 */
    double scale_factor = GLOBAL_SCALE_MOD * 200.0;
    double px = (double)x / scale_factor, pz = (double)z / scale_factor;
    double h0 = 0.50 * h_rn(g, 0, px, pz);
    double h01 = 0.25 * h_rn(g, 1, px * 2.0, pz * 2.0) + h0;
    return (h01 + (h01 / 0.75) * 0.15 * h_n(g, 2, px * 5.0, pz * 5.0) +
            (h01 / 0.75) * 0.05 * h_rn(g, 3, px * 9.0, pz * 9.0)) *
               750.0 +
           40.0;
}

static const cell_point *get_cell_points(bxo_stdgen *g, int32_t cx, int32_t cz) {
    /* stdgen.rs:116-147 */
    uint32_t slot = ((uint32_t)cx * 73856093u ^ (uint32_t)cz * 19349663u) % MEMO_SLOTS;
    memo_entry *m = &g->memo[slot];
    if (m->valid && m->cx == cx && m->cz == cz) return m->pts;
    /* get_seed, stdgen.rs:117-119: `cell.x as u64` sign-extends, then << 32; `cell.y as u64 & 0xFFFF_FFFF` */
    uint64_t s = g->seed ^ ((((uint64_t)(int64_t)cx) << 32) | (((uint64_t)(int64_t)cz) & 0xFFFFFFFFULL));
    bxo_xoshiro256ss r;
    bxo_xoshiro256ss_seed_from_u64(&r, s);
    static const int base[4][2] = {{SUPERGRID_SIZE / 4, SUPERGRID_SIZE / 4},
                                   {3 * SUPERGRID_SIZE / 4, SUPERGRID_SIZE / 4},
                                   {0, 3 * SUPERGRID_SIZE / 4},
                                   {SUPERGRID_SIZE / 2, 3 * SUPERGRID_SIZE / 4}};
    const int MOD = SUPERGRID_SIZE / 4;
    for (int i = 0; i < 4; i++) {
        int32_t xoff = (int32_t)(bxo_xoshiro256ss_next_u32(&r) % (uint32_t)MOD) - MOD / 2;
        int32_t yoff = (int32_t)(bxo_xoshiro256ss_next_u32(&r) % (uint32_t)MOD) - MOD / 2;
        m->pts[i].pos[0] = cx * SUPERGRID_SIZE + base[i][0] + xoff;
        m->pts[i].pos[1] = cz * SUPERGRID_SIZE + base[i][1] + yoff;
        m->pts[i].elevation_class = elevation_noise(g, m->pts[i].pos[0], m->pts[i].pos[1]);
    }
    m->valid = 1;
    m->cx = cx;
    m->cz = cz;
    return m->pts;
}

static cell_point find_nearest_cell_point(bxo_stdgen *g, int32_t x, int32_t z) {
    /* stdgen.rs:149-169 with num == 1: `pos / SUPERGRID_SIZE` is Rust's truncating i32 division;
     * Iterator::min_by returns the FIRST minimum in scan order (cdx outer, cdy inner, 4 points). */
    int32_t cellx = x / SUPERGRID_SIZE, cellz = z / SUPERGRID_SIZE;
    int have = 0;
    int32_t best_d = 0;
    cell_point best;
    memset(&best, 0, sizeof best);
    for (int cdx = -1; cdx <= 1; cdx++)
        for (int cdy = -1; cdy <= 1; cdy++) {
            cell_point pts[4];
            memcpy(pts, get_cell_points(g, cellx + cdx, cellz + cdy), sizeof pts);
            for (int i = 0; i < 4; i++) {
                int32_t dx = pts[i].pos[0] - x, dz = pts[i].pos[1] - z;
                int32_t dist = dx * dx + dz * dz; /* distance2, stdgen.rs:19-21 */
                if (!have || dist < best_d) {
                    have = 1;
                    best_d = dist;
                    best = pts[i];
                }
            }
        }
    return best;
}

/* Rust f64::round (half away from zero) followed by the saturating `as i32` cast. */
static int32_t round_as_i32(double h) {
    double r = round(h);
    if (!(r == r)) return 0;
    if (r >= 2147483647.0) return 2147483647;
    if (r <= -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)r;
}

void bxo_calc_voxel_params(bxo_stdgen *g, int32_t x, int32_t z, bxo_voxel_params *out) {
    /* stdgen.rs:215-276 */
    cell_point cp = find_nearest_cell_point(g, x, z);
    double cec = cp.elevation_class;
    double ec = elevation_noise(g, x, z);
    double height;
    if (ec < 0.4) {
        height = plains_height_noise(g, x, z);
    } else if (ec < 0.5) {
        double nlin = (ec - 0.4) / 0.1;
        double olin = 1.0 - nlin;
        double ph = plains_height_noise(g, x, z);
        double hh = hills_height_noise(g, x, z);
        height = olin * ph + nlin * hh;
    } else if (ec < 0.7) {
        height = hills_height_noise(g, x, z);
    } else if (ec < 0.8) {
        double nlin = (ec - 0.7) / 0.1;
        double olin = 1.0 - nlin;
        double hh = hills_height_noise(g, x, z);
        double mh = mountains_height_noise(g, x, z);
        height = olin * hh + nlin * mh;
    } else {
        height = mountains_height_noise(g, x, z);
    }
    int elevation;
    if (cec < 0.4) {
        elevation = 0;
    } else if (cec < 0.5) {
        double bnoise = bxo_super_simplex_2d(g->perm[4], (double)x, (double)z); /* height_map_gen.last() */
        elevation = bnoise < 0.0 ? 0 : 1;
    } else if (cec < 0.7) {
        elevation = 1;
    } else if (cec < 0.8) {
        double bnoise = bxo_super_simplex_2d(g->perm[4], (double)x, (double)z);
        elevation = bnoise < 0.0 ? 1 : 2;
    } else {
        elevation = 2;
    }
    out->height = round_as_i32(height);
    out->elevation = elevation;
    out->height_f64 = height;
    out->ec = ec;
    out->cec = cec;
}

void bxo_generate_chunk(bxo_stdgen *g, int32_t cx, int32_t cy, int32_t cz, const uint16_t ids[5],
                        uint32_t blocks_yzx[BXO_CHUNK_DIM3]) {
    /* stdgen.rs:304-374 */
    const uint32_t i_air = ids[0], i_grass = ids[1], i_dirt = ids[2], i_stone = ids[3], i_snow_grass = ids[4];
    const int32_t VCD = 32;
    static __thread bxo_voxel_params vparams[BXO_CHUNK_DIM2];
    for (int i = 0; i < BXO_CHUNK_DIM2; i++) {
        int32_t x = (int32_t)(i % 32) + cx * VCD;
        int32_t z = (int32_t)((i / 32) % 32) + cz * VCD;
        bxo_calc_voxel_params(g, x, z, &vparams[i]);
    }
    for (int vidx = 0; vidx < BXO_CHUNK_DIM3; vidx++) {
        int32_t xc = vidx % 32;
        int32_t zc = (vidx / 32) % 32;
        int32_t yc = (vidx / 32 / 32) % 32;
        int32_t y = cy * VCD + yc;
        const bxo_voxel_params *vp = &vparams[xc + zc * 32];
        int32_t h = vp->height;
        uint32_t id;
        if (y == h) {
            id = (vp->elevation == 2 && y > 80) ? i_snow_grass : i_grass;
        } else if (y < h - 5) {
            id = i_stone;
        } else if (y < h) {
            id = i_dirt;
        } else {
            id = i_air;
        }
        blocks_yzx[vidx] = id << 16; /* VoxelDatum::new(id, 0), lib.rs:181-185 */
    }
}
