/*
 * oracle/tables.c — TEST INFRASTRUCTURE (see bxw_oracle.h). Direction / orientation algebra, the standard
 * block registry and the five VoxelShapeDef tables, restated from the reference sources cited per function.
 */
#include "bxw_oracle.h"

#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Directions (bxw_util/src/direction.rs)
 * ---------------------------------------------------------------------------------------------- */

void bxo_dir_to_vec(int dir, int out[3]) {
    /* direction.rs:89-99 */
    out[0] = out[1] = out[2] = 0;
    out[dir >> 1] = (dir & 1) ? 1 : -1;
}

int bxo_dir_from_vec(const int v[3]) {
    /* direction.rs:76-87: exactly one component is +-1, the others 0 */
    int nz = 0, axis = -1;
    for (int a = 0; a < 3; a++) {
        if (v[a] != 0) {
            nz++;
            axis = a;
        }
    }
    if (nz != 1) return -1;
    if (v[axis] == 1) return axis * 2 + 1;
    if (v[axis] == -1) return axis * 2;
    return -1;
}

int bxo_dir_opposite(int dir) { return dir ^ 1; /* direction.rs:43-53 */ }

int bxo_lh_cross(int a, int b) {
    /* direction.rs:157-161 looks up DIRECTION_CROSS_TABLE (452-492); the reference test
     * direction_cross_verify (429-448) pins that table to -(a x b), which is what we evaluate. */
    int va[3], vb[3], c[3];
    bxo_dir_to_vec(a, va);
    bxo_dir_to_vec(b, vb);
    c[0] = -(va[1] * vb[2] - va[2] * vb[1]);
    c[1] = -(va[2] * vb[0] - va[0] * vb[2]);
    c[2] = -(va[0] * vb[1] - va[1] * vb[0]);
    return bxo_dir_from_vec(c);
}

int bxo_orient_from_index(int idx, int *right, int *up) {
    /* direction.rs:267-284 */
    if (idx < 0 || idx >= 24) return -1;
    int i = (idx + 5) % 24;
    int right_idx = i / 4;
    int up_idx = i % 4;
    if ((up_idx / 2) >= (right_idx / 2)) up_idx += 2;
    *right = right_idx;
    *up = up_idx;
    return 0;
}

int bxo_orient_to_index(int right, int up) {
    /* direction.rs:249-264 */
    int right_idx = right;
    int up_idx = up;
    if (up_idx > right_idx) up_idx -= 2;
    return (right_idx * 4 + up_idx + 24 - 5) % 24;
}

int bxo_orient_from_dirs(int right, int up, int front) {
    /* direction.rs:216-222 */
    int c = bxo_lh_cross(right, up);
    return (c >= 0 && c == front) ? 1 : 0;
}

void bxo_orient_matrix(int idx, int m[9]) {
    /* direction.rs:287-293: columns are right, up, back; back = lh_cross(up, right) (344-346) */
    int right, up;
    bxo_orient_from_index(idx, &right, &up);
    int back = bxo_lh_cross(up, right);
    int cr[3], cu[3], cb[3];
    bxo_dir_to_vec(right, cr);
    bxo_dir_to_vec(up, cu);
    bxo_dir_to_vec(back, cb);
    for (int r = 0; r < 3; r++) {
        m[r * 3 + 0] = cr[r];
        m[r * 3 + 1] = cu[r];
        m[r * 3 + 2] = cb[r];
    }
}

int bxo_orient_apply(int idx, int dir) {
    /* direction.rs:196-209 (LUT .0) : try_from_vec(M * dir) */
    int m[9], v[3], o[3];
    bxo_orient_matrix(idx, m);
    bxo_dir_to_vec(dir, v);
    for (int r = 0; r < 3; r++) o[r] = m[r * 3] * v[0] + m[r * 3 + 1] * v[1] + m[r * 3 + 2] * v[2];
    return bxo_dir_from_vec(o);
}

int bxo_orient_unapply(int idx, int dir) {
    /* direction.rs:196-209 (LUT .1) : try_from_vec(M^T * dir) */
    int m[9], v[3], o[3];
    bxo_orient_matrix(idx, m);
    bxo_dir_to_vec(dir, v);
    for (int c = 0; c < 3; c++) o[c] = m[0 * 3 + c] * v[0] + m[1 * 3 + c] * v[1] + m[2 * 3 + c] * v[2];
    return bxo_dir_from_vec(o);
}

/* ------------------------------------------------------------------------------------------------
 * Registry (voxregistry.rs:94-111, blocks/mod.rs:6-65, lib.rs:602-606)
 * ---------------------------------------------------------------------------------------------- */

static void def_single(bxo_block_def *d, uint32_t t) {
    d->present = 1;
    d->mesh_kind = 1;
    d->debug_color[0] = d->debug_color[1] = d->debug_color[2] = 1.0f; /* voxregistry.rs:130 */
    for (int i = 0; i < 6; i++) d->tex[i] = t;
}

static void def_tsb(bxo_block_def *d, uint32_t top, uint32_t side, uint32_t bottom) {
    /* lib.rs:602-606: [side, side, bottom, top, side, side] */
    def_single(d, side);
    d->tex[2] = bottom;
    d->tex[3] = top;
}

void bxo_standard_registry(bxo_block_def out[8]) {
    /* Texture ids: rank in the byte-wise sorted key list of res/manifest.toml (92 keys):
     * dbg_back 9, dbg_down 10, dbg_front 11, dbg_left 12, dbg_right 13, dbg_up 14, dirt 15, dirt_grass 16,
     * dirt_snow 18, grass_top 29, snow 55, stone 56, stone_diamond 61, table 73, wood 90. */
    memset(out, 0, 8 * sizeof(bxo_block_def));
    out[0].present = 1; /* core:void, mesh None, debug colour 0 (voxregistry.rs:101-108) */
    out[0].mesh_kind = 0;
    def_tsb(&out[1], 29, 16, 15); /* core:grass      new_tsb("grass_top","dirt_grass","dirt") */
    def_tsb(&out[2], 55, 18, 15); /* core:snow_grass new_tsb("snow","dirt_snow","dirt") */
    def_single(&out[3], 15);      /* core:dirt */
    def_single(&out[4], 56);      /* core:stone */
    def_single(&out[5], 61);      /* core:diamond_ore */
    def_single(&out[6], 0);       /* core:debug: [left,right,down,up,front,back] */
    out[6].tex[0] = 12;
    out[6].tex[1] = 13;
    out[6].tex[2] = 10;
    out[6].tex[3] = 14;
    out[6].tex[4] = 11;
    out[6].tex[5] = 9;
    def_tsb(&out[7], 73, 90, 73); /* core:table new_tsb("table","wood","table") */
}

/* ------------------------------------------------------------------------------------------------
 * Shapes (bxw_world/src/blocks/stdshapes.rs)
 * ---------------------------------------------------------------------------------------------- */

static int isign_thresh(float x) {
    /* stdshapes.rs:132,160: |x| < 0.1 -> 0 else signum */
    if (fabsf(x) < 0.1f) return 0;
    return x > 0.0f ? 1 : -1;
}

static void corner_ao_set(bxo_vsvertex *v, const float corner[3], const int inormal[3]) {
    /* stdshapes.rs:131-152 */
    int ic[3] = {isign_thresh(corner[0]), isign_thresh(corner[1]), isign_thresh(corner[2])};
    int n = 0;
    memcpy(v->ao[n++], ic, sizeof ic);
    memcpy(v->ao[n++], inormal, sizeof ic);
    for (int a = 0; a < 3; a++) {
        if (inormal[a] == 0) {
            int c[3] = {ic[0], ic[1], ic[2]};
            c[a] = 0;
            memcpy(v->ao[n++], c, sizeof c);
        }
    }
    v->n_ao = n;
}

static void set_vertex(bxo_vsvertex *v, const float off[3], float tu, float tv, float bx, float by, float bz,
                       int sign, const float corner[3], const int inormal[3]) {
    memcpy(v->offset, off, 3 * sizeof(float));
    v->texcoord[0] = tu;
    v->texcoord[1] = tv;
    v->barycentric[0] = bx;
    v->barycentric[1] = by;
    v->barycentric[2] = bz;
    v->barycentric_sign = sign;
    corner_ao_set(v, corner, inormal);
}

static void quad_side(bxo_vsside *s, int can_clip, int can_be_clipped, float cx, float cy, float cz, float rx,
                      float ry, float rz, float ux, float uy, float uz) {
    /* stdshapes.rs:154-198 quad_verts + QUAD_INDICES */
    const float c[3] = {cx, cy, cz}, r[3] = {rx, ry, rz}, u[3] = {ux, uy, uz};
    float fn[3] = {-(r[1] * u[2] - r[2] * u[1]), -(r[2] * u[0] - r[0] * u[2]), -(r[0] * u[1] - r[1] * u[0])};
    int inormal[3] = {isign_thresh(fn[0]), isign_thresh(fn[1]), isign_thresh(fn[2])};
    float o[4][3];
    for (int a = 0; a < 3; a++) {
        o[0][a] = c[a] - r[a] - u[a];
        o[1][a] = c[a] - r[a] + u[a];
        o[2][a] = c[a] + r[a] + u[a];
        o[3][a] = c[a] + r[a] - u[a];
    }
    s->can_clip = can_clip;
    s->can_be_clipped = can_be_clipped;
    s->n_verts = 4;
    set_vertex(&s->verts[0], o[0], 0.0f, 1.0f, 0.0f, 1.0f, 1.0f, -1, o[0], inormal);
    set_vertex(&s->verts[1], o[1], 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 1, o[1], inormal);
    set_vertex(&s->verts[2], o[2], 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, -1, o[2], inormal);
    set_vertex(&s->verts[3], o[3], 1.0f, 1.0f, 0.0f, 0.0f, 1.0f, 1, o[3], inormal);
    static const uint32_t qi[6] = {0, 1, 2, 2, 3, 0};
    s->n_idx = 6;
    memcpy(s->idx, qi, sizeof qi);
}

static void empty_side(bxo_vsside *s, int can_clip, int can_be_clipped) {
    memset(s, 0, sizeof *s);
    s->can_clip = can_clip;
    s->can_be_clipped = can_be_clipped;
}

/* One explicit triangle-list vertex row: offset*2, texcoord, barycentric, AO corner, AO inormal. */
typedef struct {
    signed char o[3];
    unsigned char t[2];
    unsigned char b[3];
    signed char c[3];
    signed char n[3];
} tri_row;

static void tri_side(bxo_vsside *s, int can_clip, int can_be_clipped, const tri_row *rows, int n) {
    memset(s, 0, sizeof *s);
    s->can_clip = can_clip;
    s->can_be_clipped = can_be_clipped;
    s->n_verts = n;
    for (int i = 0; i < n; i++) {
        float off[3] = {rows[i].o[0] * 0.5f, rows[i].o[1] * 0.5f, rows[i].o[2] * 0.5f};
        float corner[3] = {(float)rows[i].c[0], (float)rows[i].c[1], (float)rows[i].c[2]};
        int inormal[3] = {rows[i].n[0], rows[i].n[1], rows[i].n[2]};
        set_vertex(&s->verts[i], off, (float)rows[i].t[0], (float)rows[i].t[1], (float)rows[i].b[0],
                   (float)rows[i].b[1], (float)rows[i].b[2], 0, corner, inormal);
        s->idx[i] = (uint32_t)i;
    }
    s->n_idx = n;
}

static bxo_shape g_shapes[5];
static int g_shapes_ready = 0;

static void side_xm_cube(bxo_vsside *s, int cc, int cbc) { quad_side(s, cc, cbc, -.5f, 0, 0, 0, 0, -.5f, 0, .5f, 0); }
static void side_xp_cube(bxo_vsside *s, int cc, int cbc) { quad_side(s, cc, cbc, .5f, 0, 0, 0, 0, .5f, 0, .5f, 0); }
static void side_ym_cube(bxo_vsside *s, int cc, int cbc) { quad_side(s, cc, cbc, 0, -.5f, 0, .5f, 0, 0, 0, 0, -.5f); }
static void side_yp_cube(bxo_vsside *s, int cc, int cbc) { quad_side(s, cc, cbc, 0, .5f, 0, .5f, 0, 0, 0, 0, .5f); }
static void side_zm_cube(bxo_vsside *s, int cc, int cbc) { quad_side(s, cc, cbc, 0, 0, -.5f, .5f, 0, 0, 0, .5f, 0); }
static void side_zp_cube(bxo_vsside *s, int cc, int cbc) { quad_side(s, cc, cbc, 0, 0, .5f, -.5f, 0, 0, 0, .5f, 0); }

static void init_shapes(void) {
    memset(g_shapes, 0, sizeof g_shapes);

    /* init_no_shape, stdshapes.rs:111-129 */
    g_shapes[0].causes_ambient_occlusion = 0;
    for (int i = 0; i < 6; i++) empty_side(&g_shapes[0].sides[i], 0, 1);

    /* init_cube_shape, stdshapes.rs:200-272 */
    {
        bxo_shape *sh = &g_shapes[1];
        sh->causes_ambient_occlusion = 1;
        side_xm_cube(&sh->sides[0], 1, 1);
        side_xp_cube(&sh->sides[1], 1, 1);
        side_ym_cube(&sh->sides[2], 1, 1);
        side_yp_cube(&sh->sides[3], 1, 1);
        side_zm_cube(&sh->sides[4], 1, 1);
        side_zp_cube(&sh->sides[5], 1, 1);
    }

    /* The X- triangle shared by slope and corner (stdshapes.rs:279-309, 391-421) */
    static const tri_row left_tri[3] = {
        {{-1, -1, 1}, {0, 1}, {0, 0, 1}, {-1, -1, 1}, {-1, 0, 0}},
        {{-1, 1, 1}, {0, 0}, {1, 0, 0}, {-1, 1, 1}, {-1, 0, 0}},
        {{-1, -1, -1}, {1, 1}, {0, 1, 0}, {-1, -1, -1}, {-1, 0, 0}},
    };

    /* init_slope_shape, stdshapes.rs:274-384 */
    {
        bxo_shape *sh = &g_shapes[2];
        sh->causes_ambient_occlusion = 1;
        tri_side(&sh->sides[0], 0, 1, left_tri, 3);
        static const tri_row right_tri[3] = {
            {{1, -1, -1}, {0, 1}, {0, 0, 1}, {1, -1, -1}, {1, 0, 0}},
            {{1, 1, 1}, {1, 0}, {1, 0, 0}, {1, 1, 1}, {1, 0, 0}},
            {{1, -1, 1}, {1, 1}, {0, 1, 0}, {1, -1, 1}, {1, 0, 0}},
        };
        tri_side(&sh->sides[1], 0, 1, right_tri, 3);
        side_ym_cube(&sh->sides[2], 1, 1);
        quad_side(&sh->sides[3], 0, 0, 0, 0, 0, .5f, 0, 0, 0, .5f, .5f); /* stdshapes.rs:354-363 */
        empty_side(&sh->sides[4], 0, 1);
        side_zp_cube(&sh->sides[5], 1, 1);
    }

    /* init_corner_shape, stdshapes.rs:386-538 */
    {
        bxo_shape *sh = &g_shapes[3];
        sh->causes_ambient_occlusion = 1;
        tri_side(&sh->sides[0], 0, 1, left_tri, 3);
        empty_side(&sh->sides[1], 0, 1);
        side_ym_cube(&sh->sides[2], 1, 1);
        static const tri_row top6[6] = {
            {{1, -1, -1}, {1, 1}, {0, 0, 1}, {1, -1, -1}, {0, 0, -1}},
            {{-1, -1, -1}, {0, 1}, {1, 0, 0}, {-1, -1, -1}, {0, 0, -1}},
            {{-1, 1, 1}, {0, 0}, {0, 1, 0}, {-1, 1, 1}, {0, 1, 0}},
            {{-1, 1, 1}, {1, 0}, {0, 0, 1}, {-1, 1, 1}, {0, 1, 0}},
            {{1, -1, 1}, {1, 1}, {1, 0, 0}, {1, -1, 1}, {1, 0, 0}},
            {{1, -1, -1}, {0, 1}, {0, 1, 0}, {1, -1, -1}, {1, 0, 0}},
        };
        tri_side(&sh->sides[3], 0, 0, top6, 6);
        empty_side(&sh->sides[4], 0, 1);
        static const tri_row back_tri[3] = {
            {{-1, 1, 1}, {1, 1}, {0, 0, 1}, {-1, 1, 1}, {0, 0, 1}},
            {{-1, -1, 1}, {1, 0}, {1, 0, 0}, {-1, -1, 1}, {0, 0, 1}},
            {{1, -1, 1}, {0, 0}, {0, 1, 0}, {1, -1, 1}, {0, 0, 1}},
        };
        tri_side(&sh->sides[5], 0, 1, back_tri, 3);
    }

    /* init_inner_corner_shape, stdshapes.rs:540-700 */
    {
        bxo_shape *sh = &g_shapes[4];
        sh->causes_ambient_occlusion = 1;
        static const tri_row left_tri_ic[3] = {
            {{-1, -1, 1}, {0, 1}, {0, 0, 1}, {-1, -1, 1}, {-1, 0, 0}},
            {{-1, 1, -1}, {1, 0}, {1, 0, 0}, {-1, 1, -1}, {-1, 0, 0}},
            {{-1, -1, -1}, {1, 1}, {0, 1, 0}, {-1, -1, -1}, {-1, 0, 0}},
        };
        tri_side(&sh->sides[0], 0, 1, left_tri_ic, 3);
        side_xp_cube(&sh->sides[1], 1, 1);
        side_ym_cube(&sh->sides[2], 1, 1);
        static const tri_row top6[6] = {
            {{1, 1, -1}, {0, 0}, {0, 0, 1}, {1, 1, -1}, {0, 1, 0}},
            {{-1, 1, -1}, {1, 0}, {1, 0, 0}, {-1, 1, -1}, {0, 1, 0}},
            {{-1, -1, 1}, {1, 1}, {0, 1, 0}, {-1, -1, 1}, {0, 0, 0}},
            {{-1, -1, 1}, {0, 1}, {0, 0, 1}, {-1, -1, 1}, {0, 0, 0}},
            {{1, 1, 1}, {0, 0}, {1, 0, 0}, {1, 1, 1}, {0, 1, 0}},
            {{1, 1, -1}, {1, 0}, {0, 1, 0}, {1, 1, -1}, {0, 1, 0}},
        };
        tri_side(&sh->sides[3], 0, 0, top6, 6);
        side_zm_cube(&sh->sides[4], 1, 1);
        static const tri_row back_tri[3] = {
            {{1, 1, 1}, {0, 1}, {0, 0, 1}, {1, 1, 1}, {0, 0, 1}},
            {{-1, -1, 1}, {1, 0}, {1, 0, 0}, {-1, -1, 1}, {0, 0, 1}},
            {{1, -1, 1}, {0, 0}, {0, 1, 0}, {1, -1, 1}, {0, 0, 1}},
        };
        tri_side(&sh->sides[5], 0, 1, back_tri, 3);
    }
    g_shapes_ready = 1;
}

const bxo_shape *bxo_shape_table(int shape_index) {
    if (!g_shapes_ready) init_shapes(); /* idempotent; callers that go multi-threaded call once first */
    if (shape_index < 0 || shape_index > 4) return 0;
    return &g_shapes[shape_index];
}

int bxo_block_shape(uint32_t datum, int mesh_kind) {
    /* stdshapes.rs:51-62 with StdMeta::from_meta (26-31): shape = meta % 64 */
    if (mesh_kind == 0) return 0;
    uint32_t meta = datum & 0xFFFFu;
    uint32_t shape = meta % 64u;
    switch (shape) {
    case 0: return 1;
    case 1: return 2;
    case 2: return 3;
    case 3: return 4;
    default: return 1;
    }
}

int bxo_block_orientation(uint32_t datum, int mesh_kind) {
    /* stdshapes.rs:64-72: orientation = meta / 64, from_index(..).unwrap_or_default() */
    if (mesh_kind == 0) return 0;
    uint32_t meta = datum & 0xFFFFu;
    uint32_t o = meta / 64u;
    return o < 24u ? (int)o : 0;
}
