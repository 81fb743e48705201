/*
 * oracle/mesher.c — TEST INFRASTRUCTURE (see bxw_oracle.h). The client chunk mesher restated loop-for-loop from
 * src/client/render/voxmesh.rs (decode stage :53-80, assemble stage :94-177, packers :29-41,
 * is_chunk_trivial :16-25).
 */
#include "bxw_oracle.h"

#include <stdlib.h>
#include <string.h>

#define INFL 34
#define INFL2 (34 * 34)
#define INFL3 (34 * 34 * 34)

typedef struct {
    uint32_t datum;
    const bxo_block_def *def;
    const bxo_shape *shape;
    int orient;
} decoded_cell;

void bxo_mesh_free(bxo_mesh *m) {
    free(m->vertices);
    free(m->indices);
    memset(m, 0, sizeof *m);
}

static int push_vertex(bxo_mesh *m, const bxo_vertex *v) {
    if (m->n_vertices == m->cap_vertices) {
        size_t nc = m->cap_vertices ? m->cap_vertices * 2 : 6144; /* voxmesh.rs:89 */
        bxo_vertex *p = (bxo_vertex *)realloc(m->vertices, nc * sizeof(bxo_vertex));
        if (!p) return -1;
        m->vertices = p;
        m->cap_vertices = nc;
    }
    m->vertices[m->n_vertices++] = *v;
    return 0;
}

static int push_index(bxo_mesh *m, uint32_t i) {
    if (m->n_indices == m->cap_indices) {
        size_t nc = m->cap_indices ? m->cap_indices * 2 : 6144; /* voxmesh.rs:90 */
        uint32_t *p = (uint32_t *)realloc(m->indices, nc * sizeof(uint32_t));
        if (!p) return -1;
        m->indices = p;
        m->cap_indices = nc;
    }
    m->indices[m->n_indices++] = i;
    return 0;
}

/* Rust `f32 as u32`: truncate toward zero, saturate, NaN -> 0. */
static uint32_t f32_as_u32(float f) {
    if (!(f == f)) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}

static uint32_t as_rgba8(const float c[4]) {
    /* voxmesh.rs:29-34 */
    return ((f32_as_u32(c[0] * 255.0f) & 0xFF) << 24) | ((f32_as_u32(c[1] * 255.0f) & 0xFF) << 16) |
           ((f32_as_u32(c[2] * 255.0f) & 0xFF) << 8) | (f32_as_u32(c[3] * 255.0f) & 0xFF);
}

static uint32_t as_wxyz10(const float p[4]) {
    /* voxmesh.rs:36-41 */
    return ((f32_as_u32(p[3] * 2.0f) & 0x3) << 30) | ((f32_as_u32(p[0] * 512.0f) & 0x3FF) << 20) |
           ((f32_as_u32(p[1] * 512.0f) & 0x3FF) << 10) | (f32_as_u32(p[2] * 512.0f) & 0x3FF);
}

int bxo_is_chunk_trivial(const bxo_registry *reg, const uint32_t *rle, size_t len) {
    /* voxmesh.rs:16-25 */
    if (len == 3 && rle[0] == rle[1]) {
        uint32_t id = rle[0] >> 16;
        if (id < reg->n && reg->defs[id].present && reg->defs[id].mesh_kind == 0) return 1;
    }
    return 0;
}

static int get_block_idx(int x, int y, int z) {
    /* voxmesh.rs:83-87 */
    return (x + 1) + (z + 1) * INFL + (y + 1) * INFL2;
}

static int decode_cell(const bxo_registry *reg, uint32_t datum, decoded_cell *c) {
    /* voxmesh.rs:76-79 */
    uint32_t id = datum >> 16;
    if (id >= reg->n || !reg->defs[id].present) return -1; /* voxregistry.rs:141-143 unwrap() panics */
    c->datum = datum;
    c->def = &reg->defs[id];
    c->shape = bxo_shape_table(bxo_block_shape(datum, c->def->mesh_kind));
    c->orient = bxo_block_orientation(datum, c->def->mesh_kind);
    return 0;
}

static int assemble(const decoded_cell *vd, bxo_mesh *out) {
    /* voxmesh.rs:94-177 */
    const float AO_OCCLUSION_FACTOR = 0.88f; /* voxmesh.rs:27 */
    for (int cell_y = 0; cell_y < 32; cell_y++)
        for (int cell_z = 0; cell_z < 32; cell_z++)
            for (int cell_x = 0; cell_x < 32; cell_x++) {
                const int ipos[3] = {cell_x, cell_y, cell_z};
                const decoded_cell *c = &vd[get_block_idx(cell_x, cell_y, cell_z)];
                const int ic_vidx = cell_x + 32 * cell_z + 1024 * cell_y; /* as_blockidx, lib.rs:104-108 */
                if (c->def->mesh_kind == 0) continue;

                int mi[9];
                bxo_orient_matrix(c->orient, mi);
                for (int side_dir = 0; side_dir < 6; side_dir++) {
                    int rot_side_dir = bxo_orient_unapply(c->orient, side_dir);
                    const bxo_vsside *side = &c->shape->sides[rot_side_dir];
                    if (side->n_idx == 0) continue;
                    int ioff[3];
                    bxo_dir_to_vec(side_dir, ioff);

                    /* hidden face removal, voxmesh.rs:112-121 */
                    int touchside = bxo_dir_opposite(side_dir);
                    const decoded_cell *t =
                        &vd[get_block_idx(ipos[0] + ioff[0], ipos[1] + ioff[1], ipos[2] + ioff[2])];
                    int touchrotside = bxo_orient_unapply(t->orient, touchside);
                    const bxo_vsside *tside = &t->shape->sides[touchrotside];
                    if (side->can_be_clipped && tside->can_clip) continue;

                    uint32_t voff = (uint32_t)out->n_vertices;
                    float bsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    for (int vi = 0; vi < side->n_verts; vi++) {
                        const bxo_vsvertex *vtx = &side->verts[vi];
                        /* AO, voxmesh.rs:129-137 */
                        float ao = 1.0f;
                        for (int k = 0; k < vtx->n_ao; k++) {
                            const int *a = vtx->ao[k];
                            int p[3];
                            for (int r = 0; r < 3; r++)
                                p[r] = ipos[r] + mi[r * 3] * a[0] + mi[r * 3 + 1] * a[1] + mi[r * 3 + 2] * a[2];
                            const decoded_cell *s = &vd[get_block_idx(p[0], p[1], p[2])];
                            if (s->shape->causes_ambient_occlusion) ao *= AO_OCCLUSION_FACTOR;
                        }
                        /* voxmesh.rs:139-145 */
                        float vo[3];
                        for (int r = 0; r < 3; r++) {
                            /* nalgebra Matrix3<f32> * Vector3<f32>: column-major axpy accumulation
                             * (col0*x, then += col1*y, then += col2*z); all terms are 0 or +-0.5 so exact. */
                            float acc = (float)mi[r * 3] * vtx->offset[0];
                            acc += (float)mi[r * 3 + 1] * vtx->offset[1];
                            acc += (float)mi[r * 3 + 2] * vtx->offset[2];
                            vo[r] = acc;
                        }
                        float position[4] = {(float)ipos[0] + vo[0] + 0.5f, (float)ipos[1] + vo[1] + 0.5f,
                                             (float)ipos[2] + vo[2] + 0.5f, 1.0f};
                        uint32_t texid = c->def->tex[rot_side_dir]; /* voxmesh.rs:146 */
                        float color[4] = {c->def->debug_color[0] * ao, c->def->debug_color[1] * ao,
                                          c->def->debug_color[2] * ao, 1.0f};
                        for (int r = 0; r < 4; r++) bsum[r] += (float)vtx->barycentric_sign * color[r];
                        int32_t idx_flags = ic_vidx;
                        if (vtx->barycentric[0] > 0.1f) idx_flags |= 1 << 17;
                        if (vtx->barycentric[1] > 0.1f) idx_flags |= 1 << 18;
                        if (vtx->barycentric[2] > 0.1f) idx_flags |= 1 << 19;
                        float pdiv[4];
                        for (int r = 0; r < 4; r++) pdiv[r] = position[r] / 32.0f;
                        bxo_vertex v;
                        v.position = as_wxyz10(pdiv);
                        v.color = as_rgba8(color);
                        v.texcoord[0] = vtx->texcoord[0];
                        v.texcoord[1] = vtx->texcoord[1];
                        v.texcoord[2] = (float)texid;
                        v.index = idx_flags;
                        memset(v.barycentric_color_offset, 0, sizeof v.barycentric_color_offset);
                        if (push_vertex(out, &v)) return -3;
                    }
                    for (size_t k = voff; k < out->n_vertices; k++)
                        memcpy(out->vertices[k].barycentric_color_offset, bsum, sizeof bsum);
                    for (int k = 0; k < side->n_idx; k++)
                        if (push_index(out, side->idx[k] + voff)) return -3;
                }
            }
    return 0;
}

int bxo_mesh_from_chunk(const bxo_registry *reg, const uint32_t *const rle[27], const size_t rle_len[27],
                        bxo_mesh *out) {
    memset(out, 0, sizeof *out);
    bxo_shape_table(0);
    bxo_rle_iter its[27];
    for (int i = 0; i < 27; i++) bxo_rle_iter_init(&its[i], rle[i], rle_len[i]);
    decoded_cell *vd = (decoded_cell *)malloc(sizeof(decoded_cell) * INFL3);
    if (!vd) return -3;
    size_t n = 0;
    int rc = 0;
    /* voxmesh.rs:62-80 */
    for (int y = 0; y < INFL && !rc; y++)
        for (int z = 0; z < INFL && !rc; z++)
            for (int x = 0; x < INFL; x++) {
                int rx = x == 0 ? 0 : (x <= 32 ? 1 : 2);
                int ry = y == 0 ? 0 : (y <= 32 ? 1 : 2);
                int rz = z == 0 ? 0 : (z <= 32 ? 1 : 2);
                int cidx = rx + rz * 3 + ry * 9;
                /* BlockPosition(x-1,y-1,z-1).as_blockidx() with rem_floor, lib.rs:104-108 */
                int bx = (x - 1) & 31, by = (y - 1) & 31, bz = (z - 1) & 31;
                uint32_t bpos = (uint32_t)(bx + 32 * bz + 1024 * by);
                uint32_t datum, idx;
                if (!bxo_rle_iter_skip_until(&its[cidx], bpos, &datum, &idx)) {
                    rc = -2;
                    break;
                }
                if (decode_cell(reg, datum, &vd[n++])) {
                    rc = -1;
                    break;
                }
            }
    if (!rc) rc = assemble(vd, out);
    free(vd);
    if (rc) bxo_mesh_free(out);
    return rc;
}

int bxo_mesh_from_dense27(const bxo_registry *reg, const uint32_t *const dense[27], bxo_mesh *out) {
    memset(out, 0, sizeof *out);
    bxo_shape_table(0);
    decoded_cell *vd = (decoded_cell *)malloc(sizeof(decoded_cell) * INFL3);
    if (!vd) return -3;
    size_t n = 0;
    int rc = 0;
    for (int y = 0; y < INFL && !rc; y++)
        for (int z = 0; z < INFL && !rc; z++)
            for (int x = 0; x < INFL; x++) {
                int rx = x == 0 ? 0 : (x <= 32 ? 1 : 2);
                int ry = y == 0 ? 0 : (y <= 32 ? 1 : 2);
                int rz = z == 0 ? 0 : (z <= 32 ? 1 : 2);
                int cidx = rx + rz * 3 + ry * 9;
                int bx = (x - 1) & 31, by = (y - 1) & 31, bz = (z - 1) & 31;
                uint32_t datum = dense[cidx][bx + 32 * bz + 1024 * by];
                if (decode_cell(reg, datum, &vd[n++])) {
                    rc = -1;
                    break;
                }
            }
    if (!rc) rc = assemble(vd, out);
    free(vd);
    if (rc) bxo_mesh_free(out);
    return rc;
}
