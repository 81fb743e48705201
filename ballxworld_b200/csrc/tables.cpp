// tables.cpp — builds the pre-rotated mesher tables (see tables.hpp).
#include "tables.hpp"

#include <cstring>

namespace bxw {
namespace {

struct V3 {
    int x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline int sgn(int v) { return v > 0 ? 1 : (v < 0 ? -1 : 0); }

// A vertex of a side in block-local space. Offsets are in HALF units (+-1 == +-0.5): every vertex of every
// standard shape sits on a corner of the unit cube (stdshapes.rs:154-196 and the explicit lists).
struct LVert {
    V3 off;       // +-1 per axis
    int u, v;     // texcoord
    int bary;     // bit0 x>0.1, bit1 y>0.1, bit2 z>0.1
    int sign;     // barycentric_sign
    V3 inormal;   // second AO sample / axis mask for the edge samples (corner_ao_set, stdshapes.rs:131-152)
};

struct LSide {
    bool can_clip, can_be_clipped;
    int nv;
    LVert v[6];
    int ni;
    int idx[6];
};

struct LShape {
    LSide side[6];
};

LSide empty_side(bool cc, bool cbc) {
    LSide s{};
    s.can_clip = cc;
    s.can_be_clipped = cbc;
    return s;
}

// quad_verts(center, right, up) with QUAD_INDICES (stdshapes.rs:154-198); arguments in half units.
LSide quad(bool cc, bool cbc, V3 c, V3 r, V3 u) {
    LSide s{};
    s.can_clip = cc;
    s.can_be_clipped = cbc;
    V3 n = cross(r, u);
    // fnormal = -right x up; components are 0 or +-1 (half-unit products), the 0.1 threshold keeps them
    V3 inormal = {sgn(-n.x), sgn(-n.y), sgn(-n.z)};
    const V3 offs[4] = {c - r - u, c - r + u, c + r + u, c + r - u};
    const int tu[4] = {0, 0, 1, 1}, tv[4] = {1, 0, 0, 1};
    const int bary[4] = {0b110, 0b100, 0b101, 0b100}; // (0,1,1) (0,0,1) (1,0,1) (0,0,1)
    const int sign[4] = {-1, 1, -1, 1};
    s.nv = 4;
    for (int i = 0; i < 4; i++) s.v[i] = LVert{offs[i], tu[i], tv[i], bary[i], sign[i], inormal};
    s.ni = 6;
    const int qi[6] = {0, 1, 2, 2, 3, 0};
    std::memcpy(s.idx, qi, sizeof qi);
    return s;
}

// explicit triangle list; each row: off(3) u v barycentric-unit-axis normal(3). In every explicit list of the
// reference the AO corner argument equals sign(offset) and the barycentric vector is a unit vector.
struct TRow {
    int ox, oy, oz, u, v, baxis, nx, ny, nz;
};
LSide tris(bool cc, bool cbc, const TRow *rows, int n) {
    LSide s{};
    s.can_clip = cc;
    s.can_be_clipped = cbc;
    s.nv = n;
    s.ni = n;
    for (int i = 0; i < n; i++) {
        s.v[i] = LVert{{rows[i].ox, rows[i].oy, rows[i].oz}, rows[i].u, rows[i].v, 1 << rows[i].baxis, 0,
                       {rows[i].nx, rows[i].ny, rows[i].nz}};
        s.idx[i] = i;
    }
    return s;
}

// cube faces (stdshapes.rs:203-270), half units
LSide face_xm(bool cc, bool cbc) { return quad(cc, cbc, {-1, 0, 0}, {0, 0, -1}, {0, 1, 0}); }
LSide face_xp(bool cc, bool cbc) { return quad(cc, cbc, {1, 0, 0}, {0, 0, 1}, {0, 1, 0}); }
LSide face_ym(bool cc, bool cbc) { return quad(cc, cbc, {0, -1, 0}, {1, 0, 0}, {0, 0, -1}); }
LSide face_yp(bool cc, bool cbc) { return quad(cc, cbc, {0, 1, 0}, {1, 0, 0}, {0, 0, 1}); }
LSide face_zm(bool cc, bool cbc) { return quad(cc, cbc, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}); }
LSide face_zp(bool cc, bool cbc) { return quad(cc, cbc, {0, 0, 1}, {-1, 0, 0}, {0, 1, 0}); }

void build_shapes(LShape sh[4]) {
    // cube (stdshapes.rs:200-272)
    sh[0].side[0] = face_xm(true, true);
    sh[0].side[1] = face_xp(true, true);
    sh[0].side[2] = face_ym(true, true);
    sh[0].side[3] = face_yp(true, true);
    sh[0].side[4] = face_zm(true, true);
    sh[0].side[5] = face_zp(true, true);

    // the X- wall triangle of slope and outer corner (stdshapes.rs:279-309, 391-421)
    const TRow wall_xm[3] = {{-1, -1, 1, 0, 1, 2, -1, 0, 0}, {-1, 1, 1, 0, 0, 0, -1, 0, 0}, {-1, -1, -1, 1, 1, 1, -1, 0, 0}};

    // slope (stdshapes.rs:274-384): ramp rising towards Z+
    {
        const TRow wall_xp[3] = {{1, -1, -1, 0, 1, 2, 1, 0, 0}, {1, 1, 1, 1, 0, 0, 1, 0, 0}, {1, -1, 1, 1, 1, 1, 1, 0, 0}};
        sh[1].side[0] = tris(false, true, wall_xm, 3);
        sh[1].side[1] = tris(false, true, wall_xp, 3);
        sh[1].side[2] = face_ym(true, true);
        sh[1].side[3] = quad(false, false, {0, 0, 0}, {1, 0, 0}, {0, 1, 1}); // the ramp (stdshapes.rs:354-363)
        sh[1].side[4] = empty_side(false, true);
        sh[1].side[5] = face_zp(true, true);
    }
    // outer corner (stdshapes.rs:386-538)
    {
        const TRow top[6] = {{1, -1, -1, 1, 1, 2, 0, 0, -1}, {-1, -1, -1, 0, 1, 0, 0, 0, -1}, {-1, 1, 1, 0, 0, 1, 0, 1, 0},
                             {-1, 1, 1, 1, 0, 2, 0, 1, 0},   {1, -1, 1, 1, 1, 0, 1, 0, 0},    {1, -1, -1, 0, 1, 1, 1, 0, 0}};
        const TRow wall_zp[3] = {{-1, 1, 1, 1, 1, 2, 0, 0, 1}, {-1, -1, 1, 1, 0, 0, 0, 0, 1}, {1, -1, 1, 0, 0, 1, 0, 0, 1}};
        sh[2].side[0] = tris(false, true, wall_xm, 3);
        sh[2].side[1] = empty_side(false, true);
        sh[2].side[2] = face_ym(true, true);
        sh[2].side[3] = tris(false, false, top, 6);
        sh[2].side[4] = empty_side(false, true);
        sh[2].side[5] = tris(false, true, wall_zp, 3);
    }
    // inner corner (stdshapes.rs:540-700)
    {
        const TRow wall_xm_ic[3] = {{-1, -1, 1, 0, 1, 2, -1, 0, 0}, {-1, 1, -1, 1, 0, 0, -1, 0, 0}, {-1, -1, -1, 1, 1, 1, -1, 0, 0}};
        const TRow top[6] = {{1, 1, -1, 0, 0, 2, 0, 1, 0}, {-1, 1, -1, 1, 0, 0, 0, 1, 0}, {-1, -1, 1, 1, 1, 1, 0, 0, 0},
                             {-1, -1, 1, 0, 1, 2, 0, 0, 0}, {1, 1, 1, 0, 0, 0, 0, 1, 0},  {1, 1, -1, 1, 0, 1, 0, 1, 0}};
        const TRow wall_zp[3] = {{1, 1, 1, 0, 1, 2, 0, 0, 1}, {-1, -1, 1, 1, 0, 0, 0, 0, 1}, {1, -1, 1, 0, 0, 1, 0, 0, 1}};
        sh[3].side[0] = tris(false, true, wall_xm_ic, 3);
        sh[3].side[1] = face_xp(true, true);
        sh[3].side[2] = face_ym(true, true);
        sh[3].side[3] = tris(false, false, top, 6);
        sh[3].side[4] = face_zm(true, true);
        sh[3].side[5] = tris(false, true, wall_zp, 3);
    }
}

V3 dir_vec(int d) {
    V3 v{0, 0, 0};
    int s = (d & 1) ? 1 : -1;
    if ((d >> 1) == 0) v.x = s;
    if ((d >> 1) == 1) v.y = s;
    if ((d >> 1) == 2) v.z = s;
    return v;
}

int vec_dir(V3 v) {
    if (v.x) return v.x > 0 ? 1 : 0;
    if (v.y) return v.y > 0 ? 3 : 2;
    return v.z > 0 ? 5 : 4;
}

} // namespace

void orientation_matrix(int idx, int m[9]) {
    // from_index (direction.rs:267-284)
    int i = (idx + 5) % 24;
    int right_idx = i / 4, up_idx = i % 4;
    if ((up_idx / 2) >= (right_idx / 2)) up_idx += 2;
    V3 r = dir_vec(right_idx), u = dir_vec(up_idx);
    // back = lh_cross(up, right) = -(up x right) (direction.rs:344-346, 429-448)
    V3 c = cross(u, r);
    V3 b = {-c.x, -c.y, -c.z};
    // to_matrixi: columns right, up, back (direction.rs:287-293)
    m[0] = r.x; m[1] = u.x; m[2] = b.x;
    m[3] = r.y; m[4] = u.y; m[5] = b.y;
    m[6] = r.z; m[7] = u.z; m[8] = b.z;
}

int orientation_unapply(int idx, int dir) {
    int m[9];
    orientation_matrix(idx, m);
    V3 v = dir_vec(dir);
    // M^T * v (direction.rs:203-204)
    V3 o = {m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
            m[2] * v.x + m[5] * v.y + m[8] * v.z};
    return vec_dir(o);
}

void build_mesh_tables(MeshTables &t) {
    std::memset(&t, 0, sizeof t);
    LShape sh[4];
    build_shapes(sh);

    // class 0: VOXEL_NO_SHAPE (stdshapes.rs:111-129): never emits, can be clipped, clips nothing -> all zero.
    for (int s = 0; s < 4; s++)
        for (int o = 0; o < 24; o++) {
            int cls = 1 + s * 24 + o;
            int m[9];
            orientation_matrix(o, m);
            auto rot = [&](V3 v) {
                return V3{m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z,
                          m[6] * v.x + m[7] * v.y + m[8] * v.z};
            };
            uint32_t cc = 0, ec = 0, eu = 0, codes = 0;
            for (int d = 0; d < 6; d++) {
                int ls = orientation_unapply(o, d);
                const LSide &sd = sh[s].side[ls];
                if (sd.can_clip) cc |= 1u << d;
                if (sd.ni > 0) {
                    if (sd.can_be_clipped) ec |= 1u << d; else eu |= 1u << d;
                }
                uint32_t code = sd.nv == 3 ? 1u : sd.nv == 4 ? 2u : sd.nv == 6 ? 3u : 0u;
                codes |= code << (2 * d);
                uint32_t signs = 0;
                for (int j = 0; j < sd.nv; j++) signs |= (uint32_t)(sd.v[j].sign + 1) << (2 * j);
                t.stab[cls][d] = (uint32_t)ls | (uint32_t)sd.nv << 3 | (uint32_t)sd.ni << 6 | signs << 9;
                for (int j = 0; j < sd.nv; j++) {
                    const LVert &lv = sd.v[j];
                    V3 w = rot(lv.off); // +-1 per axis
                    unsigned long long e = 0;
                    e |= (unsigned long long)(w.x > 0) | (unsigned long long)(w.y > 0) << 1 | (unsigned long long)(w.z > 0) << 2;
                    e |= (unsigned long long)lv.u << 3 | (unsigned long long)lv.v << 4;
                    e |= (unsigned long long)lv.bary << 5;
                    e |= (unsigned long long)(lv.sign + 1) << 8;
                    // corner_ao_set(corner, inormal) (stdshapes.rs:131-152), then rotated (voxmesh.rs:130-131)
                    V3 samples[5];
                    int n = 0;
                    V3 ic = {sgn(lv.off.x), sgn(lv.off.y), sgn(lv.off.z)};
                    samples[n++] = ic;
                    samples[n++] = lv.inormal;
                    if (lv.inormal.x == 0) samples[n++] = V3{0, ic.y, ic.z};
                    if (lv.inormal.y == 0) samples[n++] = V3{ic.x, 0, ic.z};
                    if (lv.inormal.z == 0) samples[n++] = V3{ic.x, ic.y, 0};
                    e |= (unsigned long long)n << 10;
                    {
                        V3 wn = rot(lv.inormal);
                        t.vdesc[cls][d][j] = (uint16_t)((e & 0xFF) | (unsigned long long)((wn.x + 1) + 3 * (wn.y + 1) + 9 * (wn.z + 1)) << 8);
                    }
                    t.aotab[cls][d][j][5] = (short)n;
                    for (int k = 0; k < n; k++) {
                        V3 a = rot(samples[k]);
                        t.aotab[cls][d][j][k] = (short)(a.y * kClsPlaneBytes + a.z * kClsRowBytes + a.x);
                        // NB: two samples of one vertex never coincide, so the popcount of the masked neighbourhood equals
                        // the number of occluding samples
                        t.aomask[cls][d][j] |= 1u << ((a.y + 1) * 9 + (a.z + 1) * 3 + (a.x + 1));
                        unsigned long long code6 = (unsigned long long)(a.x + 1) | (unsigned long long)(a.y + 1) << 2 |
                                                   (unsigned long long)(a.z + 1) << 4;
                        e |= code6 << (13 + 6 * k);
                    }
                    t.vtab[cls][d][j] = e;
                }
            }
            t.ctab[cls] = cc | ec << 6 | eu << 12 | codes << 18;
        }
    // aolut: corner_ao_set(corner, inormal) as a neighbourhood mask; the rule commutes with the 24 rotations
    for (int cb = 0; cb < 8; cb++)
        for (int nc = 0; nc < 27; nc++) {
            const int cv[3] = {(cb & 1) ? 1 : -1, (cb & 2) ? 1 : -1, (cb & 4) ? 1 : -1};
            const int nv[3] = {nc % 3 - 1, (nc / 3) % 3 - 1, nc / 9 - 1};
            auto bit = [](int x, int y, int z) { return 1u << ((y + 1) * 9 + (z + 1) * 3 + (x + 1)); };
            uint32_t m = bit(cv[0], cv[1], cv[2]) | bit(nv[0], nv[1], nv[2]);
            if (nv[0] == 0) m |= bit(0, cv[1], cv[2]);
            if (nv[1] == 0) m |= bit(cv[0], 0, cv[2]);
            if (nv[2] == 0) m |= bit(cv[0], cv[1], 0);
            t.aolut[cb * 27 + nc] = m;
        }
    t.aomask_ok = 1;
    for (int cls = 1; cls < kNumClasses; cls++)
        for (int d = 0; d < 6; d++)
            for (int j = 0; j < (int)((t.stab[cls][d] >> 3) & 7u); j++)
                if (t.aolut[(t.vdesc[cls][d][j] & 7) * 27 + (t.vdesc[cls][d][j] >> 8)] != t.aomask[cls][d][j]) t.aomask_ok = 0;
    for (int cls = 1; cls < kNumClasses; cls++)
        for (int d = 0; d < 6; d++)
            for (int j = 0; j < (int)((t.stab[cls][d] >> 3) & 7u); j++)
                if (__builtin_popcount(t.aomask[cls][d][j]) != t.aotab[cls][d][j][5]) t.aomask_ok = 0;
    float ao = 1.0f;
    for (int k = 0; k < 8; k++) {
        t.ao_lut[k] = ao;
        volatile float next = ao * 0.88f; // one IEEE f32 multiply per occluding sample, in sequence
        ao = next;
    }
}

} // namespace bxw
