/*
 * oracle/noise.c — TEST INFRASTRUCTURE (see bxw_oracle.h).
 *
 * Restatement of the third-party arithmetic the reference generator calls (bxw_world/src/stdgen.rs:6-8, 93-96,
 * 104-112, 126, 137-138, 172, 179, 249, 259). NONE of these crates is vendored under /root/reference; the
 * versions are those pinned by the reference's Cargo.lock:
 *     noise 0.8.2 (Cargo.lock:1309-1317)  -> SuperSimplex, PermutationTable, gradient::grad2
 *     rand 0.7.3 (Cargo.lock:1638-1641)   -> SliceRandom::shuffle, UniformInt<u32>::sample_single
 *     rand_xorshift 0.2.0                 -> XorShiftRng
 *     rand_xoshiro 0.6.0 (Cargo.lock:1729-1733) -> SplitMix64, Xoshiro256StarStar
 * PARITY UNPINNED: the published algorithms are restated from the crates' documented behaviour; the reference
 * holds no test or golden vector at this boundary and the Rust toolchain is absent from this image.
 *
 * Also: the in-tree bxw_terragen SuperSimplex (bxw_terragen/src/supersimplex.rs), which IS fully specified by
 * reference source.
 */
#include "bxw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- rand_xoshiro 0.6.0 ---------------- */

#define PHI 0x9e3779b97f4a7c15ULL

uint32_t bxo_splitmix64_next_u32(bxo_splitmix64 *s) {
    /* SplitMix64::next_u32 uses Stafford's Mix4 finaliser and returns the high half */
    s->x += PHI;
    uint64_t z = s->x;
    z = (z ^ (z >> 33)) * 0x62A9D9ED799705F5ULL;
    z = (z ^ (z >> 28)) * 0xCB24D0A5C88C35B3ULL;
    return (uint32_t)(z >> 32);
}

uint64_t bxo_splitmix64_next_u64(bxo_splitmix64 *s) {
    s->x += PHI;
    uint64_t z = s->x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

void bxo_xoshiro256ss_seed_from_u64(bxo_xoshiro256ss *r, uint64_t seed) {
    /* from_splitmix!: the 32 seed bytes are filled from SplitMix64::next_u64, read back little-endian */
    bxo_splitmix64 sm = {seed};
    for (int i = 0; i < 4; i++) r->s[i] = bxo_splitmix64_next_u64(&sm);
}

static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

uint64_t bxo_xoshiro256ss_next_u64(bxo_xoshiro256ss *r) {
    uint64_t result = rotl64(r->s[1] * 5, 7) * 9;
    uint64_t t = r->s[1] << 17;
    r->s[2] ^= r->s[0];
    r->s[3] ^= r->s[1];
    r->s[1] ^= r->s[2];
    r->s[0] ^= r->s[3];
    r->s[2] ^= t;
    r->s[3] = rotl64(r->s[3], 45);
    return result;
}

uint32_t bxo_xoshiro256ss_next_u32(bxo_xoshiro256ss *r) {
    /* "the lowest bits have some linear dependencies, so we use the upper bits instead" */
    return (uint32_t)(bxo_xoshiro256ss_next_u64(r) >> 32);
}

/* ---------------- rand_xorshift 0.2.0 + rand 0.7.3 shuffle ---------------- */

typedef struct { uint32_t x, y, z, w; } xorshift_rng;

static uint32_t xorshift_next(xorshift_rng *r) {
    uint32_t x = r->x;
    uint32_t t = x ^ (x << 11);
    r->x = r->y;
    r->y = r->z;
    r->z = r->w;
    uint32_t w_ = r->w;
    r->w = w_ ^ (w_ >> 19) ^ (t ^ (t >> 8));
    return r->w;
}

static uint32_t gen_range_u32(xorshift_rng *r, uint32_t low, uint32_t high) {
    /* rand 0.7.3 UniformInt<u32>::sample_single (widening-multiply rejection, conservative zone) */
    uint32_t range = high - low;
    int lz = __builtin_clz(range);
    uint32_t zone = (range << lz) - 1u;
    for (;;) {
        uint32_t v = xorshift_next(r);
        uint64_t m = (uint64_t)v * (uint64_t)range;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo <= zone) return low + hi;
    }
}

void bxo_permutation_table(uint32_t seed, uint8_t out[256]) {
    /* noise 0.8.2 PermutationTable::new: 16-byte XorShift seed = [1,0,0,0, seed LE x3]; table = 0..=255
     * shuffled with rand 0.7.3 SliceRandom::shuffle (for i in (1..len).rev(): swap(i, gen_range(0, i+1))). */
    xorshift_rng r;
    r.x = 1u;
    r.y = seed;
    r.z = seed;
    r.w = seed;
    if (r.x == 0 && r.y == 0 && r.z == 0 && r.w == 0) r.x = r.y = r.z = r.w = 0x0BAD5EEDu; /* unreachable: x = 1 */
    for (int i = 0; i < 256; i++) out[i] = (uint8_t)i;
    for (uint32_t i = 255; i >= 1; i--) {
        uint32_t j = gen_range_u32(&r, 0, i + 1);
        uint8_t tmp = out[i];
        out[i] = out[j];
        out[j] = tmp;
    }
}

/* ---------------- noise 0.8.2 super_simplex_2d ---------------- */

static const double TO_REAL_CONSTANT_2D = -0.211324865405187;   /* (1/sqrt(2+1) - 1)/2 */
static const double TO_SIMPLEX_CONSTANT_2D = 0.366025403784439; /* (sqrt(2+1) - 1)/2 */
static const double NORM_CONSTANT_2D = 1.0 / 0.05428295288661623;

typedef struct { int lx, ly; double ox, oy; } lattice_entry;

/* LATTICE_LOOKUP_2D: for each of the 8 index groups (bit0: region_sum>=1, bit1: x-cond, bit2: y-cond) the four
 * contributing lattice points and the real-space offsets of the evaluation point from them, as the crate's
 * 15-digit decimal literals. The point structure equals the in-tree generator at
 * bxw_terragen/src/supersimplex.rs:187-234. */
#define A 0.788675134594813
#define B 0.211324865405187
#define C 0.577350269189626
#define D 1.366025403784439
#define E 0.36602540378443904
static const lattice_entry LATTICE_LOOKUP_2D[32] = {
    {0, 0, 0.0, 0.0}, {1, 1, -C, -C}, {-1, 0, A, -B}, {0, -1, -B, A},
    {0, 0, 0.0, 0.0}, {1, 1, -C, -C}, {0, 1, B, -A},  {1, 0, -A, B},
    {0, 0, 0.0, 0.0}, {1, 1, -C, -C}, {1, 0, -A, B},  {0, -1, -B, A},
    {0, 0, 0.0, 0.0}, {1, 1, -C, -C}, {2, 1, -D, -E}, {1, 0, -A, B},
    {0, 0, 0.0, 0.0}, {1, 1, -C, -C}, {-1, 0, A, -B}, {0, 1, B, -A},
    {0, 0, 0.0, 0.0}, {1, 1, -C, -C}, {0, 1, B, -A},  {1, 2, -E, -D},
    {0, 0, 0.0, 0.0}, {1, 1, -C, -C}, {1, 0, -A, B},  {0, 1, B, -A},
    {0, 0, 0.0, 0.0}, {1, 1, -C, -C}, {2, 1, -D, -E}, {1, 2, -E, -D},
};
#undef A
#undef B
#undef C
#undef D
#undef E

static void grad2(unsigned index, double g[2]) {
    /* noise 0.8.2 gradient::grad2 */
    const double DIAG = 0.70710678118654752440; /* core::f64::consts::FRAC_1_SQRT_2 */
    switch (index % 8) {
    case 0: g[0] = 1.0; g[1] = 0.0; break;
    case 1: g[0] = -1.0; g[1] = 0.0; break;
    case 2: g[0] = 0.0; g[1] = 1.0; break;
    case 3: g[0] = 0.0; g[1] = -1.0; break;
    case 4: g[0] = DIAG; g[1] = DIAG; break;
    case 5: g[0] = -DIAG; g[1] = DIAG; break;
    case 6: g[0] = DIAG; g[1] = -DIAG; break;
    default: g[0] = -DIAG; g[1] = -DIAG; break;
    }
}

double bxo_super_simplex_2d(const uint8_t perm[256], double px, double py) {
    /* Transform point from real space to simplex space */
    double to_simplex_offset = (px + py) * TO_SIMPLEX_CONSTANT_2D;
    double sx = px + to_simplex_offset, sy = py + to_simplex_offset;
    /* Base point of the simplex and barycentric coordinates in simplex space */
    double bxf = floor(sx), byf = floor(sy);
    long bxi = (long)bxf, byi = (long)byf;
    double rx = sx - bxf, ry = sy - byf;
    /* Index into the lookup table */
    double region_sum = floor(rx + ry);
    unsigned index = ((region_sum >= 1.0) ? 1u : 0u) << 2;
    index |= ((rx - ry * 0.5 + 1.0 - region_sum * 0.5 >= 1.0) ? 1u : 0u) << 3;
    index |= ((ry - rx * 0.5 + 1.0 - region_sum * 0.5 >= 1.0) ? 1u : 0u) << 4;
    /* Barycentric coordinates back to real space */
    double to_real_offset = (rx + ry) * TO_REAL_CONSTANT_2D;
    double qx = rx + to_real_offset, qy = ry + to_real_offset;

    double value = 0.0;
    for (unsigned k = index; k < index + 4; k++) {
        const lattice_entry *le = &LATTICE_LOOKUP_2D[k];
        double dx = qx + le->ox, dy = qy + le->oy;
        double attn = (2.0 / 3.0) - (dx * dx + dy * dy);
        if (attn > 0.0) {
            long lx = bxi + le->lx, ly = byi + le->ly;
            /* PermutationTable::hash([x, y]) = values[values[x & 0xff] ^ (y & 0xff)] */
            unsigned h = perm[perm[(unsigned)(lx & 0xff)] ^ (unsigned)(ly & 0xff)];
            double g[2];
            grad2(h, g);
            double a2 = attn * attn; /* powi(4) */
            value += (a2 * a2) * (g[0] * dx + g[1] * dy);
        }
    }
    return value * NORM_CONSTANT_2D;
}

/* ---------------- bxw_terragen in-tree SuperSimplex ---------------- */

#define PSIZE 2048
#define PMASK 2047

struct bxo_tg_simplex {
    uint16_t perm[PSIZE];
    double gx[PSIZE], gy[PSIZE];
};

static void tg_lattice(int idx, int *px, int *py, double *dx, double *dy) {
    /* lookup_2d::init_lut, supersimplex.rs:187-234 */
    int i = idx >> 2, k = idx & 3;
    int i1, j1, i2, j2;
    if ((i & 1) == 0) {
        i1 = (i & 2) == 0 ? -1 : 1;
        j1 = 0;
        i2 = 0;
        j2 = (i & 4) == 0 ? -1 : 1;
    } else {
        i1 = (i & 2) != 0 ? 2 : 0;
        j1 = 1;
        i2 = 1;
        j2 = (i & 4) != 0 ? 2 : 0;
    }
    int x = k == 0 ? 0 : k == 1 ? 1 : k == 2 ? i1 : i2;
    int y = k == 0 ? 0 : k == 1 ? 1 : k == 2 ? j1 : j2;
    /* lattice_point_derivatives, supersimplex.rs:125-128 */
    double ssv = (double)(x + y) * -0.211324865405187;
    *px = x;
    *py = y;
    *dx = (double)(-x) - ssv;
    *dy = (double)(-y) - ssv;
}

static const double TG_GRAD2[24][2] = {
    /* supersimplex.rs:241-266 */
    {0.130526192220052, 0.99144486137381},    {0.38268343236509, 0.923879532511287},
    {0.608761429008721, 0.793353340291235},   {0.793353340291235, 0.608761429008721},
    {0.923879532511287, 0.38268343236509},    {0.99144486137381, 0.130526192220051},
    {0.99144486137381, -0.130526192220051},   {0.923879532511287, -0.38268343236509},
    {0.793353340291235, -0.60876142900872},   {0.608761429008721, -0.793353340291235},
    {0.38268343236509, -0.923879532511287},   {0.130526192220052, -0.99144486137381},
    {-0.130526192220052, -0.99144486137381},  {-0.38268343236509, -0.923879532511287},
    {-0.608761429008721, -0.793353340291235}, {-0.793353340291235, -0.608761429008721},
    {-0.923879532511287, -0.38268343236509},  {-0.99144486137381, -0.130526192220052},
    {-0.99144486137381, 0.130526192220051},   {-0.923879532511287, 0.38268343236509},
    {-0.793353340291235, 0.608761429008721},  {-0.608761429008721, 0.793353340291235},
    {-0.38268343236509, 0.923879532511287},   {-0.130526192220052, 0.99144486137381},
};

bxo_tg_simplex *bxo_tg_simplex_new(uint64_t seed) {
    /* supersimplex.rs:20-43 */
    bxo_tg_simplex *s = (bxo_tg_simplex *)malloc(sizeof *s);
    if (!s) return 0;
    uint16_t source[PSIZE];
    for (int i = 0; i < PSIZE; i++) source[i] = (uint16_t)i;
    const double N2 = 0.05481866495625118;
    for (int i = PSIZE - 1; i >= 0; i--) {
        seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
        uint64_t r = (seed + 31ULL) % ((uint64_t)i + 1ULL);
        s->perm[i] = source[r];
        int g = s->perm[i] % 24; /* grads tiled over PSIZE: grad2[i % 24], supersimplex.rs:268-272 */
        s->gx[i] = TG_GRAD2[g][0] / N2;
        s->gy[i] = TG_GRAD2[g][1] / N2;
        source[r] = source[i];
    }
    return s;
}

void bxo_tg_simplex_free(bxo_tg_simplex *s) { free(s); }
const uint16_t *bxo_tg_simplex_perm(const bxo_tg_simplex *s) { return s->perm; }

static int32_t tg_floor_i32(double x) {
    /* math.rs:68-78 widefloor_f64_i32: truncating cast, minus one where x < trunc(x) */
    int32_t xi = (int32_t)x;
    if (x < (double)xi) xi -= 1;
    return xi;
}

double bxo_tg_simplex_noise2(const bxo_tg_simplex *sp, double x, double y) {
    /* noise2_wide + noise2_base_wide, supersimplex.rs:53-107 (one lane) */
    double s = (x + y) * 0.366025403784439;
    double xs = x + s, ys = y + s;
    int32_t xsb = tg_floor_i32(xs), ysb = tg_floor_i32(ys);
    double xsi = xs - (double)xsb, ysi = ys - (double)ysb;
    int32_t a = (int32_t)(xsi + ysi);
    double af = (double)a;
    int32_t index = (a << 2) | ((int32_t)(xsi - ysi / 2.0 + 1.0 - af / 2.0) << 3) |
                    ((int32_t)(ysi - xsi / 2.0 + 1.0 - af / 2.0) << 4);
    double ssi = (xsi + ysi) * -0.211324865405187;
    double xi = xsi + ssi, yi = ysi + ssi;
    double value = 0.0;
    for (int i = 0; i < 4; i++) {
        int cx, cy;
        double cdx, cdy;
        tg_lattice(index + i, &cx, &cy, &cdx, &cdy);
        double dx = xi + cdx, dy = yi + cdy;
        double attn = 2.0 / 3.0 - dx * dx - dy * dy;
        if (!(attn > 0.0)) continue;
        unsigned pxm = (unsigned)((xsb + cx) & PMASK), pym = (unsigned)((ysb + cy) & PMASK);
        unsigned gi = (unsigned)sp->perm[pxm] ^ pym;
        double extrapolation = sp->gx[gi] * dx + sp->gy[gi] * dy;
        double attn2 = attn * attn;
        double attn4 = attn2 * attn2;
        value = value + attn4 * extrapolation;
    }
    return value;
}

uint64_t bxo_tg_splitmix64(uint64_t seed, uint64_t *newseed) {
    /* bxw_terragen/src/math.rs:24-30 */
    uint64_t ns = seed + 0x9e3779b97f4a7c15ULL;
    uint64_t z = ns;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    if (newseed) *newseed = ns;
    return z ^ (z >> 31);
}
