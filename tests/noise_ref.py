"""Second, independently written transliteration of the generator chain (TEST INFRASTRUCTURE).

Pure Python, written from the published algorithms of the crates the reference's Cargo.lock pins and from
bxw_world/src/stdgen.rs itself; shares no code with oracle/noise.c or oracle/stdgen.c. It exists so that the two
restatements pin each other (tests/test_rng_kats.py); neither is the real `noise` crate (parity unpinned until
tools/golden_dump runs on a machine with cargo).

  rand_xoshiro 0.6.0   SplitMix64 (next_u64 / next_u32), Xoshiro256StarStar (seed_from_u64, next_u64 / next_u32)
  rand_xorshift 0.2.0  XorShiftRng (from_seed, next_u32)
  rand 0.7.3           SliceRandom::shuffle -> gen_index -> UniformInt<u32>::sample_single
  noise 0.8.2          PermutationTable::new / hash, gradient::grad2, super_simplex_2d
  bxw_world/src/stdgen.rs:80-276   CellGen
Python floats are IEEE binary64 and every operation below is a single correctly rounded operation, as in Rust.
"""
import math

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


class SplitMix64:
    """rand_xoshiro::SplitMix64. next_u64 is Vigna's splitmix64.c; next_u32 uses the `mix32` variant (murmur3-style
    64->32 finaliser constants) and keeps the high half."""

    def __init__(self, seed):
        self.x = seed & M64

    def next_u64(self):
        self.x = (self.x + 0x9E3779B97F4A7C15) & M64
        z = self.x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)

    def next_u32(self):
        self.x = (self.x + 0x9E3779B97F4A7C15) & M64
        z = self.x
        z = ((z ^ (z >> 33)) * 0x62A9D9ED799705F5) & M64
        z = ((z ^ (z >> 28)) * 0xCB24D0A5C88C35B3) & M64
        return z >> 32


def _rotl(x, k):
    return ((x << k) | (x >> (64 - k))) & M64


class Xoshiro256StarStar:
    def __init__(self, s):
        self.s = list(s)

    @classmethod
    def seed_from_u64(cls, seed):
        sm = SplitMix64(seed)  # from_splitmix!: four next_u64 words, little-endian round trip is the identity
        return cls([sm.next_u64() for _ in range(4)])

    def next_u64(self):
        s = self.s
        result = (_rotl((s[1] * 5) & M64, 7) * 9) & M64
        t = (s[1] << 17) & M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = _rotl(s[3], 45)
        return result

    def next_u32(self):
        return self.next_u64() >> 32  # the crate keeps the upper bits


class XorShiftRng:
    def __init__(self, seed_bytes):
        w = [int.from_bytes(bytes(seed_bytes[4 * i: 4 * i + 4]), "little") for i in range(4)]
        if not any(w):
            w = [0x0BAD5EED] * 4
        self.x, self.y, self.z, self.w = w

    def next_u32(self):
        x = self.x
        t = (x ^ (x << 11)) & M32
        self.x, self.y, self.z = self.y, self.z, self.w
        w = self.w
        self.w = (w ^ (w >> 19) ^ (t ^ (t >> 8))) & M32
        return self.w

    def next_u64(self):
        lo = self.next_u32()
        return (self.next_u32() << 32) | lo


def gen_range_u32(rng, low, high):
    """rand 0.7.3 UniformInt<u32>::sample_single: widening multiply, conservative rejection zone."""
    rng_range = (high - low) & M32
    lz = 32 - rng_range.bit_length()
    zone = (((rng_range << lz) & M32) - 1) & M32
    while True:
        v = rng.next_u32()
        m = v * rng_range
        if (m & M32) <= zone:
            return (low + (m >> 32)) & M32


def permutation_table(seed):
    """noise 0.8.2 PermutationTable::new(seed: u32): XorShiftRng seeded with [1,0,0,0, seed_le, seed_le, seed_le],
    values 0..=255 shuffled by rand 0.7.3 SliceRandom::shuffle."""
    sb = list(int(seed & M32).to_bytes(4, "little"))
    rng = XorShiftRng([1, 0, 0, 0] + sb * 3)
    seq = list(range(256))
    for i in range(len(seq) - 1, 0, -1):
        j = gen_range_u32(rng, 0, i + 1)
        seq[i], seq[j] = seq[j], seq[i]
    return seq


TO_REAL = -0.211324865405187
TO_SIMPLEX = 0.366025403784439
NORM_2D = 1.0 / 0.05428295288661623
DIAG = 0.7071067811865476  # core::f64::consts::FRAC_1_SQRT_2
GRAD2 = [(1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0), (DIAG, DIAG), (-DIAG, DIAG), (DIAG, -DIAG), (-DIAG, -DIAG)]


def _lattice_lookup():
    """LATTICE_LOOKUP_2D: 8 groups of 4 (lattice point, real-space offset). The crate lists decimal literals that are the
    shortest f64 representation of -p - (p.x + p.y) * TO_REAL_CONSTANT_2D evaluated in f64; derived here by that rule."""
    table = []
    for i in range(8):
        if i & 1 == 0:
            i1, j1 = (-1, 0) if i & 2 == 0 else (1, 0)
            i2, j2 = (0, -1) if i & 4 == 0 else (0, 1)
        else:
            i1, j1 = (2, 1) if i & 2 else (0, 1)
            i2, j2 = (1, 2) if i & 4 else (1, 0)
        for (px, py) in ((0, 0), (1, 1), (i1, j1), (i2, j2)):
            s = (px + py) * TO_REAL
            table.append(((px, py), (0.0 if px == 0 and py == 0 else -px - s, 0.0 if px == 0 and py == 0 else -py - s)))
    return table


LATTICE = _lattice_lookup()


def super_simplex_2d(perm, x, y):
    """noise 0.8.2 core::super_simplex::super_simplex_2d with PermutationTable::hash."""
    off = (x + y) * TO_SIMPLEX
    sx, sy = x + off, y + off
    bx, by = math.floor(sx), math.floor(sy)
    rx, ry = sx - bx, sy - by
    region_sum = float(math.floor(rx + ry))
    index = (int(region_sum >= 1.0) << 2) | (int(rx - ry * 0.5 + 1.0 - region_sum * 0.5 >= 1.0) << 3) | (
        int(ry - rx * 0.5 + 1.0 - region_sum * 0.5 >= 1.0) << 4)
    roff = (rx + ry) * TO_REAL
    qx, qy = rx + roff, ry + roff
    value = 0.0
    for (lp, lo) in LATTICE[index:index + 4]:
        dx, dy = qx + lo[0], qy + lo[1]
        attn = (2.0 / 3.0) - (dx * dx + dy * dy)
        if attn > 0.0:
            lx, ly = int(bx) + lp[0], int(by) + lp[1]
            h = perm[perm[lx & 0xFF] ^ (ly & 0xFF)]
            g = GRAD2[h % 8]
            a2 = attn * attn  # powi(4)
            value += (a2 * a2) * (g[0] * dx + g[1] * dy)
    return value * NORM_2D


def _trunc_div(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _round_half_away(v):
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


class CellGen:
    """bxw_world/src/stdgen.rs:80-276."""

    def __init__(self, seed):
        sd = SplitMix64(seed)  # set_seed, :103-114
        self.seed = seed & M64
        self.height = [permutation_table(sd.next_u32()) for _ in range(5)]
        self.elevation = permutation_table(sd.next_u32())
        self.moisture = permutation_table(sd.next_u32())
        self.cells = {}

    def elevation_noise(self, x, z):  # :171-175
        sf = 256.0 * 10.0
        return (super_simplex_2d(self.elevation, x / sf, z / sf) + 1.0) / 2.0

    def h_n(self, i, px, pz):
        return (super_simplex_2d(self.height[i], px, pz) + 1.0) / 2.0

    def h_rn(self, i, px, pz):
        return 2.0 * (0.5 - abs(0.5 - self.h_n(i, px, pz)))

    def cell_points(self, cx, cz):  # :121-147
        key = (cx, cz)
        if key in self.cells:
            return self.cells[key]
        s = self.seed ^ ((((cx & M64) << 32) & M64) | (cz & M32))
        r = Xoshiro256StarStar.seed_from_u64(s)
        pts = []
        for (bx, bz) in ((32, 32), (96, 32), (0, 96), (64, 96)):
            xo = (r.next_u32() % 32) - 16
            zo = (r.next_u32() % 32) - 16
            px, pz = cx * 128 + bx + xo, cz * 128 + bz + zo
            pts.append((px, pz, self.elevation_noise(px, pz)))
        self.cells[key] = pts
        return pts

    def params(self, x, z):  # :215-276 -> (height f64, round(height), elevation 0/1/2, ec, cec)
        cx, cz = _trunc_div(x, 128), _trunc_div(z, 128)
        best = None
        for dx in (-1, 0, 1):
            for dz in (-1, 0, 1):
                for (px, pz, e) in self.cell_points(cx + dx, cz + dz):
                    d = (px - x) ** 2 + (pz - z) ** 2
                    if best is None or d < best[0]:  # Iterator::min_by keeps the first minimum
                        best = (d, e)
        cec = best[1]
        ec = self.elevation_noise(x, z)

        def plains():
            px, pz = x / 1200.0, z / 1200.0
            return (0.75 * self.h_n(0, px, pz) + 0.25 * self.h_n(1, px * 2.0, pz * 2.0)) * 5.0 + 10.0

        def hills():
            px, pz = x / 1600.0, z / 1600.0
            return (0.60 * self.h_n(0, px, pz) + 0.25 * self.h_n(1, px * 1.5, pz * 1.5) + 0.15 * self.h_n(2, px * 3.0, pz * 3.0)) * 30.0 + 15.0

        def mountains():
            px, pz = x / 2000.0, z / 2000.0
            h0 = 0.50 * self.h_rn(0, px, pz)
            h01 = 0.25 * self.h_rn(1, px * 2.0, pz * 2.0) + h0
            return (h01 + (h01 / 0.75) * 0.15 * self.h_n(2, px * 5.0, pz * 5.0) + (h01 / 0.75) * 0.05 * self.h_rn(3, px * 9.0, pz * 9.0)) * 750.0 + 40.0

        if ec < 0.4:
            h = plains()
        elif ec < 0.5:
            nlin = (ec - 0.4) / 0.1
            olin = 1.0 - nlin
            h = olin * plains() + nlin * hills()
        elif ec < 0.7:
            h = hills()
        elif ec < 0.8:
            nlin = (ec - 0.7) / 0.1
            olin = 1.0 - nlin
            h = olin * hills() + nlin * mountains()
        else:
            h = mountains()
        if cec < 0.4:
            el = 0
        elif cec < 0.5:
            el = 0 if super_simplex_2d(self.height[4], float(x), float(z)) < 0.0 else 1
        elif cec < 0.7:
            el = 1
        elif cec < 0.8:
            el = 1 if super_simplex_2d(self.height[4], float(x), float(z)) < 0.0 else 2
        else:
            el = 2
        return h, _round_half_away(h), el, ec, cec

    def column(self, x, z, y0, n):
        """StdGenerator::generate_chunk block selection (stdgen.rs:349-373) for voxels y0..y0+n of one column, as
        block NAMES' indices: 0 void, 1 grass, 2 dirt, 3 stone, 4 snow_grass."""
        _, h, el, _, _ = self.params(x, z)
        out = []
        for y in range(y0, y0 + n):
            if y == h:
                out.append(4 if (el == 2 and y > 80) else 1)
            elif y < h - 5:
                out.append(3)
            elif y < h:
                out.append(2)
            else:
                out.append(0)
        return out
