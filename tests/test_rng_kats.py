"""Pins of the generator's third-party arithmetic that need no cargo.

1. Published known-answer vectors of the RNG crates the reference's Cargo.lock pins (the vectors are the ones in the
   crates' own test suites, which in turn come from Vigna's reference C programs): rand_xoshiro 0.6.0 SplitMix64
   (`reference`, `next_u32`) and Xoshiro256StarStar (`reference`); rand_xorshift 0.2.0 (`test_xorshift_true_values`).
   Checked against BOTH the oracle (oracle/noise.c) and the independent Python transliteration (tests/noise_ref.py).
2. The PermutationTable / shuffle / SuperSimplex / CellGen chain: oracle vs the independent transliteration, bit for bit.
   This pins the two restatements to each other, not to the real `noise 0.8.2` crate -- the generator stays
   "parity unpinned" until tools/golden_dump has run on a machine with cargo.
"""
import ctypes as C
import struct

import numpy as np

import noise_ref as R

SPLITMIX_SEED = 1477776061723855037
SPLITMIX_U64 = [
    1985237415132408290, 2979275885539914483, 13511426838097143398, 8488337342461049707, 15141737807933549159,
    17093170987380407015, 16389528042912955399, 13177319091862933652, 10841969400225389492, 17094824097954834098,
    3336622647361835228, 9678412372263018368, 11111587619974030187, 7882215801036322410, 5709234165213761869,
    7799681907651786826, 4616320717312661886,
]
SPLITMIX_U32_SEED = 10
SPLITMIX_U32 = [3930361779, 4016923089, 4113052479, 925926767, 1755287528, 802865554, 954171070, 3724185978, 173676273, 1414488795]
XOSHIRO_STATE = [1, 2, 3, 4]
XOSHIRO_U64 = [
    11520, 0, 1509978240, 1215971899390074240, 1216172134540287360, 607988272756665600, 16172922978634559625,
    8476171486693032832, 10595114339597558777, 2904607092377533576,
]
XORSHIFT_SEED = [16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1]
XORSHIFT_U32 = [2081028795, 620940381, 269070770, 16943764, 854422573, 29242889, 1550291885, 1227154591, 271695242]
XORSHIFT_U64 = [
    9247529084182843387, 8321512596129439293, 14104136531997710878, 6848554330849612046, 343577296533772213,
    17828467390962600268, 9847333257685787782, 7717352744383350108, 1133407547287910111,
]


def test_splitmix64_published_vectors(oracle):
    L = oracle.lib()
    s = C.c_uint64(SPLITMIX_SEED)
    assert [L.bxo_splitmix64_next_u64(C.byref(s)) for _ in SPLITMIX_U64] == SPLITMIX_U64
    s = C.c_uint64(SPLITMIX_U32_SEED)
    assert [L.bxo_splitmix64_next_u32(C.byref(s)) for _ in SPLITMIX_U32] == SPLITMIX_U32
    r = R.SplitMix64(SPLITMIX_SEED)
    assert [r.next_u64() for _ in SPLITMIX_U64] == SPLITMIX_U64
    r = R.SplitMix64(SPLITMIX_U32_SEED)
    assert [r.next_u32() for _ in SPLITMIX_U32] == SPLITMIX_U32


def test_xoshiro256starstar_published_vectors(oracle):
    L = oracle.lib()
    L.bxo_xoshiro256ss_next_u64.argtypes = [C.POINTER(C.c_uint64 * 4)]
    L.bxo_xoshiro256ss_next_u64.restype = C.c_uint64
    st = (C.c_uint64 * 4)(*XOSHIRO_STATE)
    assert [L.bxo_xoshiro256ss_next_u64(C.byref(st)) for _ in XOSHIRO_U64] == XOSHIRO_U64
    r = R.Xoshiro256StarStar(XOSHIRO_STATE)
    assert [r.next_u64() for _ in XOSHIRO_U64] == XOSHIRO_U64
    # seed_from_u64 = four SplitMix64::next_u64 words; next_u32 = upper half (what get_cell_points consumes)
    st = (C.c_uint64 * 4)()
    L.bxo_xoshiro256ss_seed_from_u64(C.byref(st), SPLITMIX_SEED)
    assert list(st) == SPLITMIX_U64[:4]
    a = R.Xoshiro256StarStar.seed_from_u64(SPLITMIX_SEED)
    assert a.s == SPLITMIX_U64[:4]
    assert [L.bxo_xoshiro256ss_next_u32(C.byref(st)) for _ in range(8)] == [a.next_u32() for _ in range(8)]


def test_xorshift_published_vectors():
    """rand_xorshift's own vector; the oracle's XorShift is file-local, so the permutation tables below carry it."""
    r = R.XorShiftRng(XORSHIFT_SEED)
    assert [r.next_u32() for _ in XORSHIFT_U32] == XORSHIFT_U32
    assert [r.next_u64() for _ in XORSHIFT_U64] == XORSHIFT_U64


def test_permutation_tables_agree(oracle):
    for seed in [0, 1, 2, 12345, 0xFFFFFFFF, 0x80000000, 3930361779]:
        a = list(oracle.permutation_table(seed))
        b = R.permutation_table(seed)
        assert sorted(b) == list(range(256))
        assert a == b, f"PermutationTable::new({seed}) differs"
    # CellGen::set_seed (stdgen.rs:103-114): seven tables from SplitMix64::next_u32
    for seed in [0, 0x13372137CAFE]:
        g = oracle.StdGen(seed)
        cg = R.CellGen(seed)
        tabs = cg.height + [cg.elevation, cg.moisture]
        for k in range(7):
            assert list(g.perm(k)) == tabs[k]


def test_lattice_table_matches_crate_literals():
    """The derived offsets are exactly the crate's decimal literals (LATTICE_LOOKUP_2D)."""
    lits = {0.0, -0.577350269189626, 0.788675134594813, -0.211324865405187, 0.211324865405187, -0.788675134594813,
            -1.366025403784439, -0.36602540378443904}
    seen = set()
    for (_, (ox, oy)) in R.LATTICE:
        seen.add(ox)
        seen.add(oy)
    assert seen == lits
    pts = [p for (p, _) in R.LATTICE]
    assert pts[2:4] == [(-1, 0), (0, -1)] and pts[14:16] == [(2, 1), (1, 0)] and pts[30:32] == [(2, 1), (1, 2)]


def _bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def test_super_simplex_bitwise(oracle):
    rng = np.random.default_rng(7)
    perm = R.permutation_table(987654321)
    pa = np.array(perm, dtype=np.uint8)
    pts = np.concatenate([rng.uniform(-300, 300, size=(3000, 2)), rng.uniform(-3, 3, size=(1000, 2)),
                          rng.integers(-50, 50, size=(500, 2)).astype(np.float64),
                          np.array([[0.0, 0.0], [-0.0, 0.0], [1e-300, -1e-300], [12345.678, -98765.4321]])])
    for x, y in pts:
        a = oracle.super_simplex(pa, float(x), float(y))
        b = R.super_simplex_2d(perm, float(x), float(y))
        assert _bits(a) == _bits(b), (x, y, a, b)


def test_calc_voxel_params_bitwise(oracle):
    rng = np.random.default_rng(11)
    for seed in [0, 0x13372137CAFE]:
        g = oracle.StdGen(seed)
        cg = R.CellGen(seed)
        cols = [(300, 28), (0, 0), (-1, -1), (-128, -129), (127, 128), (16383, 16383), (-5000, 7000)]
        cols += [tuple(int(v) for v in rng.integers(-20000, 20000, size=2)) for _ in range(150)]
        cols += [(int(x), 77) for x in range(-140, -116)]  # truncating pos / 128 around a negative cell edge (stdgen.rs:150)
        classes = set()
        for (x, z) in cols:
            p = g.params(x, z)
            h, hr, el, ec, cec = cg.params(x, z)
            assert _bits(p.height_f64) == _bits(h) and p.height == hr and p.elevation == el, (seed, x, z)
            assert _bits(p.ec) == _bits(ec) and _bits(p.cec) == _bits(cec)
            classes.add(el)
        assert len(classes) >= 2


def test_generate_chunk_columns(oracle):
    """Block selection (stdgen.rs:349-373) through both restatements, on columns of three chunks."""
    g = oracle.StdGen(0)
    cg = R.CellGen(0)
    ids = oracle.STANDARD_GEN_IDS  # void, grass, dirt, stone, snow_grass
    for (cx, cy, cz) in [(9, 0, 0), (-3, 1, 5), (40, 2, -17)]:
        c = g.generate_chunk(cx, cy, cz).reshape(32, 32, 32)  # [y][z][x]
        for (lx, lz) in [(0, 0), (31, 31), (12, 28), (5, 17)]:
            want = [ids[k] << 16 for k in cg.column(cx * 32 + lx, cz * 32 + lz, cy * 32, 32)]
            assert list(c[:, lz, lx]) == want
