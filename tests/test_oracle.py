"""CPU tests: the oracle against the reference's own known answers and properties."""
import ctypes as C

import numpy as np


def test_rle_compress_kats(oracle):
    # bxw_world/src/lib.rs:486-498
    assert list(oracle.compress_rle(np.zeros(32768, np.uint32))) == [0, 0, 32766]
    assert list(oracle.compress_rle(np.ones(32768, np.uint32))) == [1, 1, 32766]


def test_rle_decompress_kats(oracle):
    # lib.rs:500-522
    for v in (0, 1):
        comp = np.array([v, v, 32766], np.uint32)
        rc, dense = oracle.decompress_rle(comp)
        assert rc == 0 and (dense == v).all()
        assert (oracle.rle_iterate(comp) == v).all() and oracle.rle_iterate(comp).size == 32768


def test_rle_random_cmp(oracle):
    # lib.rs:524-546: Xoshiro256StarStar::seed_from_u64(1234).next_u32() % 16
    L = oracle.lib()
    st = (C.c_uint64 * 4)()
    L.bxo_xoshiro256ss_seed_from_u64(C.byref(st), 1234)
    data = np.array([L.bxo_xoshiro256ss_next_u32(C.byref(st)) % 16 for _ in range(32768)], np.uint32)
    comp = oracle.compress_rle(data)
    rc, dec = oracle.decompress_rle(comp)
    assert rc == 0 and np.array_equal(dec, data)
    assert np.array_equal(oracle.rle_iterate(comp), data)


def test_rle_errors(oracle):
    assert oracle.decompress_rle(np.array([0, 0], np.uint32))[0] == 4          # FinalPosMismatch(len)
    assert oracle.decompress_rle(np.array([0, 0, 32768], np.uint32))[0] == 2   # placeholder does not decode
    assert oracle.decompress_rle(np.array([0, 0, 5, 1, 1], np.uint32))[0] == 3  # MissingRleRepeatN
    assert oracle.decompress_rle(np.array([0, 0, 10, 7], np.uint32))[0] == 4   # short


def test_orientation_properties(oracle):
    # bxw_util/src/direction.rs:354-448
    L = oracle.lib()
    allowed, used = 0, set()
    for d1 in range(6):
        for d2 in range(6):
            for d3 in range(6):
                if L.bxo_orient_from_dirs(d1, d2, d3):
                    allowed += 1
                    idx = L.bxo_orient_to_index(d1, d2)
                    r, u = C.c_int(), C.c_int()
                    assert L.bxo_orient_from_index(idx, C.byref(r), C.byref(u)) == 0
                    assert (r.value, u.value) == (d1, d2)
                    assert idx not in used
                    used.add(idx)
                    assert L.bxo_orient_apply(idx, 1) == d1 and L.bxo_orient_apply(idx, 3) == d2
                    assert L.bxo_orient_apply(idx, 4) == d3  # DIR_FRONT = ZMinus
                    for d in range(6):
                        assert L.bxo_orient_unapply(idx, L.bxo_orient_apply(idx, d)) == d
    assert allowed == 24 and used == set(range(24))
    m = (C.c_int * 9)()
    L.bxo_orient_matrix(0, m)
    assert list(m) == [1, 0, 0, 0, 1, 0, 0, 0, 1]
    # direction_cross_verify: lh_cross(RIGHT, UP) = FRONT etc.
    assert L.bxo_lh_cross(1, 3) == 4 and L.bxo_lh_cross(3, 4) == 1 and L.bxo_lh_cross(4, 1) == 3
    assert sum(1 for a in range(6) for b in range(6) if L.bxo_lh_cross(a, b) >= 0) == 24


def test_unapply_table_matches_survey(oracle):
    # SURVEY.md App. C.1 rows
    L = oracle.lib()
    rows = {0: [0, 1, 2, 3, 4, 5], 1: [0, 1, 4, 5, 3, 2], 6: [5, 4, 1, 0, 2, 3], 12: [2, 3, 5, 4, 1, 0],
            17: [4, 5, 3, 2, 0, 1], 23: [0, 1, 3, 2, 5, 4]}
    for o, row in rows.items():
        assert [L.bxo_orient_unapply(o, d) for d in range(6)] == row


def test_shape_side_summary(oracle):
    # (can_clip, can_be_clipped, n_verts, n_idx) per shape and side
    from oracle_lib import lib
    import ctypes

    class V(ctypes.Structure):
        _fields_ = [("offset", ctypes.c_float * 3), ("tc", ctypes.c_float * 2), ("b", ctypes.c_float * 3),
                    ("sign", ctypes.c_int32), ("n_ao", ctypes.c_int32), ("ao", (ctypes.c_int32 * 3) * 5)]

    class S(ctypes.Structure):
        _fields_ = [("cc", ctypes.c_int32), ("cbc", ctypes.c_int32), ("nv", ctypes.c_int32), ("v", V * 6),
                    ("ni", ctypes.c_int32), ("idx", ctypes.c_uint32 * 6)]

    class Sh(ctypes.Structure):
        _fields_ = [("ao", ctypes.c_int32), ("sides", S * 6)]

    L = lib()
    L.bxo_shape_table.restype = ctypes.POINTER(Sh)
    want = {
        0: [(0, 1, 0, 0)] * 6,
        1: [(1, 1, 4, 6)] * 6,
        2: [(0, 1, 3, 3), (0, 1, 3, 3), (1, 1, 4, 6), (0, 0, 4, 6), (0, 1, 0, 0), (1, 1, 4, 6)],
        3: [(0, 1, 3, 3), (0, 1, 0, 0), (1, 1, 4, 6), (0, 0, 6, 6), (0, 1, 0, 0), (0, 1, 3, 3)],
        4: [(0, 1, 3, 3), (1, 1, 4, 6), (1, 1, 4, 6), (0, 0, 6, 6), (1, 1, 4, 6), (0, 1, 3, 3)],
    }
    for si, rows in want.items():
        sh = L.bxo_shape_table(si).contents
        assert sh.ao == (0 if si == 0 else 1)
        assert [(s.cc, s.cbc, s.nv, s.ni) for s in sh.sides] == rows
    # AO sample counts: axis-aligned quad 4, slope ramp 3, inner-corner zero-normal vertices 5
    assert L.bxo_shape_table(1).contents.sides[0].v[0].n_ao == 4
    assert L.bxo_shape_table(2).contents.sides[3].v[0].n_ao == 3
    assert L.bxo_shape_table(4).contents.sides[3].v[2].n_ao == 5


def test_lone_grass_cube_kat(oracle):
    # SURVEY.md App. C.5
    d = [np.zeros(32768, np.uint32) for _ in range(27)]
    d[13][5 + 32 * 7 + 1024 * 6] = 1 << 16
    v, ix = oracle.mesh_from_dense27(oracle.OracleRegistry(), d)
    assert len(v) == 24 and len(ix) == 36
    assert [int(p) for p in v["position"][:4]] == [0x05018080, 0x0501C080, 0x0501C070, 0x05018070]
    assert [int(p) for p in v["index"][:4]] == [0x000C18E5, 0x000818E5, 0x000A18E5, 0x000818E5]
    assert int(v["color"][0]) == 0xFFFFFFFF
    assert float(v["texcoord"][0][2]) == 16.0 and (v["barycentric_color_offset"][0] == 0).all()
    assert list(ix[:12]) == [0, 1, 2, 2, 3, 0, 4, 5, 6, 6, 7, 4]
    # texture of the top side is grass_top (29), bottom dirt (15)
    assert float(v["texcoord"][3 * 4][2]) == 29.0 and float(v["texcoord"][2 * 4][2]) == 15.0


def test_ao_colour_lut(oracle):
    # SURVEY.md App. C.5 colour LUT: k occluders -> colour
    want = {0: 0xFFFFFFFF, 1: 0xE0E0E0FF, 2: 0xC5C5C5FF, 3: 0xADADADFF}
    d = [np.zeros(32768, np.uint32) for _ in range(27)]
    c = d[13]
    c[10 + 32 * 10 + 1024 * 10] = 4 << 16            # the cube
    c[11 + 32 * 11 + 1024 * 11] = 4 << 16            # diagonal (corner) neighbour of its +x+y+z vertex
    v, _ = oracle.mesh_from_dense27(oracle.OracleRegistry(), d)
    seen = {int(x) for x in v["color"]}
    assert seen <= set(want.values()) | {0x989898FF}
    assert 0xE0E0E0FF in seen and 0xFFFFFFFF in seen


def test_mesh_rle_and_dense_entry_points_agree(oracle):
    reg = oracle.OracleRegistry()
    d = [oracle.random_chunk(dx, dy, dz, seed=3, void_threshold=200) for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
    d[0] = np.zeros(32768, np.uint32)                # uniform chunk = 3 RLE words
    d[26] = np.full(32768, 4 << 16, np.uint32)
    rle = [oracle.compress_rle(c) for c in d]
    assert len(rle[0]) == 3
    a = oracle.mesh_from_dense27(reg, d)
    b = oracle.mesh_from_chunk(reg, rle)
    assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1]) and len(a[0]) > 0
    assert oracle.is_chunk_trivial(reg, rle[0]) and not oracle.is_chunk_trivial(reg, rle[26])
    assert not oracle.is_chunk_trivial(reg, rle[13])


def test_generator_crosscheck_values(oracle):
    """values of an independent transliteration of the same published algorithms.
    They pin the two restatements to each other, NOT to the real `noise` crate (parity unpinned)."""
    assert list(oracle.permutation_table(0)[:16]) == [181, 206, 3, 127, 58, 121, 169, 47, 62, 123, 5, 142, 99, 140, 178, 171]
    g = oracle.StdGen(0)
    rows = [((300, 28), 26.719526088188, 27, 1, 0.650214821251, 0.647495037387),
            ((0, 0), 30.0, 30, 1, 0.5, 0.476404908898),
            ((-1, -1), 29.824104878788, 30, 0, 0.498994885027, 0.476404908898),
            ((1000, 1000), 30.358520602134, 30, 1, 0.586039694495, 0.589072657147),
            ((-5000, 7000), 37.846638445066, 38, 1, 0.647768485887, 0.651305073208),
            ((16383, 16383), 24.309477331092, 24, 1, 0.586773141391, 0.576854145618)]
    for (x, z), h, hr, el, ec, cec in rows:
        p = g.params(x, z)
        assert abs(p.height_f64 - h) < 1e-9 and p.height == hr and p.elevation == el
        assert abs(p.ec - ec) < 1e-9 and abs(p.cec - cec) < 1e-9


def test_noise_is_bounded_and_continuous(oracle):
    perm = oracle.permutation_table(12345)
    rng = np.random.default_rng(0)
    pts = rng.uniform(-500, 500, size=(5000, 2))
    vals = np.array([oracle.super_simplex(perm, x, y) for x, y in pts])
    assert np.abs(vals).max() <= 1.0 + 1e-9 and vals.std() > 0.1
    step = np.array([oracle.super_simplex(perm, x + 1e-7, y) for x, y in pts])
    assert np.abs(step - vals).max() < 1e-5


def test_generate_chunk_layers(oracle):
    """Block selection (stdgen.rs:349-373): grass at h, 5 dirt below, stone under, void above."""
    g = oracle.StdGen(0)
    c = g.generate_chunk(9, 0, 0).reshape(32, 32, 32)  # [y][z][x], world x 288..319, spawn column (300, 28)
    p = g.params(300, 28)
    col = c[:, 28, 300 - 288]
    h = p.height
    assert col[h] >> 16 == 1 and (col[h + 1:] == 0).all()
    assert (col[h - 5:h] >> 16 == 3).all() and (col[:h - 5] >> 16 == 4).all()


def test_terragen_supersimplex(oracle):
    """bxw_terragen/src/supersimplex.rs: permutation is a permutation; smoke points of check_working_noise."""
    L = oracle.lib()
    s = L.bxo_tg_simplex_new(7)
    perm = np.ctypeslib.as_array(L.bxo_tg_simplex_perm(s), shape=(2048,)).copy()
    assert sorted(perm.tolist()) == list(range(2048))
    for c in [0.1, 0.5, 0.9, 1.0, 2.0, 3.0, 4.9, 5.753]:
        v = L.bxo_tg_simplex_noise2(s, c, c)
        assert abs(v) <= 1.0 + 1e-9
    L.bxo_tg_simplex_free(s)
    # math.rs:24-30 splitmix64 KAT (Vigna's reference: seed 0 -> 0xE220A8397B1DCDAF)
    ns = C.c_uint64()
    assert L.bxo_tg_splitmix64(0, C.byref(ns)) == 0xE220A8397B1DCDAF and ns.value == 0x9E3779B97F4A7C15
