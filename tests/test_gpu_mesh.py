"""Parity of the CUDA mesher (through the C ABI) against the oracle: bit-exact vertices and indices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def assert_mesh_equal(got, want, what=""):
    gv, gi = got
    wv, wi = want
    assert len(gv) == len(wv) and len(gi) == len(wi), f"{what}: counts {len(gv)},{len(gi)} vs oracle {len(wv)},{len(wi)}"
    # bitwise comparison, floats included
    assert gv.tobytes() == wv.tobytes(), f"{what}: vertex bytes differ (first at {first_diff(gv, wv)})"
    assert np.array_equal(gi, wi), f"{what}: indices differ"


def first_diff(a, b):
    a8 = np.frombuffer(a.tobytes(), dtype=np.uint8).reshape(-1, 40)
    b8 = np.frombuffer(b.tobytes(), dtype=np.uint8).reshape(-1, 40)
    bad = np.nonzero((a8 != b8).any(axis=1))[0]
    if len(bad) == 0:
        return None
    k = int(bad[0])
    return k, a[k], b[k]


def neighborhood(oracle, cx, cy, cz, **kw):
    out = []
    for dy in (-1, 0, 1):
        for dz in (-1, 0, 1):
            for dx in (-1, 0, 1):
                out.append(oracle.random_chunk(cx + dx, cy + dy, cz + dz, **kw))
    return out


def test_lone_grass_cube_kat(ctx, oracle):
    """SURVEY.md App. C.5 known answer + oracle equality."""
    d = [np.zeros(32768, np.uint32) for _ in range(27)]
    d[13][5 + 32 * 7 + 1024 * 6] = 1 << 16
    v, ix = ctx.mesh_chunk27(d)
    assert len(v) == 24 and len(ix) == 36
    assert [int(p) for p in v["position"][:4]] == [0x05018080, 0x0501C080, 0x0501C070, 0x05018070]
    assert [int(p) for p in v["index"][:4]] == [0x000C18E5, 0x000818E5, 0x000A18E5, 0x000818E5]
    assert int(v["color"][0]) == 0xFFFFFFFF and float(v["texcoord"][0][2]) == 16.0
    assert list(ix[:12]) == [0, 1, 2, 2, 3, 0, 4, 5, 6, 6, 7, 4]
    assert_mesh_equal((v, ix), oracle.mesh_from_dense27(oracle.OracleRegistry(), d), "lone cube")


@pytest.mark.parametrize("pos", [(0, 0, 0), (3, -2, 5), (-7, 1, -1)])
@pytest.mark.parametrize("thr", [128, 250, 16])
def test_random_shapes_chunk27(ctx, oracle, pos, thr):
    """cfg 2 style input: hashed random ids, all shapes, all 24 orientations, incl. halo influence."""
    d = neighborhood(oracle, *pos, seed=1, void_threshold=thr)
    got = ctx.mesh_chunk27(d)
    want = oracle.mesh_from_dense27(oracle.OracleRegistry(), d)
    assert len(want[0]) > 0
    assert_mesh_equal(got, want, f"random {pos} thr={thr}")


def test_uniform_and_empty(ctx, oracle):
    reg = oracle.OracleRegistry()
    void = [np.zeros(32768, np.uint32) for _ in range(27)]
    v, ix = ctx.mesh_chunk27(void)
    assert len(v) == 0 and len(ix) == 0
    stone = [np.full(32768, 4 << 16, np.uint32) for _ in range(27)]
    v, ix = ctx.mesh_chunk27(stone)
    assert len(v) == 0 and len(ix) == 0
    # solid centre in a void neighbourhood: the full 6*1024 faces
    d = [np.zeros(32768, np.uint32) for _ in range(27)]
    d[13] = np.full(32768, 4 << 16, np.uint32)
    assert_mesh_equal(ctx.mesh_chunk27(d), oracle.mesh_from_dense27(reg, d), "solid centre")
    # void centre, solid neighbours: nothing (void emits nothing)
    d = [np.full(32768, 3 << 16, np.uint32) for _ in range(27)]
    d[13] = np.zeros(32768, np.uint32)
    v, ix = ctx.mesh_chunk27(d)
    assert len(v) == 0


def test_maximum_output_chunk(ctx, oracle):
    """3-D checkerboard of cubes: every face visible -> 16384 cubes x 24 vertices."""
    g = np.indices((32, 32, 32)).sum(axis=0) & 1
    d = [np.zeros(32768, np.uint32) for _ in range(27)]
    d[13] = np.where(g.reshape(-1) == 1, np.uint32(6 << 16), np.uint32(0)).astype(np.uint32)
    got = ctx.mesh_chunk27(d)
    assert len(got[0]) == 16384 * 24 and len(got[1]) == 16384 * 36
    assert_mesh_equal(got, oracle.mesh_from_dense27(oracle.OracleRegistry(), d), "checkerboard")


def test_odd_meta_values(ctx, oracle):
    """shape > 3 -> cube, orientation >= 24 -> default (stdshapes.rs:54-70); meta on a None-mesh block ignored."""
    rng = np.random.default_rng(7)
    d = []
    for _ in range(27):
        ids = rng.integers(0, 8, 32768).astype(np.uint32)
        meta = rng.integers(0, 65536, 32768).astype(np.uint32)
        keep = rng.random(32768) < 0.3
        d.append(np.where(keep, (ids << 16) | meta, 0).astype(np.uint32))
    assert_mesh_equal(ctx.mesh_chunk27(d), oracle.mesh_from_dense27(oracle.OracleRegistry(), d), "odd meta")


def test_unknown_block_id_is_an_error(ctx):
    import ballxworld_b200 as bxw

    d = [np.zeros(32768, np.uint32) for _ in range(27)]
    # id 200 is not registered; the reference panics (voxregistry.rs:141-143). Slot 4 is the -y neighbour: only its
    # y=31 layer is part of the 34^3 table, so the bad voxel must sit there to be seen at all.
    d[4][4 + 32 * 3 + 1024 * 31] = 200 << 16
    with pytest.raises(bxw.BxwError) as e:
        ctx.mesh_chunk27(d)
    assert e.value.code == -4


def test_region_mesh_matches_oracle(ctx, oracle):
    """Region form: upload random chunks (shell included), mesh all interior chunks, compare each with the oracle."""
    nx, ny, nz = 4, 3, 2
    ctx.region_create(10, -1, 20, nx, ny, nz)
    chunks = {}
    for ly in range(-1, ny + 1):
        for lz in range(-1, nz + 1):
            for lx in range(-1, nx + 1):
                thr = 128 if (lx + ly + lz) % 3 else 245
                c = oracle.random_chunk(10 + lx, -1 + ly, 20 + lz, seed=5, void_threshold=thr)
                chunks[(lx, ly, lz)] = c
                ctx.region_upload_chunk(lx, ly, lz, c)
    av, ai = ctx.region_mesh()
    spans = ctx.region_mesh_spans()
    V, I = ctx.region_mesh_fetch(av, ai)
    reg = oracle.OracleRegistry()
    tot_v = 0
    for iy in range(ny):
        for iz in range(nz):
            for ix in range(nx):
                s = spans[ix + nx * (iz + nz * iy)]
                d = [chunks[(ix + dx, iy + dy, iz + dz)] for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
                want = oracle.mesh_from_dense27(reg, d)
                got = (V[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])], I[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])])
                assert_mesh_equal(got, want, f"region chunk {(ix, iy, iz)}")
                tot_v += len(want[0])
    assert tot_v > 0
    # download returns what was uploaded
    assert np.array_equal(ctx.region_download_chunk(0, 0, 0), chunks[(0, 0, 0)])
    ctx.region_destroy()


def test_large_and_sparse_block_ids(ctx, oracle):
    """Block ids beyond the tables cached on chip (>= 64 for colours / textures, >= 256 for the mesh kind) and beyond 12 bits
    (the face record splits the id over two words), in a registry with holes."""
    import ballxworld_b200 as bxw

    n = 6000
    rng = np.random.default_rng(11)
    defs = (oracle.BlockDef * n)()
    used = [3, 63, 64, 70, 255, 256, 300, 4095, 4096, 4100, 5999]
    defs[0].present, defs[0].mesh_kind = 1, 0                      # void
    for i in used:
        defs[i].present, defs[i].mesh_kind = 1, 1
        for k in range(3):
            defs[i].debug_color[k] = float(rng.random())
        for k in range(6):
            defs[i].tex[k] = int(rng.integers(0, 92))
    defs[300].mesh_kind = 0                                        # a registered block without a mesh
    oreg = oracle.OracleRegistry(defs)
    ctx.set_registry(oracle.defs_to_numpy(defs))
    ctx._bound_registry = None
    try:
        chunks = []
        for c in range(27):
            r = rng.integers(0, 1 << 30, 32768)
            ids = np.array(used, dtype=np.uint32)[r % len(used)]
            shape = np.where((r >> 8) % 4 == 0, (r >> 10) % 4, 0).astype(np.uint32)
            orient = ((r >> 12) % 24).astype(np.uint32)
            d = (ids << 16) | (orient * 64 + shape)
            d[(r >> 20) % 3 == 0] = 0
            chunks.append(d.astype(np.uint32))
        wv, wi = oracle.mesh_from_dense27(oreg, chunks)
        gv, gi = ctx.mesh_chunk27(chunks)
        assert len(wv) > 50000
        assert gv.tobytes() == wv.tobytes() and np.array_equal(gi, wi)
    finally:
        bxw.standard_registry().bind(ctx)


@pytest.mark.gpu
def test_cubes_only_kernel_class_lookup_paths(ctx, oracle):
    """The cubes-only mesh_kernel finds a voxel's class three ways: one byte permute per four voxels when a 128-voxel group holds
    block ids 0..7 with meta 0, a register table for ids below 16, the shared-memory table beyond. All of them against the
    oracle (cubes with every orientation / unknown shape numbers, ids up to 40, a block without a mesh), and a hole in the
    registry among the small ids must still be the reference's panic (voxregistry.rs:141-143)."""
    import ballxworld_b200 as bxw

    n = 41
    defs = (oracle.BlockDef * n)()
    std = oracle.standard_registry_defs()
    for i in range(8):
        defs[i] = std[i]
    for i in range(8, n):
        defs[i].present = 1
        defs[i].mesh_kind = 0 if i in (12, 33) else 1
        for k in range(3):
            defs[i].debug_color[k] = [i / 41.0, 0.5, 1.0 - i / 41.0][k]
        for k in range(6):
            defs[i].tex[k] = i + k
    nx, ny, nz = 2, 1, 2
    rng = np.random.default_rng(11)

    def chunk(kind):
        c = np.zeros((32, 32, 32), np.uint32)  # [y][z][x]
        if kind == 0:  # terrain-like: ids 1..7, meta 0, in long runs along x (the byte-permute path) under a ragged surface
            h = rng.integers(8, 24, (32, 32))
            ids = rng.integers(1, 8, (32, 32, 4)).repeat(8, axis=2)
            c = np.where(np.arange(32)[:, None, None] <= h[None, :, :], ids.astype(np.uint32) << 16, 0).astype(np.uint32)
        elif kind == 1:  # cubes of ids up to 40 with orientations and unknown shape numbers (all are cubes: stdshapes.rs:59)
            ids = rng.integers(1, n, (32, 32, 32)).astype(np.uint32)
            orient = rng.integers(0, 30, (32, 32, 32)).astype(np.uint32)  # >= 24 falls back to the default orientation
            shape = np.where(rng.random((32, 32, 32)) < 0.5, 0, rng.integers(4, 64, (32, 32, 32))).astype(np.uint32)
            c = np.where(rng.random((32, 32, 32)) < 0.4, (ids << 16) | (orient * 64 + shape), 0).astype(np.uint32)
        elif kind == 2:  # mostly small ids with a few large ones sprinkled in (groups that leave the fast path half way)
            ids = np.where(rng.random((32, 32, 32)) < 0.02, rng.integers(8, n, (32, 32, 32)), rng.integers(0, 8, (32, 32, 32)))
            c = (ids.astype(np.uint32) << 16)
        return np.ascontiguousarray(c.reshape(-1))

    try:
        ctx.set_registry(oracle.defs_to_numpy(defs))
        oreg = oracle.OracleRegistry(defs)
        ctx.region_create(-3, 2, 7, nx, ny, nz)
        chunks = {}
        for ly in range(-1, ny + 1):
            for lz in range(-1, nz + 1):
                for lx in range(-1, nx + 1):
                    chunks[(lx, ly, lz)] = chunk((lx + 2 * lz + ly) % 3)
                    ctx.region_upload_chunk(lx, ly, lz, chunks[(lx, ly, lz)])
        av, ai = ctx.region_mesh()
        spans = ctx.region_mesh_spans()
        V, I = ctx.region_mesh_fetch(av, ai)
        for iz in range(nz):
            for ix in range(nx):
                s = spans[ix + nx * iz]
                d = [chunks[(ix + dx, dy, iz + dz)] for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
                wv, wi = oracle.mesh_from_dense27(oreg, d)
                assert int(s["v_count"]) == len(wv) and int(s["i_count"]) == len(wi) and len(wv) > 0
                v0, i0 = int(s["v_off"]), int(s["i_off"])
                assert V[v0:v0 + len(wv)].tobytes() == wv.tobytes(), f"vertices of chunk ({ix}, 0, {iz})"
                assert np.array_equal(I[i0:i0 + len(wi)], wi), f"indices of chunk ({ix}, 0, {iz})"
        # a hole among the small ids: block id 5 is not registered and sits in a group of small ids with meta 0
        defs[5].present = 0
        ctx.set_registry(oracle.defs_to_numpy(defs))
        bad = chunks[(0, 0, 0)].copy()
        bad[3 + 32 * 4 + 1024 * 2] = 5 << 16
        ctx.region_upload_chunk(0, 0, 0, bad)
        with pytest.raises(bxw.BxwError) as e:
            ctx.region_mesh()
        assert e.value.code == -4
    finally:
        ctx.region_destroy()
        ctx.set_registry(oracle.defs_to_numpy(oracle.standard_registry_defs()))
        ctx._bound_registry = None
