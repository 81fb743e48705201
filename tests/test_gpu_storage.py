"""GPU tests of the storage rows added in round 2: uniform-chunk elision behind a page table (the device analogue of the
reference's 3-word VChunk, bxw_world/src/lib.rs:263-294), in-place re-meshes + arena compaction (the reference drops a
chunk's old ChunkBuffers when the new one lands, src/client/render/voxrender.rs:398-413), region re-positioning, and the
getters that must not write. Everything through the C ABI, bit-exact against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _download_all(ctx, dims):
    nx, ny, nz = dims
    return {(lx, ly, lz): ctx.region_download_chunk(lx, ly, lz)
            for ly in range(-1, ny + 1) for lz in range(-1, nz + 1) for lx in range(-1, nx + 1)}


def _pages(ctx, dims):
    _, n_shared = ctx.region_page_table()
    return ctx.region_download_pages(), n_shared


def _meshes(ctx, dims):
    spans = ctx.region_mesh_spans()
    av, ai = ctx.region_arena_extent()
    V, I = ctx.region_mesh_fetch(av, ai)
    out = {}
    nx, ny, nz = dims
    for iy in range(ny):
        for iz in range(nz):
            for ix in range(nx):
                s = spans[ix + nx * (iz + nz * iy)]
                out[(ix, iy, iz)] = (V[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])].tobytes(),
                                     I[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])].tobytes())
    return out


def _oracle_mesh(oracle, chunks, pos):
    ix, iy, iz = pos
    d = [chunks[(ix + dx, iy + dy, iz + dz)] for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
    wv, wi = oracle.mesh_from_dense27(oracle.OracleRegistry(), d)
    return wv.tobytes(), wi.tobytes()


DIMS = (3, 4, 2)       # y range -1..4 chunks: buried stone below, terrain, sky above
ORIGIN = (9, 0, -1)


def test_elided_region_equals_materialised_region(ctx, oracle):
    """Same blocks, same meshes, same save blobs with and without elision; elided chunks sit on the shared pages."""
    ctx.region_create(*ORIGIN, *DIMS)
    ctx.set_option(ctx.OPT_ELIDE_UNIFORM, 0)
    ctx.region_generate(0, 64)
    ctx.region_mesh()
    pg, n_shared = _pages(ctx, DIMS)
    assert np.array_equal(pg, np.arange(pg.size, dtype=np.uint32)) and n_shared == 2
    ref_chunks = _download_all(ctx, DIMS)
    ref_meshes = _meshes(ctx, DIMS)
    lpos = [(lx, ly, lz) for ly in range(-1, DIMS[1] + 1) for lz in (0, 1) for lx in (0, 2)]
    ref_words, ref_offs = ctx.region_export_rle(lpos)
    ref_words = ref_words.copy()

    ctx.set_option(ctx.OPT_ELIDE_UNIFORM, 1)
    ctx.region_generate(0, 64)
    ctx.region_mesh()
    pg, _ = _pages(ctx, DIMS)
    elided = np.flatnonzero(pg != np.arange(pg.size))
    assert elided.size >= pg.size // 4 and set(pg[elided]) <= {pg.size, pg.size + 1}
    got_chunks = _download_all(ctx, DIMS)
    for k in ref_chunks:
        assert np.array_equal(got_chunks[k], ref_chunks[k]), k
    assert _meshes(ctx, DIMS) == ref_meshes
    words, offs = ctx.region_export_rle(lpos)
    assert np.array_equal(words, ref_words) and np.array_equal(offs, ref_offs)
    # the sky chunk is uniform void on the shared page: exported as the reference's 3-word VChunk
    w, o = ctx.region_export_rle([(1, DIMS[1], 0)])
    assert list(w) == [0, 0, 32766]
    # and the generator's blocks are the oracle's
    g = oracle.StdGen(0)
    for (lx, ly, lz) in [(0, 0, 0), (1, 3, 1), (2, -1, 0), (-1, 4, 2)]:
        assert np.array_equal(got_chunks[(lx, ly, lz)], g.generate_chunk(ORIGIN[0] + lx, ORIGIN[1] + ly, ORIGIN[2] + lz))
    ctx.region_destroy()


def test_writers_materialise_elided_chunks(ctx, oracle):
    """Edit, upload, save-blob import and halo import into chunks that are not materialised."""
    import ballxworld_b200 as bxw

    ctx.region_create(*ORIGIN, *DIMS)
    ctx.region_generate(0, 64)
    ctx.region_mesh()
    chunks = _download_all(ctx, DIMS)
    nst_x, nst_z = DIMS[0] + 2, DIMS[2] + 2
    sidx = lambda l: (l[0] + 1) + nst_x * ((l[2] + 1) + nst_z * (l[1] + 1))
    pg, _ = _pages(ctx, DIMS)
    sky, deep = (1, 3, 0), (1, 0, 1)  # all void / check it is elided; a stone chunk may or may not be elided on this terrain
    assert pg[sidx(sky)] != sidx(sky) and not chunks[sky].any()
    # 1. edit: a block placed in mid-air; a failed CAS (wrong `from`) into another elided chunk leaves it uniform
    gx, gy, gz = (ORIGIN[0] + 1) * 32 + 5, (ORIGIN[1] + 3) * 32 + 7, (ORIGIN[2] + 0) * 32 + 9
    to = bxw.VoxelDatum.new(4, bxw.StdMeta.from_parts(1, 3))  # an oriented stone slope
    ch = np.array([(gx, gy, gz, 0, to), (gx + 32, gy, gz, 12345, to)], dtype=bxw.CHANGE_DTYPE)
    nd = ctx.region_apply_changes(ch)
    assert nd == 2
    chunks[sky] = chunks[sky].copy()
    chunks[sky][5 + 32 * 9 + 1024 * 7] = to
    assert np.array_equal(ctx.region_download_chunk(*sky), chunks[sky])
    assert not ctx.region_download_chunk(2, 3, 0).any()
    pg, _ = _pages(ctx, DIMS)
    assert pg[sidx(sky)] == sidx(sky)
    assert ctx.region_remesh_dirty() == 2
    m = _meshes(ctx, DIMS)
    assert m[sky] == _oracle_mesh(oracle, chunks, sky) and len(m[sky][0]) > 0
    assert m[(2, 3, 0)] == (b"", b"")
    # 2. upload into an elided chunk of the shell, 3. save-blob import into an elided interior chunk
    up = oracle.random_chunk(3, 1, 4, seed=5, void_threshold=240)
    ctx.region_upload_chunk(0, DIMS[1], 1, up)            # y+ shell, above (0, 3, 1)
    chunks[(0, DIMS[1], 1)] = up
    imp = oracle.random_chunk(7, 7, 7, seed=6, void_threshold=250)
    ctx.region_import_rle([(2, 3, 1)], oracle.compress_rle(imp), np.array([0, len(oracle.compress_rle(imp))], np.uint64))
    chunks[(2, 3, 1)] = imp
    ctx.region_mesh()
    m = _meshes(ctx, DIMS)
    for pos in [(0, 3, 1), (2, 3, 1), (1, 3, 1), (2, 3, 0)]:
        assert m[pos] == _oracle_mesh(oracle, chunks, pos), pos
    # 4. halo import: a plane equal to the shell's own content leaves elided chunks alone; a differing voxel materialises
    import torch

    hb = ctx.region_halo_bytes()
    plane = torch.empty(hb // 4, dtype=torch.int32, device="cuda")
    before, _ = _pages(ctx, DIMS)
    # the x+ shell column holds generated terrain: its own x=0 plane is what a neighbour rank would send
    SY, SZ = DIMS[1] + 2, DIMS[2] + 2
    own = np.zeros((SY * 32, SZ * 32), np.uint32)
    for ly in range(-1, DIMS[1] + 1):
        for lz in range(-1, DIMS[2] + 1):
            c = ctx.region_download_chunk(DIMS[0], ly, lz).reshape(32, 32, 32)  # [y][z][x]
            own[(ly + 1) * 32:(ly + 2) * 32, (lz + 1) * 32:(lz + 2) * 32] = c[:, :, 0]
    plane.copy_(torch.from_numpy(own.view(np.int32).reshape(-1)))
    torch.cuda.synchronize()
    ctx.region_halo_import(1, plane.data_ptr())
    after, _ = _pages(ctx, DIMS)
    assert np.array_equal(before, after)
    own2 = own.copy()
    ty, tz = (DIMS[1] + 1) * 32 + 3, 1 * 32 + 4   # a voxel of the sky chunk (3, 4, 0) of the shell column
    assert own2[ty, tz] == 0
    own2[ty, tz] = 4 << 16
    plane.copy_(torch.from_numpy(own2.view(np.int32).reshape(-1)))
    torch.cuda.synchronize()
    ctx.region_halo_import(1, plane.data_ptr())
    after, _ = _pages(ctx, DIMS)
    changed = np.flatnonzero(before != after)
    assert list(changed) == [sidx((DIMS[0], DIMS[1], 0))] and after[changed[0]] == changed[0]
    got = ctx.region_download_chunk(DIMS[0], DIMS[1], 0)
    want = np.zeros(32768, np.uint32)
    want[0 + 32 * 4 + 1024 * 3] = 4 << 16
    assert np.array_equal(got, want)
    ctx.region_destroy()


def test_getters_do_not_write(ctx, oracle):
    """bxw_region_heightmap with the raw f64 output must leave blocks, edits and height map alone."""
    import ballxworld_b200 as bxw

    ctx.region_create(*ORIGIN, 2, 2, 1)
    ctx.region_generate(0x13372137CAFE, 64)
    # another generator call on the same context with a different seed (bxw_gen_chunk shares the context-wide tables)
    g0 = ctx.gen_chunk(0, 9, 0, 0)
    assert np.array_equal(g0, oracle.StdGen(0).generate_chunk(9, 0, 0))
    gx, gy, gz = ORIGIN[0] * 32 + 3, 40, ORIGIN[2] * 32 + 3
    before = ctx.region_download_chunk(0, 1, 0)
    old = int(before[3 + 32 * 3 + 1024 * (gy - 32)])
    ctx.region_apply_changes(np.array([(gx, gy, gz, old, 6 << 16)], dtype=bxw.CHANGE_DTYPE))
    h, e, raw = ctx.region_heightmap(raw=True)
    g = oracle.StdGen(0x13372137CAFE)
    for (x, z) in [(0, 0), (17, 40), (95, 95)]:
        p = g.params((ORIGIN[0] - 1) * 32 + x, (ORIGIN[2] - 1) * 32 + z)
        assert h[z, x] == p.height and e[z, x] == p.elevation and raw[z, x] == p.height_f64
    after = ctx.region_download_chunk(0, 1, 0)
    assert after[3 + 32 * 3 + 1024 * (gy - 32)] == 6 << 16
    before = before.copy()
    before[3 + 32 * 3 + 1024 * (gy - 32)] = 6 << 16
    assert np.array_equal(after, before)
    assert ctx.region_remesh_dirty() >= 1  # the dirty flags survived too
    ctx.region_destroy()


def test_remesh_reuses_arena_space_and_fetches_through_region(ctx, oracle):
    """Region.fetch() after remesh_dirty sees the new meshes; repeated edit cycles keep the arena bounded;
    compaction moves spans without changing a byte of any mesh."""
    import ballxworld_b200 as bxw

    reg = bxw.standard_registry()
    region = bxw.Region(ORIGIN, (3, 2, 3), reg, ctx=ctx)
    region.generate(0, 64)
    region.mesh()
    dims = (3, 2, 3)
    chunks = _download_all(ctx, dims)
    st0 = ctx.region_arena_stats()
    assert st0["vertices_garbage"] == 0 and st0["vertices_used"] == region.arena[0]
    rng = np.random.default_rng(5)
    g = oracle.StdGen(0)
    peak_used = 0
    for cycle in range(12):
        # put / remove blocks on the surface of the centre chunk column
        ch = []
        for _ in range(40):
            x, z = int(rng.integers(0, 96)), int(rng.integers(0, 96))
            gx, gz = ORIGIN[0] * 32 + x, ORIGIN[2] * 32 + z
            h = g.params(gx, gz).height
            y = int(np.clip(h + rng.integers(0, 3), 0, 63))
            key = (x // 32, y // 32, z // 32)
            vi = (x % 32) + 32 * (z % 32) + 1024 * (y % 32)
            old = int(chunks[key][vi])
            to = 0 if old else int(bxw.VoxelDatum.new(int(rng.integers(1, 8)), bxw.StdMeta.from_parts(int(rng.integers(0, 4)), int(rng.integers(0, 24)))))
            ch.append((gx, ORIGIN[1] * 32 + y, gz, old, to))
            chunks[key] = chunks[key].copy()
            chunks[key][vi] = to
        nd = region.apply_voxel_changes(ch)
        assert region.remesh_dirty() == nd
        spans = region.spans()
        V, I = region.fetch()   # sized from the refreshed extent
        for (ix, iy, iz) in [(1, 0, 1), (0, 1, 2), (2, 0, 0), (1, 1, 1)]:
            b = region.chunk_buffers(spans, V, I, ix, iy, iz)
            assert (b.vertices.tobytes(), b.indices.tobytes()) == _oracle_mesh(oracle, chunks, (ix, iy, iz)), (cycle, ix, iy, iz)
        st = ctx.region_arena_stats()
        peak_used = max(peak_used, st["vertices_used"])
    # 12 cycles re-meshed ~100 chunk meshes; without reuse/compaction the arena would have grown by several times its size
    assert peak_used < 2.2 * st0["vertices_used"], (peak_used, st0)
    before = _meshes(ctx, dims)
    av, ai = ctx.region_compact_arena()
    st = ctx.region_arena_stats()
    assert st["vertices_garbage"] == 0 and st["indices_garbage"] == 0 and st["vertices_used"] == av <= peak_used
    assert _meshes(ctx, dims) == before
    # meshing still works after a compaction (fresh arena pointers)
    nd = region.apply_voxel_changes([(ORIGIN[0] * 32 + 40, 70, ORIGIN[2] * 32 + 40, 0, 4 << 16)])
    assert region.remesh_dirty() == nd
    region.close()


def test_set_origin_moves_the_box(ctx, oracle):
    """Slab streaming: one allocation, generate -> mesh -> retire box by box (bxw_region_set_origin)."""
    ctx.region_create(9, 0, 0, 2, 2, 1)
    g = oracle.StdGen(0)
    totals = []
    for x0 in (9, 11):
        ctx.region_set_origin(x0, 0, 0)
        ctx.region_generate(0, 64)
        ctx.region_mesh()
        assert np.array_equal(ctx.region_download_chunk(1, 1, 0), g.generate_chunk(x0 + 1, 1, 0))
        chunks = _download_all(ctx, (2, 2, 1))
        m = _meshes(ctx, (2, 2, 1))
        for pos in [(0, 0, 0), (1, 1, 0)]:
            assert m[pos] == _oracle_mesh(oracle, chunks, pos)
        totals.append(sum(len(v[0]) for v in m.values()))
    assert all(t > 0 for t in totals)
    ctx.region_destroy()


def test_rle_offsets_are_validated_and_long_runs_decode(ctx, oracle):
    import ballxworld_b200 as bxw

    w = np.array([0, 0, 32766], np.uint32)
    with pytest.raises(bxw.BxwError) as e:
        ctx.rle_decompress(np.concatenate([w, w]), np.array([0, 6, 3], np.uint64))
    assert e.value.code == -1
    # more than 256 runs of length 66 (count word 64: the cooperative long-run path), ADVICE r01
    n_runs = 32768 // 66
    vals = np.arange(1, n_runs + 1, dtype=np.uint32)
    dense = np.repeat(vals, 66)
    dense = np.concatenate([dense, np.full(32768 - dense.size, 0xABCD, np.uint32)])
    words = oracle.compress_rle(dense)
    assert (words[2::3][:n_runs] == 64).all()
    got = ctx.rle_decompress(words, np.array([0, len(words)], np.uint64))
    assert np.array_equal(got[0], dense)


def test_generate_in_two_parts_equals_generate(ctx, oracle):
    """bxw_region_generate_part: boundary chunk columns first, the rest second (multi-GPU steps overlap the halo transfer with
    part 2). Same blocks, planes, summaries and meshes as one bxw_region_generate."""
    for dims in [(4, 2, 2), (1, 2, 2), (2, 1, 1)]:
        ctx.region_create(9, 0, -1, *dims)
        ctx.region_generate(0, 64)
        ctx.region_mesh()
        ref = _download_all(ctx, dims)
        ref_m = _meshes(ctx, dims)
        ref_pages = ctx.region_download_pages()
        ref_h = ctx.region_heightmap()
        ctx.region_destroy()
        ctx.region_create(9, 0, -1, *dims)
        ctx.region_generate_part(0, 64, 1)
        import ballxworld_b200 as bxw
        with pytest.raises(bxw.BxwError):
            ctx.region_generate_part(5, 64, 2)  # another seed than part 1
        ctx.region_generate_part(0, 64, 2)
        ctx.region_mesh()
        got = _download_all(ctx, dims)
        for k in ref:
            assert np.array_equal(got[k], ref[k]), (dims, k)
        assert _meshes(ctx, dims) == ref_m
        assert np.array_equal(ctx.region_download_pages(), ref_pages)
        h = ctx.region_heightmap()
        assert np.array_equal(h[0], ref_h[0]) and np.array_equal(h[1], ref_h[1])
        ctx.region_destroy()


def test_pipelined_mesher_equals_sequential(ctx, oracle):
    """BXW_OPT_PIPELINED_MESHER: emit_kernel next to mesh_kernel (two streams, record blocks consumed as they are published)
    produces the same meshes as one kernel after the other -- hashed blocks with every shape, terrain, a re-mesh pass, and a
    pass that overflows the arena and is repeated."""
    import ballxworld_b200 as bxw

    try:
        for kind in ("hashed", "terrain"):
            dims = (5, 4, 5) if kind == "hashed" else (8, 3, 8)
            ctx.region_create(9, 0, -2, *dims)
            if kind == "hashed":
                ctx.region_fill_synthetic(3, 160)
            else:
                ctx.region_generate(0, 64)
            ctx.set_option(ctx.OPT_PIPELINED_MESHER, 0)
            ctx.region_mesh()
            ref = _meshes(ctx, dims)
            ctx.set_option(ctx.OPT_PIPELINED_MESHER, 2)
            for _ in range(3):
                ctx.region_mesh()
                assert _meshes(ctx, dims) == ref, kind
            # a chunk against the oracle directly
            chunks = _download_all(ctx, dims)
            assert ref[(2, 1, 2)] == _oracle_mesh(oracle, chunks, (2, 1, 2))
            # edits + re-mesh in pipelined mode
            gx, gy, gz = 9 * 32 + 70, 40, -2 * 32 + 70
            k = (2, 1, 2)
            vi = (gx % 32) + 32 * (gz % 32) + 1024 * (gy % 32)
            old = int(chunks[k][vi])
            to = int(bxw.VoxelDatum.new(6, bxw.StdMeta.from_parts(2, 7)))
            nd = ctx.region_apply_changes(np.array([(gx, gy, gz, old, to)], dtype=bxw.CHANGE_DTYPE))
            assert ctx.region_remesh_dirty() == nd
            chunks[k] = chunks[k].copy()
            chunks[k][vi] = to
            assert _meshes(ctx, dims)[k] == _oracle_mesh(oracle, chunks, k)
            ctx.region_destroy()
        # arena too small for a pipelined pass: the pass is repeated after the arenas grew
        dims = (4, 4, 4)
        ctx.region_create(0, 0, 0, *dims)
        ctx.region_fill_synthetic(5, 250)
        ctx.set_option(ctx.OPT_PIPELINED_MESHER, 0)
        ctx.region_mesh()
        ctx.region_fill_synthetic(5, 100)   # many more sides than the arena was sized for
        ctx.region_mesh()
        ref = _meshes(ctx, dims)
        ctx.region_destroy()
        ctx.region_create(0, 0, 0, *dims)
        ctx.region_fill_synthetic(5, 250)
        ctx.set_option(ctx.OPT_PIPELINED_MESHER, 2)
        ctx.region_mesh()
        ctx.region_fill_synthetic(5, 100)
        ctx.region_mesh()
        assert _meshes(ctx, dims) == ref
        ctx.region_destroy()
    finally:
        ctx.set_option(ctx.OPT_PIPELINED_MESHER, 0)
