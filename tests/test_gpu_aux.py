"""GPU parity of the rows around the hot path: RLE codec, RLE-fed mesher drop-in, edits + dirty re-mesh, halo planes,
terragen noise, synthetic input generator. All through the C ABI, all bit-exact against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def terrain_chunks(oracle, n=6):
    g = oracle.StdGen(0)
    return [g.generate_chunk(9 + i % 3, i // 3, 0) for i in range(n)]


def test_rle_compress_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(3)
    chunks = [np.zeros(32768, np.uint32), np.ones(32768, np.uint32)]
    chunks += terrain_chunks(oracle)
    chunks.append(rng.integers(0, 16, 32768).astype(np.uint32))           # reference rle_random_cmp style
    chunks.append(rng.integers(0, 2, 32768).astype(np.uint32))            # many short pairs
    chunks.append(np.repeat(rng.integers(0, 5, 32768 // 2), 2).astype(np.uint32))  # worst case: all pairs -> 1.5x
    chunks.append(oracle.random_chunk(1, 2, 3))
    last = np.zeros(32768, np.uint32)
    last[-1] = 7                                                          # run ending at the final voxel
    chunks.append(last)
    dense = np.stack(chunks)
    words, offs = ctx.rle_compress(dense)
    assert offs[0] == 0 and offs[-1] == len(words)
    for i, c in enumerate(chunks):
        want = oracle.compress_rle(c)
        got = words[int(offs[i]): int(offs[i + 1])]
        assert np.array_equal(got, want), f"chunk {i}: {len(got)} vs {len(want)} words"
    assert list(words[:3]) == [0, 0, 32766] and list(words[3:6]) == [1, 1, 32766]  # lib.rs:486-498
    back = ctx.rle_decompress(words, offs)
    assert np.array_equal(back, dense)


def test_rle_decompress_arbitrary_valid_streams(ctx, oracle):
    """Streams an encoder would not produce but decompress_rle accepts (count 0, split runs)."""
    streams = [
        np.array([5, 5, 0, 5, 5, 32764 - 0], np.uint32),          # a a 0 | a a n
        np.array([1, 2, 2, 10, 2, 2, 32768 - 1 - 12 - 2], np.uint32),
        np.array([9, 9, 100, 9, 9, 32768 - 102 - 2], np.uint32),
    ]
    offs = np.zeros(len(streams) + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in streams])
    got = ctx.rle_decompress(np.concatenate(streams), offs)
    for i, s in enumerate(streams):
        rc, want = oracle.decompress_rle(s)
        assert rc == 0 and np.array_equal(got[i], want)
        assert np.array_equal(oracle.rle_iterate(s), want)


@pytest.mark.parametrize("bad", [[0, 0], [0, 0, 32768], [0, 0, 5, 1, 1], [0, 0, 10, 7], [3, 3, 32766, 3]])
def test_rle_decompress_errors(ctx, oracle, bad):
    import ballxworld_b200 as bxw

    w = np.array(bad, np.uint32)
    assert oracle.decompress_rle(w)[0] != 0
    with pytest.raises(bxw.BxwError) as e:
        ctx.rle_decompress(w, np.array([0, len(w)], np.uint64))
    assert e.value.code in (-6, -1)


def test_mesh_from_chunk_dropin_with_vchunks(ctx, oracle):
    """The reference-shaped call: 27 RLE VChunks in iter_neighbors order -> ChunkBuffers."""
    import ballxworld_b200 as bxw

    reg = bxw.standard_registry()
    gen = bxw.StdGenerator(0, ctx=ctx)
    cpos = bxw.ChunkPosition(9, 0, 0)
    vchunks, dense = [], []
    for npos in bxw.iter_neighbors(cpos, True):
        uc = bxw.UncompressedChunk(npos)
        gen.generate_chunk(uc, reg)
        vc = bxw.VChunk(npos)
        vc.compress(uc, ctx=ctx)
        assert np.array_equal(vc.decompress(ctx=ctx).blocks_yzx, uc.blocks_yzx)
        vchunks.append(vc)
        dense.append(uc.blocks_yzx.copy())
    assert [(c.position.x, c.position.y, c.position.z) for c in vchunks][:4] == [(8, -1, -1), (9, -1, -1), (10, -1, -1), (8, -1, 0)]
    bufs = bxw.mesh_from_chunk(reg, vchunks, (16, 16), ctx=ctx)
    wv, wi = oracle.mesh_from_chunk(oracle.OracleRegistry(), [oracle.compress_rle(d) for d in dense])
    assert len(wv) > 0 and bufs.vertices.tobytes() == wv.tobytes() and np.array_equal(bufs.indices, wi)
    # is_chunk_trivial: all-void chunk high above the terrain, vs a buried all-stone chunk
    air = bxw.VChunk(bxw.ChunkPosition(0, 30, 0), np.array([0, 0, 32766], np.uint32))
    stone = bxw.VChunk(bxw.ChunkPosition(0, -3, 0), np.array([4 << 16, 4 << 16, 32766], np.uint32))
    assert bxw.is_chunk_trivial(air, reg, ctx=ctx) and not bxw.is_chunk_trivial(stone, reg, ctx=ctx)
    assert not bxw.is_chunk_trivial(vchunks[13], reg, ctx=ctx)
    # serialize/deserialize (generation.rs:158-194)
    blob = vchunks[13].serialize()
    assert np.array_equal(bxw.VChunk.deserialize(cpos, blob).vox, vchunks[13].vox)


def test_generator_requires_standard_names(ctx):
    import ballxworld_b200 as bxw

    reg = bxw.VoxelRegistry()  # only core:void
    with pytest.raises(LookupError, match="No standard grass block definition found"):
        bxw.StdGenerator(0, ctx=ctx).generate_chunk(bxw.UncompressedChunk(), reg)
    bxw.standard_registry().bind(ctx)  # restore the session registry


def test_synthetic_fill_matches_numpy(ctx, oracle):
    ctx.region_create(-2, 1, 5, 2, 1, 2)
    ctx.region_fill_synthetic(1, 128)
    for l in [(-1, -1, -1), (0, 0, 0), (2, 1, 2), (1, 0, -1)]:
        want = oracle.random_chunk(-2 + l[0], 1 + l[1], 5 + l[2], seed=1, void_threshold=128)
        assert np.array_equal(ctx.region_download_chunk(*l), want), l
    ctx.region_fill_synthetic(9, 250)
    assert np.array_equal(ctx.region_download_chunk(0, 0, 1), oracle.random_chunk(-2, 1, 6, seed=9, void_threshold=250))
    ctx.region_destroy()


def _region_chunks(ctx, nx, ny, nz):
    return {(lx, ly, lz): ctx.region_download_chunk(lx, ly, lz)
            for ly in range(-1, ny + 1) for lz in range(-1, nz + 1) for lx in range(-1, nx + 1)}


def _check_chunks_vs_oracle(ctx, oracle, chunks, dims, which):
    nx, ny, nz = dims
    spans = ctx.region_mesh_spans()
    av = int((spans["v_off"] + spans["v_count"]).max())
    ai = int((spans["i_off"] + spans["i_count"]).max())
    V, I = ctx.region_mesh_fetch(av, ai)
    reg = oracle.OracleRegistry()
    for (ix, iy, iz) in which:
        d = [chunks[(ix + dx, iy + dy, iz + dz)] for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
        wv, wi = oracle.mesh_from_dense27(reg, d)
        s = spans[ix + nx * (iz + nz * iy)]
        gv = V[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])]
        gi = I[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])]
        assert gv.tobytes() == wv.tobytes() and np.array_equal(gi, wi), f"chunk {(ix, iy, iz)}"


def test_apply_changes_and_remesh_dirty(ctx, oracle):
    """World::apply_voxel_changes semantics (worldmgr.rs:312-357): ordered CAS, touching_chunks dirty set, re-mesh."""
    import ballxworld_b200 as bxw

    dims = (3, 2, 3)
    origin = (9, 0, -1)
    ctx.region_create(*origin, *dims)
    ctx.region_generate(0, 64)
    ctx.region_mesh()
    chunks = _region_chunks(ctx, *dims)
    g = oracle.StdGen(0)
    x0, y0, z0 = origin[0] * 32, origin[1] * 32, origin[2] * 32

    def world(lx, ly, lz, x, y, z):
        return (x0 + lx * 32 + x, y0 + ly * 32 + y, z0 + lz * 32 + z)

    def cur(pos):
        lx, ly, lz = (pos[0] - x0) // 32, (pos[1] - y0) // 32, (pos[2] - z0) // 32
        return int(chunks[(lx, ly, lz)][(pos[0] & 31) + 32 * (pos[2] & 31) + 1024 * (pos[1] & 31)])

    slope = bxw.VoxelDatum.new(6, bxw.StdMeta.from_parts(1, 7))
    corner = bxw.VoxelDatum.new(7, bxw.StdMeta.from_parts(2, 13))
    h = g.params(x0 + 40, z0 + 40).height
    p_mid = (x0 + 40, h + 1, z0 + 40)                 # above the surface, interior of chunk (1, *, 1)
    p_edge = world(1, h // 32, 1, 0, h % 32, 5)       # x == 0 of chunk 1 -> also dirties chunk 0
    p_corner = world(0, 0, 0, 0, 0, 0)                # touches the shell on three sides
    changes = [
        (*p_mid, cur(p_mid), slope),
        (*p_mid, slope, corner),                       # second change to the same voxel sees the first
        (*p_mid, slope, 5 << 16),                      # stale `from`: must NOT apply
        (*p_edge, cur(p_edge), 0),                     # dig a hole at a chunk face
        (*p_corner, cur(p_corner) ^ 0xFFFF, 1 << 16),  # wrong `from`: no-op, but still marks dirty
        (x0 - 100000, 0, 0, 0, 1 << 16),               # chunk not loaded: skipped
    ]
    nd = ctx.region_apply_changes(np.array(changes, dtype=bxw.CHANGE_DTYPE))
    exp_dirty = {(1, (h + 1) // 32, 1), (1, h // 32, 1), (0, h // 32, 1), (0, 0, 0)}
    exp_dirty = {c for c in exp_dirty if 0 <= c[0] < dims[0] and 0 <= c[1] < dims[1] and 0 <= c[2] < dims[2]}
    assert nd == len(exp_dirty)
    # expected storage
    def put(pos, v):
        lx, ly, lz = (pos[0] - x0) // 32, (pos[1] - y0) // 32, (pos[2] - z0) // 32
        chunks[(lx, ly, lz)] = chunks[(lx, ly, lz)].copy()
        chunks[(lx, ly, lz)][(pos[0] & 31) + 32 * (pos[2] & 31) + 1024 * (pos[1] & 31)] = v
    put(p_mid, corner)
    put(p_edge, 0)
    for k in [(1, (h + 1) // 32, 1), (1, h // 32, 1), (0, 0, 0)]:
        assert np.array_equal(ctx.region_download_chunk(*k), chunks[k])
    assert ctx.region_remesh_dirty() == len(exp_dirty)
    _check_chunks_vs_oracle(ctx, oracle, chunks, dims, sorted(exp_dirty))
    # untouched chunks still hold their original (valid) meshes
    others = [(ix, iy, iz) for ix in range(3) for iy in range(2) for iz in range(3) if (ix, iy, iz) not in exp_dirty]
    _check_chunks_vs_oracle(ctx, oracle, {k: v for k, v in _region_chunks(ctx, *dims).items()}, dims,
                            [c for c in others if abs(c[0] - 1) + abs(c[2] - 1) > 1][:4])
    assert ctx.region_remesh_dirty() == 0
    ctx.region_destroy()


def test_halo_planes_roundtrip(ctx, oracle):
    import torch

    nx, ny, nz = 2, 1, 2
    ctx.region_create(0, 0, 0, nx, ny, nz)
    ctx.region_fill_synthetic(4, 100)
    nb = ctx.region_halo_bytes()
    assert nb == (nz + 2) * 32 * (ny + 2) * 32 * 4
    buf = torch.zeros(nb // 4, dtype=torch.int32, device="cuda:0")
    for face, lx, x in ((0, 0, 0), (1, nx - 1, 31)):
        ctx.region_halo_export(face, buf.data_ptr())
        ctx.synchronize()
        plane = buf.cpu().numpy().view(np.uint32).reshape((ny + 2) * 32, (nz + 2) * 32)
        for ly in range(-1, ny + 1):
            for lz in range(-1, nz + 1):
                c = ctx.region_download_chunk(lx, ly, lz).reshape(32, 32, 32)  # [y][z][x]
                assert np.array_equal(plane[(ly + 1) * 32:(ly + 2) * 32, (lz + 1) * 32:(lz + 2) * 32], c[:, :, x])
    # import: a marker plane lands in the shell column next to the face
    marker = torch.arange(nb // 4, dtype=torch.int32, device="cuda:0")
    ctx.region_halo_import(0, marker.data_ptr())
    ctx.region_halo_import(1, marker.data_ptr())
    ctx.synchronize()
    m = marker.cpu().numpy().view(np.uint32).reshape((ny + 2) * 32, (nz + 2) * 32)
    lo = ctx.region_download_chunk(-1, 0, 1).reshape(32, 32, 32)
    hi = ctx.region_download_chunk(nx, 0, 1).reshape(32, 32, 32)
    assert np.array_equal(lo[:, :, 31], m[32:64, 64:96]) and np.array_equal(hi[:, :, 0], m[32:64, 64:96])
    assert not np.array_equal(lo[:, :, 30], m[32:64, 64:96])
    ctx.region_destroy()


def test_terragen_noise2_matches_oracle(ctx, oracle):
    L = oracle.lib()
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-3000, 3000, 4000), np.array([0.1, 0.5, 0.9, 1.0, 2.0, 3.0, 4.9, 5.753, 0.0, -0.0, -1.0])])
    ys = np.concatenate([rng.uniform(-3000, 3000, 4000), np.array([0.1, 0.5, 0.9, 1.0, 2.0, 3.0, 4.9, 5.753, 0.0, 7.25, -1.0])])
    for seed in (0, 7, 0x13372137CAFE):
        got = ctx.terragen_noise2(seed, xs, ys)
        s = L.bxo_tg_simplex_new(seed)
        want = np.array([L.bxo_tg_simplex_noise2(s, float(x), float(y)) for x, y in zip(xs, ys)])
        L.bxo_tg_simplex_free(s)
        assert got.tobytes() == want.tobytes(), f"seed {seed}: max abs diff {np.abs(got - want).max()}"


def test_mesh_sphere_view_distance(ctx, oracle):
    """bxw_region_mesh_sphere: the load-anchor sphere of worldmgr.rs:734-742, same meshes as the full-region pass."""
    dims = (6, 3, 6)
    origin = (8, 0, -3)
    ctx.region_create(*origin, *dims)
    ctx.region_generate(0, 64)
    av, ai = ctx.region_mesh()
    full_spans = ctx.region_mesh_spans()
    FV, FI = ctx.region_mesh_fetch(av, ai)
    anchor, radius = (10, 1, -1), 2
    n, sv, si = ctx.region_mesh_sphere(anchor, radius)
    spans = ctx.region_mesh_spans()
    SV, SI = ctx.region_mesh_fetch(sv, si)
    want = 0
    for iy in range(dims[1]):
        for iz in range(dims[2]):
            for ix in range(dims[0]):
                k = ix + dims[0] * (iz + dims[2] * iy)
                d2 = (origin[0] + ix - anchor[0]) ** 2 + (origin[1] + iy - anchor[1]) ** 2 + (origin[2] + iz - anchor[2]) ** 2
                s, f = spans[k], full_spans[k]
                if d2 <= radius * radius:
                    want += 1
                    assert s["v_count"] == f["v_count"] and s["i_count"] == f["i_count"]
                    a = SV[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])]
                    b = FV[int(f["v_off"]): int(f["v_off"]) + int(f["v_count"])]
                    assert a.tobytes() == b.tobytes()
                    assert np.array_equal(SI[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])], FI[int(f["i_off"]): int(f["i_off"]) + int(f["i_count"])])
                else:
                    assert s["v_count"] == 0 and s["i_count"] == 0
    assert n == want and want > 10
    assert ctx.region_mesh_sphere((1000, 0, 0), 3)[0] == 0
    ctx.region_destroy()


def test_region_save_blobs_roundtrip(ctx, oracle):
    """WorldBlocks::serialize_data / deserialize_data (generation.rs:158-194) served from device storage: export equals
    the oracle's compress_rle of the same chunk, import decodes on the device, marks the 27-neighbourhood dirty and the
    re-meshed result matches the oracle."""
    import ballxworld_b200 as bxw

    dims = (3, 2, 2)
    origin = (9, 0, 0)
    reg = bxw.Region(origin, dims, bxw.standard_registry(), ctx)
    reg.generate(0)
    reg.mesh()
    chunks = _region_chunks(ctx, *dims)
    lpos = [(-1, -1, -1), (0, 0, 0), (1, 1, 0), (2, 0, 1), (3, 2, 2), (1, 0, 1)]
    words, offs = ctx.region_export_rle(lpos)
    for i, k in enumerate(lpos):
        assert np.array_equal(words[int(offs[i]):int(offs[i + 1])], oracle.compress_rle(chunks[k])), k
    blobs = reg.serialize_data(lpos)
    assert blobs[1] == oracle.compress_rle(chunks[(0, 0, 0)]).astype("<u4").tobytes()
    vc = reg.get_data(1, 1, 0)
    assert vc.position == bxw.ChunkPosition(10, 1, 0) and np.array_equal(vc.decompress(ctx).blocks_yzx, chunks[(1, 1, 0)])
    # size query / capacity error
    with pytest.raises(bxw.BxwError):
        small = np.empty(2, np.uint32)
        o = np.empty(len(lpos) + 1, np.uint64)
        ctx._ck(ctx.L.bxw_region_export_rle(ctx.h, np.array(lpos, np.int32).ctypes.data, len(lpos), small.ctypes.data, 2, o.ctypes.data))

    # import: swap two chunks' contents and load a hashed chunk into a third, all as save blobs
    new = {(0, 0, 0): chunks[(2, 0, 1)], (2, 0, 1): chunks[(0, 0, 0)], (1, 0, 1): oracle.random_chunk(4, 5, 6)}
    keys = list(new)
    reg.deserialize_data(keys, [oracle.compress_rle(new[k]).astype("<u4").tobytes() for k in keys])
    for k in keys:
        chunks[k] = new[k]
        assert np.array_equal(ctx.region_download_chunk(*k), new[k])
    n = reg.remesh_dirty()
    assert n == dims[0] * dims[1] * dims[2]      # every interior chunk is within one chunk of an imported one
    _check_chunks_vs_oracle(ctx, oracle, chunks, dims, [(x, y, z) for x in range(3) for y in range(2) for z in range(2)])
    # a stream that does not decode leaves the region untouched
    with pytest.raises(bxw.BxwError):
        ctx.region_import_rle([(0, 0, 0), (1, 0, 0)], np.array([1, 1, 32766, 0, 0, 5], np.uint32), np.array([0, 3, 6], np.uint64))
    assert np.array_equal(ctx.region_download_chunk(0, 0, 0), new[(0, 0, 0)])
    assert reg.remesh_dirty() == 0
    with pytest.raises(ValueError):
        reg.deserialize_data([(0, 0, 0)], [b"abc"])
    ctx.region_destroy()


def _all_interior(dims):
    return [(x, y, z) for x in range(dims[0]) for y in range(dims[1]) for z in range(dims[2])]


def test_chunk_summaries_follow_every_writer(ctx, oracle):
    """Whole-region meshing skips chunks whose summaries prove an empty mesh (uniform chunk without a mesh, or uniform
    cubes surrounded by the same uniform cubes). Every kind of write into such a chunk must bring it back."""
    import torch
    import ballxworld_b200 as bxw

    dims = (3, 5, 2)
    origin = (9, -2, 0)
    ctx.region_create(*origin, *dims)
    ctx.region_generate(0, 64)
    ctx.region_mesh()
    chunks = _region_chunks(ctx, *dims)
    sky = [k for k in _all_interior(dims) if not chunks[k].any()]
    deep = [k for k in _all_interior(dims) if k[1] == 0]
    assert sky and all(np.all(chunks[k] == chunks[deep[0]][0]) for k in deep)     # uniform void / uniform stone exist
    _check_chunks_vs_oracle(ctx, oracle, chunks, dims, _all_interior(dims))
    spans = ctx.region_mesh_spans()
    assert all(int(spans[k[0] + dims[0] * (k[2] + dims[2] * k[1])]["v_count"]) == 0 for k in sky)
    listed, interior = ctx.region_mesh_candidates()
    assert interior == len(_all_interior(dims)) and 0 < listed <= interior - len(sky)
    ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, 0)            # every chunk read: same meshes
    ctx.region_mesh()
    assert ctx.region_mesh_candidates() == (interior, interior)
    _check_chunks_vs_oracle(ctx, oracle, chunks, dims, _all_interior(dims))
    ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, 1)
    x0, y0, z0 = origin[0] * 32, origin[1] * 32, origin[2] * 32

    def remesh_and_check():
        ctx.region_mesh()
        _check_chunks_vs_oracle(ctx, oracle, _region_chunks(ctx, *dims), dims, _all_interior(dims))
        return ctx.region_mesh_spans()

    # 1. edits strictly inside a sky chunk and inside a deep stone chunk
    ks, kd = sky[0], (1, 0, 0)
    stone = int(chunks[kd][0])
    p_sky = (x0 + ks[0] * 32 + 7, y0 + ks[1] * 32 + 9, z0 + ks[2] * 32 + 11)
    p_deep = (x0 + kd[0] * 32 + 16, y0 + kd[1] * 32 + 16, z0 + kd[2] * 32 + 16)
    ctx.region_apply_changes(np.array([(*p_sky, 0, stone), (*p_deep, stone, 0)], dtype=bxw.CHANGE_DTYPE))
    spans = remesh_and_check()
    assert int(spans[ks[0] + dims[0] * (ks[2] + dims[2] * ks[1])]["v_count"]) == 24      # one free-standing cube
    assert int(spans[kd[0] + dims[0] * (kd[2] + dims[2] * kd[1])]["v_count"]) == 24      # six walls of the hole
    # 2. dense upload into another sky chunk
    ku = sky[-1]
    assert ku != ks
    ctx.region_upload_chunk(*ku, oracle.random_chunk(3, 1, 4))
    spans = remesh_and_check()
    assert int(spans[ku[0] + dims[0] * (ku[2] + dims[2] * ku[1])]["v_count"]) > 0
    # 3. save-blob import into a deep stone chunk
    ki = (0, 0, 1)
    ctx.region_import_rle([ki], oracle.compress_rle(oracle.random_chunk(8, 8, 8)), np.array([0, len(oracle.compress_rle(oracle.random_chunk(8, 8, 8)))], np.uint64))
    spans = remesh_and_check()
    assert int(spans[ki[0] + dims[0] * (ki[2] + dims[2] * ki[1])]["v_count"]) > 0
    # 4. halo import: void replaces the x = 31 plane of the -x shell, so deep stone at ix = 0 grows -x faces
    nb = ctx.region_halo_bytes()
    ctx.region_halo_import(0, torch.zeros(nb // 4, dtype=torch.int32, device="cuda:0").data_ptr())
    ctx.synchronize()
    spans = remesh_and_check()
    assert int(spans[0 + dims[0] * (0 + dims[2] * 0)]["v_count"]) >= 4 * 32 * 32
    # 5. regenerating restores the summaries and the original meshes
    ctx.region_generate(0, 64)
    ctx.region_mesh()
    _check_chunks_vs_oracle(ctx, oracle, chunks, dims, _all_interior(dims))
    ctx.region_destroy()


def test_arena_reserve_and_regrowth(ctx, oracle):
    """bxw_region_mesh_reserve with room to spare, then with arenas far too small: the pass notices (vertex, index or
    face-record space), grows them and produces the same meshes; appended re-meshes after edits grow the arena too."""
    import ballxworld_b200 as bxw

    dims = (3, 2, 3)
    ctx.region_create(0, 0, 0, *dims)
    ctx.region_fill_synthetic(7, 140)
    av0, ai0 = ctx.region_mesh()
    spans0 = ctx.region_mesh_spans().copy()
    chunks = _region_chunks(ctx, *dims)
    ctx.region_destroy()
    for cap_v, cap_i in ((av0 * 2, ai0 * 2), (1024, 1024), (av0 // 2, ai0 * 2)):
        ctx.region_create(0, 0, 0, *dims)
        ctx.region_fill_synthetic(7, 140)
        ctx.region_mesh_reserve(cap_v, cap_i)
        av, ai = ctx.region_mesh()
        spans = ctx.region_mesh_spans()
        assert np.array_equal(spans["v_count"], spans0["v_count"]) and np.array_equal(spans["i_count"], spans0["i_count"])
        _check_chunks_vs_oracle(ctx, oracle, chunks, dims, [(0, 0, 0), (2, 1, 2), (1, 0, 1)])
        # fill a free-standing void pocket... rather: carve a block out of the middle chunk and append its re-mesh several times
        x, y, z = 32 + 5, 9, 32 + 7
        cur = int(chunks[(1, 0, 1)][(x & 31) + 32 * (z & 31) + 1024 * y])
        for k in range(6):
            nxt = (5 << 16) if cur == 0 else 0
            ctx.region_apply_changes(np.array([(x, y, z, cur, nxt)], dtype=bxw.CHANGE_DTYPE))
            assert ctx.region_remesh_dirty() >= 1
            cur = nxt
        edited = dict(chunks)
        edited[(1, 0, 1)] = chunks[(1, 0, 1)].copy()
        edited[(1, 0, 1)][(x & 31) + 32 * (z & 31) + 1024 * y] = cur
        _check_chunks_vs_oracle(ctx, oracle, edited, dims, [(1, 0, 1), (0, 0, 0)])
        ctx.region_destroy()
