"""GPU tests of the device-side load-anchor scheduler (bxw_region_update_anchors / _apply_deltas) against a Python restatement
of World::recalculate_load_deltas (tests/load_deltas_ref.py, bxw_world/src/worldmgr.rs:633-790)."""
import numpy as np
import pytest

import load_deltas_ref as R

pytestmark = pytest.mark.gpu

ORIGIN = (6, 0, -2)
DIMS = (10, 3, 9)  # interior chunks; storage box = +1 shell


def _state_sets(ctx):
    b, m = ctx.region_load_state()
    nx, ny, nz = DIMS
    SX, SY, SZ = nx + 2, ny + 2, nz + 2
    blocks, mesh = set(), set()
    for c in np.flatnonzero(b):
        sx, sz, sy = int(c % SX), int((c // SX) % SZ), int(c // (SX * SZ))
        blocks.add((ORIGIN[0] - 1 + sx, ORIGIN[1] - 1 + sy, ORIGIN[2] - 1 + sz))
    for s in np.flatnonzero(m):
        ix, iz, iy = int(s % nx), int((s // nx) % nz), int(s // (nx * nz))
        mesh.add((ORIGIN[0] + ix, ORIGIN[1] + iy, ORIGIN[2] + iz))
    return blocks, mesh


def _check_list(ctx, anchors):
    """device list == restatement: same deltas, both sub-lists sorted by distance, interleaved unload/load"""
    blocks, mesh = _state_sets(ctx)
    box_min = tuple(o - 1 for o in ORIGIN)
    box_dims = tuple(d + 2 for d in DIMS)
    un, ld = R.recalculate_load_deltas(anchors, box_min, box_dims, ORIGIN, DIMS, blocks, mesh)
    nu, nl = ctx.region_update_anchors(anchors)
    assert (nu, nl) == (len(un), len(ld))
    got = ctx.region_load_deltas()
    assert len(got) == nu + nl
    g_un = [((int(d["cx"]), int(d["cy"]), int(d["cz"])), int(d["kinds"]), int(d["min_anchor_distance"])) for d in got if d["unload"]]
    g_ld = [((int(d["cx"]), int(d["cy"]), int(d["cz"])), int(d["kinds"]), int(d["min_anchor_distance"])) for d in got if not d["unload"]]
    assert sorted(g_un) == sorted(un) and sorted(g_ld) == sorted(ld)
    assert [t[2] for t in g_un] == sorted(t[2] for t in g_un) and [t[2] for t in g_ld] == sorted(t[2] for t in g_ld)  # nearest first
    # the interleaving of worldmgr.rs:699-731
    kinds_seq = ["unload" if d["unload"] else "load" for d in got]
    assert kinds_seq == [t[0] for t in R.interleave(un, ld)]
    return un, ld


def test_load_deltas_follow_the_reference(ctx, oracle):
    import ballxworld_b200 as bxw

    reg = bxw.standard_registry()
    region = bxw.Region(ORIGIN, DIMS, reg, ctx=ctx)
    region.unload_all()
    assert _state_sets(ctx) == (set(), set())
    a1 = [(9, 1, 2, 3, 1)]
    un, ld = _check_list(ctx, a1)
    assert not un and ld and all(k == R.KIND_BLOCKS for _, k, _ in ld)  # nothing loaded yet: only block data can be requested
    st = ctx.region_apply_deltas(0, 64)
    assert st["generated"] == len(ld) and st["meshed"] == 0 and st["consumed"] == len(ld)
    # second tick: meshes of chunks whose 27 neighbourhood now holds block data
    un, ld = _check_list(ctx, a1)
    assert ld and all(k == R.KIND_MESH for _, k, _ in ld)
    st = ctx.region_apply_deltas(0, 64)
    assert st["meshed"] == len(ld)
    un, ld = _check_list(ctx, a1)
    assert not un and not ld  # steady state
    # the generated blocks and the meshes are the oracle's
    g = oracle.StdGen(0)
    blocks, mesh = _state_sets(ctx)
    spans = ctx.region_mesh_spans()
    oreg = oracle.OracleRegistry()
    checked = 0
    for p in sorted(mesh)[::7]:
        d = [g.generate_chunk(p[0] + dx, p[1] + dy, p[2] + dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
        l = tuple(p[k] - ORIGIN[k] for k in range(3))
        assert np.array_equal(ctx.region_download_chunk(*l), d[13])
        wv, wi = oracle.mesh_from_dense27(oreg, d)
        v, i = ctx.region_mesh_fetch_span(spans[l[0] + DIMS[0] * (l[2] + DIMS[2] * l[1])])
        assert v.tobytes() == wv.tobytes() and np.array_equal(i, wi)
        checked += 1
    assert checked >= 3
    # two anchors, one without meshes (a server-side anchor); the first one moves: loads, unloads, priorities from both
    a2 = [(11, 1, 2, 3, 1), (8, 1, 6, 2, 0)]
    un, ld = _check_list(ctx, a2)
    assert un and ld
    st = ctx.region_apply_deltas(0, 64)
    assert st["unmeshed"] == sum(1 for _, k, _ in un if k & R.KIND_MESH) and st["generated"] == len(ld)
    for _ in range(3):
        _check_list(ctx, a2)
        ctx.region_apply_deltas(0, 64)
    un, ld = _check_list(ctx, a2)
    assert not un and not ld
    blocks, mesh = _state_sets(ctx)
    assert all(any(sum((p[k] - a[k]) ** 2 for k in range(3)) <= a[3] ** 2 for a in a2) for p in blocks)
    assert all(sum((p[k] - a2[0][k]) ** 2 for k in range(3)) <= 9 for p in mesh)
    # max_deltas: a partial tick consumes the head of the list only (the reference stops after 4096 deltas / 256 tasks)
    a3 = [(13, 1, 4, 3, 1)]
    un, ld = _check_list(ctx, a3)
    st = ctx.region_apply_deltas(0, 64, max_deltas=10)
    assert st["consumed"] == 10
    _check_list(ctx, a3)
    # no anchors: everything unloads
    ctx.region_apply_deltas(0, 64)
    for _ in range(3):
        ctx.region_update_anchors(a3)
        ctx.region_apply_deltas(0, 64)
    un, ld = _check_list(ctx, [])
    assert un and not ld
    ctx.region_apply_deltas(0, 64)
    assert _state_sets(ctx) == (set(), set())
    region.close()


def test_anchor_sweep_keeps_the_arena_bounded(ctx):
    """The view-distance sweep of BASELINE configs[4] in miniature: an anchor walks along x, each tick loads what entered the
    sphere and unloads what left it; meshes that left give their arena space back (compaction once a quarter is garbage)."""
    import ballxworld_b200 as bxw

    reg = bxw.standard_registry()
    region = bxw.Region((0, 0, 0), (40, 3, 8), reg, ctx=ctx)
    region.unload_all()
    peak, max_live, total_loaded = 0, 0, 0
    for step in range(32):
        anchors = [(4 + step, 1, 4, 4, 1)]
        for _ in range(2):  # blocks, then meshes
            st = region.tick(anchors)
            total_loaded += st["meshed"]
        sp = region.spans()
        live = int(sp["v_count"].astype(np.int64).sum())
        stt = ctx.region_arena_stats()
        assert stt["vertices_used"] - stt["vertices_garbage"] >= live
        max_live = max(max_live, live)
        peak = max(peak, stt["vertices_used"])
        b, m = ctx.region_load_state()
        assert 0 < int(m.sum()) <= 3 * 8 * 9
    assert max_live > 0 and total_loaded > 3 * int(m.sum())
    assert peak < 2.5 * max_live + 200000, (peak, max_live)  # append-only, the arena would hold every mesh ever loaded
    region.close()
