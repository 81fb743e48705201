"""BASELINE.json configs[0] (seed 0, 16x16x8 chunks, generated then meshed) against the committed golden fixture
tests/golden/cfg1_golden.json (made by tests/golden/make_golden.py from the oracle)."""
import hashlib
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1_golden.json")))


def test_oracle_reproduces_golden_subset(oracle):
    """CPU: the oracle still produces the committed vectors (3 chunks of the sub-box + shape of the fixture)."""
    assert GOLD["total_vertices"] == 1455136 and GOLD["total_indices"] == 2182704  # SURVEY.md section 8d scale figures
    g = oracle.StdGen(GOLD["seed"])
    reg = oracle.OracleRegistry()
    keys = sorted(GOLD["sub_box"])[:: max(1, len(GOLD["sub_box"]) // 3)][:3]
    for k in keys:
        x, y, z = (int(v) for v in k.split(","))
        d = [g.generate_chunk(x + dx, y + dy, z + dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
        v, i = oracle.mesh_from_dense27(reg, d)
        e = GOLD["sub_box"][k]
        assert hashlib.sha256(d[13].astype("<u4").tobytes()).hexdigest() == e["blocks"]
        assert hashlib.sha256(v.tobytes()).hexdigest() == e["vertices"]
        assert hashlib.sha256(i.astype("<u4").tobytes()).hexdigest() == e["indices"]


@pytest.mark.gpu
def test_cuda_path_matches_golden_cfg1(ctx):
    """GPU: the whole config-1 region through the C ABI: totals, per-chunk counts, height map and sub-box hashes."""
    nx, ny, nz = GOLD["region"]
    ctx.region_create(0, 0, 0, nx, ny, nz)
    av, ai = ctx.region_generate_and_mesh(GOLD["seed"], 64)
    spans = ctx.region_mesh_spans()
    V, I = ctx.region_mesh_fetch(av, ai)
    assert int(spans["v_count"].astype(np.int64).sum()) == GOLD["total_vertices"]
    assert int(spans["i_count"].astype(np.int64).sum()) == GOLD["total_indices"]
    assert int((spans["v_count"] > 0).sum()) == GOLD["nonempty_chunks"]
    counts = np.stack([spans["v_count"].astype("<i8"), spans["i_count"].astype("<i8")], axis=-1).reshape(ny, nz, nx, 2)
    assert hashlib.sha256(np.ascontiguousarray(counts).tobytes()).hexdigest() == GOLD["counts_sha256"]
    h, e = ctx.region_heightmap()
    assert hashlib.sha256(np.ascontiguousarray(h.astype("<i4")).tobytes()).hexdigest() == GOLD["heightmap_sha256"]
    assert int(h.min()) == GOLD["height_min"] and int(h.max()) == GOLD["height_max"]
    for k, g in GOLD["sub_box"].items():
        x, y, z = (int(v) for v in k.split(","))
        s = spans[x + nx * (z + nz * y)]
        assert (int(s["v_count"]), int(s["i_count"])) == (g["v"], g["i"])
        assert hashlib.sha256(ctx.region_download_chunk(x, y, z).astype("<u4").tobytes()).hexdigest() == g["blocks"]
        assert hashlib.sha256(V[int(s["v_off"]): int(s["v_off"]) + g["v"]].tobytes()).hexdigest() == g["vertices"]
        assert hashlib.sha256(I[int(s["i_off"]): int(s["i_off"]) + g["i"]].astype("<u4").tobytes()).hexdigest() == g["indices"]
    ctx.region_destroy()


CFG3_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg3_totals.json")


@pytest.mark.gpu
def test_cuda_path_matches_oracle_totals_cfg3(ctx):
    """BASELINE.json configs[2] at full size (128x128x16 chunks, 40 GB of blocks): the device's vertex / index totals equal the
    oracle's (tests/golden/make_cfg3_totals.py), with the chunk summaries on and off, and a sample of chunks is identical
    between the two passes."""
    import ballxworld_b200 as bxw

    if not os.path.exists(CFG3_PATH):
        pytest.skip("cfg3_totals.json not generated")
    gold = json.load(open(CFG3_PATH))
    nx, ny, nz = gold["dims"]
    try:
        ctx.region_create(*gold["origin"], nx, ny, nz)
    except bxw.BxwError as e:  # needs ~50 GB of device memory
        pytest.skip(f"region does not fit this device: {e}")
    try:
        av, ai = ctx.region_generate_and_mesh(gold["seed"], 64)
        spans = ctx.region_mesh_spans().copy()
        assert int(spans["v_count"].astype(np.int64).sum()) == gold["total_vertices"]
        assert int(spans["i_count"].astype(np.int64).sum()) == gold["total_indices"]
        listed, interior = ctx.region_mesh_candidates()
        assert interior == nx * ny * nz and int((spans["v_count"] > 0).sum()) <= listed < interior // 4
        V, I = ctx.region_mesh_fetch(av, ai)
        ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, 0)
        av2, ai2 = ctx.region_mesh()
        ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, 1)
        spans2 = ctx.region_mesh_spans()
        assert np.array_equal(spans["v_count"], spans2["v_count"]) and np.array_equal(spans["i_count"], spans2["i_count"])
        V2, I2 = ctx.region_mesh_fetch(av2, ai2)
        nonempty = np.nonzero(spans["v_count"])[0]
        for k in nonempty[:: max(1, len(nonempty) // 200)]:
            a, b = spans[k], spans2[k]
            n, m = int(a["v_count"]), int(a["i_count"])
            assert V[int(a["v_off"]): int(a["v_off"]) + n].tobytes() == V2[int(b["v_off"]): int(b["v_off"]) + n].tobytes()
            assert np.array_equal(I[int(a["i_off"]): int(a["i_off"]) + m], I2[int(b["i_off"]): int(b["i_off"]) + m])
    finally:
        ctx.region_destroy()


CFG2_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg2_golden.json")


def test_oracle_reproduces_cfg2_golden_sample(oracle):
    """CPU: the oracle still produces the committed mesh-only vectors (a few chunks of both variants)."""
    gold = json.load(open(CFG2_PATH))
    reg = oracle.OracleRegistry()
    for thr, picks in (("128", [(0, 0, 0), (7, 9, 3)]), ("250", [(15, 15, 15), (4, 0, 11)])):
        gv = gold["variants"][thr]
        assert len(gv["digest8"]) == 4096
        for (x, y, z) in picks:
            d = [oracle.random_chunk(x + dx, y + dy, z + dz, seed=gold["seed"], void_threshold=int(thr))
                 for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
            v, i = oracle.mesh_from_dense27(reg, d)
            dg = hashlib.sha256(v.tobytes() + i.astype("<u4").tobytes()).digest()
            assert dg[:8].hex() == gv["digest8"][x + 16 * (z + 16 * y)]


@pytest.mark.gpu
@pytest.mark.parametrize("thr", ["128", "250"])
def test_cuda_mesher_matches_golden_cfg2_full_size(ctx, thr):
    """BASELINE.json configs[1] at its stated size: all 4096 chunks of hashed random blocks (every shape, every orientation),
    per-chunk SHA-256 of vertex || index bytes against the oracle-made fixture."""
    gold = json.load(open(CFG2_PATH))
    gv = gold["variants"][thr]
    n = 16
    ctx.region_create(0, 0, 0, n, n, n)
    ctx.region_fill_synthetic(gold["seed"], int(thr))
    av, ai = ctx.region_mesh()
    spans = ctx.region_mesh_spans()
    assert int(spans["v_count"].astype(np.int64).sum()) == gv["total_vertices"]
    assert int(spans["i_count"].astype(np.int64).sum()) == gv["total_indices"]
    counts = np.stack([spans["v_count"].astype("<i8"), spans["i_count"].astype("<i8")], axis=-1)
    assert hashlib.sha256(np.ascontiguousarray(counts).tobytes()).hexdigest() == gv["counts_sha256"]
    digests = []
    for k in range(4096):  # chunk by chunk: the dense variant's arena is 45 GB
        v, i = ctx.region_mesh_fetch_span(spans[k])
        digests.append(hashlib.sha256(v.tobytes() + i.tobytes()).digest())
    bad = [k for k in range(4096) if digests[k][:8].hex() != gv["digest8"][k]]
    assert not bad, f"{len(bad)} chunks differ, first {bad[:5]}"
    assert hashlib.sha256(b"".join(digests)).hexdigest() == gv["all_digests_sha256"]
    ctx.region_destroy()
