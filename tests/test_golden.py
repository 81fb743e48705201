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
