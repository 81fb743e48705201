#!/usr/bin/env python
"""Mesh-only benchmark (BASELINE.json configs[1], SURVEY.md section 8d cfg 2): 16x16x16 interior chunks of hashed
random blocks (all shapes, all orientations) inside an 18^3 box, meshed from GPU-resident storage.
Prints one JSON line with the mesher's achieved algorithmic GB/s against the measured HBM peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import ballxworld_b200 as bxw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="16,16,16")
    ap.add_argument("--threshold", type=int, default=128, help="void if hash byte < threshold (250 = sparse variant)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    nx, ny, nz = (int(x) for x in a.dims.split(","))
    ctx = bxw.Context(0)
    if os.environ.get("BXW_PIPE"):
        ctx.set_option(ctx.OPT_PIPELINED_MESHER, int(os.environ["BXW_PIPE"]))
    reg = bxw.standard_registry()
    reg.bind(ctx)
    ctx.region_create(0, 0, 0, nx, ny, nz)
    ctx.region_fill_synthetic(1, a.threshold)
    for _ in range(a.warmup):
        av, ai = ctx.region_mesh()
    ctx.mesh_kernel_time_ms(reset=True)
    for _ in range(a.steps):
        av, ai = ctx.region_mesh()
    ems, _ = ctx.kernel_time_ms(ctx.KERNEL_EMIT)
    ms, n = ctx.mesh_kernel_time_ms()
    spans = ctx.region_mesh_spans()
    V = int(spans["v_count"].astype(np.uint64).sum())
    I = int(spans["i_count"].astype(np.uint64).sum())
    n_chunks = nx * ny * nz
    alg = n_chunks * 131072 + 40 * V + 4 * I
    per = ms / n
    peak = 6450.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    print(json.dumps({"workload": f"mesh-only {nx}x{nz}x{ny} chunks, hashed random blocks, void threshold {a.threshold}",
                      "chunks": n_chunks, "vertices": V, "indices": I, "ms_per_launch": per,
                      "mesh_kernel_ms": (ms - ems) / n, "emit_kernel_ms": ems / n,
                      "chunks_per_s": n_chunks / (per / 1e3), "algorithmic_bytes": alg,
                      "achieved_gbs": alg / (per / 1e3) / 1e9, "peak_gbs": peak, "frac": alg / (per / 1e3) / 1e9 / peak}))
    ctx.close()


if __name__ == "__main__":
    main()
