#!/usr/bin/env python
"""Streaming edit benchmark (BASELINE.json configs[4] on one GPU, SURVEY.md section 8d cfg 5): a load anchor sweeps +x one
chunk per step; every step applies N random block edits inside the anchor sphere (radius 10, src/config.rs:35-36) through
World::apply_voxel_changes semantics and re-meshes the dirty chunks. Reports dirty chunks re-meshed per second and the
edit->mesh latency (host wall clock around apply_changes + remesh_dirty, device synchronised)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import ballxworld_b200 as bxw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="96,16,32")
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--edits", type=int, default=4096)
    ap.add_argument("--radius", type=int, default=10)
    a = ap.parse_args()
    nx, ny, nz = (int(x) for x in a.dims.split(","))
    ctx = bxw.Context(0)
    reg = bxw.standard_registry()
    region = bxw.Region((0, 0, 0), (nx, ny, nz), reg, ctx=ctx)
    region.generate(0)
    region.mesh()
    rng = np.random.default_rng(1)
    lat, dirty = [], 0
    for step in range(a.steps):
        anchor = np.array([a.radius + step, ny // 4, nz // 2])
        # random positions inside the anchor sphere, random standard blocks with random shapes/orientations
        off = rng.integers(-a.radius * 32, a.radius * 32, size=(a.edits * 2, 3))
        off = off[(off.astype(np.int64) ** 2).sum(axis=1) <= (a.radius * 32) ** 2][: a.edits]
        pos = anchor * 32 + off
        ch = np.zeros(len(pos), dtype=bxw.CHANGE_DTYPE)
        ch["x"], ch["y"], ch["z"] = pos[:, 0], pos[:, 1], pos[:, 2]
        ids = rng.integers(1, 8, len(pos)).astype(np.uint32)
        meta = (rng.integers(0, 24, len(pos)) * 64 + rng.integers(0, 4, len(pos))).astype(np.uint32)
        ch["to"] = (ids << 16) | meta
        ch["from"] = 0  # place into air; edits whose target is not air are no-ops (compare-and-swap) but still dirty the chunk
        ctx.synchronize()
        t0 = time.perf_counter()
        nd = ctx.region_apply_changes(ch)
        nr = ctx.region_remesh_dirty()
        ctx.synchronize()
        lat.append(time.perf_counter() - t0)
        dirty += nr
        assert nr == nd
    lat = np.array(lat[4:])
    print(json.dumps({"workload": f"{nx}x{nz}x{ny} chunks, {a.steps} steps x {a.edits} edits in a radius-{a.radius} sphere",
                      "dirty_chunks_remeshed": dirty, "dirty_chunks_per_s": dirty / sum(lat) * len(lat) / a.steps if len(lat) else None,
                      "edit_to_mesh_latency_ms_median": float(np.median(lat) * 1e3), "edit_to_mesh_latency_ms_p95": float(np.percentile(lat, 95) * 1e3)}))
    region.close()
    ctx.close()


if __name__ == "__main__":
    main()
