#!/usr/bin/env python
"""Latency of the per-chunk drop-ins (host pointers in, host pointers out), the form a maintainer gets by swapping the
bodies of StdGenerator::generate_chunk and mesh_from_chunk, next to the oracle on one core."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import ballxworld_b200 as bxw  # noqa: E402
import oracle_lib as O  # noqa: E402


def main():
    ctx = bxw.Context(0)
    reg = bxw.standard_registry()
    reg.bind(ctx)
    ctx.set_gen_ids(*bxw.StdGenerator.resolve_ids(reg))
    g = O.StdGen(0)
    oreg = O.OracleRegistry()
    pos = [(9 + i % 4, i // 16, (i // 4) % 4) for i in range(32)]     # surface and sky chunks
    out = np.empty(32768, np.uint32)
    for p in pos[:4]:
        ctx.gen_chunk(0, *p, out=out)
    t0 = time.perf_counter()
    for p in pos:
        ctx.gen_chunk(0, *p, out=out)
    t_gen = (time.perf_counter() - t0) / len(pos)
    t0 = time.perf_counter()
    for p in pos:
        g.generate_chunk(*p)
    t_gen_cpu = (time.perf_counter() - t0) / len(pos)
    # mesh: 27 dense neighbours per call
    nb = {}
    for p in pos[:8]:
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    q = (p[0] + dx, p[1] + dy, p[2] + dz)
                    if q not in nb:
                        nb[q] = g.generate_chunk(*q)
    sets = [[nb[(p[0] + dx, p[1] + dy, p[2] + dz)] for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)] for p in pos[:8]]
    for d in sets[:2]:
        ctx.mesh_chunk27(d)
    t0 = time.perf_counter()
    nv = 0
    for d in sets:
        v, i = ctx.mesh_chunk27(d)
        nv += len(v)
    t_mesh = (time.perf_counter() - t0) / len(sets)
    t0 = time.perf_counter()
    for d in sets:
        O.mesh_from_dense27(oreg, d)
    t_mesh_cpu = (time.perf_counter() - t0) / len(sets)
    rles = [[O.compress_rle(c) for c in d] for d in sets]
    for r in rles[:2]:
        ctx.mesh_chunk27_rle(r)
    t0 = time.perf_counter()
    for r in rles:
        ctx.mesh_chunk27_rle(r)
    t_mesh_rle = (time.perf_counter() - t0) / len(rles)
    print(json.dumps({"gen_chunk_us": t_gen * 1e6, "oracle_gen_chunk_us_1core": t_gen_cpu * 1e6,
                      "mesh_chunk27_dense_us": t_mesh * 1e6, "mesh_chunk27_rle_us": t_mesh_rle * 1e6,
                      "oracle_mesh_dense27_us_1core": t_mesh_cpu * 1e6, "vertices_per_mesh_call": nv / len(sets)}))
    ctx.close()


if __name__ == "__main__":
    main()
