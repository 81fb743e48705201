#!/usr/bin/env python
"""Experiment: two regions in flight on one GPU (two contexts, one stream each): the generation of one region overlaps the
meshing of the other. Prints ms per region pass, sequential vs overlapped."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ballxworld_b200 as bxw  # noqa: E402

dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "128,16,128").split(","))
reg = bxw.standard_registry()
ctxs, regions, streams = [], [], []
for k in range(2):
    ctx = bxw.Context(0)
    st = torch.cuda.Stream(device=0)
    ctx.set_stream(st.cuda_stream)
    ctxs.append(ctx)
    streams.append(st)
    regions.append(bxw.Region((0, 0, 0), dims, reg, ctx=ctx))


def seq(n):
    for k in range(n):
        r = regions[k & 1]
        r.generate(0, 64)
        r.mesh()


def ovl(n):  # n even; region A's mesh call blocks the host while B's generation (already queued) runs beside it
    regions[0].generate(0, 64)
    for k in range(n):
        regions[(k + 1) & 1].generate(0, 64)
        regions[k & 1].mesh()
    regions[n & 1].mesh()  # drain the extra generation so both variants end in the same state


out = {}
for name, fn in (("sequential", seq), ("overlapped", ovl)):
    fn(4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(20)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name + "_ms_per_pass"] = round(dt * 1000 / (20 if name == "sequential" else 21), 4)
print(json.dumps(out))
