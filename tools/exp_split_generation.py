#!/usr/bin/env python
"""Cost of the split generation (boundary columns first, the rest on the side stream) against one launch, one GPU, no exchange."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ballxworld_b200 as bxw  # noqa: E402

ctx = bxw.Context(0)
st = torch.cuda.Stream(device=0)
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
region = bxw.Region((0, 0, 0), (128, 16, 128), bxw.standard_registry(), ctx=ctx)


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(n):
        fn()
    b.record(st)
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n, 4)


def split():
    region.generate_part(1, 0, 64)
    region.generate_part(2, 0, 64)


out = {"one_launch_ms": timed(lambda: region.generate(0, 64)), "split_ms": timed(split),
       "part1_only_ms": None}
print(json.dumps(out))
