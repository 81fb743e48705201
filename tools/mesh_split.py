#!/usr/bin/env python
"""mesh_kernel / emit_kernel split of the mesh pass on the bench terrain (or a slab of it), candidates vs streaming, cubes-only
kernel on / off. usage: mesh_split.py [nx,ny,nz]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ballxworld_b200 as bxw  # noqa: E402

dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "128,16,128").split(","))
ctx = bxw.Context(0)
if os.environ.get("BXW_PIPE"):
    ctx.set_option(ctx.OPT_PIPELINED_MESHER, int(os.environ["BXW_PIPE"]))
reg = bxw.standard_registry()
region = bxw.Region((0, 0, 0), dims, reg, ctx=ctx)
out = {}
for elide in (1, 0):
    ctx.set_option(ctx.OPT_ELIDE_UNIFORM, elide)
    region.generate(0, 64)
    for lean in (1, 0):
        ctx.set_option(ctx.OPT_CUBES_ONLY_MESHER, lean)
        for summ in (1, 0):
            ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, summ)
            region.mesh()
            region.mesh()
            ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
            for _ in range(4):
                region.mesh()
            ems, _ = ctx.kernel_time_ms(ctx.KERNEL_EMIT)
            ms, n = ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
            out[f"elide{elide}_lean{lean}_summaries{summ}"] = {"mesh_kernel_ms": round((ms - ems) / n, 4), "emit_kernel_ms": round(ems / n, 4),
                                                               "listed": ctx.region_mesh_candidates()[0]}
print(json.dumps(out, indent=1))
