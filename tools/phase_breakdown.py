#!/usr/bin/env python
"""Per-phase SM cycles of mesh_kernel (stage / count / order / face records; emit_kernel is a separate launch) on the bench terrain (needs the profiling build:
`python -m ballxworld_b200.build --phase-timing` and BXW_CUDA_LIB=ballxworld_b200/csrc/libbxw_cuda_phase.so)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ballxworld_b200 as bxw  # noqa: E402


def show(tag, ph, mhz=1965.0):
    fetch, a, b, c, emitted, chunks = ph[:6]
    tot = fetch + a + b + c
    print(f"{tag}: chunks {chunks} (emitting {emitted}); per chunk: fetch {fetch / max(chunks, 1) / mhz:.2f} us, "
          f"A {a / max(chunks, 1) / mhz:.2f} us, B+scan {b / max(chunks, 1) / mhz:.2f} us, C {c / max(emitted, 1) / mhz:.2f} us "
          f"(shares {100 * fetch / tot:.0f}/{100 * a / tot:.0f}/{100 * b / tot:.0f}/{100 * c / tot:.0f} %); "
          f"C split per emitting chunk: tile setup {ph[6] / max(emitted, 1) / mhz:.2f}, face order {ph[7] / max(emitted, 1) / mhz:.2f}, "
          f"record hand-off {ph[8] / max(emitted, 1) / mhz:.2f}, end barrier {ph[9] / max(emitted, 1) / mhz:.2f} us")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="64,16,64")
    a = ap.parse_args()
    nx, ny, nz = (int(x) for x in a.dims.split(","))
    ctx = bxw.Context(0)
    reg = bxw.standard_registry()
    region = bxw.Region((0, 0, 0), (nx, ny, nz), reg, ctx)
    region.generate(0)
    for _ in range(2):
        region.mesh()
    show("storage (candidate list)", ctx.debug_phase_cycles())
    ctx.set_option(ctx.OPT_MESH_FROM_HEIGHT_MAP, 1)
    for _ in range(2):
        region.generate_and_mesh(0)
    ctx.set_option(ctx.OPT_MESH_FROM_HEIGHT_MAP, 0)
    show("height-map mode", ctx.debug_phase_cycles())
    ctx.region_fill_synthetic(1, 128)
    ctx.region_destroy()
    ctx.region_create(0, 0, 0, 12, 12, 12)
    ctx.region_fill_synthetic(1, 128)
    for _ in range(2):
        ctx.region_mesh()
    show("hashed random blocks (cfg 2)", ctx.debug_phase_cycles())
    ctx.close()


if __name__ == "__main__":
    main()
