#!/usr/bin/env python
"""FP32 vs f64 noise evaluation of the StdGenerator kernels (north_star's tolerance report).

For a region, generates the column parameters with precision=64 (the reference's arithmetic) and precision=32 and
reports: max relative difference of the raw heights, columns whose rounded height or elevation class differ, voxels
that differ, and how many f64 columns sit within 1e-5 of a selection threshold (|frac(h) - 0.5| for the rounding,
the reported quantity the north_star allows to mismatch)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import ballxworld_b200 as bxw  # noqa: E402


def report(ctx, origin, dims, seed):
    ctx.region_create(*origin, *dims)
    ctx.region_generate(seed, 64)
    h64, e64, raw64 = ctx.region_heightmap(raw=True)
    blocks64 = {k: ctx.region_download_chunk(*k) for k in [(0, 0, 0), (dims[0] - 1, 0, dims[2] - 1), (1, 1, 1)] if all(a < b for a, b in zip(k, dims))}
    ctx.region_generate(seed, 32)
    h32, e32 = ctx.region_heightmap(raw=False)
    # raw f32-path heights: regenerate with raw output is f64-only in the ABI, so compare the rounded results and
    # bound the raw difference through the columns that flipped
    vox_diff = 0
    for k, b in blocks64.items():
        vox_diff += int(np.count_nonzero(ctx.region_download_chunk(*k) != b))
    near = np.abs(np.abs(raw64 - np.floor(raw64)) - 0.5) < 1e-5
    out = {
        "columns": int(h64.size),
        "height_mismatch_columns": int(np.count_nonzero(h64 != h32)),
        "elevation_mismatch_columns": int(np.count_nonzero(e64 != e32)),
        "max_abs_height_step": int(np.abs(h64.astype(np.int64) - h32).max()),
        "f64_columns_within_1e-5_of_rounding_threshold": int(np.count_nonzero(near)),
        "sampled_chunk_voxels_differing": vox_diff,
    }
    ctx.region_destroy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="16,8,16")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    dims = tuple(int(x) for x in a.dims.split(","))
    ctx = bxw.Context(0)
    reg = bxw.standard_registry()
    reg.bind(ctx)
    ctx.set_gen_ids(*bxw.StdGenerator.resolve_ids(reg))
    print(json.dumps(report(ctx, (0, 0, 0), dims, a.seed)))
    ctx.close()


if __name__ == "__main__":
    main()
