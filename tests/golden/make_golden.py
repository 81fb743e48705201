#!/usr/bin/env python
"""Generates tests/golden/cfg1_golden.json from the ORACLE (oracle/, the CPU restatement of the reference path).

BASELINE.json configs[0]: StdGenerator seed 0 (src/client/world/mod.rs:33), region 16x16x8 chunks at the origin, meshed with
the client registry. The Rust reference cannot be built in this image, so these are vectors of
the restatement: they pin the CUDA path and any later change of the oracle, NOT the oracle against the real crates.
Content: totals over the 2048 chunks, per-chunk (V, I) for all chunks, SHA-256 of blocks / vertices / indices for a
4x4x8 sub-box (x 8..11, z 0..3, all y) and the height map hash."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

SEED, NX, NY, NZ = 0, 16, 8, 16
SUB = dict(x=(8, 12), z=(0, 4))


def main():
    g = O.StdGen(SEED)
    reg = O.OracleRegistry()
    chunks = {}

    def chunk(x, y, z):
        if (x, y, z) not in chunks:
            chunks[(x, y, z)] = g.generate_chunk(x, y, z)
        return chunks[(x, y, z)]

    counts = np.zeros((NY, NZ, NX, 2), dtype=np.int64)
    sub = {}
    for y in range(NY):
        for z in range(NZ):
            for x in range(NX):
                d = [chunk(x + dx, y + dy, z + dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
                if d[13].min() == d[13].max() == 0:  # is_chunk_trivial
                    continue
                v, i = O.mesh_from_dense27(reg, d)
                counts[y, z, x] = (len(v), len(i))
                if SUB["x"][0] <= x < SUB["x"][1] and SUB["z"][0] <= z < SUB["z"][1]:
                    sub[f"{x},{y},{z}"] = dict(blocks=hashlib.sha256(d[13].astype("<u4").tobytes()).hexdigest(),
                                               vertices=hashlib.sha256(v.tobytes()).hexdigest(),
                                               indices=hashlib.sha256(i.astype("<u4").tobytes()).hexdigest(), v=len(v), i=len(i))
        # drop chunk layers that are no longer needed to bound memory
        for k in [k for k in chunks if k[1] < y - 1]:
            del chunks[k]
    heights = np.array([[g.params(x, z).height for x in range(-32, (NX + 1) * 32)] for z in range(-32, (NZ + 1) * 32)], dtype="<i4")
    out = dict(seed=SEED, region=[NX, NY, NZ], total_vertices=int(counts[..., 0].sum()), total_indices=int(counts[..., 1].sum()),
               nonempty_chunks=int((counts[..., 0] > 0).sum()),
               counts_sha256=hashlib.sha256(counts.astype("<i8").tobytes()).hexdigest(),
               heightmap_sha256=hashlib.sha256(heights.tobytes()).hexdigest(), height_min=int(heights.min()), height_max=int(heights.max()),
               sub_box=sub)
    json.dump(out, open(os.path.join(HERE, "cfg1_golden.json"), "w"), indent=1, sort_keys=True)
    print({k: v for k, v in out.items() if k != "sub_box"}, len(sub), "sub-box chunks")


if __name__ == "__main__":
    main()
