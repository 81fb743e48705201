#!/usr/bin/env python
"""Generates tests/golden/cfg2_golden.json from the ORACLE (oracle/, CPU restatement of the reference mesher).

BASELINE.json configs[1] / SURVEY.md section 8d cfg 2: 16x16x16 interior chunks of hashed random blocks (all shapes, all 24
orientations; seed 1) inside an 18^3 box, dense variant (void threshold 128) and sparse variant (250). For EVERY one of the
4096 chunks of both variants: the first 8 bytes of SHA-256(vertex bytes || index bytes) and (V, I); plus the SHA-256 over
all full digests. The mesher's tables are all in-tree reference source (voxmesh.rs, stdshapes.rs, direction.rs), so these
are vectors of a loop-for-loop restatement; the reference itself holds no mesher vectors."""
import hashlib
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

N = 16
SEED = 1


def layer(args):
    thr, y = args
    reg = O.OracleRegistry()
    cache = {}

    def chunk(x, yy, z):
        k = (x, yy, z)
        if k not in cache:
            cache[k] = O.random_chunk(x, yy, z, seed=SEED, void_threshold=thr)
        return cache[k]

    out = []
    for z in range(N):
        for x in range(N):
            d = [chunk(x + dx, y + dy, z + dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
            v, i = O.mesh_from_dense27(reg, d)
            out.append((hashlib.sha256(v.tobytes() + i.astype("<u4").tobytes()).digest(), len(v), len(i)))
    return y, out


def main():
    res = {"seed": SEED, "dims": [N, N, N], "variants": {}}
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        for thr in (128, 250):
            layers = dict(pool.map(layer, [(thr, y) for y in range(N)]))
            digests, counts = [], []
            for y in range(N):  # chunk order x + nx*(z + nz*y) = the span table's order
                for (dg, v, i) in layers[y]:
                    digests.append(dg)
                    counts.append((v, i))
            counts = np.array(counts, dtype="<i8")
            res["variants"][str(thr)] = {
                "total_vertices": int(counts[:, 0].sum()), "total_indices": int(counts[:, 1].sum()),
                "counts_sha256": hashlib.sha256(counts.tobytes()).hexdigest(),
                "all_digests_sha256": hashlib.sha256(b"".join(digests)).hexdigest(),
                "digest8": [d[:8].hex() for d in digests],
            }
            print(thr, {k: v for k, v in res["variants"][str(thr)].items() if k != "digest8"}, flush=True)
    json.dump(res, open(os.path.join(HERE, "cfg2_golden.json"), "w"), separators=(",", ":"), sort_keys=True)


if __name__ == "__main__":
    main()
