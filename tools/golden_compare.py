#!/usr/bin/env python
"""Compare tools/golden_dump output (real Rust StdGenerator) with the oracle. Pins or refutes oracle/noise.c."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402  (test infrastructure; this tool is a checker, not product code)

import numpy as np  # noqa: E402

golden = json.load(open(sys.argv[1]))
bad = tot = 0
M64 = (1 << 64) - 1
for seed_s, val in golden.get("noise", {}).items():
    # raw crate samples: PermutationTable::hash and SuperSimplex::get bits (point set: the dump's xorshift64* recurrence)
    seed = int(seed_s)
    perm = O.permutation_table(seed)
    h0 = [int(perm[perm[i & 255] ^ 0]) for i in range(256)]
    hg = [int(perm[perm[((k % 16) * 17 - 100) & 255] ^ (((k // 16) * 13 - 90) & 255)]) for k in range(256)]
    st = 0x9E3779B97F4A7C15 ^ seed
    bits = []
    for _ in range(512):
        xy = []
        for _ in range(2):
            st ^= st >> 12
            st = (st ^ (st << 25)) & M64
            st ^= st >> 27
            r = (st * 0x2545F4914F6CDD1D) & M64
            xy.append(((r >> 44) - (1 << 19)) / 1024.0)
        bits.append("%016x" % np.float64(O.super_simplex(perm, xy[0], xy[1])).view(np.uint64))
    ok = h0 == val["hash_i0"] and hg == val["hash_grid"] and bits == val["get_bits"]
    print(f"noise seed {seed}: permutation hash {'ok' if h0 == val['hash_i0'] and hg == val['hash_grid'] else 'DIFFERS'}, "
          f"SuperSimplex::get {sum(a == b for a, b in zip(bits, val['get_bits']))}/512 bit-exact")
    bad += not ok
for key, val in golden.items():
    if not key.startswith("seed_"):
        continue
    seed = int(key.split("_")[1])
    g = O.StdGen(seed)
    for pos, sha in val["chunks"].items():
        x, y, z = (int(v) for v in pos.split(","))
        got = hashlib.sha256(g.generate_chunk(x, y, z).astype("<u4").tobytes()).hexdigest()
        tot += 1
        bad += got != sha
print(f"{tot - bad}/{tot} chunks match the real generator" + ("" if bad else ": oracle generator parity PINNED"))
sys.exit(1 if bad else 0)
