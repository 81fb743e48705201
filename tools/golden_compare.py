#!/usr/bin/env python
"""Compare tools/golden_dump output (real Rust StdGenerator) with the oracle. Pins or refutes oracle/noise.c."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402  (test infrastructure; this tool is a checker, not product code)

golden = json.load(open(sys.argv[1]))
bad = tot = 0
for key, val in golden.items():
    seed = int(key.split("_")[1])
    g = O.StdGen(seed)
    for pos, sha in val["chunks"].items():
        x, y, z = (int(v) for v in pos.split(","))
        got = hashlib.sha256(g.generate_chunk(x, y, z).astype("<u4").tobytes()).hexdigest()
        tot += 1
        bad += got != sha
print(f"{tot - bad}/{tot} chunks match the real generator" + ("" if bad else ": oracle generator parity PINNED"))
sys.exit(1 if bad else 0)
