#!/usr/bin/env python
"""profiles/r02_sass_evidence.md: instruction counts and selected opcodes of every kernel in libbxw_cuda.so (cuobjdump -sass)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ballxworld_b200", "csrc", "libbxw_cuda.so")
PICK = ["UBLKCP", "FENCE.VIEW.ASYNC", "UTMACMDFLUSH", "LDGSTS", "DEPBAR", "PRMT", "MATCH", "VOTE", "REDUX", "RED", "ATOMG", "POPC",
        "BAR.SYNC", "CALL", "LDS.128", "LDS.64", "STS.128", "STS.64", "LDG.E.128", "STG.E.EF", "DMUL", "DADD", "DFMA", "FMUL", "FADD", "FFMA"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
kernels, cur, k = [], None, 0
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = {"name": names[k], "n": 0, "ops": collections.Counter()}
        k += 1
        kernels.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur["n"] += 1
        op = m.group(1)
        for p in PICK:
            if op == p or op.startswith(p + "."):
                cur["ops"][p] += 1
                break


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(bxw::.*$|\(unsigned.*$|\(const.*$|\(int.*$", "", n)
    return n.replace("bxw::", "")


print("| kernel | SASS instructions | selected opcodes |\n|---|---|---|")
for c in sorted(kernels, key=lambda c: -c["n"]):
    print(f"| `{short(c['name'])}` | {c['n']} | " + ", ".join(f"{p} {c['ops'][p]}" for p in PICK if c["ops"][p]) + " |")
