#!/usr/bin/env python
"""Text summary of one kernel launch in an .ncu-rep for profiles/: headline throughputs, occupancy, stall reasons per issued
instruction, DRAM bytes and the hottest source lines. usage: ncu_summary.py <rep> <kernel regex> <title> [units for per-unit counts]"""
import csv
import io
import subprocess
import sys

rep, kre, title = sys.argv[1], sys.argv[2], sys.argv[3]
units = sys.argv[4] if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, r = rows[0], rows[2]
g = {k: r[i] for i, k in enumerate(h)}
print(f"## {title}")
print(f"  {g['Kernel Name'][:110]}  grid {g['Grid Size']} x block {g['Block Size']}")
keys = [("duration us", "gpu__time_duration.sum"), ("issue active %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L1/TEX throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("FP64 pipe %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("registers / thread", "launch__registers_per_thread"), ("dynamic smem / block", "launch__shared_mem_per_block_dynamic"),
        ("static smem / block", "launch__shared_mem_per_block_static"), ("warp instructions", "smsp__inst_executed.sum"),
        ("dram read bytes", "dram__bytes_read.sum"), ("dram write bytes", "dram__bytes_write.sum")]
for name, k in keys:
    if k in g:
        print(f"    {name:28s} {g[k]}")
st = sorted(((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k, v in g.items()
             if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and v), reverse=True)
print("    stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st if v >= 0.05))
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, kre] + ([units, "1.5"] if units else ["0", "1.5"]))
print()
