#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + hottest source lines (needs -lineinfo and --import-source on)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in d:
        print(f"{k:70s} {d[k][0]} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
data = []
cur = None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    if len(r) > 8 and r[2] == '-' and r[0].isdigit():
        try:
            data.append((int(r[6] or 0), int(r[7] or 0), cur, int(r[0]), r[1]))
        except ValueError:
            pass
tot = sum(x[0] for x in data) or 1
toti = sum(x[1] for x in data) or 1
print(f"total samples {tot}, warp instructions {toti}")
for x in sorted(data, reverse=True)[:top]:
    print(f"{100 * x[0] / tot:5.1f}% smp {100 * x[1] / toti:5.1f}% inst  {x[2]}:{x[3]}: {x[4][:110]}")
