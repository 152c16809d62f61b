#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep from `ncu --set full --import-source on`) into a small text file for profiles/.
usage: tools/ncu_summary.py report.ncu-rep n_events [out.txt]"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, n_events = sys.argv[1], float(sys.argv[2])
out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
get = {h: (val[i], units[i]) for i, h in enumerate(hdr)}
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum"]
print(f"# ncu summary of {rep.split('/')[-1]} ({n_events:.0f} events in the profiled launch)", file=out)
for k in KEYS:
    if k in get:
        print(f"{k:70s} {get[k][0]:>18s} {get[k][1]}", file=out)


def tobytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


for k, (v, u) in get.items():  # pipe utilisation, when the report carries it (ncu --set full)
    if re.search(r"sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct|sm__pipe_tensor_cycles_active_realtime|"
                 r"sm__inst_executed_pipe_fp64\.avg\.pct|sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active", k):
        print(f"{k:70s} {v:>18s} {u}", file=out)

rd = tobytes(*get["dram__bytes_read.sum"]); wr = tobytes(*get["dram__bytes_write.sum"])
print(f"dram traffic per event: read {rd / n_events:.1f} B  write {wr / n_events:.1f} B  total {(rd + wr) / n_events:.1f} B", file=out)
inst = float(get["smsp__inst_executed.sum"][0])
print(f"warp instructions per event: {inst / n_events:.1f}", file=out)

det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
print("\n# stall / scheduler notes from the details page", file=out)
for line in det.splitlines():
    if re.search(r"each warp of this workload spends|Eligible Warps Per Scheduler|Issued Warp Per Scheduler|Theoretical Occupancy|Achieved Occupancy|Executed Ipc Active", line):
        print(line.strip(), file=out)

srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(srcp)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
ist, iex, ith = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
agg = defaultdict(lambda: [0, 0, 0, ""])
tot = [0, 0]
ops = defaultdict(int)
for r in rows[hi + 1:]:
    try:
        ex, st, th = int(r[iex]), int(r[ist]), int(r[ith])
    except (ValueError, IndexError):
        continue
    if r[0] == "":
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
        if m:
            ops[m.group(2)] += ex
        continue
    a = agg[r[0]]; a[0] += ex; a[1] += st; a[2] += th; a[3] = r[1].strip()[:90]
    tot[0] += ex; tot[1] += st
print("\n# hottest source lines (share of executed instructions, share of stall samples, active lanes)", file=out)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{ln:>5s} inst {100 * a[0] / tot[0]:5.1f}%  stall {100 * a[1] / max(tot[1], 1):5.1f}%  lanes {a[2] / max(a[0], 1):4.1f} | {a[3]}", file=out)
t = sum(ops.values())
print("\n# SASS opcode mix: " + ", ".join(f"{k} {100 * v / t:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:16]), file=out)
print("# Blackwell-specific stores present: " + ", ".join(k for k in ops if k in ("UBLKCP", "STG", "UTMASTG")), file=out)
