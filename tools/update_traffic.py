#!/usr/bin/env python
"""Write profiles/traffic.json from ncu --set full captures of the CURRENT build (one report per BASELINE config,
taken at the bench's launch shape), stamped with the hash of the kernel sources so bench.py refuses stale numbers.

usage: tools/update_traffic.py name=report.ncu-rep:events_in_launch [...]      e.g.  c2=gpurun_out/prof_c2.ncu-rep:4096000
Existing entries of other configs are kept only if the stamp still matches."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

path = os.path.join(ROOT, "profiles", "traffic.json")
stamp = bench.source_stamp()
try:
    cur = json.load(open(path))
except (OSError, ValueError):
    cur = {}
if cur.get("_source_stamp") != stamp:
    cur = {}
cur["_source_stamp"] = stamp
cur["_note"] = ("per config: dram__bytes_read.sum + dram__bytes_write.sum of ONE skeleton-kernel launch at the bench's launch "
                "shape and smsp__inst_executed.sum / events (ncu --set full --clock-control none); the zero-fill of the sparse "
                "diagnostic columns (52 B/event) is issued by separate fill launches and is not included")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for arg in sys.argv[1:]:
    name, rest = arg.split("=")
    rep, events = rest.rsplit(":", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, val = rows[0], rows[1], rows[2]
    get = {h: (val[i], units[i]) for i, h in enumerate(hdr)}
    b = sum(float(get[k][0]) * UNIT.get(get[k][1], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    cur[name] = {"dram_bytes_per_launch": b, "events_in_launch": float(events),
                 "warp_inst_per_event": float(get["smsp__inst_executed.sum"][0]) / float(events),
                 "kernel": get["Kernel Name"][0], "kernel_ms": float(get["gpu__time_duration.sum"][0]) * (1e-3 if get["gpu__time_duration.sum"][1] == "us" else 1.0),
                 "report": os.path.basename(rep)}
json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(cur, indent=1, sort_keys=True))
