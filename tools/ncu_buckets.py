#!/usr/bin/env python
"""Bucket the stall samples / executed instructions of an ncu source page (csv from
`ncu -i rep --page source --csv --print-source sass,cuda`) by the chain.cuh function the line belongs to."""
import collections
import csv
import os
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
ist, iex = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "pdmpflux.jl_b200", "csrc", "chain.cuh")).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.search(r"__device__.*?\b(\w+)\s*\(", l)
    if m and not l.strip().startswith("//"):
        funcs.append((i, m.group(1)))
    if "__global__" in l:
        funcs.append((i, "kernel"))


def fn(ln):
    name = "?"
    for s, n in funcs:
        if s <= ln:
            name = n
        else:
            break
    return name


agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for r in rows[hi + 1:]:
    try:
        ex, st = int(r[iex]), int(r[ist])
    except (ValueError, IndexError):
        continue
    if r[0] != "":
        ln, txt = int(r[0]), r[1]
        ours = ln <= len(src) and src[ln - 1].strip()[:20] == txt.strip()[:20]
        name = fn(ln) if ours else "other: " + txt.strip()[:44]
        a = agg[name]
        a[0] += ex; a[1] += st; tot[0] += ex; tot[1] += st
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{n:52s} inst {100 * a[0] / tot[0]:6.2f}%  stall {100 * a[1] / tot[1]:6.2f}%")
