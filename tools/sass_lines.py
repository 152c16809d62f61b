#!/usr/bin/env python
"""Static SASS census of ONE skeleton_kernel instantiation, by source line (no GPU needed).

usage: tools/sass_lines.py TEAM SAMPLER POT PATH [lo-hi ...]     (line ranges of csrc/chain.cuh to list)
Compiles the instantiation alone (seconds), disassembles it with line info and prints registers / stack, the
instruction count per source line (static: a line inside a loop counts once) and the opcode mix of the ranges.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pdmpflux.jl_b200", "csrc")
team, sampler, pot, path = sys.argv[1:5]
nw = os.environ.get("NW", "0")
ranges = [tuple(int(x) for x in r.split("-")) for r in sys.argv[5:]]
extra = os.environ.get("EXTRA_NVCCFLAGS", "").split()
with tempfile.TemporaryDirectory() as td:
    src = os.path.join(td, "one.cu")
    open(src, "w").write('#include "chain.cuh"\nnamespace pdmpflux {\ntemplate __global__ void skeleton_kernel<%s, %s, %s, %s, %s>'
                         '(const __grid_constant__ KernelParams);\n}\n' % (team, sampler, pot, path, nw))
    cubin = os.path.join(td, "one.cubin")
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-I", CSRC,
                        "-cubin", "-o", cubin, src, "-Xptxas", "-v"] + extra, capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    for line in r.stderr.splitlines():
        if "registers" in line or "spill" in line and "skeleton" in line:
            print(line.strip())
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    if os.environ.get("KEEP_DIS"):
        open(os.environ["KEEP_DIS"], "w").write(dis)
cur = None
by_line = collections.Counter()
ops = collections.defaultdict(collections.Counter)
total = 0
for l in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        total += 1
        by_line[cur] += 1
        ops[cur][m.group(2).split(".")[0]] += 1
print("total SASS instructions:", total)
if not ranges:
    for (f, n), c in by_line.most_common(40):
        print(f"{f}:{n:5d} {c:6d}")
for lo, hi in ranges:
    tot = collections.Counter()
    n = 0
    for (f, ln), c in sorted(by_line.items()):
        if f == "chain.cuh" and lo <= ln <= hi:
            n += c
            tot.update(ops[(f, ln)])
            print(f"  {ln:5d} {c:5d}  " + " ".join(f"{k}:{v}" for k, v in ops[(f, ln)].most_common(6)))
    print(f"lines {lo}-{hi}: {n} instructions; " + " ".join(f"{k}:{v}" for k, v in tot.most_common(12)))
