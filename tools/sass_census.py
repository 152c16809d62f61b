#!/usr/bin/env python
"""SASS census of the built library (no GPU needed): which Blackwell-specific mnemonics the kernels contain.

usage: tools/sass_census.py [out.txt]        (default: profiles/r2_sass_census.txt)
Runs `cuobjdump -sass` on pdmpflux.jl_b200/lib/libpdmpflux_cuda.so and counts opcodes per kernel.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pdmpflux.jl_b200", "lib", "libpdmpflux_cuda.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_census.txt")

p = subprocess.Popen(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True)
per_kernel = collections.defaultdict(collections.Counter)
excerpt = {}
cur = None
ins = re.compile(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)")
for line in p.stdout:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = ins.match(line)
    if m and cur:
        op = m.group(1)
        per_kernel[cur][op] += 1
        for key in ("UBLKCP", "STG.E.ENL2.256", "DMMA"):
            if op.startswith(key) and (key, cur) not in excerpt and len([k for k in excerpt if k[0] == key]) < 2:
                excerpt[(key, cur)] = line.rstrip()
p.wait()

total = collections.Counter()
for c in per_kernel.values():
    total.update(c)


def family(prefix):
    n = sum(v for k, v in total.items() if k.startswith(prefix))
    ks = sum(1 for c in per_kernel.values() if any(k.startswith(prefix) for k in c))
    return n, ks


base = collections.Counter()
for k, v in total.items():
    base[k.split(".")[0]] += v
with open(out_path, "w") as f:
    f.write(f"# SASS census of pdmpflux.jl_b200/lib/libpdmpflux_cuda.so (cuobjdump -sass, sm_100a), round 2; tools/sass_census.py\n")
    f.write(f"# {len(per_kernel)} device functions, {sum(total.values())} instructions\n")
    f.write("# Blackwell / tensor-path mnemonics (count over the whole library, number of kernels that contain them):\n")
    for key in ("DMMA", "UBLKCP", "SYNCS", "STG.E.ENL2.256", "UTMA", "UTCMMA", "HMMA", "VOTE", "MATCH"):
        n, ks = family(key)
        f.write(f"{key:18s} {n:8d} instructions in {ks:4d} kernels\n")
    f.write("# (no UTMALDG / UTCMMA: FP64 has no tcgen05 kind and the history rows are 1-D contiguous -> bulk copies UBLKCP, not tensor maps)\n")
    f.write("# top opcodes: " + ", ".join(f"{k} {v}" for k, v in base.most_common(14)) + "\n")
    f.write("# kernels with DMMA:\n")
    for k, c in sorted(per_kernel.items(), key=lambda kv: -sum(v for o, v in kv[1].items() if o.startswith("DMMA")))[:6]:
        n = sum(v for o, v in c.items() if o.startswith("DMMA"))
        if n:
            f.write(f"  {n:6d}  {k}\n")
    f.write("# the BASELINE kernels: instructions, DFMA+DADD+DMUL, SHFL, VOTE, MATCH, UBLKCP, STG.256\n")
    want = {"C1 <1,zigzag,gauss_std,grid>": "skeleton_kernelILi1ELi0ELi0ELi2ELi0E", "C2 <8,zigzag,banana,brent,-1>": "skeleton_kernelILi8ELi0ELi3ELi1ELin1E",
            "C2 thread per chain <1,zigzag,banana,brent>": "skeleton_kernelILi1ELi0ELi3ELi1ELi0E", "C3 <8,bps,equicorr,grid>": "skeleton_kernelILi8ELi1ELi2ELi2ELi0E",
            "C5f <32,fecmc,gauss_std,grid>": "skeleton_kernelILi32ELi2ELi0ELi2ELi0E", "C5b <32,boomerang,gauss_std,grid>": "skeleton_kernelILi32ELi3ELi0ELi2ELi0E",
            "C4 logreg <3,split>": "logreg_zigzag_kernelILi3ELb1E"}
    for label, frag in want.items():
        for k, c in per_kernel.items():
            if frag in k:
                g = lambda pre: sum(v for o, v in c.items() if o.startswith(pre))
                f.write(f"  {label:46s} {sum(c.values()):6d}  fp64 {g('DFMA') + g('DADD') + g('DMUL'):5d}  SHFL {g('SHFL'):4d}  VOTE {g('VOTE'):3d}  "
                        f"MATCH {g('MATCH'):3d}  UBLKCP {g('UBLKCP'):2d}  STG.256 {g('STG.E.ENL2.256'):3d}  DMMA {g('DMMA'):4d}\n")
    f.write("# excerpts: the TMA row store, the 256-bit store and the FP64 MMA as they appear in the kernels\n")
    for (key, k), line in excerpt.items():
        f.write(f"{line}    <- {k[:70]}\n")
print(open(out_path).read())
