"""Scratch: tiny runs of the newer kernels for compute-sanitizer (memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pdmpflux_b200 as p
from oracle_cases import logreg_data
g = np.random.default_rng(0)
for (n, d, nch, G) in ((70, 13, 5, 4), (45, 100, 3, 10), (33, 5, 700, 3)):
    X, y, s0 = logreg_data(n, d)
    s = p.ZigZagAD(d, p.LogReg(X, y, s0), grid_size=G)
    h = p.sample_skeleton(s, 6, 0.2 * g.standard_normal((nch, d)), np.where(g.random((nch, d)) < 0.5, -1.0, 1.0), seed=1)
    print("logreg", n, d, nch, h.t[:, -1].mean(), p.RV_diagnostic(h, s.potential, B=3)[:2])
os.environ["PDMPFLUX_VBITS"] = "1"; os.environ["PDMPFLUX_SLAB_BYTES"] = str(1 << 14); os.environ["PDMPFLUX_HOST_THREADS"] = "2"
s = p.ZigZagAD(7, p.Banana(), grid_size=2, tmax=3.0, adaptive=False)
h = p.sample_skeleton(s, 80, g.standard_normal((9, 7)), np.where(g.random((9, 7)) < 0.5, -1.0, 1.0), seed=2)
print("vbits", h.t[:, -1].mean(), int(h.errored_bound.sum()))
print("rv", p.RV_diagnostic(h, p.Banana(), B=5)[:3])
