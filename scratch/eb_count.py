import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pdmpflux_b200 as p
from bench import make_sampler, CONFIGS
for cfg, nch, n in (("c2", 4096, 1001), ("c1", 4096, 1001), ("c3", 2048, 301)):
    d = CONFIGS[cfg]["d"]
    s = make_sampler(p, cfg)
    x0 = np.full((nch, d), CONFIGS[cfg]["x0"]); v0 = np.ones((nch, d)) / (d ** 0.5 if CONFIGS[cfg]["unit_v"] else 1.0)
    hb = p.sample_skeleton(s, n, x0, v0, seed=2024)
    print(cfg, "events", nch * (n - 1), "errored_bound nonzero", np.count_nonzero(hb.errored_bound), "eva nonzero rows", np.count_nonzero(np.abs(hb.error_value_ar).sum(axis=2)),
          "rejected mean", hb.rejected.mean(), "hitting_horizon mean", hb.hitting_horizon.mean())
