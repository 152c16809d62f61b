"""Scratch: run one config once (for ncu).  usage: one.py cfg team chains n_ev"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pdmpflux_b200 as p
from bench import make_sampler, CONFIGS
cfg, team, nch, n_ev = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
if team: os.environ["PDMPFLUX_TEAM"] = str(team)
d = CONFIGS[cfg]["d"]; dev = torch.device("cuda")
s = make_sampler(p, cfg)
x0 = torch.full((nch, d), CONFIGS[cfg]["x0"], dtype=torch.float64, device=dev)
v0 = torch.ones((nch, d), dtype=torch.float64, device=dev) / (d ** 0.5 if CONFIGS[cfg]["unit_v"] else 1.0)
ch = p.DeviceChains(s, x0, v0, seed=2024)
f64 = torch.float64
bufs = dict(X=torch.empty((nch, n_ev, d), dtype=f64, device=dev), V=torch.empty((nch, n_ev, d), dtype=f64, device=dev),
            t=torch.empty((nch, n_ev), dtype=f64, device=dev), horizon=torch.empty((nch, n_ev), dtype=f64, device=dev),
            ar=torch.empty((nch, n_ev), dtype=f64, device=dev), error_value_ar=torch.empty((nch, n_ev, 5), dtype=f64, device=dev),
            errored_bound=torch.empty((nch, n_ev), dtype=torch.int32, device=dev), rejected=torch.empty((nch, n_ev), dtype=torch.int32, device=dev),
            hitting_horizon=torch.empty((nch, n_ev), dtype=torch.int32, device=dev))
view = p.device_history_view(n_ev, **bufs)
st = torch.cuda.current_stream().cuda_stream
for i in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ch.advance(n_ev, view, 0, st); e1.record(); torch.cuda.synchronize()
    print(f"{cfg} team={team} chains={nch} n_ev={n_ev}: {e0.elapsed_time(e1):.2f} ms  {nch*n_ev/e0.elapsed_time(e1)/1e3:.2f} Mev/s")
ch.status()
_, _, cnt = ch.status()
import numpy as np
tot = cnt.sum(axis=1)
g = tot[: nch // 4 * 4].reshape(-1, 4)
print("per chain (3 launches): builds mean %.2f rates mean %.2f total mean %.2f max %d; max-of-4 mean %.2f; max-of-4 over CTA max %d" % (
    cnt[:, 0].mean(), cnt[:, 1].mean(), tot.mean(), tot.max(), g.max(axis=1).mean(), g.max(axis=1).max()))
