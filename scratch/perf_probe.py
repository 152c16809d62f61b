"""Scratch: device-resident throughput of the named configs (not a bench line)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pdmpflux_b200 as p

def run(name, mk, d, nch, n_ev, v_unit, reps=3, full=True, team=None):
    if team: os.environ["PDMPFLUX_TEAM"] = str(team)
    else: os.environ.pop("PDMPFLUX_TEAM", None)
    s = mk()
    dev = torch.device("cuda")
    x0 = torch.zeros((nch, d), dtype=torch.float64, device=dev)
    if "banana" in name: x0 += 1.0
    v0 = torch.ones((nch, d), dtype=torch.float64, device=dev)
    if v_unit: v0 /= d ** 0.5
    ch = p.DeviceChains(s, x0, v0, seed=2024)
    bufs = dict(X=torch.empty((nch, n_ev, d), dtype=torch.float64, device=dev), V=torch.empty((nch, n_ev, d), dtype=torch.float64, device=dev),
                t=torch.empty((nch, n_ev), dtype=torch.float64, device=dev))
    if full:
        bufs.update(horizon=torch.empty((nch, n_ev), dtype=torch.float64, device=dev), ar=torch.empty((nch, n_ev), dtype=torch.float64, device=dev),
                    error_value_ar=torch.empty((nch, n_ev, 5), dtype=torch.float64, device=dev),
                    errored_bound=torch.empty((nch, n_ev), dtype=torch.int32, device=dev), rejected=torch.empty((nch, n_ev), dtype=torch.int32, device=dev),
                    hitting_horizon=torch.empty((nch, n_ev), dtype=torch.int32, device=dev))
    view = p.device_history_view(n_ev, **bufs)
    st = torch.cuda.current_stream().cuda_stream
    ch.advance(n_ev, view, 0, st)  # warm-up / burn-in
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ch.advance(n_ev, view, 0, st); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    stt, pos, cnt = ch.status()
    ev = nch * n_ev
    bpe = 16 * d + 76 if full else 16 * d + 8
    print(f"{name:34s} team={team or 'auto':>4} chains={nch:6d} n_ev={n_ev:5d} {best*1e3:9.2f} ms  {ev/best/1e6:9.2f} Mev/s  {ev*bpe/best/1e9:8.1f} GB/s"
          f"  builds/ev={cnt[:,0].sum()/ (ev*(reps+1)):.2f} rates/ev={cnt[:,1].sum()/(ev*(reps+1)):.2f}", flush=True)
    ch.close()

which = sys.argv[1:] or ["c1", "c2", "c3", "c5"]
if "c1" in which:
    for team in (1, 8):
        run("C1 zigzagAD gauss d=10 G=10", lambda: p.ZigZagAD(10, p.GaussStd()), 10, 65536, 500, False, team=team)
    run("C1 zigzagAD gauss d=10 G=10", lambda: p.ZigZagAD(10, p.GaussStd()), 10, 4096, 2000, False, team=8)
if "c2" in which:
    for nch, team in ((4096, 32), (4096, 8), (65536, 8), (65536, 32), (65536, 1)):
        run("C2 zigzag banana d=50 brent", lambda: p.ZigZag(50, p.Banana(), grid_size=0), 50, nch, 200 if nch > 10000 else 1000, False, team=team)
if "c3" in which:
    for team in (8, 32):
        run("C3 bps equicorr d=100 G=10", lambda: p.BPS(100, p.GaussEquicorr(0.9), refresh_rate=0.1), 100, 16384, 300, True, team=team)
if "c5" in which:
    run("C5 fecmc gauss d=1000", lambda: p.ForwardECMC(1000, p.GaussStd()), 1000, 8192, 40, True, team=32)
    run("C5 boomerang gauss d=1000", lambda: p.Boomerang(1000, p.GaussStd()), 1000, 8192, 40, True, team=32)
