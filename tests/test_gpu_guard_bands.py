"""Own bounds check of the HBM write path (record!, Composites.jl:239-260): every PDMPHistory column lives inside a
larger canary-filled allocation; after record + two advance() calls the canaries in front of, behind and -- for the
columns no call was asked to write -- inside each slab must be intact, and every requested entry must have been written.
Covers the 256-bit / TMA bulk / scalar store variants (aligned and 8-byte-offset buffers, even and odd d, odd first
columns) of every kernel family and team width."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAD = 96                      # canary elements on each side
N_COLS, WRITTEN = 15, 13      # columns 13, 14 of every chain are never written
F64_CANARY = 0x7FF8DEAD0000BEEF   # a NaN payload no kernel produces
I32_CANARY = 0x5A5A5A5A
U8_CANARY = 0x5A


@pytest.fixture(scope="module")
def p():
    import ctypes

    import pdmpflux_b200
    n = ctypes.c_int(0)
    pdmpflux_b200.lib().pdmpflux_device_count(ctypes.byref(n))
    assert n.value > 0, "GPU tests need a CUDA device"
    return pdmpflux_b200


class Guarded:
    """A (n_chains, N_COLS, per) buffer carved out of a canary-filled allocation at `offset` elements past a 256-byte
    boundary."""

    def __init__(self, torch, nch, per, kind, offset):
        raw, canary = {"f64": (torch.int64, F64_CANARY), "i32": (torch.int32, I32_CANARY), "u8": (torch.uint8, U8_CANARY)}[kind]
        n = nch * N_COLS * per
        self.canary = canary
        self.big = torch.full((PAD + offset + n + PAD,), canary, dtype=raw, device="cuda")
        self.lo, self.n = PAD + offset, n
        self.raw = self.big[self.lo:self.lo + n]
        t = self.raw.view(torch.float64) if kind == "f64" else self.raw
        self.tensor = t.view(nch, N_COLS, per) if per > 1 else t.view(nch, N_COLS)
        self.shape3 = (nch, N_COLS, per)

    def check(self, name):
        outside_lo, outside_hi = self.big[:self.lo], self.big[self.lo + self.n:]
        assert bool((outside_lo == self.canary).all()), f"{name}: write in front of the buffer"
        assert bool((outside_hi == self.canary).all()), f"{name}: write behind the buffer"
        r = self.raw.view(self.shape3)
        assert bool((r[:, WRITTEN:] == self.canary).all()), f"{name}: write into a column nobody asked for"
        assert bool((r[:, :WRITTEN] != self.canary).all()), f"{name}: a requested entry was not written"


def _unit(g, nch, d):
    v = g.standard_normal((nch, d))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def _signs(g, nch, d):
    return np.where(g.random((nch, d)) < 0.5, -1.0, 1.0)


def _cases(p):
    g = np.random.default_rng(5)
    Xl = g.standard_normal((70, 5)) / np.sqrt(5); yl = (g.random(70) < 0.5).astype(float)
    return [
        # name, sampler, d, n_chains, velocity maker, team, sticky
        ("zz_brent_banana_t1", lambda: p.ZigZag(50, p.Banana(), grid_size=0), 50, 37, _signs, 1, False),
        ("zz_brent_banana_t4", lambda: p.ZigZag(50, p.Banana(), grid_size=0), 50, 37, _signs, 4, False),
        ("zz_brent_banana_t8", lambda: p.ZigZag(50, p.Banana(), grid_size=0), 50, 37, _signs, 8, False),
        ("zz_brent_odd_d_t8", lambda: p.ZigZag(33, p.GaussDiag(np.linspace(0.5, 2, 33)), grid_size=0), 33, 11, _signs, 8, False),
        ("zz_brent_odd_d_t1", lambda: p.ZigZag(33, p.GaussDiag(np.linspace(0.5, 2, 33)), grid_size=0), 33, 70, _signs, 1, False),
        ("zz_brent_t32", lambda: p.ZigZag(300, p.GaussStd(), grid_size=0), 300, 5, _signs, 32, False),
        ("zz_grid_t1", lambda: p.ZigZagAD(10, p.GaussStd()), 10, 70, _signs, 1, False),
        ("zz_grid_d3_t1", lambda: p.ZigZagAD(3, p.GaussStd()), 3, 9, _signs, 1, False),
        ("zz_grid_t8_equi", lambda: p.ZigZagAD(33, p.GaussEquicorr(0.5)), 33, 13, _signs, 8, False),
        ("zz_generic", lambda: p.ZigZag(6, p.Banana(), grid_size=8), 6, 5, _signs, None, False),   # finite differences: generic path
        ("bps_t8", lambda: p.BPS(100, p.GaussEquicorr(0.9)), 100, 19, _unit, 8, False),
        ("bps_t1", lambda: p.BPS(7, p.GaussStd(), refresh_rate=0.5), 7, 33, _unit, 1, False),
        ("fecmc_t32", lambda: p.ForwardECMC(1000, p.GaussStd()), 1000, 5, _unit, 32, False),
        ("fecmc_t8_global_scratch", lambda: p.ForwardECMC(200, p.GaussStd()), 200, 21, _unit, 8, False),
        ("fecmc_small", lambda: p.ForwardECMC(6, p.Banana(), ran_p=True, mix_p=0.7), 6, 9, _unit, None, False),
        ("boom_fd_t32", lambda: p.Boomerang(1000, p.GaussStd()), 1000, 3, _unit, 32, False),
        ("boom_t8", lambda: p.Boomerang(21, p.GaussDiag(np.linspace(0.5, 2, 21)), AD_backend="ForwardDiff"), 21, 7, _unit, 8, False),
        ("sticky_t8", lambda: p.StickyZigZagAD(7, p.GaussStd(), np.full(7, 0.7)), 7, 9, _signs, 8, True),
        ("sticky_t1", lambda: p.StickyZigZagAD(3, p.GaussStd(), np.full(3, 1.5)), 3, 5, _signs, 1, True),
        ("speedup_t8", lambda: p.SpeedUpZigZagAD(9, p.GaussDiag(np.linspace(0.5, 2, 9))), 9, 6, _signs, 8, False),
        ("logreg", lambda: p.ZigZagAD(5, p.LogReg(Xl, yl, 10.0), grid_size=6), 5, 9, _signs, None, False),
    ]


CASE_NAMES = ["zz_brent_banana_t1", "zz_brent_banana_t4", "zz_brent_banana_t8", "zz_brent_odd_d_t8", "zz_brent_odd_d_t1",
              "zz_brent_t32", "zz_grid_t1", "zz_grid_d3_t1", "zz_grid_t8_equi", "zz_generic", "bps_t8", "bps_t1",
              "fecmc_t32", "fecmc_t8_global_scratch", "fecmc_small", "boom_fd_t32", "boom_t8", "sticky_t8", "sticky_t1",
              "speedup_t8", "logreg"]


@pytest.mark.parametrize("offset", [0, 1], ids=["aligned", "offset8"])
@pytest.mark.parametrize("name", CASE_NAMES)
def test_history_writes_stay_inside_their_buffers(p, name, offset):
    import torch
    case = {c[0]: c for c in _cases(p)}[name]
    _, make, d, nch, vel, team, sticky = case
    g = np.random.default_rng(11)
    x0 = g.standard_normal((nch, d)); v0 = vel(g, nch, d)
    if team:
        os.environ["PDMPFLUX_TEAM"] = str(team)
    try:
        s = make()
        ch = p.DeviceChains(s, x0, v0, seed=9)
    finally:
        os.environ.pop("PDMPFLUX_TEAM", None)
    bufs = {"X": Guarded(torch, nch, d, "f64", offset), "V": Guarded(torch, nch, d, "f64", offset),
            "t": Guarded(torch, nch, 1, "f64", offset), "horizon": Guarded(torch, nch, 1, "f64", offset),
            "ar": Guarded(torch, nch, 1, "f64", offset), "error_value_ar": Guarded(torch, nch, 5, "f64", offset),
            "errored_bound": Guarded(torch, nch, 1, "i32", offset), "rejected": Guarded(torch, nch, 1, "i32", offset),
            "hitting_horizon": Guarded(torch, nch, 1, "i32", offset)}
    if sticky:
        bufs["is_active"] = Guarded(torch, nch, d, "u8", offset)
    view = p.device_history_view(N_COLS, **{k: b.tensor for k, b in bufs.items()})
    ch.record(view, 0)
    ch.advance(5, view, 1)      # columns 1..5: the second call starts at an odd column
    ch.advance(7, view, 6)      # columns 6..12
    torch.cuda.synchronize()
    st, _, _ = ch.status()
    assert (st == 0).all(), st
    for k, b in bufs.items():
        b.check(f"{name}/{k}")
    t = bufs["t"].tensor.cpu().numpy()[:, :WRITTEN]
    assert np.all(np.diff(t, axis=1) >= 0.0) and np.all(t[:, 0] == 0.0)
    ch.close()
