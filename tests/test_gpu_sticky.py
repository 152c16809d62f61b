"""Sticky Zig-Zag on the device against the literal numpy restatement of src/StickySamplingLoop.jl
(oracle/pdmp_oracle_np.py: StickyChain): free-running parity with injected draws (exact freeze / thaw sequences),
teacher-forced one-step parity from arbitrary frozen states, the sticky sample_from_skeleton, and the stationary law."""
import math
import os

import numpy as np
import pytest

import pdmp_oracle_np as onp
from conftest import record_parity_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def p():
    import ctypes

    import pdmpflux_b200
    n = ctypes.c_int(0)
    pdmpflux_b200.lib().pdmpflux_device_count(ctypes.byref(n))
    assert n.value > 0, "GPU tests need a CUDA device"
    return pdmpflux_b200


def relerr(a, b):
    scale = max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


CASES = [
    # name, d, oracle potential, device potential, kappa, config kwargs, n_sk
    ("sticky_gauss3", 3, lambda: onp.GaussStd(), lambda p: p.GaussStd(), 1.5, dict(), 1200),
    ("sticky_diag7_unsigned", 7, lambda: onp.GaussDiag(np.linspace(0.5, 2.0, 7)), lambda p: p.GaussDiag(np.linspace(0.5, 2.0, 7)),
     0.7, dict(signed_bound=False, grid_size=6), 1000),
    ("sticky_gauss5_nonadaptive", 5, lambda: onp.GaussStd(), lambda p: p.GaussStd(), 2.0, dict(adaptive=False, tmax=0.4), 800),
    ("sticky_banana4_short", 4, lambda: onp.Banana(), lambda p: p.Banana(), 1.0, dict(), 150),
    ("sticky_equicorr40", 40, lambda: onp.GaussEquicorr(40, 0.5), lambda p: p.GaussEquicorr(0.5), 0.3, dict(grid_size=5), 500),
]


def _oracle(case, n_chains=1):
    name, d, opot, _, kappa, kw, n_sk = case
    import zlib
    g = np.random.default_rng(zlib.crc32(name.encode()))
    x0 = 0.7 * g.standard_normal((n_chains, d))
    v0 = np.where(g.random((n_chains, d)) < 0.5, -1.0, 1.0)
    E = g.standard_exponential((n_chains, 30 * n_sk)); U = g.random((n_chains, 12 * n_sk))
    hs = []
    for c in range(n_chains):
        s = onp.Sampler(d, opot(), onp.Config(sampler=onp.ZIGZAG, **kw).normalised(d))
        hs.append(onp.sample_skeleton_sticky(s, np.full(d, kappa), n_sk, x0[c], v0[c], onp.Tape(E[c], U[c], np.zeros(1))))
    return x0, v0, E, U, hs


@pytest.mark.parametrize("team", [1, 8, 32])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_sticky_free_running_parity(p, case, team):
    name, d, _, dpot, kappa, kw, n_sk = case
    if team == 1 and d > 16:
        pytest.skip("thread-per-chain is for small d")
    x0, v0, E, U, hs = _oracle(case, n_chains=2)
    s = p.StickyZigZag(d, dpot(p), np.full(d, kappa), AD_backend="ForwardDiff", **kw)
    os.environ["PDMPFLUX_TEAM"] = str(team)
    try:
        h = p.sample_skeleton(s, n_sk, x0, v0, tape=(E, U, np.zeros((2, 1))))
    finally:
        os.environ.pop("PDMPFLUX_TEAM")
    worst = 0.0
    for c, r in enumerate(hs):
        act = h.is_active[c].T.astype(bool)
        assert np.array_equal(act, r.is_active), "freeze / thaw sequence differs"
        assert np.array_equal(h.V[c].T, r.V)                      # velocities only flip sign: bit-exact
        e = max(relerr(h.X[c].T, r.X), relerr(h.t[c], r.t), relerr(h.horizon[c], r.horizon), relerr(h.ar[c], r.ar))
        worst = max(worst, e)
        assert np.array_equal(h.rejected[c], r.rejected) and np.array_equal(h.hitting_horizon[c], r.hitting_horizon)
        assert np.array_equal(h.errored_bound[c], r.errored_bound)
        assert list(h.tape_pos[c, :2]) == list(r.tape_pos[:2])
        assert (~r.is_active).any() and (r.is_active[:, 1:] != r.is_active[:, :-1]).any()
    record_parity_error(f"sticky_free_running/{name}/team{team}", max_rel=worst, events=n_sk - 1, tol=1e-10)
    assert worst < 1e-10, worst


def test_sticky_sample_from_skeleton_and_law(p):
    """sample_from_skeleton(::StickyPDMP, ...) (src/sample.jl:516-561): frozen coordinates do not move; and the fraction
    of time a coordinate of a standard Gaussian slab spends frozen follows the sticky Zig-Zag's stationary law up to the
    reference's thaw-clock quirk (checked against the oracle's own number on the same Philox-free tape)."""
    d, kappa, n_sk = 2, 0.8, 3000
    g = np.random.default_rng(5)
    x0 = g.standard_normal(d); v0 = np.array([1.0, -1.0])
    E = g.standard_exponential((1, 40 * n_sk)); U = g.random((1, 12 * n_sk))
    so = onp.Sampler(d, onp.GaussStd(), onp.Config(sampler=onp.ZIGZAG).normalised(d))
    r = onp.sample_skeleton_sticky(so, np.full(d, kappa), n_sk, x0, v0, onp.Tape(E[0], U[0], np.zeros(1)))
    s = p.StickyZigZagAD(d, p.GaussStd(), np.full(d, kappa))
    h = p.sample_skeleton(s, n_sk, x0, v0, tape=(E, U, np.zeros((1, 1))))
    assert np.array_equal(h.is_active, r.is_active) and relerr(h.X, r.X) < 1e-10 and relerr(h.t, r.t) < 1e-10
    N = 5000
    out = p.sample_from_skeleton(s, N, h)
    # independent numpy evaluation with the masked velocity
    dt = h.t[-1] / N
    tm = dt * np.arange(1, N + 1)
    idx = np.searchsorted(h.t, tm, side="right") - 1
    vu = np.where(h.is_active, h.V, 0.0)
    ref = h.X[:, idx] + vu[:, idx] * (tm - h.t[idx])
    assert out.shape == (d, N) and np.allclose(out, ref, rtol=1e-12, atol=1e-13)
    frozen_frac = float(np.mean(out == 0.0))
    seg = np.diff(r.t)
    oracle_frac = float((seg[None, :] * (~r.is_active[:, :-1])).sum() / (d * seg.sum()))
    assert abs(frozen_frac - oracle_frac) < 0.02 and 0.05 < frozen_frac < 0.6
    with pytest.raises(p.UnsupportedError):
        p.sample_skeleton(s, 3.0, x0, v0, seed=1)       # the time-horizon variant is not offered for sticky samplers
    with pytest.raises(p.DimensionMismatch):
        p.StickyZigZag(3, p.GaussStd(), np.ones(2))
    # Philox mode: deterministic, sharding-independent, finite
    xb = g.standard_normal((64, d)); vb = np.where(g.random((64, d)) < 0.5, -1.0, 1.0)
    a = p.sample_skeleton(s, 300, xb, vb, seed=9)
    b = p.sample_skeleton(s, 300, xb[10:20], vb[10:20], seed=9, chain_offset=10)
    assert np.array_equal(a.X[10:20], b.X) and np.array_equal(a.is_active[10:20], b.is_active)
    assert np.isfinite(a.X).all() and (np.diff(a.t, axis=1) > 0).all() and (a.is_active == 0).any()
