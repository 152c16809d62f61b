"""Speed-Up Zig-Zag on the device against the literal numpy restatement of
src/Samplers/SpeedUpZigZagSamplers.jl (oracle/pdmp_oracle_np.py: SpeedUpSampler; its d/dt is taken by complex-step
differentiation of the reference's own flow expressions, independently of the device's hand-derived formulas)."""
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
    # name, d, oracle potential, device potential, config kwargs, n_sk
    ("suzz_gauss4", 4, lambda: onp.GaussStd(), lambda p: p.GaussStd(), dict(), 600),
    ("suzz_diag9_unsigned", 9, lambda: onp.GaussDiag(np.linspace(0.5, 2.0, 9)), lambda p: p.GaussDiag(np.linspace(0.5, 2.0, 9)),
     dict(signed_bound=False, grid_size=6), 500),
    ("suzz_banana5", 5, lambda: onp.Banana(), lambda p: p.Banana(), dict(grid_size=8), 400),
    ("suzz_equicorr33_scalar", 33, lambda: onp.GaussEquicorr(33, 0.4), lambda p: p.GaussEquicorr(0.4),
     dict(vectorized_bound=False, adaptive=False, tmax=0.2), 300),
    ("suzz_gauss6_brent", 6, lambda: onp.GaussStd(), lambda p: p.GaussStd(), dict(grid_size=0), 300),
]


@pytest.mark.parametrize("team", [1, 8, 32])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_speedup_one_step_parity(p, case, team):
    """Teacher-forced: every event of the oracle's skeleton is regenerated on the GPU from the previous state."""
    name, d, opot, dpot, kw, n_sk = case
    if team == 1 and d > 16:
        pytest.skip("thread-per-chain is for small d")
    import zlib
    g = np.random.default_rng(zlib.crc32(name.encode()))
    x0 = 0.6 * g.standard_normal(d); v0 = np.where(g.random(d) < 0.5, -1.0, 1.0)
    E = g.standard_exponential(12 * n_sk); U = g.random(6 * n_sk)
    so = onp.SpeedUpSampler(d, opot(), onp.Config(sampler=onp.ZIGZAG, **kw))
    tape = onp.Tape(E, U, np.zeros(1))
    ch = onp.Chain(so, x0, v0, tape)
    states, cursors = [(x0.copy(), v0.copy(), 0.0, ch.state.horizon)], [list(tape.pos)]
    recs = []
    for k in range(1, n_sk):
        st = ch.get_event_state()
        states.append((st.x.copy(), st.v.copy(), float(st.t), float(st.horizon)))
        cursors.append(list(tape.pos))
        recs.append((float(st.ar), int(st.rejected), int(st.hitting_horizon), int(st.errored_bound)))
    pos = np.array(cursors)
    use = np.diff(pos, axis=0)
    wE, wU = int(use[:, 0].max()) + 1, int(use[:, 1].max()) + 1
    idx = lambda start, w: start[:, None] + np.arange(w)[None, :]
    pad = lambda a, w: np.concatenate([a, np.ones(w)])
    tE, tU = pad(E, wE)[idx(pos[:-1, 0], wE)], pad(U, wU)[idx(pos[:-1, 1], wU)]
    X = np.array([s[0] for s in states]); V = np.array([s[1] for s in states])
    T = np.array([s[2] for s in states]); H = np.array([s[3] for s in states])
    s = p.SpeedUpZigZag(d, dpot(p), AD_backend="ForwardDiff", **kw)
    os.environ["PDMPFLUX_TEAM"] = str(team)
    try:
        h = p.sample_skeleton(s, 2, X[:-1], V[:-1], tape=(tE, tU, np.zeros((n_sk - 1, 1))), t0=T[:-1], horizon0=H[:-1], batch=True)
    finally:
        os.environ.pop("PDMPFLUX_TEAM")
    ex = (np.abs(h.X[:, 1] - X[1:]).max(axis=1) / np.maximum(np.abs(X[1:]).max(axis=1), 1e-300)).max()
    dt_o = T[1:] - T[:-1]
    et = (np.abs((h.t[:, 1] - h.t[:, 0]) - dt_o) / dt_o).max()
    eh = (np.abs(h.horizon[:, 1] - H[1:]) / H[1:]).max()
    ea = np.abs(h.ar[:, 1] - np.array([r[0] for r in recs])).max()
    record_parity_error(f"speedup_one_step/{name}/team{team}", x=ex, t=et, horizon=eh, ar=ea, tol=1e-10)
    assert max(ex, et, eh, ea) < 1e-10, dict(x=ex, t=et, horizon=eh, ar=ea)
    assert np.array_equal(h.V[:, 1], V[1:])
    assert np.array_equal(h.rejected[:, 1], [r[1] for r in recs])
    assert np.array_equal(h.hitting_horizon[:, 1], [r[2] for r in recs])
    assert np.array_equal(h.errored_bound[:, 1], [r[3] for r in recs])
    assert np.array_equal(h.tape_pos[:, :2], use[:, :2])


def test_speedup_sample_from_skeleton_and_free_running(p):
    d, n_sk = 3, 500
    g = np.random.default_rng(8)
    x0 = g.standard_normal(d); v0 = np.array([1.0, -1.0, 1.0])
    E = g.standard_exponential((1, 12 * n_sk)); U = g.random((1, 6 * n_sk))
    so = onp.SpeedUpSampler(d, onp.GaussStd(), onp.Config(sampler=onp.ZIGZAG))
    r = onp.sample_skeleton(so, n_sk, x0, v0, onp.Tape(E[0], U[0], np.zeros(1)))
    s = p.SpeedUpZigZagAD(d, p.GaussStd())
    h = p.sample_skeleton(s, n_sk, x0, v0, tape=(E, U, np.zeros((1, 1))))
    e = max(relerr(h.X, r.X), relerr(h.t, r.t))
    record_parity_error("speedup_free_running/gauss3", max_rel=e, events=n_sk - 1, tol=1e-9)
    assert e < 1e-9 and np.array_equal(h.V, r.V)
    # sample_from_skeleton interpolates with the sampler's own (nonlinear) flow (src/sample.jl:475-513 uses sampler.flow)
    N = 777
    out = p.sample_from_skeleton(s, N, h)
    dt = h.t[-1] / N
    ref = np.empty((d, N))
    for j in range(N):
        tm = (j + 1) * dt
        i = np.searchsorted(h.t, tm, side="right") - 1
        ref[:, j] = so.flow(h.X[:, i], h.V[:, i], tm - h.t[i])[0]
    assert out.shape == (d, N) and np.allclose(out, ref, rtol=1e-11, atol=1e-12)
    # Philox mode: finite, deterministic across sharding
    xb = g.standard_normal((40, d)); vb = np.where(g.random((40, d)) < 0.5, -1.0, 1.0)
    a = p.sample_skeleton(s, 200, xb, vb, seed=4)
    b = p.sample_skeleton(s, 200, xb[5:9], vb[5:9], seed=4, chain_offset=5)
    assert np.isfinite(a.X).all() and np.array_equal(a.X[5:9], b.X) and (np.diff(a.t, axis=1) > 0).all()
