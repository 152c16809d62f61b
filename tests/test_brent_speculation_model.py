"""Host model of the speculative Brent search of the team kernels (chain.cuh: build_bound_brent, NW = -1) against the
oracle's one-at-a-time recurrence (oracle/pdmp_oracle_np.py: brent_minimum, the restatement of Optim's Brent used by
upper_bound_constant, src/UpperBound.jl:18-36).

The kernel predicts the abscissae of the next 40 iterations from the bracket alone (x <- x + g (E - x)), evaluates the
rate at all of them, validates every iteration side by side and commits the prefix that is as assumed; the rest is
replayed one iteration at a time.  This file restates that control flow in plain Python floats (same operations in the
same order, no fused multiply-add) and checks that it returns the same minimum after the same number of iterations as
the serial recurrence -- on increasing rates (one pass), decreasing rates (the search walks left, ~70 iterations) and
non-monotone ones (direction changes, failed golden steps).  CPU only; the CUDA transcription is held to the oracle
by tests/test_gpu_parity.py (cases *_brent, zz_saddle20_brent)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pdmp_oracle_np as o  # noqa: E402

SQRT_EPS, EPS = o.SQRT_EPS, o.EPS
G = 0.5 * (3.0 - np.sqrt(5.0))
TEAM, S = 8, 5
N = TEAM * S


def golden(lo, hi, x):
    tol = SQRT_EPS * abs(x) + EPS
    mid = (hi + lo) / 2
    og = (hi - x) if x < mid else (lo - x)
    gstp = G * og
    u = x + (gstp if abs(gstp) >= tol else (tol if gstp > 0 else -tol))
    return u, tol, mid, og, gstp


def parabola(st, tol):
    lo, hi, x, w, v, fx, fw, fv, stp, old = st
    xw, xv = x - w, x - v
    r = xw * (fx - fv)
    q = xv * (fx - fw)
    pp = xv * q - xw * r
    q = 2 * (q - r)
    if q > 0:
        pp = -pp
    q = abs(q)
    return (abs(old) > tol) and (abs(pp) < abs(q * old / 2)) and (pp < q * (hi - x)) and (pp < q * (x - lo)), pp, q


def serial_step(st, f):
    """one iteration of the recurrence; (state, stopped)"""
    lo, hi, x, w, v, fx, fw, fv, stp, old = st
    u, tol, mid, og, gstp = golden(lo, hi, x)
    tol2 = 2 * tol
    if abs(x - mid) <= tol2 - (hi - lo) / 2:
        return st, True
    para, pp, q = parabola(st, tol)
    new_old, new_stp = og, gstp
    if para:
        sp = pp / q
        xt = x + sp
        if (xt - lo) < tol2 or (hi - xt) < tol2:
            sp = tol if x < mid else -tol
        new_old, new_stp = stp, sp
        u = x + (sp if abs(sp) >= tol else (tol if sp > 0 else -tol))
    old, stp = new_old, new_stp
    fu = f(u)
    if fu < fx:
        if u < x:
            hi = x
        else:
            lo = x
        v, fv, w, fw, x, fx = w, fw, x, fx, u, fu
    else:
        if u < x:
            lo = u
        else:
            hi = u
        if fu <= fw or w == x:
            v, fv, w, fw = w, fw, u, fu
        elif fu <= fv or v == x or v == w:
            v, fv = u, fu
    return (lo, hi, x, w, v, fx, fw, fv, stp, old), False


def start(f, h):
    x = 0.0 + G * (h - 0.0)
    fx = f(x)
    return (0.0, h, x, x, x, fx, fx, fx, 0.0, 0.0)


def serial(f, h):
    st, n = start(f, h), 0
    while n < 1000:
        st, stopped = serial_step(st, f)
        if stopped:
            break
        n += 1
    return st[5], n


def speculative(f, h, stats):
    st, it, done = start(f, h), 0, False
    while it < 1000 and not done:
        lo, hi, x, w, v, fx, fw, fv, stp, old = st
        right = x < (hi + lo) / 2
        E = hi if right else lo
        X = [x]
        for _ in range(N):
            X.append(X[-1] + G * (E - X[-1]))           # the three-operation recurrence every lane runs
        W = [w] + X[:-1]
        V = [v, w] + X[:-2]
        F, stop, as_assumed, TOL, U = [0.0] * N, [False] * N, [False] * N, [0.0] * N, [0.0] * N
        for k in range(N):                              # "lane k % 8, block k // 8": its own iteration in full
            moving = (lo if right else hi) if k == 0 else W[k]
            lo_, hi_ = (moving, hi) if right else (lo, moving)
            U[k], TOL[k], mid, og, gstp = golden(lo_, hi_, X[k])
            stop[k] = abs(X[k] - mid) <= 2 * TOL[k] - (hi_ - lo_) / 2
            as_assumed[k] = ((X[k] < mid) == right) and abs(gstp) >= TOL[k]
            F[k] = f(U[k])
        good = [False] * N
        for k in range(N):
            fxk = F[k - 1] if k >= 1 else fx
            fwk = F[k - 2] if k >= 2 else (fx if k == 1 else fw)
            fvk = F[k - 3] if k >= 3 else (fx if k == 2 else (fw if k == 1 else fv))
            moving = (lo if right else hi) if k == 0 else W[k]
            lo_, hi_ = (moving, hi) if right else (lo, moving)
            oldk = old if k == 0 else E - W[k]
            para, _, _ = parabola((lo_, hi_, X[k], W[k], V[k], fxk, fwk, fvk, 0.0, oldk), TOL[k])
            good[k] = (not para) and F[k] < fxk and as_assumed[k]
        nterm = next((k for k in range(N) if stop[k]), N)
        nbad = next((k for k in range(N) if not good[k]), N)
        nc = min(nterm, nbad)
        it += nc
        stats["passes"] += 1
        if nterm <= nbad and nterm < N:
            if nterm >= 1:
                fx = F[nterm - 1]
            st, done = (lo, hi, x, w, v, fx, fw, fv, stp, old), True
        else:
            if nc >= 1:
                xn = X[nc] if nc < N else U[N - 1]
                wn = X[nc - 1]
                vn = X[nc - 2] if nc >= 2 else w
                f1 = F[nc - 1]
                f2 = F[nc - 2] if nc >= 2 else fx
                f3 = F[nc - 3] if nc >= 3 else (fx if nc == 2 else fw)
                lo2, hi2 = (wn, hi) if right else (lo, wn)
                o2 = E - wn
                st = (lo2, hi2, xn, wn, vn, f1, f2, f3, G * o2, o2)
            if nc < N:
                for _ in range(TEAM):
                    st, stopped = serial_step(st, f)
                    if stopped:
                        done = True
                        break
                    it += 1
                    stats["replayed"] += 1
    return st[5], it


def test_serial_model_is_the_oracles_brent():
    g = np.random.default_rng(0)
    s = o.Sampler(20, o.GaussDiag(np.linspace(0.5, 2.0, 20)), o.Config(sampler=o.ZIGZAG, grid_size=0))
    for _ in range(20):
        x0, v0, h = g.standard_normal(20), np.where(g.random(20) < 0.5, -1.0, 1.0), float(g.uniform(0.05, 2.0))
        f = lambda t: -s.rate(x0, v0, t)  # noqa: E731
        assert serial(f, h)[0] == o.brent_minimum(f, 0.0, h)


@pytest.mark.parametrize("name", ["banana50", "gauss_increasing", "negative_curvature", "saddle"])
def test_speculative_search_equals_serial_recurrence(name):
    d = {"banana50": 50, "gauss_increasing": 33, "negative_curvature": 20, "saddle": 20}[name]
    pot = {"banana50": lambda: o.Banana(), "gauss_increasing": lambda: o.GaussDiag(np.linspace(0.5, 2.0, d)),
           "negative_curvature": lambda: o.GaussDiag(-np.linspace(0.5, 2.0, d)),
           "saddle": lambda: o.GaussDiag(np.where(np.arange(d) % 2 == 0, 1.0, -1.0) * np.linspace(0.5, 2.0, d))}[name]()
    s = o.Sampler(d, pot, o.Config(sampler=o.ZIGZAG, grid_size=0))
    g = np.random.default_rng(7)
    stats = {"passes": 0, "replayed": 0}
    n_bounds = 60
    for _ in range(n_bounds):
        x0, v0, h = g.standard_normal(d), np.where(g.random(d) < 0.5, -1.0, 1.0), float(g.uniform(0.05, 2.0))
        f = lambda t: -s.rate(x0, v0, t)  # noqa: E731
        a, na = serial(f, h)
        b, nb = speculative(f, h, stats)
        assert a == b and na == nb, (name, a, b, na, nb)
    if name in ("banana50", "gauss_increasing"):   # increasing rates: one pass and the +-tol step before the stopping rule
        assert stats["passes"] == n_bounds and stats["replayed"] <= 2 * n_bounds
    else:                                          # the replay branches are actually exercised
        assert stats["replayed"] > n_bounds
