"""Drivers with the reference's names: sample_skeleton, sample_from_skeleton, sample (src/sample.jl:27-58,
253-284, 475-513), plus the batched (many independent chains) forms the GPU path exists for."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _lib
from .history import PDMPHistory, PDMPHistoryBatch
from .samplers import AbstractPDMP


def _ptr(a):
    return None if a is None else a.ctypes.data


def _make_tape(tape, n_chains):
    if tape is None:
        return None, None
    E, U, N = (np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in tape)
    for a in (E, U, N):
        if a.shape[0] != n_chains:
            raise _lib.DimensionMismatch("tape streams must have one row per chain")
    t = _lib.Tape(_ptr(E), _ptr(U), _ptr(N), E.shape[1], U.shape[1], N.shape[1], 0)
    return t, (E, U, N)


def _init_arrays(sampler, xinit, vinit):
    scalar = np.isscalar(xinit) and np.isscalar(vinit)
    if scalar:  # src/sample.jl:289-319: scalar initial values are wrapped into 1-vectors
        if not (math.isfinite(xinit) and math.isfinite(vinit)):
            raise _lib.ArgumentError("Initial position and velocity must be finite numbers")
        xinit, vinit = [float(xinit)], [float(vinit)]
    x = np.ascontiguousarray(xinit, dtype=np.float64)
    v = np.ascontiguousarray(vinit, dtype=np.float64)
    batched = x.ndim == 2
    x2, v2 = np.atleast_2d(x), np.atleast_2d(v)
    if x2.shape[1] != sampler.dim or v2.shape[1] != sampler.dim or x2.shape != v2.shape:
        raise _lib.DimensionMismatch(
            f"xinit and vinit must have the same dimension as pdmp.dim ({sampler.dim}). Current dimensions: "
            f"xinit ({x2.shape[1]}), vinit ({v2.shape[1]})")
    return x2, v2, batched


def sample_skeleton(sampler: AbstractPDMP, n_sk, xinit, vinit, *, seed=None, verbose=True, tape=None,
                    chain_offset=0, batch=None, t0=None, horizon0=None, event0=0):
    """sample_skeleton(sampler, n_sk, xinit, vinit; seed, verbose) (src/sample.jl:253-284).

    xinit/vinit of shape (d,) run one chain and return a `PDMPHistory` (reference semantics); shape (C, d)
    runs C independent chains and returns a `PDMPHistoryBatch` (chain c = Philox stream chain_offset + c).
    `tape=(E, U, N)` injects the random draws (parity testing); otherwise draws are Philox(seed, chain, event).
    `t0`, `horizon0` (per chain) and `event0` resume from a saved state (the last column of an earlier
    history) instead of init_state; column 0 of the result is then that state.
    """
    if isinstance(n_sk, (float, np.floating)):  # Julia dispatches on Float64: the time-horizon method
        return sample_skeleton_until(sampler, float(n_sk), xinit, vinit, seed=seed, verbose=verbose, tape=tape,
                                     chain_offset=chain_offset, batch=batch)
    if not isinstance(n_sk, (int, np.integer)):
        raise TypeError("n_sk must be an integer (number of skeleton points) or a float (time horizon T)")
    if n_sk <= 0:
        raise _lib.ArgumentError(f"n_sk must be positive. Current value: {n_sk}")
    x, v, batched = _init_arrays(sampler, xinit, vinit)
    if batch is not None:
        batched = batch
    n_chains = x.shape[0]
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    hb = PDMPHistoryBatch(n_chains, int(n_sk), sampler.dim, sticky=getattr(sampler, "kappa", None) is not None)
    view = _lib.History(_ptr(hb.X), _ptr(hb.V), _ptr(hb.t), _ptr(hb.horizon), _ptr(hb.ar), _ptr(hb.error_value_ar),
                        _ptr(hb.errored_bound), _ptr(hb.rejected), _ptr(hb.hitting_horizon), _ptr(hb.status),
                        _ptr(hb.tape_pos), _ptr(hb.counters), int(n_sk), 0, _ptr(hb.is_active))
    t, keep = _make_tape(tape, n_chains)
    t0a = None if t0 is None else np.ascontiguousarray(np.broadcast_to(t0, (n_chains,)), dtype=np.float64)
    h0a = None if horizon0 is None else np.ascontiguousarray(np.broadcast_to(horizon0, (n_chains,)), dtype=np.float64)
    rc = _lib.lib().pdmpflux_sample_skeleton_resume(sampler._handle, n_chains, int(n_sk), _ptr(x), _ptr(v),
                                                    _ptr(t0a), _ptr(h0a), int(event0),
                                                    C.c_uint64(int(seed) & (2**64 - 1)), int(chain_offset),
                                                    C.byref(t) if t is not None else None, C.byref(view), None)
    sampler.state = hb.status.copy()
    _lib.check(rc, hb.status)
    return hb if batched else hb.chain(0)


def sample_skeleton_until(sampler: AbstractPDMP, T, xinit, vinit, *, seed=None, verbose=True, tape=None, chain_offset=0,
                          batch=None, init_capacity=1024):
    """sample_skeleton(sampler, T::Float64, xinit, vinit; seed, init_capacity) (src/sample.jl:323-439): advance every
    chain to time T; the skeleton ends with the point at exactly t = T (t[end] == T).  The number of events is not
    known a priori: like the reference the history capacity doubles until every chain fits (the run is
    deterministic, so it is simply repeated with the larger capacity).  Returns a `PDMPHistory` for a (d,) init, a
    list of ragged `PDMPHistory` (one per chain) for a (C, d) init."""
    T = float(T)
    if not math.isfinite(T) or T < 0:
        raise _lib.ArgumentError(f"T must be finite and non-negative. Current value: {T}")
    if getattr(sampler, "kappa", None) is not None:
        raise _lib.UnsupportedError("the time-horizon sample_skeleton is not available for StickyZigZag on the device path")
    x, v, batched = _init_arrays(sampler, xinit, vinit)
    if batch is not None:
        batched = batch
    n_chains = x.shape[0]
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    cap = max(1, int(init_capacity))
    t, keep = _make_tape(tape, n_chains)
    while True:
        hb = PDMPHistoryBatch(n_chains, cap, sampler.dim)
        ncols = np.zeros(n_chains, dtype=np.int64)
        view = _lib.History(_ptr(hb.X), _ptr(hb.V), _ptr(hb.t), _ptr(hb.horizon), _ptr(hb.ar), _ptr(hb.error_value_ar),
                            _ptr(hb.errored_bound), _ptr(hb.rejected), _ptr(hb.hitting_horizon), _ptr(hb.status),
                            _ptr(hb.tape_pos), _ptr(hb.counters), cap, 0)
        rc = _lib.lib().pdmpflux_sample_skeleton_until(sampler._handle, n_chains, T, cap, _ptr(x), _ptr(v),
                                                       C.c_uint64(int(seed) & (2**64 - 1)), int(chain_offset),
                                                       C.byref(t) if t is not None else None, C.byref(view), _ptr(ncols),
                                                       None)
        if rc == _lib.ERR_CAPACITY:
            cap *= 2
            continue
        sampler.state = hb.status.copy()
        _lib.check(rc, hb.status)
        break
    out = []
    for c in range(n_chains):
        n = int(ncols[c])
        out.append(PDMPHistory(hb.X[c, :n].T, hb.V[c, :n].T, hb.t[c, :n], hb.horizon[c, :n], hb.ar[c, :n],
                               hb.errored_bound[c, :n], hb.error_value_ar[c, :n].T, hb.rejected[c, :n],
                               hb.hitting_horizon[c, :n]))
    return out if batched else out[0]


def _as_batch_arrays(history):
    if isinstance(history, PDMPHistoryBatch):
        return history.X, history.V, history.t
    # PDMPHistory: X is (d, n) like the Julia matrix -> chain-major slab (1, n, d)
    return (np.ascontiguousarray(history.X.T)[None], np.ascontiguousarray(history.V.T)[None],
            np.ascontiguousarray(history.t)[None])


def sample_from_skeleton(sampler: AbstractPDMP, N, history, dt=None, *, discard_vt=True):
    """The three methods of the reference, dispatched like Julia on the argument types:
      sample_from_skeleton(sampler, N::Int, history)            N equidistant samples, dt = t[end]/N   (sample.jl:475-513)
      sample_from_skeleton(sampler, dt::Float64, history)       samples at j*dt, j = 1..floor(t[end]/dt) (sample.jl:573-646)
      sample_from_skeleton(sampler, N::Int, dt::Float64, history) -> call as (sampler, N, history, dt): first N skeleton
                                                                points, samples at dt:dt:t[N]           (sample.jl:649-682)
    Returns a (d, M) array (or (2d+1, M) with discard_vt=False) for a `PDMPHistory`; (C, M, d) for a batch."""
    if dt is not None or isinstance(N, (float, np.floating)):
        return _sample_from_skeleton_dt(sampler, N, history, dt, discard_vt)
    if N <= 0:
        raise _lib.ArgumentError(f"N must be positive. Current value: {N}")
    X, V, t = _as_batch_arrays(history)
    n_chains, n_sk, d = X.shape
    ld = d if discard_vt else 2 * d + 1
    out = np.empty((n_chains, int(N), ld))
    if getattr(sampler, "kappa", None) is not None:   # sample_from_skeleton(::StickyPDMP, ...) (src/sample.jl:516-561)
        act = history.is_active if isinstance(history, PDMPHistoryBatch) else np.ascontiguousarray(history.is_active.T)[None]
        act = np.ascontiguousarray(act, dtype=np.uint8)
        _lib.check(_lib.lib().pdmpflux_sample_from_skeleton_sticky(d, n_sk, n_chains, _ptr(X), _ptr(V), _ptr(t), _ptr(act),
                                                                   int(N), int(bool(discard_vt)), _ptr(out), 0, None))
        return out if isinstance(history, PDMPHistoryBatch) else out[0].T
    _lib.check(_lib.lib().pdmpflux_sample_from_skeleton(sampler.flow_kind, d, n_sk, n_chains, _ptr(X), _ptr(V), _ptr(t),
                                                        int(N), int(bool(discard_vt)), _ptr(out), 0, None))
    return out if isinstance(history, PDMPHistoryBatch) else out[0].T


def _sample_from_skeleton_dt(sampler, N, history, dt, discard_vt):
    X, V, t = _as_batch_arrays(history)
    n_chains, ld_sk, d = X.shape
    if dt is None:            # (sampler, dt, history)
        dt, n_sk = float(N), ld_sk
    else:                     # (sampler, N, dt, history): only the first N skeleton points
        dt, n_sk = float(dt), int(N)
        if not 1 <= n_sk <= ld_sk:
            raise _lib.ArgumentError(f"N must be in 1..{ld_sk}. Current value: {N}")
    if not (dt > 0 and math.isfinite(dt)):
        raise _lib.ArgumentError(f"dt must be positive. Current value: {dt}")
    t_end = t[:, n_sk - 1]
    # floor(T_end / dt) (sample.jl:609) and length(dt:dt:t[end]) (:655) agree up to the range's rounding guard
    n_out = np.floor(t_end / dt).astype(np.int64)
    if not np.all(n_out == n_out[0]):
        raise _lib.ArgumentError("chains have different t[end]: call per chain (ragged output)")
    M = int(n_out[0])
    ld = d if discard_vt else 2 * d + 1
    out = np.empty((n_chains, M, ld))
    if M > 0:
        _lib.check(_lib.lib().pdmpflux_sample_from_skeleton_dt(sampler.flow_kind, d, n_sk, ld_sk, n_chains, _ptr(X), _ptr(V),
                                                               _ptr(t), dt, M, int(bool(discard_vt)), _ptr(out), 0, None))
    return out if isinstance(history, PDMPHistoryBatch) else out[0].T


def RV_diagnostic(history, potential, *, B=0, flow_kind=0):
    """RV_diagnostic(history, U; B) (src/diagnostic.jl:37-75): realised volatility of U along the skeleton, on the
    device.  `potential` is the descriptor whose U(x) plugin is evaluated (the reference takes a Julia closure).
    `history`: a `PDMPHistory` -> float; a `PDMPHistoryBatch` or a list of (ragged) `PDMPHistory` -> array (C,).
    B = 0 picks floor(sqrt(n)) like the reference; B < 0 raises ArgumentError.  flow_kind = 1 interpolates with the
    Boomerang rotation (the online variant's sampler.flow) instead of the offline diagnostic's straight lines."""
    if B < 0:
        raise _lib.ArgumentError(f"B must be non-negative. Current value: {B}")
    ncols = None
    if isinstance(history, (list, tuple)):
        if not history:
            return np.zeros(0)
        d = history[0].X.shape[0]
        ncols = np.array([h.t.shape[0] for h in history], dtype=np.int64)
        ld = max(1, int(ncols.max()))
        X = np.zeros((len(history), ld, d)); V = np.zeros((len(history), ld, d)); t = np.zeros((len(history), ld))
        for c, h in enumerate(history):
            n = int(ncols[c])
            X[c, :n] = h.X.T; V[c, :n] = h.V.T; t[c, :n] = h.t
        t_end = np.array([h.t[-1] if h.t.shape[0] else 0.0 for h in history])
    else:
        X, V, t = _as_batch_arrays(history)
        t_end = t[:, -1] if t.shape[1] else np.zeros(t.shape[0])
    n_chains, ld, d = X.shape
    if ld == 0:
        out = np.zeros(n_chains)      # N == 0 -> 0.0 (diagnostic.jl:40)
        return out if isinstance(history, (list, tuple, PDMPHistoryBatch)) else float(out[0])
    bad = ~(np.isfinite(t_end) & (t_end >= 0.0))
    if bad.any():
        raise _lib.ArgumentError(f"history.t[end] must be finite and non-negative. Current value: {t_end[bad][0]}")
    h = potential._create(d)
    try:
        rv = np.empty(n_chains)
        _lib.check(_lib.lib().pdmpflux_rv_diagnostic(h, int(flow_kind), ld, ld, n_chains, _ptr(ncols), int(B), _ptr(X),
                                                     _ptr(V), _ptr(t), _ptr(rv), 0, None))
    finally:
        _lib.lib().pdmpflux_potential_destroy(h)
    return rv if isinstance(history, (list, tuple, PDMPHistoryBatch)) else float(rv[0])


def sample_skeleton_with_diagnostic(sampler: AbstractPDMP, T, xinit, vinit, potential=None, *, B=1000, seed=None,
                                    verbose=True, init_capacity=1024, tape=None):
    """sample_skeleton_with_diagnostic(sampler, T, xinit, vinit, U; B, seed) (src/sample.jl:75-236): the time-horizon
    skeleton plus the realised volatility of U over B blocks of [0, T].  The reference accumulates the increments
    online through sampler.flow; the sum telescopes to the per-boundary form, so the value is computed here from the
    finished skeleton on the device with the sampler's own flow (equal to the online value up to rounding, and to
    RV_diagnostic for the straight-line samplers -- the reference's own 1e-10 check, test/test_diagnostics.jl:126-143).
    Returns (history, rv); for a (C, d) init (list of histories, array of rv)."""
    if B <= 0:
        raise _lib.ArgumentError(f"B must be positive. Current value: {B}")
    hist = sample_skeleton_until(sampler, T, xinit, vinit, seed=seed, verbose=verbose, tape=tape,
                                 init_capacity=init_capacity)
    pot = sampler.potential if potential is None else potential
    if float(T) == 0.0:   # the initial point alone (sample.jl:111-114)
        return hist, (np.zeros(len(hist)) if isinstance(hist, list) else 0.0)
    return hist, RV_diagnostic(hist, pot, B=B, flow_kind=sampler.flow_kind)


def sample(sampler: AbstractPDMP, N_sk, N_samples, xinit, vinit, *, seed=None, verbose=True, discard_vt=True):
    """sample(sampler, N_sk, N_samples, xinit, vinit; seed) = sample_from_skeleton o sample_skeleton
    (src/sample.jl:27-58)."""
    history = sample_skeleton(sampler, N_sk, xinit, vinit, seed=seed, verbose=verbose)
    return sample_from_skeleton(sampler, N_samples, history, discard_vt=discard_vt)


def skeleton_moments(sampler: AbstractPDMP, history, burn_in_cols=0):
    """Per-chain time averages of x and x^2 over the skeleton (closed-form segment integrals on the device).
    Returns (mean[C, d], second_moment[C, d], T[C]).  No reference equivalent (ESS inputs)."""
    X, V, t = _as_batch_arrays(history)
    n_chains, n_sk, d = X.shape
    m1 = np.empty((n_chains, d)); m2 = np.empty((n_chains, d)); T = np.empty(n_chains)
    _lib.check(_lib.lib().pdmpflux_skeleton_moments(sampler.flow_kind, d, n_sk, n_chains, int(burn_in_cols), _ptr(X),
                                                    _ptr(V), _ptr(t), _ptr(m1), _ptr(m2), _ptr(T), 0, None))
    return m1 / T[:, None], m2 / T[:, None], T


def ess_from_chain_means(mean, second):
    """Cross-chain ESS per coordinate: with m_{c,i} chain c's time average of x_i and
    Var_pi(x_i) the pooled marginal variance, ESS_i per chain = Var_pi(x_i) / Var_c(m_{c,i}); the total is C
    times that.  Returns (ess_x[d], ess_x2[d]) totals over all chains."""
    C_ = mean.shape[0]
    pooled_mean = mean.mean(axis=0)
    pooled_var = second.mean(axis=0) - pooled_mean**2
    ess_x = C_ * pooled_var / mean.var(axis=0, ddof=1)
    return ess_x, pooled_var
