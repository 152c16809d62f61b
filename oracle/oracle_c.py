"""
TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_build/libpdmp_oracle.so (the C restatement).

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs only.  PARITY UNPINNED (see
pdmp_oracle.c header): no Julia here, no golden skeleton vectors upstream.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpdmp_oracle.so")

ZIGZAG, BPS, FECMC, BOOMERANG = 0, 1, 2, 3
GAUSS_STD, GAUSS_DIAG, GAUSS_EQUICORR, BANANA, BANANA_README, LOGREG, GAUSS_DENSE = range(7)
DERIV_JVP, DERIV_FD = 0, 1
ST_OK, ST_TAPE_EXHAUSTED, ST_NOT_PROBVEC, ST_ITER_LIMIT = 0, 1, 2, 3


class Cfg(C.Structure):
    _fields_ = [
        ("sampler", C.c_int32), ("potential", C.c_int32), ("dim", C.c_int32), ("grid_size", C.c_int32),
        ("vectorized_bound", C.c_int32), ("signed_bound", C.c_int32), ("adaptive", C.c_int32),
        ("deriv_mode", C.c_int32),
        ("gaussian_velocity", C.c_int32), ("ran_p", C.c_int32), ("switch_", C.c_int32), ("positive", C.c_int32),
        ("tmax", C.c_double), ("refresh_rate", C.c_double), ("mix_p", C.c_double), ("speed_factor", C.c_double),
        ("pot_params", C.POINTER(C.c_double)), ("n_pot_params", C.c_int64),
    ]


class Hist(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("X", "V", "t", "horizon", "ar", "error_value_ar", "errored_bound", "rejected", "hitting_horizon",
                 "tape_pos")]


def build(force=False):
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < max(os.path.getmtime(os.path.join(_HERE, f))
                                        for f in ("pdmp_oracle.c", "pdmp_draws.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.pdmp_oracle_bound.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def make_cfg(sampler, potential, dim, pot_params=None, *, grid_size=10, tmax=2.0, refresh_rate=0.0,
             vectorized_bound=True, signed_bound=True, adaptive=True, deriv_mode=DERIV_JVP,
             gaussian_velocity=False, ran_p=False, mix_p=0.5, switch=True, positive=True, speed_factor=1.0):
    pp = np.ascontiguousarray(pot_params if pot_params is not None else np.zeros(1), dtype=np.float64)
    c = Cfg(sampler, potential, dim, grid_size, int(vectorized_bound), int(signed_bound), int(adaptive),
            deriv_mode, int(gaussian_velocity), int(ran_p), int(switch), int(positive),
            float(tmax), float(refresh_rate), float(mix_p), float(speed_factor), _dp(pp), pp.size)
    c._keep = pp
    return c


class SkeletonResult:
    pass


def sample_skeleton(cfg: Cfg, n_sk, xinit, vinit, *, tape=None, seed=0, chain_offset=0, nthreads=1,
                    store=True):
    """xinit/vinit: (C, d) arrays (row c = chain c).  tape = (E, U, N) each (C, n*) or None -> Philox."""
    x = np.ascontiguousarray(np.atleast_2d(xinit), dtype=np.float64)
    v = np.ascontiguousarray(np.atleast_2d(vinit), dtype=np.float64)
    nch, d = x.shape
    assert d == cfg.dim and v.shape == x.shape
    r = SkeletonResult()
    if store:
        r.X = np.full((nch, n_sk, d), np.nan)
        r.V = np.full((nch, n_sk, d), np.nan)
        r.error_value_ar = np.zeros((nch, n_sk, 5))
    else:
        r.X = r.V = r.error_value_ar = None
    r.t = np.full((nch, n_sk), np.nan)
    r.horizon = np.full((nch, n_sk), np.nan)
    r.ar = np.full((nch, n_sk), np.nan)
    r.errored_bound = np.zeros((nch, n_sk), dtype=np.int32)
    r.rejected = np.zeros((nch, n_sk), dtype=np.int32)
    r.hitting_horizon = np.zeros((nch, n_sk), dtype=np.int32)
    r.status = np.zeros(nch, dtype=np.int32)
    r.counters = np.zeros((nch, 2), dtype=np.int64)
    r.tape_used = np.zeros((nch, 3), dtype=np.int64)
    r.tape_pos = np.zeros((nch, n_sk, 3), dtype=np.int64)

    def vp(a):
        return a.ctypes.data if a is not None else None

    h = Hist(vp(r.X), vp(r.V), vp(r.t), vp(r.horizon), vp(r.ar), vp(r.error_value_ar), vp(r.errored_bound),
             vp(r.rejected), vp(r.hitting_horizon), vp(r.tape_pos))
    if tape is not None:
        E, U, N = (np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in tape)
        assert E.shape[0] == nch and U.shape[0] == nch and N.shape[0] == nch
        mode = 0
    else:
        E = U = N = np.zeros((nch, 1))
        mode = 1
    rc = lib().pdmp_oracle_sample_skeleton(
        C.byref(cfg), C.c_int64(nch), C.c_int64(n_sk), _dp(x), _dp(v), C.c_int(mode), C.c_uint64(seed),
        C.c_int64(chain_offset), _dp(E), C.c_int64(E.shape[1]), _dp(U), C.c_int64(U.shape[1]), _dp(N),
        C.c_int64(N.shape[1]), C.byref(h), r.status.ctypes.data_as(C.c_void_p),
        r.counters.ctypes.data_as(C.c_void_p), r.tape_used.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("pdmp_oracle_sample_skeleton: invalid arguments")
    return r


def sample_skeleton_until(cfg: Cfg, T, cap, xinit, vinit, *, tape=None, seed=0, chain_offset=0, nthreads=1):
    """Time-horizon variant (src/sample.jl:323-439).  Returns (result arrays with `cap` columns, ncols[C])."""
    x = np.ascontiguousarray(np.atleast_2d(xinit), dtype=np.float64)
    v = np.ascontiguousarray(np.atleast_2d(vinit), dtype=np.float64)
    nch, d = x.shape
    r = SkeletonResult()
    r.X = np.full((nch, cap, d), np.nan); r.V = np.full((nch, cap, d), np.nan)
    r.error_value_ar = np.zeros((nch, cap, 5))
    r.t = np.full((nch, cap), np.nan); r.horizon = np.full((nch, cap), np.nan); r.ar = np.full((nch, cap), np.nan)
    r.errored_bound = np.zeros((nch, cap), dtype=np.int32); r.rejected = np.zeros((nch, cap), dtype=np.int32)
    r.hitting_horizon = np.zeros((nch, cap), dtype=np.int32)
    r.status = np.zeros(nch, dtype=np.int32)
    r.ncols = np.zeros(nch, dtype=np.int64)
    h = Hist(r.X.ctypes.data, r.V.ctypes.data, r.t.ctypes.data, r.horizon.ctypes.data, r.ar.ctypes.data,
             r.error_value_ar.ctypes.data, r.errored_bound.ctypes.data, r.rejected.ctypes.data,
             r.hitting_horizon.ctypes.data, None)
    if tape is not None:
        E, U, N = (np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in tape)
        mode = 0
    else:
        E = U = N = np.zeros((nch, 1)); mode = 1
    rc = lib().pdmp_oracle_sample_skeleton_until(
        C.byref(cfg), C.c_int64(nch), C.c_int64(cap), C.c_double(T), _dp(x), _dp(v), C.c_int(mode), C.c_uint64(seed),
        C.c_int64(chain_offset), _dp(E), C.c_int64(E.shape[1]), _dp(U), C.c_int64(U.shape[1]), _dp(N),
        C.c_int64(N.shape[1]), C.byref(h), r.status.ctypes.data_as(C.c_void_p), r.ncols.ctypes.data_as(C.c_void_p),
        C.c_int(nthreads))
    if rc != 0:
        raise ValueError("pdmp_oracle_sample_skeleton_until: invalid arguments")
    return r


def bound(cfg: Cfg, x, v, horizon):
    G = max(cfg.grid_size, 2)
    x = np.ascontiguousarray(x, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    grid = np.zeros(G); box = np.zeros(G - 1); cum = np.zeros(G)
    step = lib().pdmp_oracle_bound(C.byref(cfg), _dp(x), _dp(v), C.c_double(horizon), _dp(grid), _dp(box), _dp(cum))
    nb = G if cfg.grid_size else 2
    return grid[:nb], box[:nb - 1], cum[:nb], step


def sample_from_skeleton(flow_kind, X, V, t, N, discard_vt=True):
    """X, V: (n_sk, d) (row k = event k); returns (N, d) or (N, 2d+1)."""
    X = np.ascontiguousarray(X, dtype=np.float64); V = np.ascontiguousarray(V, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    n_sk, d = X.shape
    out = np.empty((N, d if discard_vt else 2 * d + 1))
    rc = lib().pdmp_oracle_sample_from_skeleton(C.c_int(flow_kind), C.c_int(d), C.c_int64(n_sk), _dp(X), _dp(V),
                                                _dp(t), C.c_int64(N), C.c_int(int(discard_vt)), _dp(out))
    if rc != 0:
        raise ValueError("N must be positive")
    return out


def draws(seed, chain, event, n):
    E = np.zeros(n); U = np.zeros(n); N = np.zeros(n)
    lib().pdmp_oracle_draws(C.c_uint64(seed), C.c_uint64(chain), C.c_uint64(event), C.c_int32(n), _dp(E), _dp(U), _dp(N))
    return E, U, N


def philox_raw(c, k):
    out = (C.c_uint32 * 4)()
    lib().pdmp_oracle_philox_raw(*(C.c_uint32(int(a)) for a in c), *(C.c_uint32(int(a)) for a in k), out)
    return [int(a) for a in out]


def num_threads():
    return int(lib().pdmp_oracle_num_threads())
