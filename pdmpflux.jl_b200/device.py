"""Device-resident form of the sampling loop: a handle on the per-chain PDMPState array living in HBM
(`pdmpflux_chains_*` in include/pdmpflux_cuda.h).  Buffers are passed as raw device pointers, so any
allocator works (bench.py uses torch tensors: `t.data_ptr()`); this module itself does not import torch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_FIELDS = ("X", "V", "t", "horizon", "ar", "error_value_ar", "errored_bound", "rejected", "hitting_horizon")


def _dev_ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


def device_history_view(n_cols, **buffers):
    """Build a `pdmpflux_history` over device buffers (torch tensors or integer device pointers).  Keyword names
    are the PDMPHistory fields; omitted fields are not stored."""
    unknown = set(buffers) - set(_FIELDS) - {"is_active"}
    if unknown:
        raise _lib.ArgumentError(f"unknown history fields: {sorted(unknown)}")
    ptrs = [_dev_ptr(buffers.get(f)) for f in _FIELDS]
    return _lib.History(*ptrs, None, None, None, int(n_cols), 1, _dev_ptr(buffers.get("is_active")))


class DeviceChains:
    """n_chains PDMPStates on the current CUDA device (the analogue of `sampler.state`, src/sample.jl:281)."""

    def __init__(self, sampler, xinit, vinit, *, seed=0, chain_offset=0, tape=None):
        self.sampler = sampler
        on_device = hasattr(xinit, "data_ptr")
        if on_device:
            n_chains, d = xinit.shape
            xp, vp = xinit.data_ptr(), vinit.data_ptr()
            self._keep = (xinit, vinit)
        else:
            x = np.ascontiguousarray(np.atleast_2d(xinit), dtype=np.float64)
            v = np.ascontiguousarray(np.atleast_2d(vinit), dtype=np.float64)
            n_chains, d = x.shape
            xp, vp = x.ctypes.data, v.ctypes.data
            self._keep = (x, v)
        if d != sampler.dim:
            raise _lib.DimensionMismatch(f"xinit has dimension {d}, sampler.dim = {sampler.dim}")
        self.n_chains = n_chains
        t = None
        if tape is not None:
            E, U, N = tape
            if hasattr(E, "data_ptr"):
                t = _lib.Tape(E.data_ptr(), U.data_ptr(), N.data_ptr(), E.shape[1], U.shape[1], N.shape[1], 1)
            else:
                E, U, N = (np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in (E, U, N))
                t = _lib.Tape(E.ctypes.data, U.ctypes.data, N.ctypes.data, E.shape[1], U.shape[1], N.shape[1], 0)
            self._tape_keep = (E, U, N)
        h = C.c_void_p()
        _lib.check(_lib.lib().pdmpflux_chains_create(sampler._handle, n_chains, xp, vp, int(on_device),
                                                     C.c_uint64(int(seed) & (2**64 - 1)), int(chain_offset),
                                                     C.byref(t) if t is not None else None, C.byref(h)))
        self._h = h

    def record(self, view, col=0, stream=None):
        _lib.check(_lib.lib().pdmpflux_chains_record(self._h, C.byref(view), int(col), stream))

    def advance(self, n_events, view=None, col0=0, stream=None):
        _lib.check(_lib.lib().pdmpflux_chains_advance(self._h, int(n_events), C.byref(view) if view is not None else None,
                                                      int(col0), stream))

    def status(self):
        st = np.zeros(self.n_chains, dtype=np.int32)
        pos = np.zeros((self.n_chains, 3), dtype=np.int64)
        cnt = np.zeros((self.n_chains, 2), dtype=np.int64)
        rc = _lib.lib().pdmpflux_chains_status(self._h, st.ctypes.data, pos.ctypes.data, cnt.ctypes.data)
        _lib.check(rc, st)
        return st, pos, cnt

    def enable_moments(self):
        """Accumulate int x dt and int x^2 dt per chain inside the kernel from now on (no skeleton needed)."""
        _lib.check(_lib.lib().pdmpflux_chains_enable_moments(self._h))

    def moments(self):
        """(int x dt, int x^2 dt) accumulated since enable_moments(), each (n_chains, d)."""
        d = self.sampler.dim
        m1 = np.empty((self.n_chains, d)); m2 = np.empty((self.n_chains, d))
        _lib.check(_lib.lib().pdmpflux_chains_get_moments(self._h, m1.ctypes.data, m2.ctypes.data, 0))
        return m1, m2

    def set_stop_time(self, T):
        _lib.check(_lib.lib().pdmpflux_chains_set_stop_time(self._h, float(T)))

    def get_state(self):
        d = self.sampler.dim
        x = np.empty((self.n_chains, d)); v = np.empty((self.n_chains, d))
        t = np.empty(self.n_chains); h = np.empty(self.n_chains)
        _lib.check(_lib.lib().pdmpflux_chains_get_state(self._h, x.ctypes.data, v.ctypes.data, t.ctypes.data,
                                                        h.ctypes.data, 0))
        return x, v, t, h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().pdmpflux_chains_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
