"""PDMPHistory (Composites.jl:138-164) for one chain, and its chain-major batch container."""
from __future__ import annotations

import numpy as np


class PDMPHistory:
    """One chain.  Field names, shapes and dtypes follow the reference struct: X, V are (d, n) (column k =
    event k, like the Julia Matrix), t/horizon/ar (n,), error_value_ar (5, n), errored_bound / rejected /
    hitting_horizon int32 (n,), is_active (d, n) all True.  `.x` / `.v` give lists of column copies
    (Composites.jl:225-234)."""

    def __init__(self, X, V, t, horizon, ar, errored_bound, error_value_ar, rejected, hitting_horizon, is_active=None):
        self.X, self.V, self.t, self.horizon, self.ar = X, V, t, horizon, ar
        self.errored_bound, self.error_value_ar = errored_bound, error_value_ar
        self.rejected, self.hitting_horizon = rejected, hitting_horizon
        self._is_active = is_active   # (d, n) bool, only stored by the sticky samplers

    @property
    def is_active(self):
        return np.ones(self.X.shape, dtype=bool) if self._is_active is None else self._is_active

    @property
    def x(self):
        return [self.X[:, k].copy() for k in range(self.X.shape[1])]

    @property
    def v(self):
        return [self.V[:, k].copy() for k in range(self.V.shape[1])]

    def __len__(self):
        return self.t.shape[0]


class PDMPHistoryBatch:
    """C chains, chain-major: X[c] is chain c's (n_sk, d) slab, i.e. the memory of a Julia Matrix(d, n_sk)."""

    def __init__(self, n_chains, n_sk, d, alloc=np.empty, sticky=False):
        self.n_chains, self.n_sk, self.d = n_chains, n_sk, d
        self.is_active = alloc((n_chains, n_sk, d), dtype=np.uint8) if sticky else None
        self.X = alloc((n_chains, n_sk, d), dtype=np.float64)
        self.V = alloc((n_chains, n_sk, d), dtype=np.float64)
        self.t = alloc((n_chains, n_sk), dtype=np.float64)
        self.horizon = alloc((n_chains, n_sk), dtype=np.float64)
        self.ar = alloc((n_chains, n_sk), dtype=np.float64)
        self.error_value_ar = alloc((n_chains, n_sk, 5), dtype=np.float64)
        self.errored_bound = alloc((n_chains, n_sk), dtype=np.int32)
        self.rejected = alloc((n_chains, n_sk), dtype=np.int32)
        self.hitting_horizon = alloc((n_chains, n_sk), dtype=np.int32)
        self.status = np.zeros(n_chains, dtype=np.int32)
        self.tape_pos = np.zeros((n_chains, 3), dtype=np.int64)
        self.counters = np.zeros((n_chains, 2), dtype=np.int64)

    def chain(self, c=0) -> PDMPHistory:
        return PDMPHistory(self.X[c].T, self.V[c].T, self.t[c], self.horizon[c], self.ar[c], self.errored_bound[c],
                           self.error_value_ar[c].T, self.rejected[c], self.hitting_horizon[c],
                           None if self.is_active is None else self.is_active[c].T.astype(bool))

    def __len__(self):
        return self.n_chains
