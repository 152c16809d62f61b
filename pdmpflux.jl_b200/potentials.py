"""Device potential descriptors: the second positional argument of the sampler constructors.

The reference takes a Julia closure (`grad U`, or `U` for the *AD constructors, ADBackend.jl:30-142); an
arbitrary closure cannot run on the device, so a descriptor names one of the hand-written device plugins
(pdmpflux.jl_b200/csrc/potentials.cuh) and carries its parameters.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

GAUSS_STD, GAUSS_DIAG, GAUSS_EQUICORR, BANANA, BANANA_README_SCALAR, LOGREG = range(6)


class Potential:
    kind = None

    def params(self, dim):
        return np.zeros(0)

    def _create(self, dim):
        p = np.ascontiguousarray(self.params(dim), dtype=np.float64)
        h = C.c_void_p()
        _lib.check(_lib.lib().pdmpflux_potential_create(self.kind, int(dim), p.ctypes.data if p.size else None,
                                                        p.size, C.byref(h)))
        return h


class GaussStd(Potential):
    """U(x) = |x|^2 / 2 (README.md:36-38)."""
    kind = GAUSS_STD


class GaussDiag(Potential):
    """U(x) = sum_i p_i x_i^2 / 2."""
    kind = GAUSS_DIAG

    def __init__(self, precisions):
        self.p = np.asarray(precisions, dtype=np.float64)

    def params(self, dim):
        if self.p.shape != (dim,):
            raise _lib.DimensionMismatch(f"GaussDiag has {self.p.size} precisions but dim = {dim}")
        return self.p


class GaussEquicorr(Potential):
    """'Slanted' Gaussian: Sigma = (1 - rho) I + rho 1 1^T."""
    kind = GAUSS_EQUICORR

    def __init__(self, rho):
        self.rho = float(rho)

    def params(self, dim):
        return np.array([self.rho])


class Banana(Potential):
    """U = (x1^2 + (x2 - x1^2 + 1)^2 + sum_{i>=3} x_i^2) / 2 (test/test_config.jl:33-36)."""
    kind = BANANA


class BananaReadmeScalar(Potential):
    """README.md:62-65 verbatim: the manual 'gradient' returns a scalar broadcast to every coordinate.
    Not a gradient field; chains can drift to a region with zero rate and never produce another event
    (the library then reports STEP_LIMIT instead of hanging like the reference would)."""
    kind = BANANA_README_SCALAR


class LogReg(Potential):
    """Bayesian logistic regression posterior, prior N(0, sigma0^2 I) (BASELINE.json config 4)."""
    kind = LOGREG

    def __init__(self, X, y, sigma0=10.0):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.sigma0 = float(sigma0)

    def params(self, dim):
        n, d = self.X.shape
        if d != dim or self.y.shape != (n,):
            raise _lib.DimensionMismatch("LogReg design matrix does not match dim")
        return np.concatenate([[float(n), self.sigma0], self.X.ravel(), self.y])
