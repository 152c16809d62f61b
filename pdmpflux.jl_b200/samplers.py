"""Sampler constructors with the reference's names, keyword arguments and defaults.

ZigZag / ZigZagAD          src/Samplers/ZigZagSamplers.jl:58-60, :118-119
BPS / BPSAD                src/Samplers/BouncyParticleSamplers.jl:21-24, :86-87
ForwardECMC / ...AD        src/Samplers/ForwardEventChainMonteCarlo.jl:301-303, :367-369
Boomerang / BoomerangAD    src/Samplers/BoomerangSamplers.jl:21-23, :79-80
StickyZigZag / ...AD       src/Samplers/StickyZigZagSamplers.jl:60-111, :117-127 (third positional argument: kappa)
SpeedUpZigZag / ...AD      src/Samplers/SpeedUpZigZagSamplers.jl:58-116, :119-129

The second positional argument is a device potential descriptor (potentials.py) instead of a Julia closure.
"""
from __future__ import annotations

import ctypes as C
import warnings

from . import _lib
from .potentials import Potential

ZIGZAG, BPS_KIND, FECMC, BOOMERANG, STICKY_ZIGZAG, SPEEDUP_ZIGZAG = 0, 1, 2, 3, 4, 5
DERIV_JVP, DERIV_FD = 0, 1

_EXACT_AD = {"ForwardDiff", "Zygote", "ReverseDiff", "Enzyme", "PolyesterForwardDiff"}


def _deriv_mode(ad_backend):
    """Which d/dt the grid bound uses (UpperBound.jl:98-121, :209-227; _pdmp_ad_backend AbstractPDMP.jl:18-28).
    Every exact AD backend computes the same analytic derivative -> JVP mode."""
    if ad_backend is None or ad_backend in ("", "Undefined", "FiniteDiff"):
        return DERIV_FD
    if ad_backend in _EXACT_AD:
        return DERIV_JVP
    raise _lib.ArgumentError(f"Unsupported AD_backend: {ad_backend}")


class AbstractPDMP:
    """Common fields of the reference's sampler structs (AbstractPDMP.jl:31-56) that have a device meaning."""
    _kind = None
    flow_kind = 0  # 0: x + v t ; 1: rotation (Boomerang)

    def __init__(self, dim, potential, *, grid_size, tmax, refresh_rate, vectorized_bound, signed_bound, adaptive,
                 AD_backend, gaussian_velocity=False, ran_p=False, mix_p=0.5, switch=True, positive=True,
                 speed_factor=1.0, max_steps=0):
        if not isinstance(potential, Potential):
            raise _lib.UnsupportedError(
                "the GPU path needs a device potential descriptor (pdmpflux_b200.GaussStd(), Banana(), ...); "
                "arbitrary Python/Julia closures cannot run on the device and there is no CPU fallback")
        dim = int(dim)
        if dim <= 0:
            raise _lib.ArgumentError(f"dimension dim must be positive. Current value: {dim}")
        if grid_size < 0:
            raise _lib.ArgumentError(f"grid_size must be non-negative. Current value: {grid_size}")
        self.dim = dim
        self.potential = potential
        self.AD_backend = AD_backend
        cfg = _lib.Config(int(grid_size), int(vectorized_bound), int(signed_bound), int(adaptive),
                          _deriv_mode(AD_backend), int(gaussian_velocity), int(ran_p), int(switch), int(positive),
                          int(max_steps), float(tmax), float(refresh_rate), float(mix_p), float(speed_factor))
        self._pot_handle = potential._create(dim)
        h = C.c_void_p()
        try:
            if self._kind == STICKY_ZIGZAG:
                _lib.check(_lib.lib().pdmpflux_sampler_create_sticky(dim, self._pot_handle, C.byref(cfg),
                                                                      self.kappa.ctypes.data, C.byref(h)))
            else:
                _lib.check(_lib.lib().pdmpflux_sampler_create(self._kind, dim, self._pot_handle, C.byref(cfg), C.byref(h)))
        except Exception:
            _lib.lib().pdmpflux_potential_destroy(self._pot_handle)
            self._pot_handle = None
            raise
        self._handle = h
        out = _lib.Config()
        _lib.check(_lib.lib().pdmpflux_sampler_get_config(h, C.byref(out)))
        # fields after the constructor rewrites, named as in the reference structs
        self.grid_size, self.tmax, self.refresh_rate = out.grid_size, out.tmax, out.refresh_rate
        self.vectorized_bound, self.signed_bound = bool(out.vectorized_bound), bool(out.signed_bound)
        self.adaptive = bool(out.adaptive)
        self.mix_p, self.ran_p, self.switch, self.positive = out.mix_p, bool(out.ran_p), bool(out.switch_), bool(out.positive)
        self.speed_factor = out.speed_factor
        self.state = None  # final per-chain status of the last sample_skeleton call (src/sample.jl:281)

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().pdmpflux_sampler_destroy(self._handle)
            if getattr(self, "_pot_handle", None):
                _lib.lib().pdmpflux_potential_destroy(self._pot_handle)
        except Exception:
            pass


class ZigZag(AbstractPDMP):
    _kind = ZIGZAG

    def __init__(self, dim, potential, *, grid_size=10, tmax=2.0, refresh_rate=0.0, vectorized_bound=True,
                 signed_bound=True, adaptive=True, AD_backend="FiniteDiff", max_steps=0):
        if signed_bound and not vectorized_bound:
            warnings.warn("Signed bound is not compatible with non-vectorized bound for ZigZag, switching to unsigned bound")
        super().__init__(dim, potential, grid_size=grid_size, tmax=tmax, refresh_rate=refresh_rate,
                         vectorized_bound=vectorized_bound, signed_bound=signed_bound, adaptive=adaptive,
                         AD_backend=AD_backend, max_steps=max_steps)


def ZigZagAD(dim, potential, *, refresh_rate=0.0, grid_size=10, tmax=2.0, vectorized_bound=True, signed_bound=True,
             adaptive=True, AD_backend="ForwardDiff", max_steps=0):
    return ZigZag(dim, potential, refresh_rate=refresh_rate, grid_size=grid_size, tmax=tmax,
                  vectorized_bound=vectorized_bound, signed_bound=signed_bound, adaptive=adaptive,
                  AD_backend=AD_backend, max_steps=max_steps)


class BPS(AbstractPDMP):
    _kind = BPS_KIND

    def __init__(self, dim, potential, *, grid_size=10, tmax=1.0, refresh_rate=0.1, vectorized_bound=False,
                 signed_bound=True, adaptive=True, AD_backend="ForwardDiff", Gaussian_velocity=False, max_steps=0):
        super().__init__(dim, potential, grid_size=grid_size, tmax=tmax, refresh_rate=refresh_rate,
                         vectorized_bound=False, signed_bound=signed_bound, adaptive=adaptive, AD_backend=AD_backend,
                         gaussian_velocity=Gaussian_velocity, max_steps=max_steps)


def BPSAD(dim, potential, *, refresh_rate=0.0, grid_size=10, tmax=2.0, vectorized_bound=True, signed_bound=True,
          adaptive=True, AD_backend="ForwardDiff", max_steps=0):
    return BPS(dim, potential, refresh_rate=refresh_rate, grid_size=grid_size, tmax=tmax, signed_bound=signed_bound,
               adaptive=adaptive, AD_backend=AD_backend, max_steps=max_steps)


class ForwardECMC(AbstractPDMP):
    _kind = FECMC

    def __init__(self, dim, potential, *, grid_size=10, tmax=2.0, signed_bound=True, adaptive=True, ran_p=False,
                 mix_p=0.5, switch=True, positive=True, AD_backend="ForwardDiff", speed_factor=1.0, normal=False,
                 max_steps=0):
        if int(dim) < 2:
            raise _lib.ArgumentError(f"The dimension must be at least 2 to use the ForwardEventChain. Got dimension {dim}")
        if normal:
            raise _lib.UnsupportedError("ForwardECMC(normal=true) throws upstream (ForwardEventChainMonteCarlo.jl:227); not ported")
        super().__init__(dim, potential, grid_size=grid_size, tmax=tmax, refresh_rate=0.0, vectorized_bound=False,
                         signed_bound=signed_bound, adaptive=adaptive, AD_backend=AD_backend, ran_p=ran_p, mix_p=mix_p,
                         switch=switch, positive=positive, speed_factor=speed_factor, max_steps=max_steps)


def ForwardECMCAD(dim, potential, *, grid_size=10, tmax=2.0, signed_bound=True, adaptive=True,
                  AD_backend="ForwardDiff", ran_p=False, mix_p=0.5, switch=True, positive=True, speed_factor=1.0,
                  max_steps=0):
    return ForwardECMC(dim, potential, grid_size=grid_size, tmax=tmax, signed_bound=signed_bound, adaptive=adaptive,
                       ran_p=ran_p, mix_p=mix_p, switch=switch, positive=positive, AD_backend=AD_backend,
                       speed_factor=speed_factor, max_steps=max_steps)


class Boomerang(AbstractPDMP):
    _kind = BOOMERANG
    flow_kind = 1

    def __init__(self, dim, potential, *, grid_size=10, tmax=1.0, refresh_rate=0.1, vectorized_bound=False,
                 signed_bound=True, adaptive=True, AD_backend="FiniteDiff", max_steps=0):
        super().__init__(dim, potential, grid_size=grid_size, tmax=tmax, refresh_rate=refresh_rate,
                         vectorized_bound=False, signed_bound=signed_bound, adaptive=adaptive, AD_backend=AD_backend,
                         max_steps=max_steps)


def BoomerangAD(dim, potential, *, refresh_rate=0.0, grid_size=10, tmax=2.0, vectorized_bound=True, signed_bound=True,
                adaptive=True, AD_backend="ForwardDiff", max_steps=0):
    return Boomerang(dim, potential, refresh_rate=refresh_rate, grid_size=grid_size, tmax=tmax,
                     signed_bound=signed_bound, adaptive=adaptive, AD_backend=AD_backend, max_steps=max_steps)


class StickyZigZag(ZigZag):
    """StickyZigZag(dim, grad U, kappa; kw...) (StickyZigZagSamplers.jl:60-111): the Zig-Zag closures plus thawing rates
    `kappa` (prior inclusion); coordinates stick to their axis when they cross it and thaw at rate kappa_i
    (src/StickySamplingLoop.jl)."""
    _kind = STICKY_ZIGZAG

    def __init__(self, dim, potential, kappa, *, refresh_rate=0.0, grid_size=10, tmax=2.0, vectorized_bound=True,
                 signed_bound=True, adaptive=True, AD_backend="FiniteDiff", max_steps=0):
        import numpy as np
        self.kappa = np.ascontiguousarray(kappa, dtype=np.float64)
        if self.kappa.shape != (int(dim),):
            raise _lib.DimensionMismatch(f"kappa must have length dim ({dim}). Current length: {self.kappa.size}")
        super().__init__(dim, potential, grid_size=grid_size, tmax=tmax, refresh_rate=refresh_rate,
                         vectorized_bound=vectorized_bound, signed_bound=signed_bound, adaptive=adaptive,
                         AD_backend=AD_backend, max_steps=max_steps)


def StickyZigZagAD(dim, potential, kappa, *, refresh_rate=0.0, grid_size=10, tmax=2.0, vectorized_bound=True,
                   signed_bound=True, adaptive=True, AD_backend="ForwardDiff", max_steps=0):
    return StickyZigZag(dim, potential, kappa, refresh_rate=refresh_rate, grid_size=grid_size, tmax=tmax,
                        vectorized_bound=vectorized_bound, signed_bound=signed_bound, adaptive=adaptive,
                        AD_backend=AD_backend, max_steps=max_steps)


class SpeedUpZigZag(AbstractPDMP):
    """SpeedUpZigZag(dim, grad U; kw...) (SpeedUpZigZagSamplers.jl:58-116): Zig-Zag with the position-dependent speed
    sqrt(1 + |x|^2) -- the closed-form nonlinear flow of :71-79 and the effective gradient of :81-83."""
    _kind = SPEEDUP_ZIGZAG
    flow_kind = 2

    def __init__(self, dim, potential, *, grid_size=10, tmax=2.0, refresh_rate=0.0, vectorized_bound=True,
                 signed_bound=True, adaptive=True, AD_backend="FiniteDiff", max_steps=0):
        if signed_bound and not vectorized_bound:
            warnings.warn("Signed bound is not compatible with non-vectorized bound for ZigZag, switching to unsigned bound")
        super().__init__(dim, potential, grid_size=grid_size, tmax=tmax, refresh_rate=refresh_rate,
                         vectorized_bound=vectorized_bound, signed_bound=signed_bound, adaptive=adaptive,
                         AD_backend=AD_backend, max_steps=max_steps)


def SpeedUpZigZagAD(dim, potential, *, refresh_rate=0.0, grid_size=10, tmax=2.0, vectorized_bound=True, signed_bound=True,
                    adaptive=True, AD_backend="ForwardDiff", max_steps=0):
    return SpeedUpZigZag(dim, potential, refresh_rate=refresh_rate, grid_size=grid_size, tmax=tmax,
                         vectorized_bound=vectorized_bound, signed_bound=signed_bound, adaptive=adaptive,
                         AD_backend=AD_backend, max_steps=max_steps)
