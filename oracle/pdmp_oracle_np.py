"""
TEST INFRASTRUCTURE ONLY -- numpy restatement of PDMPFlux.jl's grid-based Poisson-thinning path.

This file is an *oracle*: only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline leg
may import it.  The product (`libpdmpflux_cuda.so` + `pdmpflux_b200`) never does.

PARITY UNPINNED: Julia is not installed in this image and the reference ships no golden skeleton
vectors, so this restatement cannot be checked against the reference's own
output here.  It is pinned instead by (i) hand-derived known-answer tests (tests/test_oracle_kat.py),
(ii) agreement with an independent C restatement (oracle/pdmp_oracle.c) and (iii) the statistical
envelopes the reference's tests assert.  `tools/record_tape.jl` records a draw tape + PDMPHistory
from the real package for anyone with Julia.

The style is deliberately literal: array expressions mirror the Julia broadcasts line by line,
including the reference's quirks (see SURVEY.md section 8a "gotchas").  All citations are relative to
/root/reference/.

Random draws never come from an RNG here: a `Tape` supplies three typed streams (E = randexp,
U = rand, N = randn) which are consumed in the order the reference consumes them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)
EPS = float(np.finfo(np.float64).eps)

ZIGZAG, BPS, FECMC, BOOMERANG = 0, 1, 2, 3
DERIV_JVP, DERIV_FD = 0, 1


# --------------------------------------------------------------------------------------------
# draw tape
# --------------------------------------------------------------------------------------------
class TapeExhausted(RuntimeError):
    pass


class Tape:
    """Three typed draw streams consumed in reference order ('RNG draw order')."""

    def __init__(self, E, U, N):
        self.E = np.asarray(E, dtype=np.float64)
        self.U = np.asarray(U, dtype=np.float64)
        self.N = np.asarray(N, dtype=np.float64)
        self.pos = [0, 0, 0]

    def randexp(self):
        if self.pos[0] >= self.E.size:
            raise TapeExhausted("E")
        r = self.E[self.pos[0]]
        self.pos[0] += 1
        return float(r)

    def rand(self):
        if self.pos[1] >= self.U.size:
            raise TapeExhausted("U")
        r = self.U[self.pos[1]]
        self.pos[1] += 1
        return float(r)

    def randn(self, n):
        if self.pos[2] + n > self.N.size:
            raise TapeExhausted("N")
        r = self.N[self.pos[2]:self.pos[2] + n].copy()
        self.pos[2] += n
        return r


def make_tape(seed, nE, nU, nN):
    g = np.random.default_rng(seed)
    return Tape(g.standard_exponential(nE), g.random(nU), g.standard_normal(nN))


# --------------------------------------------------------------------------------------------
# potentials: grad(x) and hvp(x, v) = H(x) v      
# --------------------------------------------------------------------------------------------
class GaussStd:
    """U = sum(x.^2)/2   (README.md:36-38)"""
    kind = 0

    def value(self, x):
        return float(np.dot(x, x)) / 2.0

    def grad(self, x):
        return x.copy()

    def hvp(self, x, v):
        return v.copy()


class GaussDiag:
    """U = sum(p .* x.^2)/2"""
    kind = 1

    def __init__(self, prec):
        self.p = np.asarray(prec, dtype=np.float64)

    def value(self, x):
        return float(np.sum(self.p * x * x)) / 2.0

    def grad(self, x):
        return self.p * x

    def hvp(self, x, v):
        return self.p * v


class GaussEquicorr:
    """'slanted' Gaussian: Sigma = (1-rho) I + rho 1 1^T; P = alpha I - beta 1 1^T (Sherman-Morrison).
    Not defined in the reference; defined here."""
    kind = 2

    def __init__(self, d, rho):
        self.alpha = 1.0 / (1.0 - rho)
        self.beta = rho / ((1.0 - rho) * (1.0 - rho + d * rho))

    def value(self, x):
        s = float(np.sum(x))
        return (self.alpha * float(np.dot(x, x)) - self.beta * s * s) / 2.0

    def grad(self, x):
        return self.alpha * x - self.beta * np.sum(x)

    def hvp(self, x, v):
        return self.alpha * v - self.beta * np.sum(v)


class Banana:
    """U = (x1^2 + (x2 - x1^2 + 1)^2 + sum_{i>=3} x_i^2)/2   (test/test_config.jl:33-36)"""
    kind = 3

    def value(self, x):
        r = x[1] - x[0] * x[0] + 1.0
        return (x[0] * x[0] + r * r + float(np.dot(x[2:], x[2:]))) / 2.0

    def grad(self, x):
        g = x.copy()
        r = x[1] - x[0] * x[0] + 1.0
        g[0] = x[0] - 2.0 * x[0] * r
        g[1] = r
        return g

    def hvp(self, x, v):
        h = v.copy()
        r = x[1] - x[0] * x[0] + 1.0
        h[0] = (1.0 - 2.0 * r + 4.0 * x[0] * x[0]) * v[0] - 2.0 * x[0] * v[1]
        h[1] = -2.0 * x[0] * v[0] + v[1]
        return h


class BananaReadmeScalar:
    """README.md:62-65: the manual 'gradient' returns a scalar that `.*` broadcasts to every coordinate."""
    kind = 4

    value = Banana.value  # README.md:56-60: the same U; only the hand-written 'gradient' differs

    def grad(self, x):
        s = x[0] + (x[1] - (x[0] * x[0] - 1.0)) + np.sum(x[2:])
        return np.full_like(x, s)

    def hvp(self, x, v):
        s = v[0] + (v[1] - 2.0 * x[0] * v[0]) + np.sum(v[2:])
        return np.full_like(x, s)


class LogReg:
    """U(theta) = sum_j [log(1+exp(z_j)) - y_j z_j] + |theta|^2/(2 sigma0^2), z = X theta.
    Not in the reference; defined by BASELINE.json config 4 / SURVEY.md section 8d (C4)."""
    kind = 5

    def __init__(self, X, y, sigma0):
        self.X = np.asarray(X, dtype=np.float64)  # n x d, row-major
        self.y = np.asarray(y, dtype=np.float64)
        self.inv_s2 = 1.0 / (sigma0 * sigma0)

    def value(self, x):
        z = self.X @ x
        return float(np.sum(np.logaddexp(0.0, z) - self.y * z) + 0.5 * self.inv_s2 * np.dot(x, x))

    @staticmethod
    def _sigmoid(z):
        return 1.0 / (1.0 + np.exp(-z))

    def grad(self, x):
        z = self.X @ x
        return self.X.T @ (self._sigmoid(z) - self.y) + x * self.inv_s2

    def hvp(self, x, v):
        z = self.X @ x
        s = self._sigmoid(z)
        return self.X.T @ (s * (1.0 - s) * (self.X @ v)) + v * self.inv_s2


# --------------------------------------------------------------------------------------------
# config / sampler definition
# --------------------------------------------------------------------------------------------
@dataclass
class Config:
    sampler: int = ZIGZAG
    grid_size: int = 10
    tmax: float = 2.0
    refresh_rate: float = 0.0
    vectorized_bound: bool = True
    signed_bound: bool = True
    adaptive: bool = True
    deriv_mode: int = DERIV_JVP
    gaussian_velocity: bool = False  # BPS
    ran_p: bool = False              # FECMC
    mix_p: float = 0.5
    switch: bool = True
    positive: bool = True
    speed_factor: float = 1.0

    def normalised(self, dim):
        """Constructor-side rewrites the reference applies (ZigZagSamplers.jl:73-78,
        BouncyParticleSamplers.jl:29-37, ForwardEventChainMonteCarlo.jl:306-323, BoomerangSamplers.jl:27-36)."""
        c = Config(**self.__dict__)
        if c.tmax == 0.0:
            c.tmax = 1.0
            c.adaptive = True
        if c.sampler == ZIGZAG:
            if c.signed_bound and not c.vectorized_bound:
                c.signed_bound = False
        else:
            c.vectorized_bound = False
        if c.sampler == FECMC:
            c.refresh_rate = 0.0
            if dim == 2:
                c.mix_p = 0.0
        return c


class Sampler:
    """The closures of src/Samplers/*.jl for one (sampler kind, potential)."""

    def __init__(self, dim, pot, cfg: Config):
        self.dim = dim
        self.pot = pot
        self.cfg = cfg.normalised(dim)
        self.kind = self.cfg.sampler

    # flow: ZigZagSamplers.jl:80, BouncyParticleSamplers.jl:32, ForwardEventChainMonteCarlo.jl:316,
    #       BoomerangSamplers.jl:31
    def flow(self, x, v, t):
        if self.kind == BOOMERANG:
            return x * math.cos(t) + v * math.sin(t), -x * math.sin(t) + v * math.cos(t)
        return x + v * t, v

    # unsigned scalar rate `sampler.rate`
    def rate(self, x0, v0, t):
        xt, vt = self.flow(x0, v0, t)
        g = self.pot.grad(xt)
        if self.kind == ZIGZAG:  # ZigZagSamplers.jl:83-86
            return float(np.sum(np.maximum(0.0, g * vt)))
        if self.kind == FECMC:   # ForwardEventChainMonteCarlo.jl:20-23
            return max(0.0, float(np.dot(g, vt)))
        # BPS :39-42 / Boomerang :38-41
        return max(0.0, float(np.dot(g, vt))) + self.cfg.refresh_rate

    # ---- (value, d/dt) of the function the bound is built from, analytic (ForwardDiff-equivalent) ----
    def _dflow(self, xt, vt):
        """d/dt of (xt, vt) along the flow."""
        if self.kind == BOOMERANG:
            return vt, -xt
        return vt, np.zeros_like(vt)

    def bound_func_scalar(self, x0, v0, t, signed):
        """Returns (val, dval/dt) of `signed_rate` or `rate` (AbstractPDMP.jl:104-112, 127-130)."""
        xt, vt = self.flow(x0, v0, t)
        dx, dv = self._dflow(xt, vt)
        g = self.pot.grad(xt)
        hd = self.pot.hvp(xt, dx)
        if self.kind == ZIGZAG:
            # unsigned scalar ZigZag bound (vectorized_bound=false): sum(max.(0, g.*v))
            y = g * vt
            dy = hd * vt + g * dv
            mask = ~(0.0 > y)
            return float(np.sum(np.maximum(0.0, y))), float(np.sum(np.where(mask, dy, 0.0)))
        y = float(np.dot(g, vt))
        dy = float(np.dot(hd, vt) + np.dot(g, dv))
        extra = 0.0 if self.kind == FECMC else self.cfg.refresh_rate
        if signed:
            return y + extra, dy
        return max(0.0, y) + extra, (dy if not (0.0 > y) else 0.0)

    def bound_func_vect(self, x0, v0, t, signed):
        """ZigZag `signed_rate_vect` / `rate_vect` (ZigZagSamplers.jl:88-98) and its d/dt."""
        xt, vt = self.flow(x0, v0, t)
        g = self.pot.grad(xt)
        hv = self.pot.hvp(xt, vt)
        y = g * vt
        dy = hv * vt
        if signed:
            return y, dy
        mask = ~(0.0 > y)
        return np.maximum(0.0, y), np.where(mask, dy, 0.0)

    # ---- velocity jumps ----
    def velocity_jump(self, x, v, tape: Tape):
        k = self.kind
        if k == ZIGZAG:
            return self._jump_zigzag(x, v, tape)
        if k == BPS:
            return self._jump_bps(x, v, tape)
        if k == FECMC:
            return self._jump_fecmc(x, v, tape)
        return self._jump_boomerang(x, v, tape)

    def _jump_zigzag(self, x, v, tape):
        # ZigZagSamplers.jl:101-107 + Distributions.jl DiscreteNonParametric rand (linear CDF scan)
        lam = np.maximum(0.0, self.pot.grad(x) * v)
        p = lam / np.sum(lam)
        sp = float(np.sum(p))
        if not (np.all(p >= 0.0) and abs(sp - 1.0) <= SQRT_EPS * max(abs(sp), 1.0)):
            # Categorical's constructor (isprobvec) rejects NaN / non-normalised p (sum(lam) == 0)
            raise FloatingPointError("ZigZag jump: rate vector is not a probability vector")
        u = tape.rand()
        n = len(p)
        cp = p[0]
        i = 0
        while cp <= u and i < n - 1:
            i += 1
            cp += p[i]
        v = v.copy()
        v[i] *= -1
        return v

    def _jump_bps(self, x, v, tape):
        # BouncyParticleSamplers.jl:50-74
        g = self.pot.grad(x)
        bounce_rate = max(0.0, float(np.dot(g, v)))
        bounce_prob = bounce_rate / (bounce_rate + self.cfg.refresh_rate)
        u = tape.rand()
        if u < bounce_prob:
            gg = float(np.dot(g, g))
            if gg == 0:
                return v
            scale = 2 * float(np.dot(v, g)) / gg
            return v - scale * g
        vn = tape.randn(self.dim)
        if self.cfg.gaussian_velocity:
            return vn
        return vn / math.sqrt(float(np.dot(vn, vn)))

    def _jump_boomerang(self, x, v, tape):
        # BoomerangSamplers.jl:49-67 (jump uses grad U - x although the rate uses grad U)
        g = self.pot.grad(x) - x
        bounce_rate = max(0.0, float(np.dot(g, v)))
        with np.errstate(invalid="ignore", divide="ignore"):
            bounce_prob = np.float64(bounce_rate) / np.float64(bounce_rate + self.cfg.refresh_rate)
        u = tape.rand()
        if u < bounce_prob:
            e = g / math.sqrt(float(np.dot(g, g)))
            return v - 2 * float(np.dot(v, e)) * e
        return tape.randn(self.dim)

    def _orthogonal_switch(self, vo, n, tape):
        # ForwardEventChainMonteCarlo.jl:60-88; randn(key, 2, dim) is column-major: g[1,1], g[2,1], g[1,2] ...
        d = self.dim
        g = tape.randn(2 * d)
        g1 = g[0::2].copy()
        g2 = g[1::2].copy()
        g1 = g1 - np.dot(g1, n) * n
        g2 = g2 - np.dot(g2, n) * n
        e1 = g1 / math.sqrt(float(np.dot(g1, g1)))
        e2 = g2 - np.dot(g2, e1) * e1
        e2 = e2 / math.sqrt(float(np.dot(e2, e2)))
        vr = vo - np.dot(vo, e1) * e1 - np.dot(vo, e2) * e2
        vo_new = vr + e2 * np.dot(e1, vo) + e1 * np.dot(e2, vo)
        if self.cfg.ran_p:
            th = tape.rand() * 2 * math.pi
            vo_new = (vr + (math.cos(th) * e1 + math.sin(th) * e2) * np.dot(e1, vo)
                      + (math.sin(th) * e1 - math.cos(th) * e2) * np.dot(e2, vo))
        if self.cfg.positive:
            vo_new = vo_new * np.sign(np.dot(vo, vo_new))
        return vo_new

    def _full_refresh(self, n, tape):
        # ForwardEventChainMonteCarlo.jl:105-113
        w = tape.randn(self.dim)
        w = w / math.sqrt(float(np.dot(w, w)))
        return w - np.dot(w, n) * n

    def _jump_fecmc(self, x, v, tape):
        # ForwardEventChainMonteCarlo.jl:132-176 (speed_factor == 1) and :178-218 (speed-up variant)
        d = self.dim
        sf = self.cfg.speed_factor
        u = tape.rand()
        if sf != 1.0:
            rho = sf * -math.sqrt(1 - u ** (2 / (d - 1)))
        else:
            rho = -math.sqrt(1 - u ** (2 / (d - 1)))
        n = self.pot.grad(x).copy()
        ng = math.sqrt(float(np.dot(n, n)))
        if ng == 0:
            n[:] = 0.0
        else:
            n /= ng
        vp = np.dot(v, n) * n
        vo = v - vp
        if math.sqrt(float(np.dot(vo, vo))) < 1e-10:
            vo = tape.randn(d)
            vo = vo - np.dot(vo, n) * n
        u2 = tape.rand()
        rad = math.sqrt(sf * sf - rho * rho) if sf != 1.0 else math.sqrt(1 - rho * rho)
        if u2 >= self.cfg.mix_p:
            return vo / math.sqrt(float(np.dot(vo, vo))) * rad + rho * n
        prop = self._orthogonal_switch(vo, n, tape) if self.cfg.switch else self._full_refresh(n, tape)
        with np.errstate(invalid="ignore", divide="ignore"):
            return prop / np.float64(math.sqrt(float(np.dot(prop, prop)))) * rad + rho * n


# --------------------------------------------------------------------------------------------
# upper bounds  (src/UpperBound.jl)
# --------------------------------------------------------------------------------------------
@dataclass
class BoundBox:  # Composites.jl:15-20
    grid: np.ndarray
    box_max: np.ndarray
    cum_sum: np.ndarray
    step_size: float


def grid_times(horizon, G):
    """range(0, stop=horizon, length=G) (UpperBound.jl:94,204).  Julia evaluates the nodes in
    TwicePrecision, i.e. (almost) correctly rounded k*horizon/(G-1) with the last node == horizon.
    Restated as t_k = RN(k*horizon/(G-1)) (exact rational arithmetic, one rounding), t_{G-1} = horizon."""
    from fractions import Fraction
    t = np.empty(G)
    if not math.isfinite(horizon):
        t[:] = np.arange(G) * (horizon / (G - 1))
        return t
    hf = Fraction(horizon)
    for k in range(G):
        t[k] = float(hf * k / (G - 1))
    t[G - 1] = horizon
    return t


def brent_minimum(f, lo, hi):
    """Optim.jl `optimize(f, lo, hi, Brent())`, defaults rel_tol=sqrt(eps), abs_tol=eps, 1000 iterations.
    Third-party (Optim.jl compat "1.9.4, 2", not vendored): restated from the published algorithm,
    call site src/UpperBound.jl:24.  Returns the minimum value found."""
    golden = 0.5 * (3.0 - math.sqrt(5.0))
    x = lo + golden * (hi - lo)
    fx = f(x)
    step = 0.0
    old_step = 0.0
    w = x
    vv = x
    fw = fx
    fv = fx
    it = 0
    while it < 1000:
        p = 0.0
        q = 0.0
        tol = SQRT_EPS * abs(x) + EPS
        mid = (hi + lo) / 2
        if abs(x - mid) <= 2 * tol - (hi - lo) / 2:
            break
        it += 1
        if abs(old_step) > tol:
            r = (x - w) * (fx - fv)
            q = (x - vv) * (fx - fw)
            p = (x - vv) * q - (x - w) * r
            q = 2 * (q - r)
            if q > 0:
                p = -p
            else:
                q = -q
        if abs(p) < abs(q * old_step / 2) and p < q * (hi - x) and p < q * (x - lo):
            old_step = step
            step = p / q
            xt = x + step
            if (xt - lo) < 2 * tol or (hi - xt) < 2 * tol:
                step = tol if x < mid else -tol
        else:
            old_step = (hi - x) if x < mid else (lo - x)
            step = golden * old_step
        if abs(step) >= tol:
            u = x + step
        else:
            u = x + (tol if step > 0 else -tol)
        fu = f(u)
        if fu < fx:
            if u < x:
                hi = x
            else:
                lo = x
            vv, fv = w, fw
            w, fw = x, fx
            x, fx = u, fu
        else:
            if u < x:
                lo = u
            else:
                hi = u
            if fu <= fw or w == x:
                vv, fv = w, fw
                w, fw = u, fu
            elif fu <= fv or vv == x or vv == w:
                vv, fv = u, fu
    return fx


def upper_bound_constant(func, start, horizon, refresh_rate=0.0):
    """UpperBound.jl:18-36"""
    m = brent_minimum(lambda t: -func(t), start, horizon)
    t = np.array([start, horizon])
    box_max = np.array([-m])
    box_max[0] += refresh_rate
    cum_sum = np.zeros(2)
    cum_sum[1] = box_max[0] * (horizon - start)
    return BoundBox(t, box_max, cum_sum, horizon - start)


def finite_difference_derivative(func, x, start, horizon):
    """UpperBound.jl:50-76 (func returns a scalar or a vector)."""
    fx = func(x)
    h = SQRT_EPS * max(1.0, abs(x))
    x_minus = max(start, x - h)
    x_plus = min(horizon, x + h)
    if x_plus == x_minus:
        return fx - fx
    if x_minus == x:
        return (func(x_plus) - fx) / (x_plus - x)
    elif x_plus == x:
        return (fx - func(x_minus)) / (x - x_minus)
    return (func(x_plus) - func(x_minus)) / (x_plus - x_minus)


def upper_bound_grid(func_vd, horizon, n_grid, refresh_rate, deriv_mode):
    """UpperBound.jl:92-137.  func_vd(t) -> (value, analytic d/dt)."""
    t = grid_times(horizon, n_grid)
    step_size = t[1] - t[0]
    values = np.array([func_vd(tk)[0] for tk in t])
    if deriv_mode == DERIV_JVP:
        grads = np.array([func_vd(tk)[1] for tk in t])
    else:
        grads = np.array([finite_difference_derivative(lambda s: func_vd(s)[0], float(tk), 0.0, horizon)
                          for tk in t])
    with np.errstate(invalid="ignore", divide="ignore"):
        pos = (values[:-1] - values[1:] + grads[1:] * step_size) / (grads[1:] - grads[:-1])
    pos = np.where(np.isnan(pos), 0.0, pos)
    pos = np.minimum(np.maximum(pos, 0.0), step_size)
    inter = values[:-1] + grads[:-1] * pos
    box_max = np.maximum(values[:-1], values[1:])
    box_max = np.maximum(box_max, inter)
    box_max = np.maximum(box_max, 0.0)
    box_max = box_max + refresh_rate
    cum_sum = np.zeros(n_grid)
    cum_sum[1:] = np.cumsum(box_max) * step_size
    return BoundBox(t, box_max, cum_sum, step_size)


def upper_bound_grid_vect(func_vd, horizon, n_grid, deriv_mode):
    """UpperBound.jl:203-247.  func_vd(t) -> (d-vector value, d-vector analytic d/dt).
    NOTE the reference's intersection position is an ABSOLUTE time (uses t[k+1], t[k]) that is then
    clamped to [0, step] and used as an OFFSET from the left node; reproduced literally."""
    t = grid_times(horizon, n_grid)
    step_size = t[1] - t[0]
    values = np.stack([func_vd(tk)[0] for tk in t], axis=1)  # d x G
    if deriv_mode == DERIV_JVP:
        grads = np.stack([func_vd(tk)[1] for tk in t], axis=1)
    else:
        grads = np.stack([finite_difference_derivative(lambda s: func_vd(s)[0], float(tk), 0.0, horizon)
                          for tk in t], axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        pos = (values[:, :-1] - values[:, 1:] + grads[:, 1:] * t[None, 1:] - grads[:, :-1] * t[None, :-1]) \
            / (grads[:, 1:] - grads[:, :-1])
    pos = np.where(np.isnan(pos), 0.0, pos)
    pos = np.minimum(np.maximum(pos, 0.0), step_size)
    inter = values[:, :-1] + grads[:, :-1] * pos
    box_max = np.maximum(values[:, :-1], values[:, 1:])
    box_max = np.maximum(box_max, inter)
    box_max = np.maximum(box_max, 0.0)
    cum_sum = np.zeros(values.shape)
    cum_sum[:, 1:] = np.cumsum(box_max, axis=1) * step_size
    return BoundBox(t, np.sum(box_max, axis=0), np.sum(cum_sum, axis=0), step_size)


def next_event(bb: BoundBox, exp_rv):
    """UpperBound.jl:264-273; searchsortedfirst = first index with cum_sum[idx] >= exp_rv (1-based)."""
    cs = bb.cum_sum
    n = len(cs)
    idx = int(np.searchsorted(cs, exp_rv, side="left")) + 1  # 1-based
    if idx > n:
        return math.inf, float(bb.box_max[-1])
    tp = bb.grid[idx - 2] + (exp_rv - cs[idx - 2]) / (cs[idx - 1] - cs[idx - 2]) * bb.step_size
    return float(tp), float(bb.box_max[idx - 2])


# --------------------------------------------------------------------------------------------
# state, history, thinning loop
# --------------------------------------------------------------------------------------------
@dataclass
class State:  # Composites.jl:59-83 (non-sticky subset)
    x: np.ndarray
    v: np.ndarray
    t: float
    horizon: float
    adaptive: bool
    accept: bool = False
    upper_bound: BoundBox | None = None
    tp: float = 0.0
    ts: float = 0.0
    exp_rv: float = 0.0
    lambda_bar: float = 0.0
    lambda_t: float = 0.0
    ar: float = 0.0
    errored_bound: int = 0
    error_value_ar: np.ndarray = field(default_factory=lambda: np.zeros(5))
    rejected: int = 0
    hitting_horizon: int = 0
    # instrumentation (not in the reference)
    n_bound_builds: int = 0
    n_rate_evals: int = 0


class History:  # Composites.jl:138-164
    def __init__(self, d, n):
        self.X = np.full((d, n), np.nan, order="F")
        self.V = np.full((d, n), np.nan, order="F")
        self.t = np.full(n, np.nan)
        self.horizon = np.full(n, np.nan)
        self.ar = np.full(n, np.nan)
        self.errored_bound = np.zeros(n, dtype=np.int32)
        self.error_value_ar = np.zeros((5, n), order="F")
        self.rejected = np.zeros(n, dtype=np.int32)
        self.hitting_horizon = np.zeros(n, dtype=np.int32)

    def record(self, k, s: State):  # Composites.jl:239-260
        self.X[:, k] = s.x
        self.V[:, k] = s.v
        self.t[k] = s.t
        self.horizon[k] = s.horizon
        self.ar[k] = s.ar
        self.errored_bound[k] = s.errored_bound
        self.error_value_ar[:, k] = s.error_value_ar
        self.rejected[k] = s.rejected
        self.hitting_horizon[k] = s.hitting_horizon


class Chain:
    """init_state + get_event_state! for one chain (AbstractPDMP.jl:93-153, SamplingLoopInplace.jl:27-217)."""

    def __init__(self, sampler: Sampler, xinit, vinit, tape: Tape):
        self.s = sampler
        self.tape = tape
        cfg = sampler.cfg
        if len(xinit) != sampler.dim or len(vinit) != sampler.dim:
            raise ValueError("DimensionMismatch")
        if cfg.grid_size < 0 or cfg.grid_size == 1:
            raise ValueError("grid_size must be 0 or >= 2")
        self.signed = cfg.signed_bound
        self.bound_refresh = cfg.refresh_rate if cfg.signed_bound else 0.0  # AbstractPDMP.jl:104-112
        self.state = State(np.array(xinit, dtype=np.float64), np.array(vinit, dtype=np.float64),
                           0.0, cfg.tmax, cfg.adaptive)

    # AbstractPDMP.jl:121-136
    def upper_bound_func(self, x, v, horizon):
        s, cfg = self.s, self.s.cfg
        self.state.n_bound_builds += 1
        if cfg.grid_size == 0:
            return upper_bound_constant(lambda t: s.rate(x, v, t), 0.0, horizon)
        if not cfg.vectorized_bound:
            return upper_bound_grid(lambda t: s.bound_func_scalar(x, v, t, self.signed), horizon,
                                    cfg.grid_size, self.bound_refresh, cfg.deriv_mode)
        return upper_bound_grid_vect(lambda t: s.bound_func_vect(x, v, t, self.signed), horizon,
                                     cfg.grid_size, cfg.deriv_mode)

    def get_event_state(self):  # SamplingLoopInplace.jl:27-39
        st = self.state
        st.errored_bound = 0
        st.rejected = 0
        st.hitting_horizon = 0
        st.error_value_ar = np.zeros(5)
        while not st.accept:
            self.one_step_of_thinning()
        st.accept = False
        return st

    def one_step_of_thinning(self):  # :65-85
        st = self.state
        ub = self.upper_bound_func(st.x, st.v, st.horizon)
        e = self.tape.randexp()
        tp, lb = next_event(ub, e)
        st.tp, st.exp_rv, st.lambda_bar, st.upper_bound = tp, e, lb, ub
        if tp > st.horizon:
            self.move_to_horizon()
        else:
            self.moves_until_horizon()

    def move_to_horizon(self):  # :87-101
        st = self.state
        st.x, st.v = self.s.flow(st.x, st.v, st.horizon)
        st.ts += st.horizon
        st.hitting_horizon += 1
        st.horizon = st.horizon * 1.01 if st.adaptive else st.horizon

    def moves_until_horizon(self):  # :103-111
        st = self.state
        st.accept = False
        while st.tp < st.horizon and not st.accept:
            self.ac_step()

    def ac_step(self):  # :113-129
        st = self.state
        st.n_rate_evals += 1
        lt = self.s.rate(st.x, st.v, st.tp)
        with np.errstate(invalid="ignore", divide="ignore"):
            ar = float(np.float64(lt) / np.float64(st.lambda_bar))
        st.lambda_t, st.ar = lt, ar
        if ar > 1.0:
            self.erroneous_acceptance_rate()
        else:
            self.ac_step_with_proxy()

    def erroneous_acceptance_rate(self):  # :131-151
        st = self.state
        horizon = st.horizon / 2
        ub = self.upper_bound_func(st.x, st.v, horizon)
        e = self.tape.randexp()
        tp, lb = next_event(ub, e)
        st.horizon = horizon if st.adaptive else st.horizon
        st.tp, st.exp_rv, st.lambda_bar, st.upper_bound = tp, e, lb, ub
        st.errored_bound += 1
        st.error_value_ar[st.errored_bound % 5] = st.ar  # 1-based `% 5 + 1`

    def ac_step_with_proxy(self):  # :153-168
        st = self.state
        accept = self.tape.rand() < st.ar
        st.accept = accept
        if accept:
            self.if_accept()
        else:
            self.if_reject()
        if (not st.accept) and (st.tp > st.horizon):  # min(tp, tt=Inf)
            self.move_to_horizon2()

    def if_accept(self):  # :170-186
        st = self.state
        st.x, st.v = self.s.flow(st.x, st.v, st.tp)
        st.v = self.s.velocity_jump(st.x, st.v, self.tape)
        st.t = st.t + st.tp + st.ts
        st.ts = 0.0
        st.tp = 0.0
        st.accept = True

    def if_reject(self):  # :188-203
        st = self.state
        e = st.exp_rv + self.tape.randexp()
        tp, lb = next_event(st.upper_bound, e)
        st.horizon = st.horizon / 1.04 if st.adaptive else st.horizon
        st.tp, st.exp_rv, st.lambda_bar = tp, e, lb
        st.rejected += 1

    def move_to_horizon2(self):  # :205-217
        st = self.state
        st.x, st.v = self.s.flow(st.x, st.v, st.horizon)
        st.ts += st.horizon
        st.hitting_horizon += 1


def sample_skeleton(sampler: Sampler, n_sk, xinit, vinit, tape: Tape):
    """src/sample.jl:253-284: column 0 = initial state, columns 1..n_sk-1 = successive events."""
    if n_sk <= 0:
        raise ValueError("n_sk must be positive")
    ch = Chain(sampler, xinit, vinit, tape)
    h = History(sampler.dim, n_sk)
    h.record(0, ch.state)
    for k in range(1, n_sk):
        st = ch.get_event_state()
        h.record(k, st)
    h.final_state = ch.state
    h.tape_pos = list(tape.pos)
    return h


def sample_from_skeleton(sampler_kind, N, X, V, t, discard_vt=True):
    """src/sample.jl:475-513 (uses only sampler.flow)."""
    if N <= 0:
        raise ValueError("N must be positive")
    d, Nh = X.shape
    dt = t[-1] / N
    out = np.empty((d if discard_vt else 2 * d + 1, N), order="F")
    i = 0
    for j in range(1, N + 1):
        tm = j * dt
        while i < Nh - 1 and t[i + 1] <= tm:
            i += 1
        tau = tm - t[i]
        if sampler_kind == BOOMERANG:
            xn = X[:, i] * math.cos(tau) + V[:, i] * math.sin(tau)
            vn = -X[:, i] * math.sin(tau) + V[:, i] * math.cos(tau)
        else:
            xn = X[:, i] + V[:, i] * tau
            vn = V[:, i]
        out[:d, j - 1] = xn
        if not discard_vt:
            out[d:2 * d, j - 1] = vn
            out[2 * d, j - 1] = tm
    return out


def rv_diagnostic(X, V, t, U, B=0):
    """RV_diagnostic(history, U; B) (src/diagnostic.jl:37-75): realised volatility of U along the skeleton,
    sampled at B+1 equidistant boundaries with the *linear* interpolation of _history_position_linear!
    (src/diagnostic.jl:23-35; is_active all true for the non-sticky samplers).  X, V are (d, n) like the Julia
    matrices, U a callable on a d-vector."""
    N = t.shape[0]
    if N == 0:
        return 0.0
    T = float(t[-1])
    if not math.isfinite(T) or T < 0.0:
        raise ValueError("history.t[end] must be finite and non-negative")
    if B == 0:
        B = max(1, int(math.floor(math.sqrt(X.shape[1]))))
    elif B < 0:
        raise ValueError("B must be non-negative")
    if T == 0.0:
        return 0.0
    boundaries = grid_times(T, B + 1)      # range(0.0, T; length=B+1)
    x_left = X[:, 0].copy()
    RV = 0.0
    i = 0
    for b in range(1, B + 1):
        tb = boundaries[b]
        while i < N - 1 and t[i + 1] <= tb:
            i += 1
        tau = tb - t[i]
        x_right = X[:, i] + V[:, i] * tau
        inc = U(x_right) - U(x_left)
        RV += inc * inc
        x_left = x_right
    return RV / T


# --------------------------------------------------------------------------------------------
# Sticky Zig-Zag: literal restatement of src/StickySamplingLoop.jl:13-164 on top of the
# masked-velocity variants of the thinning steps (src/SamplingLoopInplace.jl:13-25, 87-217) and the
# StickyZigZag closures (src/Samplers/StickyZigZagSamplers.jl:69-101, identical to Zig-Zag's).  Oracle only so
# far: the CUDA path for it is the next row of the scope table.

# --------------------------------------------------------------------------------------------
# Speed-Up Zig-Zag (src/Samplers/SpeedUpZigZagSamplers.jl:35-130): Zig-Zag with the position-dependent speed
# s(x) = sqrt(1 + |x|^2), i.e. the closed-form nonlinear flow of :71-79 and the effective gradient
# grad U_eff = s grad U - grad s (:81-83).  Everything else (bounds, thinning loop, categorical flip) is the generic
# machinery with these closures.  The reference obtains d/dt of the rate by ForwardDiff through the flow; here it is
# taken by complex-step differentiation of the very same expressions (exact to rounding for analytic functions), so
# that the oracle's derivative is independent of the hand-derived formulas of the device code.
# --------------------------------------------------------------------------------------------
class SpeedUpSampler(Sampler):
    def __init__(self, dim, pot, cfg: Config):
        super().__init__(dim, pot, cfg)
        if self.kind != ZIGZAG:
            raise ValueError("SpeedUpSampler restates SpeedUpZigZag only")

    def flow(self, x, v, t):  # SpeedUpZigZagSamplers.jl:71-79 (literal)
        d = self.dim
        y = x - v[0] * x[0] * v
        c = v[0] * np.dot(y, v)
        a = (1 + np.dot(y, y)) / d - (c ** 2) / (d ** 2)
        Y_0 = x[0] + (c / d)
        b_t = (Y_0 + np.sqrt(Y_0 ** 2 + a)) * np.exp(math.sqrt(d) * v[0] * t)
        X_1 = (b_t ** 2 - a) / (2 * b_t) - (c / d)
        return y + v[0] * X_1 * v, v

    def grad_eff(self, x):  # :81-83
        speed = np.sqrt(1.0 + np.dot(x, x))
        return speed * self.pot.grad(x) - x / speed

    def rate(self, x0, v0, t):  # :86-89
        xt, vt = self.flow(x0, v0, t)
        return float(np.sum(np.maximum(0.0, self.grad_eff(xt) * vt)))

    def _signed_vect_complex(self, x0, v0, t):
        h = 1e-30
        xt, vt = self.flow(x0.astype(complex), v0.astype(complex), complex(t, h))
        y = self.grad_eff(xt) * vt
        return y.real.copy(), y.imag / h

    def bound_func_vect(self, x0, v0, t, signed):  # :91-99 + d/dt
        y, dy = self._signed_vect_complex(x0, v0, t)
        if signed:
            return y, dy
        mask = ~(0.0 > y)
        return np.maximum(0.0, y), np.where(mask, dy, 0.0)

    def bound_func_scalar(self, x0, v0, t, signed):  # vectorized_bound = false: the scalar unsigned rate
        y, dy = self._signed_vect_complex(x0, v0, t)
        mask = ~(0.0 > y)
        return float(np.sum(np.maximum(0.0, y))), float(np.sum(np.where(mask, dy, 0.0)))

    def velocity_jump(self, x, v, tape: Tape):  # :102-108
        lam = np.maximum(0.0, self.grad_eff(x) * v)
        p = lam / np.sum(lam)
        sp = float(np.sum(p))
        if not (np.all(p >= 0.0) and abs(sp - 1.0) <= SQRT_EPS * max(abs(sp), 1.0)):
            raise FloatingPointError("SpeedUpZigZag jump: rate vector is not a probability vector")
        u = tape.rand()
        n = len(p)
        cp = p[0]
        i = 0
        while cp <= u and i < n - 1:
            i += 1
            cp += p[i]
        v = v.copy()
        v[i] *= -1
        return v


# --------------------------------------------------------------------------------------------
class StickyHistory(History):
    """PDMPHistory with the is_active BitMatrix that record! stores (Composites.jl:239-260)."""

    def __init__(self, d, n):
        super().__init__(d, n)
        self.is_active = np.ones((d, n), dtype=bool, order="F")

    def record(self, k, s):
        super().record(k, s)
        self.is_active[:, k] = s.is_active


class StickyChain(Chain):
    """get_event_state!(state, ::StickyPDMP) (SamplingLoopInplace.jl:49-63) and its loop body
    (StickySamplingLoop.jl:30-164).  `kappa[i]` is the thawing rate of coordinate i.  Quirks kept as they are:
    the velocity jump after an accepted event sees the full velocity, frozen coordinates included (if_accept!,
    SamplingLoopInplace.jl:178); thawing adds tt but not the time already spent (ts) to the clock
    (StickySamplingLoop.jl:160-161); axis crossings are only looked for at the start of an outer step (:52-60)."""

    def __init__(self, sampler: Sampler, kappa, xinit, vinit, tape: Tape):
        super().__init__(sampler, xinit, vinit, tape)
        if sampler.cfg.sampler != ZIGZAG:
            raise ValueError("StickyChain restates StickyZigZag only")
        self.kappa = np.asarray(kappa, dtype=np.float64)
        st = self.state
        st.is_active = np.ones(sampler.dim, dtype=bool)   # PDMPState ctor: trues(d), tt = Inf (Composites.jl:95-99)
        st.tt = math.inf
        st.stick_or_thaw_event = False

    def _active_velocity(self):  # SamplingLoopInplace.jl:13-25
        st = self.state
        return st.v if st.is_active.all() else np.where(st.is_active, st.v, 0.0)

    def get_event_state(self):  # SamplingLoopInplace.jl:49-63
        st = self.state
        st.errored_bound = 0
        st.rejected = 0
        st.hitting_horizon = 0
        st.error_value_ar = np.zeros(5)
        while not st.accept and not st.stick_or_thaw_event:
            self.one_step_of_thinning_or_sticking_or_thawing()
        st.accept = False
        st.stick_or_thaw_event = False
        return st

    def one_step_of_thinning_or_sticking_or_thawing(self):  # StickySamplingLoop.jl:30-67
        st = self.state
        v_used = self._active_velocity()
        ub = self.upper_bound_func(st.x, v_used, st.horizon)
        e = self.tape.randexp()
        tp, lb = next_event(ub, e)
        rate_thawing = 0.0
        for i in range(len(st.is_active)):
            if not st.is_active[i]:
                rate_thawing += self.kappa[i]
        tt = math.inf if rate_thawing == 0 else self.tape.randexp() / rate_thawing
        st.tp, st.exp_rv, st.lambda_bar, st.upper_bound, st.tt = tp, e, lb, ub, tt
        event_time = min(tp, st.horizon, tt)
        xe, _ = self.s.flow(st.x, v_used, event_time)
        crossed = bool(np.any(st.x * xe < 0))
        if crossed:
            self.move_to_axes_and_stick()
        elif min(tp, tt) > st.horizon:
            self.move_to_horizon()
        else:
            self.moves_until_horizon_or_axes()

    def move_to_axes_and_stick(self):  # StickySamplingLoop.jl:73-107
        st = self.state
        t_togo, i = math.inf, -1
        for j in range(len(st.x)):
            if st.is_active[j]:
                dj = st.x[j] * st.v[j]
                if dj < 0:
                    tj = -dj
                    if tj < t_togo:
                        t_togo, i = tj, j
        if t_togo == math.inf:
            raise RuntimeError("erronous t_togo, although no axis is crossed")
        v_used = self._active_velocity()
        st.x, _ = self.s.flow(st.x, v_used, t_togo)
        st.is_active = st.is_active.copy()
        st.is_active[i] = False
        st.t += t_togo + st.ts
        st.ts = 0.0
        st.stick_or_thaw_event = True

    def move_to_horizon(self):  # SamplingLoopInplace.jl:87-101 (masked branch)
        st = self.state
        if st.is_active.all():
            st.x, st.v = self.s.flow(st.x, st.v, st.horizon)
        else:
            st.x, _ = self.s.flow(st.x, self._active_velocity(), st.horizon)
        st.ts += st.horizon
        st.hitting_horizon += 1
        st.horizon = st.horizon * 1.01 if st.adaptive else st.horizon

    def moves_until_horizon_or_axes(self):  # StickySamplingLoop.jl:121-132
        st = self.state
        while min(st.tp, st.tt) < st.horizon and not st.accept and not st.stick_or_thaw_event:
            if st.tp < st.tt:
                self.ac_step()
            else:
                self.thaw_one_coordinate()

    def thaw_one_coordinate(self):  # StickySamplingLoop.jl:138-164
        st = self.state
        st.x, _ = self.s.flow(st.x, self._active_velocity(), st.tt)
        total = 0.0
        for j in range(len(st.is_active)):
            if not st.is_active[j]:
                total += self.kappa[j]
        u = self.tape.rand() * total
        acc, i = 0.0, -1
        for j in range(len(st.is_active)):
            if not st.is_active[j]:
                acc += self.kappa[j]
                if acc >= u:
                    i = j
                    break
        st.is_active = st.is_active.copy()
        st.is_active[i] = True
        st.t += st.tt
        st.ts = 0.0
        st.stick_or_thaw_event = True

    # --- masked variants of the shared steps (SamplingLoopInplace.jl:113-217) ---
    def ac_step(self):  # :113-129
        st = self.state
        st.n_rate_evals += 1
        lt = self.s.rate(st.x, self._active_velocity(), st.tp)
        with np.errstate(invalid="ignore", divide="ignore"):
            ar = float(np.float64(lt) / np.float64(st.lambda_bar))
        st.lambda_t, st.ar = lt, ar
        if ar > 1.0:
            self.erroneous_acceptance_rate()
        else:
            self.ac_step_with_proxy()

    def erroneous_acceptance_rate(self):  # :131-151
        st = self.state
        horizon = st.horizon / 2
        ub = self.upper_bound_func(st.x, self._active_velocity(), horizon)
        e = self.tape.randexp()
        tp, lb = next_event(ub, e)
        st.horizon = horizon if st.adaptive else st.horizon
        st.tp, st.exp_rv, st.lambda_bar, st.upper_bound = tp, e, lb, ub
        st.errored_bound += 1
        st.error_value_ar[st.errored_bound % 5] = st.ar

    def ac_step_with_proxy(self):  # :153-168
        st = self.state
        accept = self.tape.rand() < st.ar
        st.accept = accept
        if accept:
            self.if_accept()
        else:
            self.if_reject()
        if (not st.accept) and (min(st.tp, st.tt) > st.horizon):
            self.move_to_horizon2()

    def if_accept(self):  # :170-186
        st = self.state
        if st.is_active.all():
            st.x, st.v = self.s.flow(st.x, st.v, st.tp)
        else:
            st.x, _ = self.s.flow(st.x, self._active_velocity(), st.tp)
        st.v = self.s.velocity_jump(st.x, st.v, self.tape)   # the FULL velocity (quirk, see class docstring)
        st.t = st.t + st.tp + st.ts
        st.ts = 0.0
        st.tp = 0.0
        st.accept = True

    def move_to_horizon2(self):  # :205-217
        st = self.state
        if st.is_active.all():
            st.x, st.v = self.s.flow(st.x, st.v, st.horizon)
        else:
            st.x, _ = self.s.flow(st.x, self._active_velocity(), st.horizon)
        st.ts += st.horizon
        st.hitting_horizon += 1


def sample_skeleton_sticky(sampler: Sampler, kappa, n_sk, xinit, vinit, tape: Tape):
    """sample_skeleton for a StickyZigZag sampler (src/sample.jl:253-284 with the StickyPDMP dispatch): every accepted
    flip, every sticking and every thawing is a skeleton point."""
    if n_sk <= 0:
        raise ValueError("n_sk must be positive")
    ch = StickyChain(sampler, kappa, xinit, vinit, tape)
    h = StickyHistory(sampler.dim, n_sk)
    h.record(0, ch.state)
    for k in range(1, n_sk):
        h.record(k, ch.get_event_state())
    h.final_state = ch.state
    h.tape_pos = list(tape.pos)
    return h
