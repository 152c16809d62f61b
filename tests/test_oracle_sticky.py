"""Sticky Zig-Zag oracle (the next row of the scope table): checks of the literal numpy restatement of
src/StickySamplingLoop.jl.  The reference's tests hold no golden vectors for it; the restatement is checked through
structural invariants and through the known stationary law of the sticky Zig-Zag process (Bierkens, Grazzi, van der
Meulen, Schauer 2021): mu(dx) ~ exp(-U(x)) prod_i (dx_i + delta_0(dx_i) / kappa_i)."""
import math

import numpy as np

import pdmp_oracle_np as onp


def _run(chain_cls, d, kappa, n_sk, seed, pot=None):
    cfg = onp.Config(sampler=onp.ZIGZAG).normalised(d)
    s = onp.Sampler(d, pot or onp.GaussStd(), cfg)
    tape = onp.make_tape(seed, 40 * n_sk, 20 * n_sk, 1)
    g = np.random.default_rng(seed)
    ch = chain_cls(s, np.full(d, kappa), g.standard_normal(d), np.where(g.random(d) < 0.5, -1.0, 1.0), tape)
    h = onp.StickyHistory(d, n_sk)
    h.record(0, ch.state)
    for k in range(1, n_sk):
        h.record(k, ch.get_event_state())
    return h


class _RepairedClock(onp.StickyChain):
    """thaw_one_coordinate drops the time already spent in horizon moves (`state.t += state.tt`,
    StickySamplingLoop.jl:160-161); adding it back gives the process whose stationary law is known."""

    def thaw_one_coordinate(self):
        ts = self.state.ts
        super().thaw_one_coordinate()
        self.state.t += ts


def test_sticky_structure():
    h = onp.sample_skeleton_sticky(onp.Sampler(3, onp.GaussDiag(np.array([0.5, 1.0, 2.0])), onp.Config().normalised(3)),
                                   np.full(3, 1.5), 1500, np.array([0.3, -0.7, 1.1]), np.array([1.0, -1.0, 1.0]),
                                   onp.make_tape(3, 60000, 30000, 1))
    dt = np.diff(h.t)
    assert np.all(dt > 0) and h.is_active[:, 0].all()
    assert np.all(np.abs(h.V) == 1.0)                       # velocities only ever flip sign
    changed = h.is_active[:, 1:] != h.is_active[:, :-1]
    assert changed.sum(axis=0).max() == 1                   # one coordinate sticks or thaws per event, never two
    assert changed.any() and (~h.is_active).any()
    # a coordinate that is frozen over a whole segment sits exactly on its axis
    frozen_seg = ~h.is_active[:, :-1] & ~h.is_active[:, 1:]
    assert np.all(h.X[:, :-1][frozen_seg] == 0.0) and np.all(h.X[:, 1:][frozen_seg] == 0.0)
    # active coordinates move in straight lines between events with their recorded velocity -- except across a
    # thawing event, whose clock forgets the horizon moves made before it (StickySamplingLoop.jl:160-161)
    thawed = (h.is_active[:, 1:] & ~h.is_active[:, :-1]).any(axis=0)
    act = h.is_active[:, :-1] & ~thawed[None, :]
    pred = h.X[:, :-1] + np.where(h.is_active[:, :-1], h.V[:, :-1], 0.0) * dt
    assert np.allclose(pred[act], h.X[:, 1:][act], rtol=0, atol=1e-9)
    assert thawed.any()


def test_sticky_stationary_law_and_the_thaw_clock_quirk():
    kappa, n_sk = 0.5, 12000
    expected = (1 / kappa) / (1 / kappa + math.sqrt(2 * math.pi))     # point mass at 0 for a standard Gaussian slab
    fr = {}
    for name, cls in (("literal", onp.StickyChain), ("repaired", _RepairedClock)):
        h = _run(cls, 1, kappa, n_sk, 1)
        dt = np.diff(h.t)
        frozen = ~h.is_active[0, :-1]
        fr[name] = dt[frozen].sum() / dt.sum()
        if name == "repaired":
            assert abs(dt[frozen].mean() - 1 / kappa) < 0.15          # frozen spells are Exp(kappa)
            x0, v0 = h.X[0, :-1], np.where(h.is_active[0, :-1], h.V[0, :-1], 0.0)
            m2 = x0 ** 2 * dt + x0 * v0 * dt ** 2 + v0 ** 2 * dt ** 3 / 3
            assert abs(m2[~frozen].sum() / dt[~frozen].sum() - 1.0) < 0.08   # the slab is N(0, 1)
    assert abs(fr["repaired"] - expected) < 0.03
    # the reference as written loses the horizon moves made while frozen: its clock under-counts frozen time
    assert fr["literal"] < fr["repaired"] - 0.1
