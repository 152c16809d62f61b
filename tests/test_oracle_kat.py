"""Known-answer tests pinning the CPU oracles (hand-derived; the reference has no golden skeleton vectors,
SURVEY.md section 8c).  Each KAT cites what it is derived from."""
import math

import numpy as np
import pytest

import oracle_c as oc
import pdmp_oracle_np as onp


def test_philox_random123_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert oc.philox_raw((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oc.philox_raw((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oc.philox_raw((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_spec_ranges_and_moments():
    E, U, N = oc.draws(2024, 3, 7, 20000)
    assert E.min() > 0 and np.isfinite(E).all() and U.min() >= 0 and U.max() < 1
    assert abs(E.mean() - 1) < 0.03 and abs(U.mean() - 0.5) < 0.01
    assert abs(N.mean()) < 0.03 and abs(N.var() - 1) < 0.03
    E2, _, _ = oc.draws(2024, 4, 7, 16)
    assert not np.allclose(E[:16], E2)


@pytest.mark.parametrize("impl", ["np", "c"])
def test_vect_bound_1d_gaussian_closed_form(impl):
    """UpperBound.jl:203-247 on func(t) = (x + t v) v, x=0.3, v=1: val = 0.3+t, grad = 1 everywhere, so the
    tangent intersection is 0/0 -> NaN -> 0 and box_k = max(val_l, val_r, val_l, 0) = val_r."""
    x, v, h, G = np.array([0.3]), np.array([1.0]), 0.8, 3
    if impl == "np":
        s = onp.Sampler(1, onp.GaussStd(), onp.Config(sampler=0, grid_size=G))
        bb = onp.upper_bound_grid_vect(lambda t: s.bound_func_vect(x, v, t, True), h, G, 0)
        grid, box, cum, step = bb.grid, bb.box_max, bb.cum_sum, bb.step_size
    else:
        grid, box, cum, step = oc.bound(oc.make_cfg(0, 0, 1, grid_size=G), x, v, h)
    assert np.allclose(grid, [0, 0.4, 0.8], rtol=0, atol=1e-16) and step == pytest.approx(0.4, rel=1e-15)
    assert np.allclose(box, [0.7, 1.1], rtol=1e-15)
    assert np.allclose(cum, [0, 0.7 * 0.4, (0.7 + 1.1) * 0.4], rtol=1e-15)


@pytest.mark.parametrize("impl", ["np", "c"])
def test_vect_bound_absolute_time_quirk(impl):
    """UpperBound.jl:229-237: the intersection is an absolute time clamped to [0, step] and used as an offset.
    Banana d=2 at x=(1,0), v=(1,1) along the line: x1(t)=1+t, x2(t)=t, r = t-(1+t)^2+1 = -t - t^2,
    g1 v1 = x1 - 2 x1 r, g2 v2 = r.  Hand evaluation for G=3, h=1 (nodes 0, .5, 1)."""
    x, v, h, G = np.array([1.0, 0.0]), np.array([1.0, 1.0]), 1.0, 3
    t = np.array([0.0, 0.5, 1.0])
    x1 = 1 + t
    r = -t - t * t
    val = np.stack([x1 - 2 * x1 * r, r])                       # g .* v
    # d/dt: g1' = 1 - 2 r - 2 x1 r' with r' = -1 - 2t ; g2' = r'
    rp = -1 - 2 * t
    grad = np.stack([1 - 2 * r - 2 * x1 * rp, rp])
    step = 0.5
    pos = (val[:, :-1] - val[:, 1:] + grad[:, 1:] * t[1:] - grad[:, :-1] * t[:-1]) / (grad[:, 1:] - grad[:, :-1])
    pos = np.clip(np.where(np.isnan(pos), 0, pos), 0, step)
    inter = val[:, :-1] + grad[:, :-1] * pos
    box = np.maximum.reduce([val[:, :-1], val[:, 1:], inter, np.zeros_like(inter)])
    cum = np.concatenate([np.zeros((2, 1)), np.cumsum(box, axis=1) * step], axis=1)
    if impl == "np":
        s = onp.Sampler(2, onp.Banana(), onp.Config(sampler=0, grid_size=G))
        bb = onp.upper_bound_grid_vect(lambda tt: s.bound_func_vect(x, v, tt, True), h, G, 0)
        got_box, got_cum = bb.box_max, bb.cum_sum
    else:
        _, got_box, got_cum, _ = oc.bound(oc.make_cfg(0, 3, 2, grid_size=G), x, v, h)
    assert np.allclose(got_box, box.sum(axis=0), rtol=1e-14)
    assert np.allclose(got_cum, cum.sum(axis=0), rtol=1e-14)
    # second cell: the tangents cross at absolute t in (0.5, 1) > step, so pos saturates at `step`
    assert pos[0, 1] == step


def test_scalar_bound_bps_double_refresh():
    """BouncyParticleSamplers.jl:44-47 + UpperBound.jl:131 + AbstractPDMP.jl:104-112: with signed_bound the
    refresh rate enters the bound twice.  Std Gaussian, x=0, |v|=1: signed_rate(t) = t + r, so
    box_k = t_{k+1} + 2 r."""
    d, r, h, G = 3, 0.25, 0.9, 4
    v = np.ones(d) / math.sqrt(d)
    grid, box, cum, step = oc.bound(oc.make_cfg(1, 0, d, grid_size=G, refresh_rate=r, tmax=1.0), np.zeros(d), v, h)
    assert np.allclose(box, grid[1:] + 2 * r, rtol=1e-14)
    grid, box, cum, step = oc.bound(oc.make_cfg(1, 0, d, grid_size=G, refresh_rate=r, tmax=1.0, signed_bound=False),
                                    np.zeros(d), v, h)
    assert np.allclose(box, grid[1:] + r, rtol=1e-14)  # unsigned: refresh once (inside `rate`), bound_refresh = 0


def test_next_event_inversion():
    bb = onp.BoundBox(np.array([0.0, 0.5, 1.0]), np.array([2.0, 4.0]), np.array([0.0, 1.0, 3.0]), 0.5)
    assert onp.next_event(bb, 0.5) == (0.25, 2.0)
    assert onp.next_event(bb, 1.0) == (0.5, 2.0)            # searchsortedfirst: first cum >= e
    assert onp.next_event(bb, 2.0) == (0.75, 4.0)
    tp, lb = onp.next_event(bb, 3.5)
    assert math.isinf(tp) and lb == 4.0


def test_brent_restatement():
    f = lambda t: (t - 0.3) ** 2 + 1.0
    assert onp.brent_minimum(f, 0.0, 1.0) == pytest.approx(1.0, abs=1e-14)
    # increasing rate on [0, 2]: Brent never evaluates the end point and undershoots the sup by O(sqrt(eps))
    calls = []
    g = lambda t: (calls.append(t), -(0.3 + t))[1]
    m = -onp.brent_minimum(g, 0.0, 2.0)
    assert 0 < (2.3 - m) / 2.3 < 1e-7 and 30 <= len(calls) <= 45
    # C and numpy restatements agree on the constant bound
    x, v = np.array([0.3, -0.2, 0.5]), np.array([1.0, -1.0, 1.0])
    _, box, cum, step = oc.bound(oc.make_cfg(0, 0, 3, grid_size=0), x, v, 1.7)
    s = onp.Sampler(3, onp.GaussStd(), onp.Config(sampler=0, grid_size=0))
    bb = onp.upper_bound_constant(lambda t: s.rate(x, v, t), 0.0, 1.7)
    assert box[0] == bb.box_max[0] and cum[1] == bb.cum_sum[1] and step == 1.7


def test_finite_difference_derivative_branches():
    f = lambda t: t * t
    # interior: central; t=0: forward; t=horizon: backward (UpperBound.jl:60-75)
    assert onp.finite_difference_derivative(f, 0.5, 0.0, 1.0) == pytest.approx(1.0, rel=1e-7)
    h = onp.SQRT_EPS
    assert onp.finite_difference_derivative(f, 0.0, 0.0, 1.0) == pytest.approx(h, rel=1e-12)
    assert onp.finite_difference_derivative(f, 1.0, 0.0, 1.0) == pytest.approx(2.0 - h, rel=1e-7)
    assert onp.finite_difference_derivative(f, 0.0, 0.0, 0.0) == 0.0  # collapsed interval


def test_sample_from_skeleton_linear_path_kat():
    """test/test_diagnostics.jl:100-124 hand-built history: X=[0,.5,1], V=1, t=[0,.5,1] is the straight line x(t)=t."""
    X = np.array([[0.0, 0.5, 1.0]]); V = np.ones((1, 3)); t = np.array([0.0, 0.5, 1.0])
    out = onp.sample_from_skeleton(0, 4, X, V, t)
    assert np.allclose(out, [[0.25, 0.5, 0.75, 1.0]], rtol=0, atol=1e-16)
    outc = oc.sample_from_skeleton(0, X.T, V.T, t, 4)
    assert np.array_equal(outc.T, out)
    full = oc.sample_from_skeleton(0, X.T, V.T, t, 4, discard_vt=False)
    assert np.allclose(full[:, 1], 1.0) and np.allclose(full[:, 2], [0.25, 0.5, 0.75, 1.0])
    # rotation flow (BoomerangSamplers.jl:31): a quarter turn from (1, 0)
    o = oc.sample_from_skeleton(1, np.array([[1.0], [0.0]]), np.array([[0.0], [-1.0]]), np.array([0.0, math.pi / 2]), 1)
    assert abs(o[0, 0]) < 1e-15


def test_skeleton_structure_and_reference_envelopes():
    """Philox-seeded C oracle run checked against the statistical envelopes and invariants the reference's own
    tests assert: test_property_based.jl:87-100 (1-D Gaussian, grid_size=0: |mean|<0.2, 0.8<var<1.2 at 5000
    events), test_comprehensive.jl:147 (diff(t) > 0), :172-180 (ZigZag keeps |v_i| = 1)."""
    cfg = oc.make_cfg(0, 0, 1, grid_size=0)
    r = oc.sample_skeleton(cfg, 5000, np.zeros((1, 1)), np.ones((1, 1)), seed=42)
    assert r.status[0] == 0 and np.all(np.diff(r.t[0]) > 0) and np.all(np.abs(r.V[0]) == 1.0)
    assert r.t[0, 0] == 0 and r.horizon[0, 0] == 2.0 and r.ar[0, 0] == 0
    s = oc.sample_from_skeleton(0, r.X[0], r.V[0], r.t[0], 5000)
    assert abs(s.mean()) < 0.2 and 0.8 < s.var() < 1.2
    # README example shape (README.md:40-44), shortened: d=10 ZigZag(AD) std Gaussian
    cfg = oc.make_cfg(0, 0, 10)
    r = oc.sample_skeleton(cfg, 20000, np.zeros((1, 10)), np.ones((1, 10)), seed=2024)
    s = oc.sample_from_skeleton(0, r.X[0], r.V[0], r.t[0], 20000)
    assert np.all(np.abs(s.mean(axis=0)) < 0.2) and np.all((0.8 < s.var(axis=0)) & (s.var(axis=0) < 1.25))
    assert np.all(np.abs(r.V[0]) == 1.0)


def test_all_samplers_target_standard_gaussian():
    """ZigZag / BPS / ForwardECMC must leave N(0, I) invariant (statistical check of jump kernels + thinning).
    Boomerang is excluded: the reference's rate uses grad U while its jump uses grad U - x
    (BoomerangSamplers.jl:38-41 vs :51-52), so for U = |x|^2/2 every event is a refresh and the invariant law
    is not N(0, I); the code, not the docs, is what is restated."""
    d, n = 4, 30000
    for sampler, kw in ((0, dict()), (1, dict(tmax=1.0, refresh_rate=0.5)), (2, dict())):
        v0 = np.ones((1, d)) / (1.0 if sampler in (0, 3) else math.sqrt(d))
        r = oc.sample_skeleton(oc.make_cfg(sampler, 0, d, **kw), n, np.zeros((1, d)), v0, seed=5)
        assert r.status[0] == 0
        s = oc.sample_from_skeleton(1 if sampler == 3 else 0, r.X[0], r.V[0], r.t[0], n)
        assert np.all(np.abs(s.mean(axis=0)) < 0.15), (sampler, s.mean(axis=0))
        assert np.all(np.abs(s.var(axis=0) - 1) < 0.2), (sampler, s.var(axis=0))


def test_error_paths():
    with pytest.raises(ValueError):
        oc.sample_skeleton(oc.make_cfg(0, 0, 2), 0, np.zeros((1, 2)), np.ones((1, 2)), seed=1)      # n_sk <= 0
    with pytest.raises(ValueError):
        oc.sample_skeleton(oc.make_cfg(0, 0, 2, grid_size=-1), 5, np.zeros((1, 2)), np.ones((1, 2)), seed=1)
    with pytest.raises(ValueError):
        oc.sample_skeleton(oc.make_cfg(2, 0, 1), 5, np.zeros((1, 1)), np.ones((1, 1)), seed=1)      # FECMC dim < 2
    # tape exhaustion is reported, not silently ignored
    r = oc.sample_skeleton(oc.make_cfg(0, 0, 2), 50, np.zeros((1, 2)), np.ones((1, 2)),
                           tape=(np.ones((1, 3)), np.full((1, 3), 0.5), np.zeros((1, 1))))
    assert r.status[0] == oc.ST_TAPE_EXHAUSTED
    # Categorical(p) with sum(lambda) == 0 throws in the reference (ZigZagSamplers.jl:103-104)
    with pytest.raises(FloatingPointError):
        s = onp.Sampler(2, onp.GaussStd(), onp.Config(sampler=0))
        s.velocity_jump(np.zeros(2), np.ones(2), onp.make_tape(0, 4, 4, 4))


def test_time_horizon_variant_oracle():
    """sample_skeleton(sampler, T) (src/sample.jl:323-439): t[end] == T exactly (test_samplers.jl:56-62), the last
    column is the flow of the previous state, event columns agree with the n_sk variant on the same draws."""
    d, T = 3, 7.5
    cfg = oc.make_cfg(0, 0, d)
    x0 = np.zeros((2, d)); v0 = np.ones((2, d))
    r = oc.sample_skeleton_until(cfg, T, 400, x0, v0, seed=3)
    full = oc.sample_skeleton(cfg, 400, x0, v0, seed=3)
    for c in range(2):
        n = r.ncols[c]
        assert 2 < n < 400 and r.t[c, n - 1] == T and np.all(np.diff(r.t[c, :n]) > 0)
        assert np.array_equal(r.X[c, :n - 1], full.X[c, :n - 1]) and np.array_equal(r.t[c, :n - 1], full.t[c, :n - 1])
        tau = T - r.t[c, n - 2]
        assert np.allclose(r.X[c, n - 1], r.X[c, n - 2] + r.V[c, n - 2] * tau, rtol=0, atol=1e-15)
        assert np.array_equal(r.V[c, n - 1], r.V[c, n - 2]) and r.ar[c, n - 1] == 0 and full.t[c, n - 1] > T
    # T == 0: the initial point alone; too small a capacity is reported
    assert oc.sample_skeleton_until(cfg, 0.0, 4, x0, v0, seed=3).ncols.tolist() == [1, 1]
    assert oc.sample_skeleton_until(cfg, T, 3, x0, v0, seed=3).ncols.tolist() == [-1, -1]


def test_rv_diagnostic_reference_kat():
    """The reference's own known answer for this path (test/test_diagnostics.jl:100-124): hand-built straight-line
    history x(t) = t on [0, 1], U = x_1^2 / 2, B = 4 -> sum_k (U(k/4) - U((k-1)/4))^2; B = 0 -> positive;
    B < 0 -> ArgumentError.  Pins the interpolation + blocking of the oracle's RV_diagnostic restatement."""
    X = np.array([[0.0, 0.5, 1.0]]); V = np.ones((1, 3)); t = np.array([0.0, 0.5, 1.0])
    U = lambda x: x[0] ** 2 / 2
    expected = sum((U([k / 4]) - U([(k - 1) / 4])) ** 2 for k in range(1, 5))
    assert onp.rv_diagnostic(X, V, t, U, B=4) == pytest.approx(expected, rel=1e-15)
    assert onp.rv_diagnostic(X, V, t, onp.GaussStd().value, B=4) == pytest.approx(expected, rel=1e-15)
    assert onp.rv_diagnostic(X, V, t, U, B=0) > 0.0
    with pytest.raises(ValueError):
        onp.rv_diagnostic(X, V, t, U, B=-1)
    assert onp.rv_diagnostic(X[:, :0], V[:, :0], t[:0], U) == 0.0        # empty history (diagnostic.jl:40)
    assert onp.rv_diagnostic(X[:, :1], V[:, :1], t[:1], U, B=3) == 0.0   # T == 0 (diagnostic.jl:53)


def test_potential_values_match_their_gradients():
    """U plugins used by RV_diagnostic: central differences of value() reproduce grad() (the README-scalar banana keeps
    the true U; only its hand-written gradient is wrong, README.md:56-65)."""
    g = np.random.default_rng(3)
    d = 6
    for pot in (onp.GaussStd(), onp.GaussDiag(np.linspace(0.5, 2, d)), onp.GaussEquicorr(d, 0.7), onp.Banana()):
        x = g.standard_normal(d)
        num = np.array([(pot.value(x + 1e-6 * e) - pot.value(x - 1e-6 * e)) / 2e-6 for e in np.eye(d)])
        assert np.allclose(num, pot.grad(x), rtol=1e-7, atol=1e-8)
    assert onp.BananaReadmeScalar().value(np.arange(1.0, 5.0)) == onp.Banana().value(np.arange(1.0, 5.0))
