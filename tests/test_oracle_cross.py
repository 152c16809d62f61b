"""The two independent CPU restatements (numpy: literal array expressions; C: scalar loops) must agree
on identical draw tapes -- guards against a transcription slip in either.

Two comparisons:
  * free-running: same init, same tape, whole skeleton.  Rounding differences (np.dot vs scalar loop) are
    amplified by chaotic dynamics (banana / reflections), by Brent and by sqrt(eps) finite differences, so
    only benign cases are held to 1e-12 over the whole run.
  * teacher-forced ("one-step"): for every event k the numpy chain is restarted from the C oracle's state
    k-1 and tape cursor and must reproduce state k.  This isolates the per-event transition from the
    accumulated divergence and is applied to every case and every event.
"""
import numpy as np
import pytest

import oracle_c as oc
import pdmp_oracle_np as onp
from oracle_cases import CASES, case_inputs, pot_params, tier_tolerance


def np_potential(kind, pp, d):
    return {0: lambda: onp.GaussStd(), 1: lambda: onp.GaussDiag(pp), 2: lambda: onp.GaussEquicorr(d, pp[0]),
            3: lambda: onp.Banana(), 4: lambda: onp.BananaReadmeScalar(),
            5: lambda: onp.LogReg(pp[2:2 + int(pp[0]) * d].reshape(int(pp[0]), d), pp[2 + int(pp[0]) * d:], pp[1])}[kind]()


def relerr(a, b):
    scale = max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


FREE_RUNNING_EXACT = {"zz_gauss10", "zz_gauss10_fd", "zz_gauss_unsigned", "zz_gauss_scalar", "zz_gauss1d_g2",
                      "zz_diag70", "fecmc_d2"}


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_numpy_vs_c_one_step(case):
    name, sampler, pk, pp, d, kw, n_sk = case
    n_sk = min(n_sk, 300 if d <= 100 else 100)  # the numpy restatement is slow
    pp = pot_params(pp, d)
    x0, v0, (E, U, N) = case_inputs(name, sampler, d, n_sk)
    r = oc.sample_skeleton(oc.make_cfg(sampler, pk, d, pp, **kw), n_sk, x0, v0, tape=(E, U, N))
    assert r.status[0] == oc.ST_OK
    tol = tier_tolerance(kw)
    s = onp.Sampler(d, np_potential(pk, pp, d), onp.Config(sampler=sampler, **kw))
    tape = onp.Tape(E[0], U[0], N[0])
    ch = onp.Chain(s, x0[0], v0[0], tape)
    worst = 0.0
    for k in range(1, n_sk):
        st = ch.state
        st.x = r.X[0, k - 1].copy(); st.v = r.V[0, k - 1].copy()
        st.t = float(r.t[0, k - 1]); st.horizon = float(r.horizon[0, k - 1])
        tape.pos = [int(a) for a in r.tape_pos[0, k - 1]]
        st = ch.get_event_state()
        e = max(relerr(st.x, r.X[0, k]), relerr(st.v, r.V[0, k]), abs(st.t - r.t[0, k]) / max(abs(st.t), 1e-300),
                abs(st.horizon - r.horizon[0, k]) / st.horizon, abs(st.ar - r.ar[0, k]))
        worst = max(worst, e)
        assert e < tol, (k, e)
        assert np.array_equal(np.sign(st.v), np.sign(r.V[0, k])), k
        assert (st.rejected, st.hitting_horizon, st.errored_bound) == \
            (r.rejected[0, k], r.hitting_horizon[0, k], r.errored_bound[0, k]), k
        assert np.allclose(st.error_value_ar, r.error_value_ar[0, k], rtol=tol, atol=0)
        assert tape.pos == [int(a) for a in r.tape_pos[0, k]], k


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in FREE_RUNNING_EXACT], ids=lambda c: c[0])
def test_numpy_vs_c_free_running(case):
    name, sampler, pk, pp, d, kw, n_sk = case
    n_sk = min(n_sk, 400)
    pp = pot_params(pp, d)
    x0, v0, (E, U, N) = case_inputs(name, sampler, d, n_sk)
    s = onp.Sampler(d, np_potential(pk, pp, d), onp.Config(sampler=sampler, **kw))
    h = onp.sample_skeleton(s, n_sk, x0[0], v0[0], onp.Tape(E[0], U[0], N[0]))
    r = oc.sample_skeleton(oc.make_cfg(sampler, pk, d, pp, **kw), n_sk, x0, v0, tape=(E, U, N))
    assert r.status[0] == oc.ST_OK
    for a, b in ((h.X.T, r.X[0]), (h.V.T, r.V[0]), (h.t, r.t[0]), (h.horizon, r.horizon[0]), (h.ar, r.ar[0])):
        assert relerr(a, b) < 1e-12
    assert np.array_equal(np.sign(h.V.T), np.sign(r.V[0]))
    assert np.array_equal(h.rejected, r.rejected[0])
    assert np.array_equal(h.hitting_horizon, r.hitting_horizon[0])
    assert np.array_equal(h.errored_bound, r.errored_bound[0])
    assert list(r.tape_used[0]) == list(h.tape_pos)
    assert r.counters[0, 0] == h.final_state.n_bound_builds
    assert r.counters[0, 1] == h.final_state.n_rate_evals
