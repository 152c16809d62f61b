#!/usr/bin/env python
"""Generate the committed golden fixtures tests/golden/*.npz from the literal numpy restatement of the reference
(oracle/pdmp_oracle_np.py).  The reference itself (Julia) cannot run in this image and ships no golden skeleton
vectors, so these pin the *restatement*: any later change to the oracles or to the CUDA path is checked
against them (tests/test_golden.py).  Each fixture holds the inputs (initial state, typed draw tape, config), the
numpy oracle's PDMPHistory columns and the tape cursor after every event (for teacher-forced one-step checks).

    python tests/golden/make_golden.py        # rewrites every fixture
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import pdmp_oracle_np as onp  # noqa: E402
from oracle_cases import CASES, case_inputs, pot_params  # noqa: E402

GOLDEN = {  # case name -> number of skeleton points stored
    "zz_gauss10": 400, "zz_gauss_unsigned": 200, "zz_gauss1d_g2": 150, "zz_banana50_brent": 80, "zz_banana5_grid_fd": 120,
    "zz_equicorr33": 150, "bps_equi100": 120, "bps_gv_nonadapt": 200, "fecmc_ranp": 200, "fecmc_full_speed": 200,
    "boom20_fd": 150, "boom_equi64": 100, "zz_logreg5_n40": 150, "zz_logreg13_unsigned": 120,
}


def np_potential(kind, pp, d):
    return {0: lambda: onp.GaussStd(), 1: lambda: onp.GaussDiag(pp), 2: lambda: onp.GaussEquicorr(d, pp[0]),
            3: lambda: onp.Banana(), 4: lambda: onp.BananaReadmeScalar(),
            5: lambda: onp.LogReg(pp[2:2 + int(pp[0]) * d].reshape(int(pp[0]), d), pp[2 + int(pp[0]) * d:], pp[1])}[kind]()


def main():
    for case in CASES:
        name, sampler, pk, pp, d, kw, _ = case
        if name not in GOLDEN:
            continue
        n_sk = GOLDEN[name]
        ppv = pot_params(pp, d)
        x0, v0, (E, U, N) = case_inputs(name, sampler, d, n_sk)
        s = onp.Sampler(d, np_potential(pk, ppv, d), onp.Config(sampler=sampler, **kw))
        tape = onp.Tape(E[0], U[0], N[0])
        ch = onp.Chain(s, x0[0], v0[0], tape)
        h = onp.History(d, n_sk)
        h.record(0, ch.state)
        pos = np.zeros((n_sk, 3), dtype=np.int64)
        for k in range(1, n_sk):
            h.record(k, ch.get_event_state())
            pos[k] = tape.pos
        used = pos[-1]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), config=json.dumps(dict(sampler=sampler, potential=pk, dim=d, kwargs=kw)),
            pot_params=np.zeros(0) if ppv is None else ppv, x0=x0[0], v0=v0[0], E=E[0, :used[0] + 8], U=U[0, :used[1] + 8],
            N=N[0, :used[2] + 8], X=h.X.T.copy(), V=h.V.T.copy(), t=h.t, horizon=h.horizon, ar=h.ar,
            errored_bound=h.errored_bound, error_value_ar=h.error_value_ar.T.copy(), rejected=h.rejected,
            hitting_horizon=h.hitting_horizon, tape_pos=pos)
        print(f"{name:24s} n_sk={n_sk:4d} T={h.t[-1]:9.3f} draws used {used.tolist()}")


if __name__ == "__main__":
    main()
