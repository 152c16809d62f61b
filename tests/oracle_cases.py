"""Shared case table for oracle cross-checks and CUDA parity tests (test infrastructure)."""
import numpy as np

ZIGZAG, BPS, FECMC, BOOMERANG = 0, 1, 2, 3
GAUSS_STD, GAUSS_DIAG, GAUSS_EQUICORR, BANANA, BANANA_README, LOGREG = range(6)

# name, sampler, potential, pot_params, dim, config kwargs, n_sk
CASES = [
    ("zz_gauss10", ZIGZAG, GAUSS_STD, None, 10, dict(), 2000),
    ("zz_gauss10_fd", ZIGZAG, GAUSS_STD, None, 10, dict(deriv_mode=1), 1000),
    ("zz_gauss_unsigned", ZIGZAG, GAUSS_STD, None, 5, dict(signed_bound=False), 1000),
    ("zz_gauss_scalar", ZIGZAG, GAUSS_STD, None, 5, dict(vectorized_bound=False), 1000),
    ("zz_gauss_scalar_fd", ZIGZAG, GAUSS_STD, None, 3, dict(vectorized_bound=False, deriv_mode=1, grid_size=4), 600),
    ("zz_gauss1d_g2", ZIGZAG, GAUSS_STD, None, 1, dict(grid_size=2, tmax=0.0), 600),
    ("zz_banana50_brent", ZIGZAG, BANANA, None, 50, dict(grid_size=0), 1000),
    # more Brent brackets: rates that fall towards the far end of the bracket (the search walks left, ~70 iterations) and
    # fixed horizons
    ("zz_diag33_brent", ZIGZAG, GAUSS_DIAG, "linspace", 33, dict(grid_size=0), 600),
    ("zz_gauss12_brent_nonadaptive", ZIGZAG, GAUSS_STD, None, 12, dict(grid_size=0, adaptive=False, tmax=0.7), 600),
    # the ends of the transposed team search's range: 3 coordinates (five of a team's eight lanes own nothing) and 64
    # (eight owned coordinates per lane, more sign changes per bracket than the 12-slot list holds now and then)
    ("zz_gauss3_brent", ZIGZAG, GAUSS_STD, None, 3, dict(grid_size=0), 400),
    ("zz_diag64_brent", ZIGZAG, GAUSS_DIAG, "linspace", 64, dict(grid_size=0, tmax=4.0), 300),
    # a saddle (half of the curvatures negative): coordinate rates that fall along the flow, so the total rate is not
    # monotone on the bracket and Brent's search changes direction, fails golden steps and ends at either end -- every
    # branch of the speculative search is exercised.  (Not a distribution; the short run stays finite.)
    ("zz_saddle20_brent", ZIGZAG, GAUSS_DIAG, "mixed_sign", 20, dict(grid_size=0, tmax=0.5), 250),
    ("zz_banana5_grid_fd", ZIGZAG, BANANA, None, 5, dict(grid_size=8, deriv_mode=1), 1000),
    ("zz_banana40_jvp", ZIGZAG, BANANA, None, 40, dict(grid_size=6), 800),
    ("zz_readme_scalar", ZIGZAG, BANANA_README, None, 6, dict(grid_size=0), 7),
    ("zz_equicorr33", ZIGZAG, GAUSS_EQUICORR, [0.5], 33, dict(grid_size=5, adaptive=False, tmax=0.3), 800),
    ("zz_diag70", ZIGZAG, GAUSS_DIAG, "linspace", 70, dict(grid_size=10), 800),
    ("bps_equi100", BPS, GAUSS_EQUICORR, [0.9], 100, dict(tmax=1.0, refresh_rate=0.1), 1000),
    ("bps_gv_nonadapt", BPS, GAUSS_STD, None, 7,
     dict(tmax=1.0, refresh_rate=0.5, gaussian_velocity=True, adaptive=False, signed_bound=False), 1000),
    ("bps_brent", BPS, BANANA, None, 12, dict(tmax=1.0, refresh_rate=0.2, grid_size=0), 600),
    ("bps_fd", BPS, BANANA, None, 9, dict(tmax=1.0, refresh_rate=0.1, deriv_mode=1), 600),
    ("fecmc1000", FECMC, GAUSS_STD, None, 1000, dict(), 300),
    ("fecmc_ranp", FECMC, BANANA, None, 6, dict(ran_p=True, mix_p=0.7), 1000),
    ("fecmc_full_speed", FECMC, GAUSS_STD, None, 4, dict(switch=False, speed_factor=1.5, positive=False), 1000),
    ("fecmc_d2", FECMC, GAUSS_STD, None, 2, dict(), 500),
    ("fecmc_unsigned65", FECMC, GAUSS_DIAG, "linspace", 65, dict(signed_bound=False, grid_size=7), 500),
    ("boom20_fd", BOOMERANG, GAUSS_STD, None, 20, dict(tmax=1.0, refresh_rate=0.1, deriv_mode=1), 1000),
    ("boom_banana_jvp", BOOMERANG, BANANA, None, 4, dict(tmax=1.0, refresh_rate=0.1), 1000),
    ("boom_diag_brent", BOOMERANG, GAUSS_DIAG, "linspace", 6, dict(tmax=1.0, refresh_rate=0.1, grid_size=0), 500),
    ("boom_equi64", BOOMERANG, GAUSS_EQUICORR, [0.3], 64, dict(tmax=1.0, refresh_rate=0.3), 500),
    # BASELINE.json config 5b at full dimension: Boomerang d = 1000, finite-difference bound, every event a 1000-normal refresh
    ("boom1000_fd", BOOMERANG, GAUSS_STD, None, 1000, dict(tmax=1.0, refresh_rate=0.1, deriv_mode=1), 120),
    # Bayesian logistic regression (BASELINE.json config 4 in miniature): n rows, prior N(0, 10^2 I)
    ("zz_logreg5_n40", ZIGZAG, LOGREG, "logreg:40", 5, dict(grid_size=6), 300),
    ("zz_logreg100_n300", ZIGZAG, LOGREG, "logreg:300", 100, dict(grid_size=10), 120),
    ("zz_logreg13_unsigned", ZIGZAG, LOGREG, "logreg:70", 13, dict(grid_size=4, signed_bound=False, adaptive=False, tmax=0.05), 200),
]


def logreg_data(n, d, seed=2024, sigma0=10.0):
    """Synthetic design of SURVEY.md 8d (C4): rows ~ N(0, I/d), theta* ~ N(0, I), y ~ Bernoulli(sigma(x.theta*))."""
    g = np.random.default_rng([seed, n, d])
    X = g.standard_normal((n, d)) / np.sqrt(d)
    theta = g.standard_normal(d)
    y = (g.random(n) < 1.0 / (1.0 + np.exp(-X @ theta))).astype(np.float64)
    return X, y, sigma0


def pot_params(pp, d):
    if isinstance(pp, str) and pp == "linspace":
        return np.linspace(0.5, 2.0, d)
    if isinstance(pp, str) and pp == "mixed_sign":   # saddle: every other coordinate has negative curvature
        return np.where(np.arange(d) % 2 == 0, 1.0, -1.0) * np.linspace(0.5, 2.0, d)
    if isinstance(pp, str) and pp.startswith("logreg:"):
        X, y, s0 = logreg_data(int(pp.split(":")[1]), d)
        return np.concatenate([[float(X.shape[0]), s0], X.ravel(), y])
    return None if pp is None else np.asarray(pp, dtype=np.float64)


def case_inputs(name, sampler, d, n_sk, n_chains=1, seed=0):
    """Deterministic initial conditions + draw tapes for a case: (x0[C,d], v0[C,d], (E,U,N)[C,*])."""
    import zlib
    g = np.random.default_rng([zlib.crc32(name.encode()), seed])
    x0 = 0.5 * g.standard_normal((n_chains, d))
    if name.startswith("zz_readme"):
        x0 = np.ones((n_chains, d))
    if sampler == ZIGZAG:
        v0 = np.where(g.random((n_chains, d)) < 0.5, -1.0, 1.0)
    else:
        v0 = g.standard_normal((n_chains, d))
        if sampler in (BPS, FECMC):
            v0 /= np.linalg.norm(v0, axis=1, keepdims=True)
    nE, nU = 8 * n_sk + 64, 4 * n_sk + 64
    nN = {ZIGZAG: 1, BPS: d * (n_sk // 2 + 8), FECMC: 2 * d * n_sk + 64, BOOMERANG: d * n_sk + 64}[sampler]
    E = g.standard_exponential((n_chains, nE))
    U = g.random((n_chains, nU))
    N = g.standard_normal((n_chains, nN))
    return x0, v0, (E, U, N)


def tier_tolerance(kw, base=1e-10):
    """Parity tiers.  Everything with an analytic derivative -- grid bounds AND the constant
    (Brent) bound -- is held to `base`, the north_star's 1e-10: the Brent recurrence is evaluated without
    multiply-add contraction, i.e. with the reference's roundings, and the achieved one-step errors are <= 5.1e-14
    (profiles/r2_parity_errors.json).  Only the sqrt(eps) finite-difference derivative mode keeps a looser tier:
    dividing O(1e-16) summation-order differences by h = sqrt(eps) leaves ~1e-8 in the derivative (measured worst
    case 1.6e-8, zz_banana5_grid_fd; the two CPU restatements differ by the same amount), so it is held to 1e-6 and
    documented as a deviation from the 1e-10 target."""
    if kw.get("deriv_mode", 0) == 1:
        return 1e-6
    return base
