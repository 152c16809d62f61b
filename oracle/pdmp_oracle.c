/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of PDMPFlux.jl's grid-based Poisson-thinning path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load
 * this library.  It is never linked into, nor called by, libpdmpflux_cuda.so or pdmpflux_b200.
 *
 * PARITY UNPINNED: the reference is pure Julia, Julia is not installed here, and the reference has
 * no golden skeleton vectors.  This restatement is pinned by hand-derived
 * known-answer tests and by agreement with the independent numpy restatement
 * (oracle/pdmp_oracle_np.py); see tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line (relative to /root/reference/) it follows.
 * Quirks of the reference are preserved on purpose (SURVEY.md 8a "gotchas").
 *
 * Draws come either from an injected tape (three typed streams per chain: E, U, N) or from the
 * counter-based specification in pdmp_draws.h.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "pdmp_draws.h"

enum { S_ZIGZAG = 0, S_BPS = 1, S_FECMC = 2, S_BOOMERANG = 3 };
enum { P_GAUSS_STD = 0, P_GAUSS_DIAG = 1, P_GAUSS_EQUICORR = 2, P_BANANA = 3, P_BANANA_README = 4,
       P_LOGREG = 5, P_GAUSS_DENSE = 6 };
enum { D_JVP = 0, D_FD = 1 };
enum { ST_OK = 0, ST_TAPE_EXHAUSTED = 1, ST_NOT_PROBVEC = 2, ST_ITER_LIMIT = 3 };

typedef struct {
    int32_t sampler, potential, dim, grid_size;
    int32_t vectorized_bound, signed_bound, adaptive, deriv_mode;
    int32_t gaussian_velocity, ran_p, switch_, positive;
    double tmax, refresh_rate, mix_p, speed_factor;
    const double* pot_params;
    int64_t n_pot_params;
} pdmp_oracle_cfg;

typedef struct {
    int mode; /* 0 tape, 1 philox */
    const double *E, *U, *N;
    int64_t nE, nU, nN, pE, pU, pN;
    uint64_t seed, chain, event;
    uint32_t sE, sU, sN;
    int exhausted;
} draws_t;

static double draw_exp(draws_t* d) {
    if (d->mode) return pdmp_draw_exp(d->seed, d->chain, d->event, d->sE++);
    if (d->pE >= d->nE) { d->exhausted = 1; return 1.0; }
    return d->E[d->pE++];
}
static double draw_uniform(draws_t* d) {
    if (d->mode) return pdmp_draw_uniform(d->seed, d->chain, d->event, d->sU++);
    if (d->pU >= d->nU) { d->exhausted = 1; return 0.5; }
    return d->U[d->pU++];
}
static double draw_normal(draws_t* d) {
    if (d->mode) return pdmp_draw_normal(d->seed, d->chain, d->event, d->sN++);
    if (d->pN >= d->nN) { d->exhausted = 1; return 1.0; }
    return d->N[d->pN++];
}

/* ------------------------------------------------------------------------------------------ */
/* potentials: gradient and Hessian-vector product                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const pdmp_oracle_cfg* c;
    int d;
    double alpha, beta;          /* equicorrelated */
    int64_t n; double inv_s2;    /* logreg */
    const double *X, *y;
    double* zbuf;                /* logreg scratch (n) x2 */
} pot_t;

static void pot_init(pot_t* p, const pdmp_oracle_cfg* c) {
    memset(p, 0, sizeof(*p));
    p->c = c; p->d = c->dim;
    if (c->potential == P_GAUSS_EQUICORR) {
        double rho = c->pot_params[0];
        p->alpha = 1.0 / (1.0 - rho);
        p->beta = rho / ((1.0 - rho) * (1.0 - rho + c->dim * rho));
    } else if (c->potential == P_LOGREG) {
        p->n = (int64_t)c->pot_params[0];
        double s0 = c->pot_params[1];
        p->inv_s2 = 1.0 / (s0 * s0);
        p->X = c->pot_params + 2;
        p->y = p->X + p->n * c->dim;
    }
}
static void pot_free(pot_t* p) { free(p->zbuf); }

static void pot_grad(const pot_t* p, const double* x, double* g) {
    int d = p->d;
    switch (p->c->potential) {
    case P_GAUSS_STD: /* README.md:36-38, U = sum(x.^2)/2 */
        for (int i = 0; i < d; ++i) g[i] = x[i];
        break;
    case P_GAUSS_DIAG:
        for (int i = 0; i < d; ++i) g[i] = p->c->pot_params[i] * x[i];
        break;
    case P_GAUSS_EQUICORR: {
        double s = 0; for (int i = 0; i < d; ++i) s += x[i];
        for (int i = 0; i < d; ++i) g[i] = p->alpha * x[i] - p->beta * s;
    } break;
    case P_BANANA: { /* test/test_config.jl:33-36 */
        double r = x[1] - x[0] * x[0] + 1.0;
        for (int i = 2; i < d; ++i) g[i] = x[i];
        g[0] = x[0] - 2.0 * x[0] * r;
        g[1] = r;
    } break;
    case P_BANANA_README: { /* README.md:62-65, scalar broadcast */
        double s2 = 0; for (int i = 2; i < d; ++i) s2 += x[i];
        double s = x[0] + (x[1] - (x[0] * x[0] - 1.0)) + s2;
        for (int i = 0; i < d; ++i) g[i] = s;
    } break;
    case P_GAUSS_DENSE:
        for (int i = 0; i < d; ++i) {
            double s = 0; const double* row = p->c->pot_params + (size_t)i * d;
            for (int j = 0; j < d; ++j) s += row[j] * x[j];
            g[i] = s;
        }
        break;
    case P_LOGREG: {
        for (int i = 0; i < d; ++i) g[i] = 0;
        for (int64_t j = 0; j < p->n; ++j) {
            const double* row = p->X + j * d; double z = 0;
            for (int i = 0; i < d; ++i) z += row[i] * x[i];
            double r = 1.0 / (1.0 + exp(-z)) - p->y[j];
            for (int i = 0; i < d; ++i) g[i] += row[i] * r;
        }
        for (int i = 0; i < d; ++i) g[i] += x[i] * p->inv_s2;
    } break;
    }
}

static void pot_hvp(const pot_t* p, const double* x, const double* v, double* h) {
    int d = p->d;
    switch (p->c->potential) {
    case P_GAUSS_STD: for (int i = 0; i < d; ++i) h[i] = v[i]; break;
    case P_GAUSS_DIAG: for (int i = 0; i < d; ++i) h[i] = p->c->pot_params[i] * v[i]; break;
    case P_GAUSS_EQUICORR: {
        double s = 0; for (int i = 0; i < d; ++i) s += v[i];
        for (int i = 0; i < d; ++i) h[i] = p->alpha * v[i] - p->beta * s;
    } break;
    case P_BANANA: {
        double r = x[1] - x[0] * x[0] + 1.0;
        for (int i = 2; i < d; ++i) h[i] = v[i];
        h[0] = (1.0 - 2.0 * r + 4.0 * x[0] * x[0]) * v[0] - 2.0 * x[0] * v[1];
        h[1] = -2.0 * x[0] * v[0] + v[1];
    } break;
    case P_BANANA_README: {
        double s2 = 0; for (int i = 2; i < d; ++i) s2 += v[i];
        double s = v[0] + (v[1] - 2.0 * x[0] * v[0]) + s2;
        for (int i = 0; i < d; ++i) h[i] = s;
    } break;
    case P_GAUSS_DENSE:
        for (int i = 0; i < d; ++i) {
            double s = 0; const double* row = p->c->pot_params + (size_t)i * d;
            for (int j = 0; j < d; ++j) s += row[j] * v[j];
            h[i] = s;
        }
        break;
    case P_LOGREG: {
        for (int i = 0; i < d; ++i) h[i] = 0;
        for (int64_t j = 0; j < p->n; ++j) {
            const double* row = p->X + j * d; double z = 0, w = 0;
            for (int i = 0; i < d; ++i) { z += row[i] * x[i]; w += row[i] * v[i]; }
            double s = 1.0 / (1.0 + exp(-z));
            double r = s * (1.0 - s) * w;
            for (int i = 0; i < d; ++i) h[i] += row[i] * r;
        }
        for (int i = 0; i < d; ++i) h[i] += v[i] * p->inv_s2;
    } break;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* per-chain workspace                                                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const pdmp_oracle_cfg* c;
    pot_t pot;
    draws_t dr;
    int d, G;
    int signed_bound;
    double bound_refresh;
    /* PDMPState (Composites.jl:59-83) */
    double *x, *v;
    double t, horizon, tp, ts, exp_rv, lambda_bar, lambda_t, ar;
    int accept, adaptive;
    int errored_bound, rejected, hitting_horizon;
    double error_value_ar[5];
    /* BoundBox (Composites.jl:15-20) */
    double *grid, *box_max, *cum_sum; double step_size; int nb; /* nb = number of grid nodes (2 for constant) */
    /* scratch */
    double *xt, *vt, *g, *h, *w1, *w2, *w3, *w4;
    double *vals, *grads; /* (d x G) for the vectorised bound, G for the scalar bound */
    int status;
    int64_t n_bound_builds, n_rate_evals;
} chain_t;

static double dot(const double* a, const double* b, int d) {
    double s = 0; for (int i = 0; i < d; ++i) s += a[i] * b[i]; return s;
}

/* flow: ZigZagSamplers.jl:80, BouncyParticleSamplers.jl:32, ForwardEventChainMonteCarlo.jl:316,
 *       BoomerangSamplers.jl:31 */
static void flow(const chain_t* ch, const double* x, const double* v, double t, double* xo, double* vo) {
    int d = ch->d;
    if (ch->c->sampler == S_BOOMERANG) {
        double ct = cos(t), st = sin(t);
        for (int i = 0; i < d; ++i) {
            double xi = x[i], vi = v[i];
            xo[i] = xi * ct + vi * st;
            vo[i] = -xi * st + vi * ct;
        }
    } else {
        for (int i = 0; i < d; ++i) { double vi = v[i]; xo[i] = x[i] + vi * t; vo[i] = vi; }
    }
}

/* `sampler.rate` (always unsigned): ZigZagSamplers.jl:83-86, BouncyParticleSamplers.jl:39-42,
 * ForwardEventChainMonteCarlo.jl:20-23, BoomerangSamplers.jl:38-41 */
static double rate_unsigned(chain_t* ch, const double* x0, const double* v0, double t) {
    int d = ch->d;
    flow(ch, x0, v0, t, ch->xt, ch->vt);
    pot_grad(&ch->pot, ch->xt, ch->g);
    if (ch->c->sampler == S_ZIGZAG) {
        double s = 0;
        for (int i = 0; i < d; ++i) { double y = ch->g[i] * ch->vt[i]; s += (y > 0.0 ? y : 0.0); }
        return s;
    }
    double y = dot(ch->g, ch->vt, d);
    double r = (y > 0.0 ? y : 0.0);
    if (ch->c->sampler != S_FECMC) r += ch->c->refresh_rate;
    return r;
}

/* value (and analytic d/dt) of the scalar function the grid bound is built from:
 * `signed_rate` / `rate` per AbstractPDMP.jl:104-112, differentiated as in SURVEY.md Appendix A. */
static double bound_scalar(chain_t* ch, const double* x0, const double* v0, double t, double* dval) {
    int d = ch->d, smp = ch->c->sampler;
    flow(ch, x0, v0, t, ch->xt, ch->vt);
    pot_grad(&ch->pot, ch->xt, ch->g);
    if (dval) pot_hvp(&ch->pot, ch->xt, ch->vt, ch->h); /* dx_t/dt = v_t for both flows */
    if (smp == S_ZIGZAG) { /* scalar unsigned ZigZag bound: sum(max.(0, g.*v)) */
        double s = 0, ds = 0;
        for (int i = 0; i < d; ++i) {
            double y = ch->g[i] * ch->vt[i];
            s += (y > 0.0 ? y : 0.0);
            if (dval && !(0.0 > y)) ds += ch->h[i] * ch->vt[i];
        }
        if (dval) *dval = ds;
        return s;
    }
    double y = dot(ch->g, ch->vt, d);
    double dy = 0;
    if (dval) {
        dy = dot(ch->h, ch->vt, d);
        if (smp == S_BOOMERANG) dy -= dot(ch->g, ch->xt, d); /* dv_t/dt = -x_t */
    }
    double extra = (smp == S_FECMC) ? 0.0 : ch->c->refresh_rate;
    if (ch->signed_bound) { if (dval) *dval = dy; return y + extra; }
    if (dval) *dval = (0.0 > y) ? 0.0 : dy;
    return (y > 0.0 ? y : 0.0) + extra;
}

/* ZigZag `signed_rate_vect` / `rate_vect` (ZigZagSamplers.jl:88-98) */
static void bound_vect(chain_t* ch, const double* x0, const double* v0, double t, double* val, double* dval) {
    int d = ch->d;
    flow(ch, x0, v0, t, ch->xt, ch->vt);
    pot_grad(&ch->pot, ch->xt, ch->g);
    if (dval) pot_hvp(&ch->pot, ch->xt, ch->vt, ch->h);
    for (int i = 0; i < d; ++i) {
        double y = ch->g[i] * ch->vt[i];
        if (ch->signed_bound) {
            val[i] = y; if (dval) dval[i] = ch->h[i] * ch->vt[i];
        } else {
            val[i] = (y > 0.0 ? y : 0.0);
            if (dval) dval[i] = (0.0 > y) ? 0.0 : ch->h[i] * ch->vt[i];
        }
    }
}

/* range(0, stop=h, length=G) (UpperBound.jl:94,204): nodes ~ correctly rounded k*h/(G-1), last == h */
static void grid_times(double h, int G, double* t) {
    double m = (double)(G - 1);
    for (int k = 0; k < G; ++k) {
        double kk = (double)k;
        double p = kk * h, e = fma(kk, h, -p);
        double q = p / m;
        double r = fma(-q, m, p) + e;
        t[k] = q + r / m;
    }
    t[G - 1] = h;
}

/* Optim.jl Brent (third party, not vendored; compat "1.9.4, 2"); call site UpperBound.jl:24 */
static double brent_max_rate(chain_t* ch, const double* x0, const double* v0, double lo, double hi) {
    const double sqrt_eps = 1.4901161193847656e-08, eps = 2.220446049250313e-16;
    const double golden = 0.5 * (3.0 - sqrt(5.0));
    double x = lo + golden * (hi - lo);
    double fx = -rate_unsigned(ch, x0, v0, x);
    double step = 0, old_step = 0, w = x, vv = x, fw = fx, fv = fx;
    for (int it = 0; it < 1000;) {
        double p = 0, q = 0;
        double tol = sqrt_eps * fabs(x) + eps;
        double mid = (hi + lo) / 2;
        if (fabs(x - mid) <= 2 * tol - (hi - lo) / 2) break;
        ++it;
        if (fabs(old_step) > tol) {
            double r = (x - w) * (fx - fv);
            q = (x - vv) * (fx - fw);
            p = (x - vv) * q - (x - w) * r;
            q = 2 * (q - r);
            if (q > 0) p = -p; else q = -q;
        }
        if (fabs(p) < fabs(q * old_step / 2) && p < q * (hi - x) && p < q * (x - lo)) {
            old_step = step;
            step = p / q;
            double xt = x + step;
            if ((xt - lo) < 2 * tol || (hi - xt) < 2 * tol) step = (x < mid) ? tol : -tol;
        } else {
            old_step = (x < mid) ? hi - x : lo - x;
            step = golden * old_step;
        }
        double u = (fabs(step) >= tol) ? x + step : x + ((step > 0) ? tol : -tol);
        double fu = -rate_unsigned(ch, x0, v0, u);
        if (fu < fx) {
            if (u < x) hi = x; else lo = x;
            vv = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
        } else {
            if (u < x) lo = u; else hi = u;
            if (fu <= fw || w == x) { vv = w; fv = fw; w = u; fw = fu; }
            else if (fu <= fv || vv == x || vv == w) { vv = u; fv = fu; }
        }
    }
    return -fx;
}

/* finite_difference_derivative (UpperBound.jl:50-76), scalar func */
static double fd_scalar(chain_t* ch, const double* x0, const double* v0, double t, double fx, double horizon) {
    const double sqrt_eps = 1.4901161193847656e-08;
    double h = sqrt_eps * fmax(1.0, fabs(t));
    double xm = fmax(0.0, t - h), xp = fmin(horizon, t + h);
    if (xp == xm) return fx - fx;
    if (xm == t) return (bound_scalar(ch, x0, v0, xp, NULL) - fx) / (xp - t);
    if (xp == t) return (fx - bound_scalar(ch, x0, v0, xm, NULL)) / (t - xm);
    return (bound_scalar(ch, x0, v0, xp, NULL) - bound_scalar(ch, x0, v0, xm, NULL)) / (xp - xm);
}
/* same, vector func; out[i] */
static void fd_vect(chain_t* ch, const double* x0, const double* v0, double t, const double* fx, double horizon,
                    double* out) {
    const double sqrt_eps = 1.4901161193847656e-08;
    int d = ch->d;
    double h = sqrt_eps * fmax(1.0, fabs(t));
    double xm = fmax(0.0, t - h), xp = fmin(horizon, t + h);
    if (xp == xm) { for (int i = 0; i < d; ++i) out[i] = fx[i] - fx[i]; return; }
    if (xm == t) {
        bound_vect(ch, x0, v0, xp, ch->w1, NULL);
        for (int i = 0; i < d; ++i) out[i] = (ch->w1[i] - fx[i]) / (xp - t);
    } else if (xp == t) {
        bound_vect(ch, x0, v0, xm, ch->w1, NULL);
        for (int i = 0; i < d; ++i) out[i] = (fx[i] - ch->w1[i]) / (t - xm);
    } else {
        bound_vect(ch, x0, v0, xp, ch->w1, NULL);
        bound_vect(ch, x0, v0, xm, ch->w2, NULL);
        for (int i = 0; i < d; ++i) out[i] = (ch->w1[i] - ch->w2[i]) / (xp - xm);
    }
}

static double clamp_pos(double pos, double step) {
    if (isnan(pos)) pos = 0.0;                 /* replace(NaN => 0.0) */
    pos = fmin(fmax(pos, 0.0), step);          /* clamp.(pos, 0, step) */
    return pos;
}

/* state.upper_bound_func: AbstractPDMP.jl:121-136 dispatching to UpperBound.jl:18-36 / 92-137 / 203-247 */
static void build_bound(chain_t* ch, const double* x0, const double* v0, double horizon) {
    const pdmp_oracle_cfg* c = ch->c;
    int d = ch->d, G = c->grid_size;
    ch->n_bound_builds++;
    if (G == 0) { /* upper_bound_constant */
        double m = brent_max_rate(ch, x0, v0, 0.0, horizon);
        ch->grid[0] = 0.0; ch->grid[1] = horizon;
        ch->box_max[0] = m + 0.0;
        ch->cum_sum[0] = 0.0; ch->cum_sum[1] = ch->box_max[0] * (horizon - 0.0);
        ch->step_size = horizon - 0.0; ch->nb = 2;
        return;
    }
    grid_times(horizon, G, ch->grid);
    double step = ch->grid[1] - ch->grid[0];
    ch->step_size = step; ch->nb = G;
    if (!c->vectorized_bound) { /* upper_bound_grid */
        double* val = ch->vals; double* gr = ch->grads;
        for (int k = 0; k < G; ++k) {
            double dv = 0;
            val[k] = bound_scalar(ch, x0, v0, ch->grid[k], c->deriv_mode == D_JVP ? &dv : NULL);
            gr[k] = dv;
        }
        if (c->deriv_mode == D_FD)
            for (int k = 0; k < G; ++k) gr[k] = fd_scalar(ch, x0, v0, ch->grid[k], val[k], horizon);
        double cs = 0; ch->cum_sum[0] = 0.0;
        for (int k = 0; k < G - 1; ++k) {
            double pos = (val[k] - val[k + 1] + gr[k + 1] * step) / (gr[k + 1] - gr[k]);
            pos = clamp_pos(pos, step);
            double inter = val[k] + gr[k] * pos;
            double b = fmax(val[k], val[k + 1]);
            b = fmax(b, inter); b = fmax(b, 0.0);
            b += ch->bound_refresh;
            ch->box_max[k] = b;
            cs += b;
            ch->cum_sum[k + 1] = cs * step;
        }
        return;
    }
    /* upper_bound_grid_vect: values/grads are d x G, column k at vals + k*d */
    for (int k = 0; k < G; ++k)
        bound_vect(ch, x0, v0, ch->grid[k], ch->vals + (size_t)k * d,
                   c->deriv_mode == D_JVP ? ch->grads + (size_t)k * d : NULL);
    if (c->deriv_mode == D_FD)
        for (int k = 0; k < G; ++k)
            fd_vect(ch, x0, v0, ch->grid[k], ch->vals + (size_t)k * d, horizon, ch->grads + (size_t)k * d);
    for (int k = 0; k < G; ++k) { ch->cum_sum[k] = 0.0; if (k < G - 1) ch->box_max[k] = 0.0; }
    for (int i = 0; i < d; ++i) {
        double cs = 0;
        for (int k = 0; k < G - 1; ++k) {
            double vl = ch->vals[(size_t)k * d + i], vr = ch->vals[(size_t)(k + 1) * d + i];
            double gl = ch->grads[(size_t)k * d + i], gr = ch->grads[(size_t)(k + 1) * d + i];
            /* absolute-time intersection (UpperBound.jl:229) clamped to [0, step] (:235), used as offset (:237) */
            double pos = (vl - vr + gr * ch->grid[k + 1] - gl * ch->grid[k]) / (gr - gl);
            pos = clamp_pos(pos, step);
            double inter = vl + gl * pos;
            double b = fmax(vl, vr); b = fmax(b, inter); b = fmax(b, 0.0);
            ch->box_max[k] += b;            /* sum over coordinates (:246) */
            cs += b;
            ch->cum_sum[k + 1] += cs * step; /* per-coordinate cumsum*step, summed over coordinates */
        }
    }
}

/* next_event: UpperBound.jl:264-273 */
static void next_event(const chain_t* ch, double e, double* tp, double* lb) {
    int n = ch->nb, idx = 0; /* 0-based first index with cum_sum[idx] >= e */
    while (idx < n && ch->cum_sum[idx] < e) ++idx;
    if (idx >= n) { *tp = INFINITY; *lb = ch->box_max[n - 2]; return; }
    *tp = ch->grid[idx - 1] + (e - ch->cum_sum[idx - 1]) / (ch->cum_sum[idx] - ch->cum_sum[idx - 1]) * ch->step_size;
    *lb = ch->box_max[idx - 1];
}

/* ------------------------------------------------------------------------------------------ */
/* velocity jumps                                                                              */
/* ------------------------------------------------------------------------------------------ */
static void jump_zigzag(chain_t* ch) { /* ZigZagSamplers.jl:101-107 + Distributions.jl categorical scan */
    int d = ch->d;
    pot_grad(&ch->pot, ch->x, ch->g);
    double s = 0;
    for (int i = 0; i < d; ++i) { double y = ch->g[i] * ch->v[i]; ch->w1[i] = (y > 0.0 ? y : 0.0); s += ch->w1[i]; }
    double sp = 0; int ok = 1;
    for (int i = 0; i < d; ++i) { ch->w1[i] = ch->w1[i] / s; if (!(ch->w1[i] >= 0.0)) ok = 0; sp += ch->w1[i]; }
    if (!ok || !(fabs(sp - 1.0) <= 1.4901161193847656e-08 * fmax(fabs(sp), 1.0))) { ch->status = ST_NOT_PROBVEC; return; }
    double u = draw_uniform(&ch->dr);
    double cp = ch->w1[0]; int i = 0;
    while (cp <= u && i < d - 1) { ++i; cp += ch->w1[i]; }
    ch->v[i] *= -1;
}

static void jump_bps(chain_t* ch) { /* BouncyParticleSamplers.jl:50-74 */
    int d = ch->d;
    pot_grad(&ch->pot, ch->x, ch->g);
    double gv = dot(ch->g, ch->v, d);
    double bounce = (gv > 0.0 ? gv : 0.0);
    double prob = bounce / (bounce + ch->c->refresh_rate);
    double u = draw_uniform(&ch->dr);
    if (u < prob) {
        double gg = dot(ch->g, ch->g, d);
        if (gg == 0) return;
        double scale = 2 * dot(ch->v, ch->g, d) / gg;
        for (int i = 0; i < d; ++i) ch->v[i] = ch->v[i] - scale * ch->g[i];
    } else {
        for (int i = 0; i < d; ++i) ch->v[i] = draw_normal(&ch->dr);
        if (!ch->c->gaussian_velocity) {
            double nv = sqrt(dot(ch->v, ch->v, d));
            for (int i = 0; i < d; ++i) ch->v[i] = ch->v[i] / nv;
        }
    }
}

static void jump_boomerang(chain_t* ch) { /* BoomerangSamplers.jl:49-67 */
    int d = ch->d;
    pot_grad(&ch->pot, ch->x, ch->g);
    for (int i = 0; i < d; ++i) ch->g[i] = ch->g[i] - ch->x[i];
    double gv = dot(ch->g, ch->v, d);
    double bounce = (gv > 0.0 ? gv : 0.0);
    double prob = bounce / (bounce + ch->c->refresh_rate);
    double u = draw_uniform(&ch->dr);
    if (u < prob) {
        double ng = sqrt(dot(ch->g, ch->g, d));
        for (int i = 0; i < d; ++i) ch->w1[i] = ch->g[i] / ng;
        double ve = dot(ch->v, ch->w1, d);
        for (int i = 0; i < d; ++i) ch->v[i] = ch->v[i] - 2 * ve * ch->w1[i];
    } else {
        for (int i = 0; i < d; ++i) ch->v[i] = draw_normal(&ch->dr);
    }
}

static double sgn(double x) { return (x > 0) - (x < 0) + (isnan(x) ? x : 0.0); }

static void jump_fecmc(chain_t* ch) { /* ForwardEventChainMonteCarlo.jl:132-218, 60-113 */
    const pdmp_oracle_cfg* c = ch->c;
    int d = ch->d;
    double sf = c->speed_factor;
    double u = draw_uniform(&ch->dr);
    double rho = -sqrt(1 - pow(u, 2.0 / (d - 1)));
    if (sf != 1.0) rho = sf * rho;
    double* n = ch->g;
    pot_grad(&ch->pot, ch->x, n);
    double ng = sqrt(dot(n, n, d));
    if (ng == 0) { for (int i = 0; i < d; ++i) n[i] = 0.0; }
    else { for (int i = 0; i < d; ++i) n[i] /= ng; }
    double vn = dot(ch->v, n, d);
    double* vo = ch->w1;
    for (int i = 0; i < d; ++i) vo[i] = ch->v[i] - vn * n[i];
    if (sqrt(dot(vo, vo, d)) < 1e-10) {
        for (int i = 0; i < d; ++i) vo[i] = draw_normal(&ch->dr);
        double a = dot(vo, n, d);
        for (int i = 0; i < d; ++i) vo[i] = vo[i] - a * n[i];
    }
    double u2 = draw_uniform(&ch->dr);
    double rad = (sf != 1.0) ? sqrt(sf * sf - rho * rho) : sqrt(1 - rho * rho);
    if (u2 >= c->mix_p) {
        double nvo = sqrt(dot(vo, vo, d));
        for (int i = 0; i < d; ++i) ch->v[i] = vo[i] / nvo * rad + rho * n[i];
        return;
    }
    double* prop = ch->w2;
    if (c->switch_) { /* _orthogonal_switch */
        double *g1 = ch->w3, *g2 = ch->w4;
        for (int i = 0; i < d; ++i) { g1[i] = draw_normal(&ch->dr); g2[i] = draw_normal(&ch->dr); } /* randn(key,2,dim) column-major */
        double a1 = dot(g1, n, d), a2 = dot(g2, n, d);
        for (int i = 0; i < d; ++i) { g1[i] = g1[i] - a1 * n[i]; g2[i] = g2[i] - a2 * n[i]; }
        double n1 = sqrt(dot(g1, g1, d));
        for (int i = 0; i < d; ++i) g1[i] = g1[i] / n1;          /* e1 */
        double b = dot(g2, g1, d);
        for (int i = 0; i < d; ++i) g2[i] = g2[i] - b * g1[i];
        double n2 = sqrt(dot(g2, g2, d));
        for (int i = 0; i < d; ++i) g2[i] /= n2;                 /* e2 */
        double c1 = dot(vo, g1, d), c2 = dot(vo, g2, d);
        /* v_r = vo - c1 e1 - c2 e2 ; vo_new = v_r + e2 c1 + e1 c2 */
        if (c->ran_p) {
            double th = draw_uniform(&ch->dr) * 2 * 3.14159265358979323846;
            double ct = cos(th), st = sin(th);
            for (int i = 0; i < d; ++i) {
                double vr = vo[i] - c1 * g1[i] - c2 * g2[i];
                prop[i] = vr + (ct * g1[i] + st * g2[i]) * c1 + (st * g1[i] - ct * g2[i]) * c2;
            }
        } else {
            for (int i = 0; i < d; ++i) {
                double vr = vo[i] - c1 * g1[i] - c2 * g2[i];
                prop[i] = vr + g2[i] * c1 + g1[i] * c2;
            }
        }
        if (c->positive) {
            double s = sgn(dot(vo, prop, d));
            for (int i = 0; i < d; ++i) prop[i] *= s;
        }
    } else { /* _full_refresh */
        for (int i = 0; i < d; ++i) prop[i] = draw_normal(&ch->dr);
        double nw = sqrt(dot(prop, prop, d));
        for (int i = 0; i < d; ++i) prop[i] = prop[i] / nw;
        double a = dot(prop, n, d);
        for (int i = 0; i < d; ++i) prop[i] = prop[i] - a * n[i];
    }
    double np_ = sqrt(dot(prop, prop, d));
    for (int i = 0; i < d; ++i) ch->v[i] = prop[i] / np_ * rad + rho * n[i];
}

static void velocity_jump(chain_t* ch) {
    switch (ch->c->sampler) {
    case S_ZIGZAG: jump_zigzag(ch); break;
    case S_BPS: jump_bps(ch); break;
    case S_FECMC: jump_fecmc(ch); break;
    default: jump_boomerang(ch); break;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* thinning loop: SamplingLoopInplace.jl                                                       */
/* ------------------------------------------------------------------------------------------ */
static void flow_inplace(chain_t* ch, double t) { flow(ch, ch->x, ch->v, t, ch->x, ch->v); }

static void move_to_horizon(chain_t* ch) { /* :87-101 */
    flow_inplace(ch, ch->horizon);
    ch->ts += ch->horizon;
    ch->hitting_horizon += 1;
    ch->horizon = ch->adaptive ? ch->horizon * 1.01 : ch->horizon;
}
static void move_to_horizon2(chain_t* ch) { /* :205-217 */
    flow_inplace(ch, ch->horizon);
    ch->ts += ch->horizon;
    ch->hitting_horizon += 1;
}
static void erroneous_acceptance_rate(chain_t* ch) { /* :131-151 */
    double horizon = ch->horizon / 2;
    build_bound(ch, ch->x, ch->v, horizon);
    double e = draw_exp(&ch->dr);
    double tp, lb; next_event(ch, e, &tp, &lb);
    ch->horizon = ch->adaptive ? horizon : ch->horizon;
    ch->tp = tp; ch->exp_rv = e; ch->lambda_bar = lb;
    ch->errored_bound += 1;
    ch->error_value_ar[ch->errored_bound % 5] = ch->ar;
}
static void if_accept(chain_t* ch) { /* :170-186 */
    flow_inplace(ch, ch->tp);
    velocity_jump(ch);
    ch->t = ch->t + ch->tp + ch->ts;
    ch->ts = 0.0; ch->tp = 0.0; ch->accept = 1;
}
static void if_reject(chain_t* ch) { /* :188-203 */
    double e = ch->exp_rv + draw_exp(&ch->dr);
    double tp, lb; next_event(ch, e, &tp, &lb);
    ch->horizon = ch->adaptive ? ch->horizon / 1.04 : ch->horizon;
    ch->tp = tp; ch->exp_rv = e; ch->lambda_bar = lb;
    ch->rejected += 1;
}
static void ac_step_with_proxy(chain_t* ch) { /* :153-168 */
    int accept = draw_uniform(&ch->dr) < ch->ar;
    ch->accept = accept;
    if (accept) if_accept(ch); else if_reject(ch);
    if (!ch->accept && ch->tp > ch->horizon) move_to_horizon2(ch); /* min(tp, tt = Inf) */
}
static void ac_step(chain_t* ch) { /* :113-129 */
    ch->n_rate_evals++;
    double lt = rate_unsigned(ch, ch->x, ch->v, ch->tp);
    double ar = lt / ch->lambda_bar;
    ch->lambda_t = lt; ch->ar = ar;
    if (ar > 1.0) erroneous_acceptance_rate(ch); else ac_step_with_proxy(ch);
}
static void one_step_of_thinning(chain_t* ch) { /* :65-85 */
    build_bound(ch, ch->x, ch->v, ch->horizon);
    double e = draw_exp(&ch->dr);
    double tp, lb; next_event(ch, e, &tp, &lb);
    ch->tp = tp; ch->exp_rv = e; ch->lambda_bar = lb;
    if (tp > ch->horizon) move_to_horizon(ch);
    else { /* moves_until_horizon! :103-111 */
        ch->accept = 0;
        while (ch->tp < ch->horizon && !ch->accept && ch->status == ST_OK && !ch->dr.exhausted) ac_step(ch);
    }
}
static void get_event_state(chain_t* ch, int64_t max_inner) { /* :27-39 */
    ch->errored_bound = 0; ch->rejected = 0; ch->hitting_horizon = 0;
    for (int j = 0; j < 5; ++j) ch->error_value_ar[j] = 0.0;
    int64_t it = 0;
    while (!ch->accept) {
        one_step_of_thinning(ch);
        if (ch->dr.exhausted) { ch->status = ST_TAPE_EXHAUSTED; return; }
        if (ch->status != ST_OK) return;
        if (++it > max_inner) { ch->status = ST_ITER_LIMIT; return; }
    }
    ch->accept = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* public entry points (ctypes)                                                                */
/* ------------------------------------------------------------------------------------------ */
static void normalise_cfg(pdmp_oracle_cfg* c) {
    /* constructor rewrites: ZigZagSamplers.jl:73-78, BouncyParticleSamplers.jl:29-37,
     * ForwardEventChainMonteCarlo.jl:306-323, BoomerangSamplers.jl:27-36 */
    if (c->tmax == 0.0) { c->tmax = 1.0; c->adaptive = 1; }
    if (c->sampler == S_ZIGZAG) { if (c->signed_bound && !c->vectorized_bound) c->signed_bound = 0; }
    else c->vectorized_bound = 0;
    if (c->sampler == S_FECMC) { c->refresh_rate = 0.0; if (c->dim == 2) c->mix_p = 0.0; }
}

typedef struct {
    double *X, *V, *t, *horizon, *ar, *error_value_ar;
    int32_t *errored_bound, *rejected, *hitting_horizon;
    int64_t* tape_pos; /* [C][n_sk][3]: draws consumed (E,U,N) up to and including event k (test aid) */
} pdmp_oracle_hist; /* chain-major: chain c at X + c*d*n_sk etc. (each slab is a Julia Matrix(d, n_sk)) */

static void record(const pdmp_oracle_hist* h, int64_t c, int64_t n_sk, int64_t k, const chain_t* ch) { /* Composites.jl:239-260 */
    int d = ch->d;
    if (h->X) memcpy(h->X + ((size_t)c * n_sk + k) * d, ch->x, sizeof(double) * d);
    if (h->V) memcpy(h->V + ((size_t)c * n_sk + k) * d, ch->v, sizeof(double) * d);
    size_t o = (size_t)c * n_sk + k;
    if (h->t) h->t[o] = ch->t;
    if (h->horizon) h->horizon[o] = ch->horizon;
    if (h->ar) h->ar[o] = ch->ar;
    if (h->errored_bound) h->errored_bound[o] = ch->errored_bound;
    if (h->error_value_ar) for (int j = 0; j < 5; ++j) h->error_value_ar[o * 5 + j] = ch->error_value_ar[j];
    if (h->rejected) h->rejected[o] = ch->rejected;
    if (h->hitting_horizon) h->hitting_horizon[o] = ch->hitting_horizon;
    if (h->tape_pos) { h->tape_pos[o * 3] = ch->dr.pE; h->tape_pos[o * 3 + 1] = ch->dr.pU; h->tape_pos[o * 3 + 2] = ch->dr.pN; }
}

/*
 * sample_skeleton (src/sample.jl:253-284) for n_chains independent chains.
 * xinit/vinit: d x n_chains column-major.  Tape streams (draw_mode 0): chain c reads E + c*nE etc.
 * status[c]: ST_*; counters[2c..2c+1] = (bound builds, rate evals); tape_used[3c..] = consumed draws.
 * Returns 0, or -1 on invalid arguments.
 */
int pdmp_oracle_sample_skeleton(const pdmp_oracle_cfg* cfg_in, int64_t n_chains, int64_t n_sk,
                                const double* xinit, const double* vinit, int draw_mode, uint64_t seed,
                                int64_t chain_offset, const double* tapeE, int64_t nE, const double* tapeU,
                                int64_t nU, const double* tapeN, int64_t nN, const pdmp_oracle_hist* hist,
                                int32_t* status, int64_t* counters, int64_t* tape_used, int nthreads) {
    pdmp_oracle_cfg cfg = *cfg_in;
    normalise_cfg(&cfg);
    if (n_sk <= 0 || cfg.dim <= 0 || cfg.grid_size < 0 || cfg.grid_size == 1) return -1;
    if (cfg.sampler == S_FECMC && cfg.dim < 2) return -1;
    int d = cfg.dim, G = cfg.grid_size > 2 ? cfg.grid_size : 2;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < n_chains; ++c) {
        chain_t ch; memset(&ch, 0, sizeof(ch));
        ch.c = &cfg; ch.d = d; ch.G = cfg.grid_size;
        pot_init(&ch.pot, &cfg);
        size_t nvec = 10;
        double* buf = (double*)calloc((size_t)d * nvec + (size_t)2 * d * G + 3 * (size_t)G + 8, sizeof(double));
        double* p = buf;
        ch.x = p; p += d; ch.v = p; p += d; ch.xt = p; p += d; ch.vt = p; p += d; ch.g = p; p += d; ch.h = p; p += d;
        ch.w1 = p; p += d; ch.w2 = p; p += d; ch.w3 = p; p += d; ch.w4 = p; p += d;
        ch.vals = p; p += (size_t)d * G; ch.grads = p; p += (size_t)d * G;
        ch.grid = p; p += G; ch.box_max = p; p += G; ch.cum_sum = p; p += G;
        ch.signed_bound = cfg.signed_bound;
        ch.bound_refresh = cfg.signed_bound ? cfg.refresh_rate : 0.0; /* AbstractPDMP.jl:104-112 */
        memcpy(ch.x, xinit + (size_t)c * d, sizeof(double) * d);
        memcpy(ch.v, vinit + (size_t)c * d, sizeof(double) * d);
        ch.t = 0.0; ch.horizon = cfg.tmax; ch.adaptive = cfg.adaptive;
        ch.dr.mode = draw_mode; ch.dr.seed = seed; ch.dr.chain = (uint64_t)(chain_offset + c);
        if (!draw_mode) {
            ch.dr.E = tapeE + (size_t)c * nE; ch.dr.nE = nE;
            ch.dr.U = tapeU + (size_t)c * nU; ch.dr.nU = nU;
            ch.dr.N = tapeN + (size_t)c * nN; ch.dr.nN = nN;
        }
        record(hist, c, n_sk, 0, &ch);
        int64_t k = 1;
        for (; k < n_sk; ++k) {
            ch.dr.event = (uint64_t)k; ch.dr.sE = ch.dr.sU = ch.dr.sN = 0;
            get_event_state(&ch, 1000000);
            if (ch.status != ST_OK) break;
            record(hist, c, n_sk, k, &ch);
        }
        if (status) status[c] = ch.status;
        if (counters) { counters[2 * c] = ch.n_bound_builds; counters[2 * c + 1] = ch.n_rate_evals; }
        if (tape_used) { tape_used[3 * c] = ch.dr.pE; tape_used[3 * c + 1] = ch.dr.pU; tape_used[3 * c + 2] = ch.dr.pN; }
        pot_free(&ch.pot);
        free(buf);
    }
    return 0;
}

/*
 * sample_skeleton(sampler, T::Float64, ...) (src/sample.jl:323-439) for n_chains chains: events with t <= T are
 * recorded; the first event beyond T is replaced by the point at t = T obtained by flowing the PREVIOUS state
 * (x_prev, v_prev, t_prev) for T - t_prev, with zeroed statistics (:385-420).  `cap` columns per chain;
 * ncols[c] = columns used, or -1 when cap was too small.
 */
int pdmp_oracle_sample_skeleton_until(const pdmp_oracle_cfg* cfg_in, int64_t n_chains, int64_t cap, double T,
                                      const double* xinit, const double* vinit, int draw_mode, uint64_t seed,
                                      int64_t chain_offset, const double* tapeE, int64_t nE, const double* tapeU,
                                      int64_t nU, const double* tapeN, int64_t nN, const pdmp_oracle_hist* hist,
                                      int32_t* status, int64_t* ncols, int nthreads) {
    pdmp_oracle_cfg cfg = *cfg_in;
    normalise_cfg(&cfg);
    if (cap <= 0 || cfg.dim <= 0 || cfg.grid_size < 0 || cfg.grid_size == 1 || !(T >= 0) || isinf(T)) return -1;
    int d = cfg.dim, G = cfg.grid_size > 2 ? cfg.grid_size : 2;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < n_chains; ++c) {
        chain_t ch; memset(&ch, 0, sizeof(ch));
        ch.c = &cfg; ch.d = d; ch.G = cfg.grid_size;
        pot_init(&ch.pot, &cfg);
        double* buf = (double*)calloc((size_t)d * 12 + (size_t)2 * d * G + 3 * (size_t)G + 8, sizeof(double));
        double* p = buf;
        ch.x = p; p += d; ch.v = p; p += d; ch.xt = p; p += d; ch.vt = p; p += d; ch.g = p; p += d; ch.h = p; p += d;
        ch.w1 = p; p += d; ch.w2 = p; p += d; ch.w3 = p; p += d; ch.w4 = p; p += d;
        double* x_prev = p; p += d; double* v_prev = p; p += d;
        ch.vals = p; p += (size_t)d * G; ch.grads = p; p += (size_t)d * G;
        ch.grid = p; p += G; ch.box_max = p; p += G; ch.cum_sum = p; p += G;
        ch.signed_bound = cfg.signed_bound;
        ch.bound_refresh = cfg.signed_bound ? cfg.refresh_rate : 0.0;
        memcpy(ch.x, xinit + (size_t)c * d, sizeof(double) * d);
        memcpy(ch.v, vinit + (size_t)c * d, sizeof(double) * d);
        ch.t = 0.0; ch.horizon = cfg.tmax; ch.adaptive = cfg.adaptive;
        ch.dr.mode = draw_mode; ch.dr.seed = seed; ch.dr.chain = (uint64_t)(chain_offset + c);
        if (!draw_mode) {
            ch.dr.E = tapeE + (size_t)c * nE; ch.dr.nE = nE;
            ch.dr.U = tapeU + (size_t)c * nU; ch.dr.nU = nU;
            ch.dr.N = tapeN + (size_t)c * nN; ch.dr.nN = nN;
        }
        int64_t k = 0;
        record(hist, c, cap, 0, &ch);
        k = 1;
        int overflow = 0;
        uint64_t ev = 0;
        while (ch.t < T) {
            memcpy(x_prev, ch.x, sizeof(double) * d);
            memcpy(v_prev, ch.v, sizeof(double) * d);
            double t_prev = ch.t;
            ch.dr.event = ++ev; ch.dr.sE = ch.dr.sU = ch.dr.sN = 0;
            get_event_state(&ch, 1000000);
            if (ch.status != ST_OK) break;
            if (k >= cap) { overflow = 1; break; }
            if (ch.t <= T) {
                record(hist, c, cap, k, &ch);
                ++k;
            } else { /* overshoot: the t = T point by flow from the previous state */
                double tau = T - t_prev;
                flow(&ch, x_prev, v_prev, tau, ch.x, ch.v);
                ch.t = T;
                ch.ar = 0.0; ch.errored_bound = 0; ch.rejected = 0; ch.hitting_horizon = 0;
                for (int j = 0; j < 5; ++j) ch.error_value_ar[j] = 0.0;
                record(hist, c, cap, k, &ch);
                ++k;
                break;
            }
        }
        if (status) status[c] = ch.status;
        if (ncols) ncols[c] = overflow ? -1 : k;
        pot_free(&ch.pot);
        free(buf);
    }
    return 0;
}

/* One bound build + inversion, exposed for known-answer tests: fills grid/box_max/cum_sum (G entries,
 * or 2 for grid_size 0) and returns step_size. */
double pdmp_oracle_bound(const pdmp_oracle_cfg* cfg_in, const double* x, const double* v, double horizon,
                         double* grid, double* box_max, double* cum_sum) {
    pdmp_oracle_cfg cfg = *cfg_in; normalise_cfg(&cfg);
    int d = cfg.dim, G = cfg.grid_size > 2 ? cfg.grid_size : 2;
    chain_t ch; memset(&ch, 0, sizeof(ch));
    ch.c = &cfg; ch.d = d; pot_init(&ch.pot, &cfg);
    double* buf = (double*)calloc((size_t)d * 10 + (size_t)2 * d * G + 3 * (size_t)G + 8, sizeof(double));
    double* p = buf;
    ch.x = p; p += d; ch.v = p; p += d; ch.xt = p; p += d; ch.vt = p; p += d; ch.g = p; p += d; ch.h = p; p += d;
    ch.w1 = p; p += d; ch.w2 = p; p += d; ch.w3 = p; p += d; ch.w4 = p; p += d;
    ch.vals = p; p += (size_t)d * G; ch.grads = p; p += (size_t)d * G;
    ch.grid = p; p += G; ch.box_max = p; p += G; ch.cum_sum = p; p += G;
    ch.signed_bound = cfg.signed_bound; ch.bound_refresh = cfg.signed_bound ? cfg.refresh_rate : 0.0;
    build_bound(&ch, x, v, horizon);
    for (int k = 0; k < ch.nb; ++k) { grid[k] = ch.grid[k]; cum_sum[k] = ch.cum_sum[k]; if (k < ch.nb - 1) box_max[k] = ch.box_max[k]; }
    double s = ch.step_size;
    pot_free(&ch.pot); free(buf);
    return s;
}

/* sample_from_skeleton: src/sample.jl:475-513, one chain; flow_kind 0 linear / 1 rotation. out is d x N
 * (or (2d+1) x N when !discard_vt), column-major. */
int pdmp_oracle_sample_from_skeleton(int flow_kind, int d, int64_t n_sk, const double* X, const double* V,
                                     const double* t, int64_t N, int discard_vt, double* out) {
    if (N <= 0) return -1;
    double dt = t[n_sk - 1] / (double)N;
    int64_t i = 0; int ld = discard_vt ? d : 2 * d + 1;
    for (int64_t j = 1; j <= N; ++j) {
        double tm = (double)j * dt;
        while (i < n_sk - 1 && t[i + 1] <= tm) ++i;
        double tau = tm - t[i];
        const double *x0 = X + (size_t)i * d, *v0 = V + (size_t)i * d;
        double* o = out + (size_t)(j - 1) * ld;
        if (flow_kind == 1) {
            double ct = cos(tau), st = sin(tau);
            for (int a = 0; a < d; ++a) { o[a] = x0[a] * ct + v0[a] * st; if (!discard_vt) o[d + a] = -x0[a] * st + v0[a] * ct; }
        } else {
            for (int a = 0; a < d; ++a) { o[a] = x0[a] + v0[a] * tau; if (!discard_vt) o[d + a] = v0[a]; }
        }
        if (!discard_vt) o[2 * d] = tm;
    }
    return 0;
}

/* draw specification exposed for tests (device Philox must reproduce these bit-for-bit up to libm ulp) */
void pdmp_oracle_draws(uint64_t seed, uint64_t chain, uint64_t event, int32_t n, double* E, double* U, double* N) {
    for (int32_t s = 0; s < n; ++s) {
        E[s] = pdmp_draw_exp(seed, chain, event, (uint32_t)s);
        U[s] = pdmp_draw_uniform(seed, chain, event, (uint32_t)s);
        N[s] = pdmp_draw_normal(seed, chain, event, (uint32_t)s);
    }
}
void pdmp_oracle_philox_raw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    pdmp_philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}
int pdmp_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
