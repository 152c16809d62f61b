// RV_diagnostic on the device: realised volatility of the potential along each chain's skeleton.
//
// Follows RV_diagnostic (src/diagnostic.jl:37-75): B+1 equidistant boundaries on [0, t[end]], the position at a
// boundary by the linear interpolation of _history_position_linear! (src/diagnostic.jl:23-35; the offline diagnostic is
// linear for every sampler, flow_kind 0), RV = sum_b (U(x(t_b)) - U(x(t_{b-1})))^2 / t[end], with x(t_0) := X[:, 1].
// flow_kind 1 evaluates the boundaries with the Boomerang rotation instead, which is what the online variant
// sample_skeleton_with_diagnostic (src/sample.jl:75-236) accumulates through sampler.flow.
//
// One CTA per chain; a warp per boundary (lanes stride the coordinates: X/V rows are contiguous, so the reads are
// coalesced), U values parked in a global scratch row, then one block reduction of the squared increments.
// U plugins: the Gaussians, the banana, and the logistic-regression posterior (a pass over X per boundary).
#include "common.cuh"
#include <math_constants.h>

namespace pdmpflux {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// U(x) for x_j = x0[j] * ca + v0[j] * cb, all lanes of the warp cooperate and return the value
__device__ double potential_value(int kind, const PotParams& pp, int d, const double* __restrict__ x0,
                                  const double* __restrict__ v0, double ca, double cb, int lane) {
    double sq = 0.0, s1 = 0.0;
    const int skip = (kind == PDMPFLUX_BANANA || kind == PDMPFLUX_BANANA_README_SCALAR) ? 2 : 0;
    for (int j = skip + lane; j < d; j += 32) {
        const double xj = x0[j] * ca + v0[j] * cb;
        if (kind == PDMPFLUX_GAUSS_DIAG) sq += __ldg(pp.vec + j) * xj * xj;
        else sq += xj * xj;
        s1 += xj;
    }
    sq = warp_sum(sq);
    switch (kind) {
        case PDMPFLUX_GAUSS_STD:
        case PDMPFLUX_GAUSS_DIAG: return 0.5 * sq;
        case PDMPFLUX_GAUSS_EQUICORR: { s1 = warp_sum(s1); return 0.5 * (pp.alpha * sq - pp.beta * s1 * s1); }
        default: {  // banana: (x1^2 + (x2 - x1^2 + 1)^2 + sum_{i>=3} x_i^2) / 2   (test/test_config.jl:33-36)
            const double x1 = x0[0] * ca + v0[0] * cb, x2 = x0[1] * ca + v0[1] * cb;
            const double r = x2 - x1 * x1 + 1.0;
            return 0.5 * (x1 * x1 + r * r + sq);
        }
    }
}

// Logistic-regression posterior (BASELINE.json config 4): U(theta) = sum_r [log(1 + e^{z_r}) - y_r z_r] + |theta|^2 / (2 s0^2),
// z = X theta.  theta (this boundary's position) is staged in the warp's slice of shared memory; lanes stride the rows
// of X (which stays in L2), log(1 + e^z) in the overflow-free form max(z, 0) + log1p(e^{-|z|}).
__device__ double logreg_value(const PotParams& pp, int d, const double* __restrict__ x0, const double* __restrict__ v0,
                               double ca, double cb, int lane, double* theta) {
    double sq = 0.0;
    for (int j = lane; j < d; j += 32) {
        const double xj = x0[j] * ca + v0[j] * cb;
        theta[j] = xj;
        sq += xj * xj;
    }
    __syncwarp();
    double acc = 0.0;
    for (int64_t r = lane; r < pp.n; r += 32) {
        const double* row = pp.vec + r * d;
        double z = 0.0;
        for (int j = 0; j < d; ++j) z = fma(row[j], theta[j], z);
        acc += fmax(z, 0.0) + log1p(exp(-fabs(z))) - pp.vec2[r] * z;
    }
    __syncwarp();
    return warp_sum(acc) + 0.5 * pp.inv_s2 * warp_sum(sq);
}

__global__ void __launch_bounds__(256) rv_kernel(int kind, PotParams pp, int flow_kind, int d, int64_t ld_sk, int64_t n_sk,
                                                 const int64_t* __restrict__ ncols, int64_t B_req,
                                                 const double* __restrict__ X, const double* __restrict__ V,
                                                 const double* __restrict__ T, double* __restrict__ uval, int64_t ld_u,
                                                 double* __restrict__ rv) {
    const int64_t c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int64_t n = ncols ? ncols[c] : n_sk;
    const double* t = T + c * ld_sk;
    __shared__ double red[8];
    extern __shared__ double theta_smem[];  // LOGREG only: 8 warps x d
    if (n <= 0) { if (threadIdx.x == 0) rv[c] = 0.0; return; }            // diagnostic.jl:40
    const double Tend = t[n - 1];
    if (!(Tend >= 0.0) || Tend == CUDART_INF) { if (threadIdx.x == 0) rv[c] = CUDART_NAN; return; }  // :43-45 (caller raises)
    if (Tend == 0.0) { if (threadIdx.x == 0) rv[c] = 0.0; return; }      // :53
    int64_t B = B_req;
    if (B == 0) { B = (int64_t)floor(sqrt((double)n)); if (B < 1) B = 1; }  // :47-48
    // range(0.0, T; length=B+1): nodes k T / B in twice-precision, the last one exactly T (same scheme as the bounds' grid)
    const double m = (double)B, gc = Tend / m, grem = fma(-gc, m, Tend) / m;
    double* u = uval + c * ld_u;
    for (int64_t b = warp; b <= B; b += nwarp) {
        const double kk = (double)b;
        const double tb = (b >= B) ? Tend : fma(kk, gc, kk * grem);
        int64_t lo = 0;
        double tau = 0.0;
        if (b > 0) {  // largest i with t[i] <= tb (the reference's pointer walk, :62-64); boundary 0 is X[:, 1] itself (:56)
            int64_t hi = n - 1;
            while (lo < hi) {
                const int64_t mid = (lo + hi + 1) >> 1;
                if (t[mid] <= tb) lo = mid; else hi = mid - 1;
            }
            tau = tb - t[lo];
        }
        double ca = 1.0, cb = tau;
        if (flow_kind == 1) sincos(tau, &cb, &ca);
        const double* xs = X + (c * ld_sk + lo) * d;
        const double* vs = V + (c * ld_sk + lo) * d;
        const double val = kind == PDMPFLUX_LOGREG ? logreg_value(pp, d, xs, vs, ca, cb, lane, theta_smem + warp * d)
                                                   : potential_value(kind, pp, d, xs, vs, ca, cb, lane);
        if (lane == 0) u[b] = val;
    }
    __syncthreads();  // makes this CTA's global writes visible to itself
    double acc = 0.0;
    for (int64_t b = 1 + threadIdx.x; b <= B; b += blockDim.x) {
        const double inc = u[b] - u[b - 1];
        acc += inc * inc;
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += red[w];
        rv[c] = s / Tend;
    }
}

}  // namespace

cudaError_t launch_rv_diagnostic(int kind, const PotParams& pp, int flow_kind, int d, int64_t ld_sk, int64_t n_sk,
                                 int64_t n_chains, const int64_t* ncols, int64_t B, const double* X, const double* V,
                                 const double* t, double* uval, int64_t ld_u, double* rv, cudaStream_t stream) {
    const size_t smem = kind == PDMPFLUX_LOGREG ? sizeof(double) * 8 * (size_t)d : 0;
    rv_kernel<<<(unsigned)n_chains, 256, smem, stream>>>(kind, pp, flow_kind, d, ld_sk, n_sk, ncols, B, X, V, t, uval, ld_u, rv);
    return cudaGetLastError();
}

}  // namespace pdmpflux
