// Zig-Zag on a Bayesian logistic-regression posterior (BASELINE.json config 4): one CTA per chain.
//
//   U(theta) = sum_r [log(1 + exp(z_r)) - y_r z_r] + |theta|^2 / (2 sigma0^2),  z = X theta,  X: n x d row-major
//   grad U   = X^T (sigma(z) - y) + theta / sigma0^2
//   H v      = X^T (sigma'(z) .* (X v)) + v / sigma0^2
//
// The potential is not part of the reference (PDMPFlux.jl takes an arbitrary Julia closure); the sampler logic is the
// reference's: vectorised signed/unsigned grid bound upper_bound_grid_vect (UpperBound.jl:203-247) with analytic
// derivatives, next_event (:264-273), the thinning loop (SamplingLoopInplace.jl:27-217), the Zig-Zag flip
// (ZigZagSamplers.jl:101-107) and record! (Composites.jl:239-260).
//
// Affine trick: along the flow line X (x + t v) = z + t w with z = X x, w = X v, so one pass over the
// rows of X yields the gradient and the Hessian-vector product at ALL grid times:
//     [G | HV] (d x 2G)  =  X^T (n x d)^T  .  [sigma(z + t_k w) - y | sigma'(z + t_k w) .* w]_k (n x 2G)
// Per row tile (64 rows, staged in shared memory) the CTA computes z, w (DMMA, N padded to 8), the 2G residual columns
// (exp), and accumulates the d x 2G product with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64 = SASS DMMA; tcgen05 has
// no FP64 kind).  X (n d 8 bytes, 80 MB for C4) stays resident in the 126 MB L2 across chains.
#include "common.cuh"
#include "philox.cuh"

namespace pdmpflux {

constexpr int kLrThreads = 128;
constexpr int kLrRows = 32;    // rows of X per tile (two tiles in flight per CTA, three CTAs per SM at d = 100)
constexpr int kLrMaxG = 12;    // 2G <= 24 residual columns = 3 n-tiles

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- TMA bulk load global -> shared with an mbarrier (SASS: UBLKCP.S.G + SYNCS) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct LrShared {  // offsets (in doubles) into dynamic shared memory
    int bar, x, v, xv, Xt, ybuf, rs, acc, z, w, lam, box, cum, red, ncs, tile, total;
};
__host__ __device__ inline LrShared lr_layout(int d, int G) {
    LrShared L;
    const int dp = (d + 7) / 8 * 8;
    const int dm = dp + 8;  // accumulator rows (m-tiles of 8, +1 spare)
    // residual-column stride: >= 2G and == 4 (mod 16) so the B fragments (4 rows x 8 columns) hit 32 distinct banks
    int ncs = 4;
    while (ncs < 2 * G) ncs += 16;
    L.ncs = ncs;
    L.tile = kLrRows * d + 32;                 // one X tile, rows contiguous (+ slack for fragment over-reads)
    int o = 0;
    L.bar = o; o += 2;                          // two mbarriers (8 bytes each)
    L.Xt = o; o += 2 * L.tile;                  // double-buffered X tiles (16-byte aligned: tile is even)
    L.ybuf = o; o += 2 * kLrRows;
    L.x = o; o += dp;
    L.v = o; o += dp;
    L.xv = o; o += (dp + 4) * 8;                // [k][8]: (x_k, v_k, 0...) B operand of the z/w product
    L.rs = o; o += kLrRows * ncs + 32;          // residual columns
    L.acc = o; o += dm * 24;                    // X^T R accumulators [i][24]
    L.z = o; o += kLrRows;
    L.w = o; o += kLrRows;
    L.lam = o; o += dp;
    L.box = o; o += kLrMaxG + 4;
    L.cum = o; o += kLrMaxG + 4;
    L.red = o; o += 64;
    L.total = o;
    return L;
}
size_t logreg_smem_bytes(int d, int G) { return sizeof(double) * (size_t)lr_layout(d, G).total; }

// One pass over all rows of X: acc[i][c] = sum_r X[r][i] * R[r][c] for nt times tt[0..nt) (c = k: sigma - y,
// c = nt + k: sigma' * w when want_h).  All threads of the CTA participate.  X tiles and y arrive by TMA bulk copies
// into a two-deep ring; `phase` carries the mbarrier parities across calls.
__device__ void lr_pass(const KernelParams& p, double* sm, const LrShared& L, const double* tt, int nt, bool want_h,
                        uint32_t (&phase)[2]) {
    const int d = p.d, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t n = p.pot.n;
    const int ncs = L.ncs;
    const int ncols = want_h ? 2 * nt : nt;
    const int n_nt = (ncols + 7) / 8;                // n-tiles of the main product
    const int n_mt = (d + 7) / 8;                    // m-tiles (coordinates)
    const int kz = (d + 3) / 4;                      // k-steps of the z/w product
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
    double acc[4][3][2];                             // up to 4 m-tiles per warp x 3 n-tiles
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int64_t ntiles = (n + kLrRows - 1) / kLrRows;
    auto issue = [&](int64_t tile) {  // thread 0: TMA the tile's rows of X and y into ring slot tile & 1
        const int slot = (int)(tile & 1);
        const int64_t r0 = tile * kLrRows;
        const int rows = (int)min((int64_t)kLrRows, n - r0);
        const uint32_t xb = ((uint32_t)rows * d * 8u) & ~15u, yb = ((uint32_t)rows * 8u) & ~15u;
        mbar_expect_tx(&bar[slot], xb + yb);
        if (xb) bulk_load(sm + L.Xt + slot * L.tile, p.pot.vec + r0 * d, xb, &bar[slot]);
        if (yb) bulk_load(sm + L.ybuf + slot * kLrRows, p.pot.vec2 + r0, yb, &bar[slot]);
    };
    __syncthreads();  // the ring is free (previous pass fully consumed)
    if (tid == 0) issue(0);
    for (int64_t tile = 0; tile < ntiles; ++tile) {
        const int slot = (int)(tile & 1);
        const int64_t r0 = tile * kLrRows;
        const int rows = (int)min((int64_t)kLrRows, n - r0);
        if (tid == 0 && tile + 1 < ntiles) issue(tile + 1);  // slot (tile+1)&1 was released by the barrier below
        mbar_wait(&bar[slot], phase[slot]);
        phase[slot] ^= 1u;
        double* Xs = sm + L.Xt + slot * L.tile;
        double* ys = sm + L.ybuf + slot * kLrRows;
        if (tid == 0) {  // odd tails that a 16-byte granular bulk copy cannot carry
            if ((rows * d) & 1) Xs[rows * d - 1] = __ldg(p.pot.vec + r0 * d + rows * d - 1);
            if (rows & 1) ys[rows - 1] = __ldg(p.pot.vec2 + r0 + rows - 1);
        }
        if ((rows & 1) || ((rows * d) & 1)) __syncthreads();
        // ---- z = X x, w = X v for the tile: DMMA with B = [x v 0 ...] (k x 8); one m-tile (8 rows) per warp ----
        {
            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;  // two accumulator pairs: halves the dependent MMA chain
            const double* arow = Xs + (warp * 8 + gid) * d + tig;
            const double* bcol = sm + L.xv + tig * 8 + gid;
            int ks = 0;
            for (; ks + 1 < kz; ks += 2) {
                dmma(c0, c1, arow[4 * ks], bcol[32 * ks]);
                dmma(e0, e1, arow[4 * ks + 4], bcol[32 * ks + 32]);
            }
            if (ks < kz) dmma(c0, c1, arow[4 * ks], bcol[32 * ks]);
            if (tig == 0) { sm[L.z + warp * 8 + gid] = c0 + e0; sm[L.w + warp * 8 + gid] = c1 + e1; }
        }
        __syncthreads();
        // ---- residual columns: thread -> (row, quarter of the times) ----
        {
            const int row = tid & (kLrRows - 1), part = tid / kLrRows;  // 4 parts
            const double z = sm[L.z + row], w = sm[L.w + row], yy = ys[row];
            const bool live = row < rows;
            for (int k = part; k < nt; k += kLrThreads / kLrRows) {
                const double eta = z + tt[k] * w;
                const double sg = 1.0 / (1.0 + exp(-eta));
                sm[L.rs + row * ncs + k] = live ? sg - yy : 0.0;
                if (want_h) sm[L.rs + row * ncs + nt + k] = live ? sg * (1.0 - sg) * w : 0.0;
            }
        }
        __syncthreads();
        // ---- acc += Xtile^T . R : A[m][k] = Xt[row0 + k][i0 + m], B[k][n] = R[row0 + k][n0 + n] ----
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int mt = warp + 4 * a;
            if (mt < n_mt) {
                const double* ap = Xs + tig * d + mt * 8 + gid;
                const double* bp = sm + L.rs + tig * ncs + gid;
#pragma unroll
                for (int ks = 0; ks < kLrRows / 4; ++ks) {
                    const double av = ap[4 * ks * d];
#pragma unroll
                    for (int b = 0; b < 3; ++b)
                        if (b < n_nt) dmma(acc[a][b][0], acc[a][b][1], av, bp[4 * ks * ncs + 8 * b]);
                }
            }
        }
        __syncthreads();  // tile consumed: its ring slot may be refilled, z / w / rs may be overwritten
    }
    // ---- accumulator fragments -> shared memory acc[i][c] ----
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int mt = warp + 4 * a;
        if (mt < n_mt) {
#pragma unroll
            for (int b = 0; b < 3; ++b)
                if (b < n_nt) {
                    const int i = mt * 8 + gid, c = 8 * b + 2 * tig;
                    sm[L.acc + i * 24 + c] = acc[a][b][0];
                    sm[L.acc + i * 24 + c + 1] = acc[a][b][1];
                }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double lr_block_sum(double v, double* red) {  // all threads get the sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kLrThreads / 32; ++k) s += red[k];
    return s;
}

__global__ void __launch_bounds__(kLrThreads) logreg_zigzag_kernel(const __grid_constant__ KernelParams p) {
    extern __shared__ __align__(128) double sm[];
    const int d = p.d, G = p.G, tid = threadIdx.x;
    const int64_t c = blockIdx.x;
    const LrShared L = lr_layout(d, G);
    const double inv_s2 = p.pot.inv_s2;

    // ---- PDMPState ----
    for (int i = tid; i < (d + 7) / 8 * 8; i += kLrThreads) {
        sm[L.x + i] = i < d ? p.sx[c * d + i] : 0.0;
        sm[L.v + i] = i < d ? p.sv[c * d + i] : 0.0;
    }
    for (int e = tid; e < 2 * L.tile; e += kLrThreads) sm[L.Xt + e] = 0.0;            // incl. the slack fragment over-reads touch
    for (int e = tid; e < kLrRows * L.ncs + 32; e += kLrThreads) sm[L.rs + e] = 0.0;
    fence_async_smem();  // order these generic-proxy writes before the TMA (async-proxy) writes into the same buffers
    uint32_t phase[2] = {0u, 0u};
    if (tid == 0) {
        uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    double t = p.st[c], horizon = p.shorizon[c], ar = p.sar[c];
    int status = p.status[c];
    int64_t n_builds = p.counters[2 * c], n_rates = p.counters[2 * c + 1];
    DrawKey key;
    const uint64_t gchain = (uint64_t)(p.chain_offset + c);
    key.k0 = (uint32_t)p.seed; key.k1 = (uint32_t)(p.seed >> 32);
    key.chain_lo = (uint32_t)gchain; key.chain_hi8 = (uint32_t)(gchain >> 32) << 8;
    uint32_t sE = 0, sU = 0;
    int64_t pE = p.tape_pos[3 * c], pU = p.tape_pos[3 * c + 1];
    const double* tE = p.tE + c * p.nE;
    const double* tU = p.tU + c * p.nU;
    bool exhausted = false;
    auto rand_exp = [&]() -> double {
        if (p.draw_mode) return draw_exp(key, sE++);
        if (pE >= p.nE) { exhausted = true; return 1.0; }
        return __ldg(tE + pE++);
    };
    auto rand_uniform = [&]() -> double {
        if (p.draw_mode) return draw_uniform(key, sU++);
        if (pU >= p.nU) { exhausted = true; return 0.5; }
        return __ldg(tU + pU++);
    };
    __syncthreads();

    auto record = [&](int64_t col, int eb, int rej, int hh, const double* eva) {  // Composites.jl:239-260
        const int64_t o = c * p.ld_cols + col;
        const int64_t orow = c * p.ld_rows + (col - p.col0) + p.col0_rows;
        for (int i = tid; i < d; i += kLrThreads) {
            if (p.X) p.X[orow * d + i] = sm[L.x + i];
            if (p.V) p.V[orow * d + i] = sm[L.v + i];
        }
        if (tid == 0) {
            if (p.T) p.T[o] = t;
            if (p.H) p.H[o] = horizon;
            if (p.AR) p.AR[o] = ar;
            if (!p.sparse_cols || eb != 0) {
                if (p.EB) p.EB[o] = eb;
                if (p.EVA) for (int k = 0; k < 5; ++k) p.EVA[o * 5 + k] = eva[k];
            }
            if ((!p.sparse_cols || rej != 0) && p.REJ) p.REJ[o] = rej;
            if ((!p.sparse_cols || hh != 0) && p.HH) p.HH[o] = hh;
        }
    };

    if (p.n_events == 0) {
        const double zero5[5] = {0, 0, 0, 0, 0};
        record(p.col0, 0, 0, 0, zero5);
        if (tid == 0) p.ncols[c] += 1;
        return;
    }

    // grid nodes, UpperBound.jl:204 (same construction as chain.cuh:make_grid)
    double gc = 0, grem = 0, gh = 0, step = 0;
    auto grid_t = [&](int k) -> double {
        if (k >= G - 1) return gh;
        const double kk = (double)k;
        return fma(kk, gc, kk * grem);
    };
    // upper_bound_grid_vect, UpperBound.jl:203-247 (analytic derivative): fills box / cum in shared memory
    auto build_bound = [&](double h) {
        ++n_builds;
        const double m = (double)(G - 1);
        gc = h / m; grem = fma(-gc, m, h) * p.inv_gm1; gh = h;
        step = grid_t(1);
        // B operand of the z/w product: [k][8] = (x_k, v_k, 0, ...)
        for (int e = tid; e < ((d + 7) / 8 * 8 + 4) * 8; e += kLrThreads) {
            const int k = e >> 3, col = e & 7;
            sm[L.xv + e] = (k < d && col == 0) ? sm[L.x + k] : ((k < d && col == 1) ? sm[L.v + k] : 0.0);
        }
        double tt[kLrMaxG];
#pragma unroll
        for (int k = 0; k < kLrMaxG; ++k) tt[k] = grid_t(min(k, G - 1));
        lr_pass(p, sm, L, tt, G, true, phase);
        // per-coordinate cells (thread i = coordinate i), QUIRK-preserving tangent formula (UpperBound.jl:229-241)
        double bpart[kLrMaxG];
#pragma unroll
        for (int k = 0; k < kLrMaxG; ++k) bpart[k] = 0.0;
        for (int i = tid; i < d; i += kLrThreads) {
            const double xi = sm[L.x + i], vi = sm[L.v + i];
            double vl = 0, gl = 0;
#pragma unroll
            for (int k = 0; k < kLrMaxG; ++k)
                if (k < G) {
                    const double g = sm[L.acc + i * 24 + k] + (xi + tt[k] * vi) * inv_s2;
                    const double hv = sm[L.acc + i * 24 + G + k] + vi * inv_s2;
                    double val = g * vi, dval = hv * vi;
                    if (!p.signed_bound) { dval = (0.0 > val) ? 0.0 : dval; val = (val > 0.0 ? val : 0.0); }
                    if (k > 0) {
                        double pos = (vl - val + dval * tt[k] - gl * tt[k - 1]) / (dval - gl);
                        if (pos != pos) pos = 0.0;
                        pos = fmin(fmax(pos, 0.0), step);
                        const double inter = vl + gl * pos;
                        bpart[k - 1] += fmax(fmax(fmax(vl, val), inter), 0.0);
                    }
                    vl = val; gl = dval;
                }
        }
        double cs = 0.0;
        if (tid == 0) sm[L.cum] = 0.0;
        for (int k = 0; k < G - 1; ++k) {
            const double b = lr_block_sum(bpart[k], sm + L.red);
            cs += b;
            if (tid == 0) { sm[L.box + k] = b; sm[L.cum + k + 1] = cs * step; }
        }
        __syncthreads();
    };
    auto next_event = [&](double e, double& tp_out, double& lb_out) {  // UpperBound.jl:264-273
        int idx = 0;
        while (idx < G && sm[L.cum + idx] < e) ++idx;
        if (idx >= G) { tp_out = CUDART_INF; lb_out = sm[L.box + G - 2]; return; }
        if (idx == 0) { tp_out = CUDART_NAN; lb_out = sm[L.box]; return; }
        tp_out = grid_t(idx - 1) + (e - sm[L.cum + idx - 1]) / (sm[L.cum + idx] - sm[L.cum + idx - 1]) * step;
        lb_out = sm[L.box + idx - 1];
    };
    // sampler.rate at tp (ZigZagSamplers.jl:83-86); leaves lambda_i = max(0, g_i v_i) in shared memory for the jump.
    // Uses the xv operand staged by the last build_bound (x, v unchanged since).
    auto rate_at = [&](double tp) -> double {
        double tt1[1] = {tp};
        lr_pass(p, sm, L, tt1, 1, false, phase);
        double part = 0.0;
        for (int i = tid; i < d; i += kLrThreads) {
            const double vi = sm[L.v + i];
            const double g = sm[L.acc + i * 24] + (sm[L.x + i] + tp * vi) * inv_s2;
            const double y = g * vi;
            const double lam = (y > 0.0 ? y : 0.0);
            sm[L.lam + i] = lam;
            part += lam;
        }
        return lr_block_sum(part, sm + L.red);
    };
    auto flow = [&](double tt) {  // ZigZagSamplers.jl:80
        for (int i = tid; i < d; i += kLrThreads) sm[L.x + i] += sm[L.v + i] * tt;
        __syncthreads();
    };

    int64_t n_rec = 0;
    if (p.use_t_stop && status == 0 && !(t < p.t_stop)) status = PDMPFLUX_CHAIN_DONE;
    if (status == 0) {
        for (int64_t ev = 0; ev < p.n_events; ++ev) {
            key.event = (uint32_t)(p.event0 + ev + 1);
            sE = sU = 0;
            int eb = 0, rej = 0, hh = 0, steps = 0;
            double eva[5] = {0, 0, 0, 0, 0};
            double ts = 0.0, tp = 0.0, lambda_bar = 0.0, exp_rv = 0.0;
            bool accept = false;
            while (!accept && status == 0) {  // get_event_state!, SamplingLoopInplace.jl:27-39
                if (++steps > p.max_steps) { status = PDMPFLUX_CHAIN_STEP_LIMIT; break; }
                build_bound(horizon);        // one_step_of_thinning!, :65-85
                double e = rand_exp();
                next_event(e, tp, lambda_bar);
                exp_rv = e;
                if (exhausted) { status = PDMPFLUX_CHAIN_TAPE_EXHAUSTED; break; }
                if (tp > horizon) {          // move_to_horizon!, :87-101
                    flow(horizon);
                    ts += horizon; hh += 1;
                    horizon = p.adaptive ? horizon * 1.01 : horizon;
                    continue;
                }
                while (tp < horizon && !accept && status == 0) {  // moves_until_horizon!, :103-111
                    ++n_rates;
                    const double lt = rate_at(tp);   // ac_step!, :113-129
                    ar = lt / lambda_bar;
                    if (ar > 1.0) {                  // erroneous_acceptance_rate!, :131-151
                        const double h2 = horizon / 2;
                        build_bound(h2);
                        e = rand_exp();
                        next_event(e, tp, lambda_bar);
                        exp_rv = e;
                        horizon = p.adaptive ? h2 : horizon;
                        eb += 1;
                        eva[eb % 5] = ar;
                    } else if (rand_uniform() < ar) {  // if_accept!, :170-186
                        if (p.use_t_stop && t + tp + ts > p.t_stop) {  // time-horizon variant, src/sample.jl:385-420
                            flow(p.t_stop - (t + ts));
                            t = p.t_stop;
                            ar = 0.0; eb = 0; rej = 0; hh = 0;
                            for (int k = 0; k < 5; ++k) eva[k] = 0.0;
                            status = PDMPFLUX_CHAIN_DONE;
                            accept = true;
                            break;
                        }
                        const double uS = rand_uniform() * lt;  // categorical draw against the rates just computed
                        flow(tp);
                        if (tid == 0) {                          // first index with cumulative lambda > u S
                            double cp = 0.0;
                            int m = d - 1;
                            for (int i = 0; i < d; ++i) { cp += sm[L.lam + i]; if (cp > uS) { m = i; break; } }
                            sm[L.v + m] = -sm[L.v + m];
                        }
                        __syncthreads();
                        t = t + tp + ts;
                        ts = 0.0; tp = 0.0;
                        accept = true;
                    } else {                           // if_reject!, :188-203
                        const double e3 = exp_rv + rand_exp();
                        next_event(e3, tp, lambda_bar);
                        horizon = p.adaptive ? horizon / 1.04 : horizon;
                        exp_rv = e3;
                        rej += 1;
                        if (tp > horizon) { flow(horizon); ts += horizon; hh += 1; }  // move_to_horizon2!, :205-217
                    }
                    if (exhausted) status = PDMPFLUX_CHAIN_TAPE_EXHAUSTED;
                }
            }
            if (status != 0 && status != PDMPFLUX_CHAIN_DONE) break;
            record(p.col0 + ev, eb, rej, hh, eva);
            ++n_rec;
            if (status == PDMPFLUX_CHAIN_DONE) break;
            if (p.use_t_stop && !(t < p.t_stop)) { status = PDMPFLUX_CHAIN_DONE; break; }
        }
    }
    __syncthreads();
    for (int i = tid; i < d; i += kLrThreads) {
        p.sx[c * d + i] = sm[L.x + i];
        p.sv[c * d + i] = sm[L.v + i];
    }
    if (tid == 0) {
        p.st[c] = t; p.shorizon[c] = horizon; p.sar[c] = ar; p.status[c] = status;
        p.counters[2 * c] = n_builds; p.counters[2 * c + 1] = n_rates;
        p.tape_pos[3 * c] = pE; p.tape_pos[3 * c + 1] = pU;
        p.ncols[c] += n_rec;
    }
}

cudaError_t launch_logreg_zigzag(const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(logreg_zigzag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    logreg_zigzag_kernel<<<grid, kLrThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace pdmpflux
