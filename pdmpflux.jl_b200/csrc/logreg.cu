// Zig-Zag on a Bayesian logistic-regression posterior (BASELINE.json config 4): four chains per CTA.
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
// rows of X yields the gradient and the Hessian-vector product at ALL grid times of a bound:
//     [G | HV] (d x 2G)  =  X^T  .  [sigma(z + t_k w) - y | sigma'(z + t_k w) .* w]_k (n x 2G)
// and the rate at a proposal time is the same pass with one column.
//
// Execution: a CTA of 4 warps owns 4 chains.  Between passes warp c runs chain c's thinning state machine; whenever
// the chains need gradients they post a request (a bound: G times with Hessian columns, or a rate: 1 time) and the
// whole CTA makes ONE pass over X serving all four:
//   * 32-row tiles of X (rows are contiguous in row-major X) and of y arrive by TMA bulk copies
//     (cp.async.bulk + mbarrier expect-tx, SASS UBLKCP.S.G) into a two-deep ring;
//   * z, w of all four chains: one DMMA product with B = [x1 v1 x2 v2 x3 v3 x4 v4] -- exactly the 8 columns of
//     mma.sync.m8n8k4.f64, no padding;
//   * residual columns (exp) for every requested time of every chain, packed side by side (up to 4 * 2G = 80);
//   * acc (d x columns) += Xtile^T . R with FP64 tensor-core MMAs (SASS DMMA; tcgen05 has no FP64 kind).
// Four chains share every byte of X read from L2 and fill the MMA tile widths that a single chain would pad
// (z/w: 2 of 8 columns, a rate request: 1 of 8).  X (80 MB at C4 size) stays resident in the 126 MB L2.
#include "common.cuh"
#include "philox.cuh"

namespace pdmpflux {

constexpr int kLrThreads = 128;
constexpr int kLrChains = 4;   // chains per CTA = warps per CTA
constexpr int kLrRows = 32;    // rows of X per tile
constexpr int kLrMaxG = 12;    // grid_size limit: 2G <= 24 columns per chain
constexpr int kLrMaxNt = 8;    // n-tiles (of 8 residual columns) one pass can carry: requests beyond wait a round
constexpr int kLrMaxCols = 8 * kLrMaxNt;
constexpr int kLrNcs = 68;     // residual-column stride in shared memory (== 4 mod 16: conflict-free B fragments)

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
    int bar, Xt, rs, ybuf, xv, zw, lam, box, cum, sched, tile, acc_stride, total;
};
// request schedule of one pass (shared memory, written by the chain warps, read by everybody)
struct LrSched {
    int n_cols;                    // packed residual columns
    int col0[kLrChains];           // first column of chain c
    int want[kLrChains];           // times chain c asks for (0: none, 1: rate, G: bound)
    int nt[kLrChains];             // times granted to chain c this pass (0 when it has to wait for the next one)
    double tp[kLrChains];          // rate request: the proposal time
    double gc[kLrChains], grem[kLrChains], gh[kLrChains];  // bound request: its time grid (see grid_t)
};
__host__ __device__ inline LrShared lr_layout(int d) {
    LrShared L;
    const int dp = (d + 7) / 8 * 8;
    L.tile = kLrRows * d + 32;                  // one X tile, rows contiguous (+ slack for fragment over-reads)
    L.acc_stride = kLrMaxCols;
    int o = 0;
    L.bar = o; o += 2;                          // two mbarriers (8 bytes each)
    L.Xt = o; o += 2 * L.tile;                  // double-buffered X tiles (16-byte aligned: tile is even)
    L.rs = o; o += kLrRows * kLrNcs + 32;       // residual columns
    // after a pass the accumulators (dp+8 rows x kLrMaxCols) overlay [Xt ring | rs]; make sure they fit
    const int need = (dp + 8) * kLrMaxCols;
    if (o - L.Xt < need) o = L.Xt + need;
    L.ybuf = o; o += 2 * kLrRows;
    L.xv = o; o += (dp + 4) * 8;                // [k][8]: columns (2c, 2c+1) = (x, v) of chain c -- state AND B operand
    L.zw = o; o += kLrRows * 8;                 // [row][8]: (z, w) of chain c in columns (2c, 2c+1)
    L.lam = o; o += kLrChains * dp;
    L.box = o; o += kLrChains * 16;
    L.cum = o; o += kLrChains * 16;
    L.sched = o; o += (int)((sizeof(LrSched) + 7) / 8);
    L.total = o;
    return L;
}
size_t logreg_smem_bytes(int d, int G) { (void)G; return sizeof(double) * (size_t)lr_layout(d).total; }

// One pass over all rows of X serving every posted request:
//   acc[i][col] = sum_r X[r][i] * R[r][col]   (stored to shared memory, row stride L.acc_stride, overlaying the tile ring)
template <int NT>
__device__ __noinline__ void lr_pass(const KernelParams& p, double* sm, const LrShared& L, uint32_t (&phase)[2]) {
    const int d = p.d, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t n = p.pot.n;
    const LrSched* S = reinterpret_cast<const LrSched*>(sm + L.sched);
    const int n_mt = (d + 7) / 8;                    // m-tiles (coordinates), <= 16
    const int kz = (d + 3) / 4;                      // k-steps of the z/w product
    // warp -> m-tiles {mrot, mrot + 4, ..}; rotated by CTA parity so that the warp carrying the odd extra tile sits on
    // a different SM sub-partition in the two co-resident CTAs
    const int mrot = (warp + 2 * (int)(blockIdx.x & 1)) & 3;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
    double acc[4][NT][2];                            // up to 4 m-tiles per warp x NT n-tiles
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

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
    // The accumulators of the previous pass overlaid the ring: every warp is done with them (caller's barrier), but the
    // overlay was written through the generic proxy, so order it before the TMA (async proxy) refills the ring.
    fence_async_smem();
    __syncthreads();
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
        // ---- (z, w) of the four chains for the tile: DMMA with B = [x1 v1 .. x4 v4] (k x 8); 8 rows per warp ----
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
            double* zr = sm + L.zw + (warp * 8 + gid) * 8 + 2 * tig;  // C[row][2 tig], C[row][2 tig + 1] = (z, w) of chain tig
            zr[0] = c0 + e0;
            zr[1] = c1 + e1;
        }
        __syncthreads();
        // ---- residual columns: warp c evaluates chain c's requested times, lane = row of the tile ----
        {
            const int nt = S->nt[warp];
            if (nt) {
                const bool live = lane < rows;
                const double z = sm[L.zw + lane * 8 + 2 * warp], w = sm[L.zw + lane * 8 + 2 * warp + 1], yy = ys[lane];
                double* rp = sm + L.rs + lane * kLrNcs + S->col0[warp];
                if (nt == 1) {  // rate at the proposal time
                    const double sg = 1.0 / (1.0 + exp(-(z + S->tp[warp] * w)));
                    rp[0] = live ? sg - yy : 0.0;
                } else {        // bound: residual and Hessian-weight columns at the G grid times
                    const double gc = S->gc[warp], grem = S->grem[warp], gh = S->gh[warp];
                    if (fabs(z) + fabs(gh * w) < 600.0) {
                        // eta_k = z + t_k w on a uniform grid: exp(-eta_k) is a geometric progression -- two exps
                        // per row instead of G (the accumulated rounding, <= G ulp, is far inside the parity tolerance)
                        double q = exp(-z);
                        const double rho = exp(-fma(1.0, gc, grem) * w);
                        for (int k = 0; k < nt; ++k) {
                            const double sg = 1.0 / (1.0 + q);
                            rp[k] = live ? sg - yy : 0.0;
                            rp[nt + k] = live ? q * sg * sg * w : 0.0;  // sigma (1 - sigma) w with 1 - sigma = q sigma
                            q *= rho;
                        }
                    } else {  // extreme logits: evaluate every node on its own
                        for (int k = 0; k < nt; ++k) {
                            const double kk = (double)k, tk = (k >= nt - 1) ? gh : fma(kk, gc, kk * grem);
                            const double sg = 1.0 / (1.0 + exp(-(z + tk * w)));
                            rp[k] = live ? sg - yy : 0.0;
                            rp[nt + k] = live ? sg * (1.0 - sg) * w : 0.0;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- acc += Xtile^T . R : A[m][k] = Xt[row0 + k][i0 + m], B[k][n] = R[row0 + k][n0 + n] ----
        {
            const double* ap = Xs + tig * d + mrot * 8 + gid;  // m-tiles mrot, mrot + 4, ... (the last may not exist)
            const double* bp = sm + L.rs + tig * kLrNcs + gid;
            const int n_a = (n_mt - mrot + 3) >> 2;  // m-tiles of this warp
#pragma unroll
            for (int ks = 0; ks < kLrRows / 4; ++ks) {
                double bv[NT];
#pragma unroll
                for (int b = 0; b < NT; ++b) bv[b] = bp[4 * ks * kLrNcs + 8 * b];
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (a < n_a) {
                        const double av = ap[4 * ks * d + 32 * a];
#pragma unroll
                        for (int b = 0; b < NT; ++b) dmma(acc[a][b][0], acc[a][b][1], av, bv[b]);
                    }
            }
        }
        __syncthreads();  // tile consumed: its ring slot may be refilled, zw / rs may be overwritten
    }
    // ---- accumulator fragments -> shared memory acc[i][col], overlaying the (now idle) tile ring ----
    double* A = sm + L.Xt;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int mt = mrot + 4 * a;
        if (mt < n_mt) {
#pragma unroll
            for (int b = 0; b < NT; ++b) {
                const int i = mt * 8 + gid, c = 8 * b + 2 * tig;
                A[i * L.acc_stride + c] = acc[a][b][0];
                A[i * L.acc_stride + c + 1] = acc[a][b][1];
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kLrThreads) logreg_zigzag_kernel(const __grid_constant__ KernelParams p) {
    extern __shared__ __align__(128) double sm[];
    const int d = p.d, G = p.G, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const LrShared L = lr_layout(d);
    const int dp = (d + 7) / 8 * 8;
    const double inv_s2 = p.pot.inv_s2;
    LrSched* S = reinterpret_cast<LrSched*>(sm + L.sched);
    const int64_t c_raw = (int64_t)blockIdx.x * kLrChains + w;  // warp w runs chain c
    const bool valid = c_raw < p.n_chains;
    const int64_t c = valid ? c_raw : 0;

    // ---- shared state: xv[k][8] holds (x, v) of chain w in columns (2w, 2w+1); it is also the z/w B operand ----
    for (int e = tid; e < (dp + 4) * 8; e += kLrThreads) sm[L.xv + e] = 0.0;
    for (int e = tid; e < 2 * L.tile + kLrRows * kLrNcs + 32; e += kLrThreads) sm[L.Xt + e] = 0.0;  // ring, slack, rs
    __syncthreads();
    if (valid)
        for (int i = lane; i < d; i += 32) {
            sm[L.xv + i * 8 + 2 * w] = p.sx[c * d + i];
            sm[L.xv + i * 8 + 2 * w + 1] = p.sv[c * d + i];
        }
    fence_async_smem();  // order these generic-proxy writes before the TMA (async-proxy) writes into the same buffers
    uint32_t phase[2] = {0u, 0u};
    if (tid == 0) {
        uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    // ---- PDMPState scalars of chain w (warp-uniform registers) ----
    double t = p.st[c], horizon = p.shorizon[c], ar = p.sar[c];
    int status = valid ? p.status[c] : 0;
    int64_t n_builds = p.counters[2 * c], n_rates = p.counters[2 * c + 1];
    DrawKey key;
    const uint64_t gchain = (uint64_t)(p.chain_offset + c);
    key.k0 = (uint32_t)p.seed; key.k1 = (uint32_t)(p.seed >> 32);
    key.chain_lo = (uint32_t)gchain; key.chain_hi8 = (uint32_t)(gchain >> 32) << 8;
    key.event = 0;
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
    auto xr = [&](int i) -> double& { return sm[L.xv + i * 8 + 2 * w]; };
    auto vr = [&](int i) -> double& { return sm[L.xv + i * 8 + 2 * w + 1]; };
    double* lam = sm + L.lam + w * dp;
    double* box = sm + L.box + w * 16;
    double* cum = sm + L.cum + w * 16;

    int eb = 0, rej = 0, hh = 0;
    double eva[5] = {0, 0, 0, 0, 0};
    auto record = [&](int64_t col) {  // Composites.jl:239-260 (one warp writes its chain's row)
        const int64_t o = c * p.ld_cols + col;
        const int64_t orow = c * p.ld_rows + (col - p.col0) + p.col0_rows;
        for (int i = lane; i < d; i += 32) {
            if (p.X) p.X[orow * d + i] = xr(i);
            if (p.V) p.V[orow * d + i] = vr(i);
        }
        if (lane == 0) {
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
    __syncthreads();

    if (p.n_events == 0) {
        if (valid) {
            record(p.col0);
            if (lane == 0) p.ncols[c] += 1;
        }
        return;
    }

    // grid nodes of the current bound, UpperBound.jl:204 (same construction as chain.cuh:make_grid)
    double gc = 0, grem = 0, gh = 0, step = 0;
    auto grid_t = [&](int k) -> double {
        if (k >= G - 1) return gh;
        const double kk = (double)k;
        return fma(kk, gc, kk * grem);
    };
    auto next_event = [&](double e, double& tp_out, double& lb_out) {  // UpperBound.jl:264-273
        int idx = 0;
        for (int k = 0; k < G; ++k) idx += (cum[k] < e) ? 1 : 0;  // searchsortedfirst on a non-decreasing vector
        if (idx >= G) { tp_out = CUDART_INF; lb_out = box[G - 2]; return; }
        if (idx == 0) { tp_out = CUDART_NAN; lb_out = box[0]; return; }
        tp_out = grid_t(idx - 1) + (e - cum[idx - 1]) / (cum[idx] - cum[idx - 1]) * step;
        lb_out = box[idx - 1];
    };
    auto flow = [&](double tt) {  // ZigZagSamplers.jl:80
        for (int i = lane; i < d; i += 32) xr(i) += vr(i) * tt;
        __syncwarp();
    };

    // ---- flattened thinning state machine of chain w (same structure as chain.cuh:run_events) ----
    int64_t ev = 0;
    int steps = 0;
    double ts = 0.0, tp = 0.0, lambda_bar = 0.0, exp_rv = 0.0, hbound = 0.0;
    bool need_build = true, half = false;
    if (p.use_t_stop && status == 0 && !(t < p.t_stop)) status = PDMPFLUX_CHAIN_DONE;
    bool live = valid && status == 0;
    key.event = (uint32_t)(p.event0 + 1);
    enum { REQ_NONE = 0, REQ_BOUND = 1, REQ_RATE = 2 };
    unsigned round = 0;

    while (true) {
        // ---- 1. every chain posts its request ----
        int req = REQ_NONE;
        if (live) {
            if (need_build) {
                if (!half && steps >= p.max_steps) { status = PDMPFLUX_CHAIN_STEP_LIMIT; live = false; }
                else {
                    hbound = half ? horizon / 2 : horizon;  // erroneous_acceptance_rate! rebuilds over half the horizon
                    const double m = (double)(G - 1);
                    gc = hbound / m; grem = fma(-gc, m, hbound) * p.inv_gm1; gh = hbound;
                    step = grid_t(1);
                    req = REQ_BOUND;
                }
            } else if (!(tp < horizon)) need_build = true;  // moves_until_horizon! falls through: fresh outer step next round
            else req = REQ_RATE;
        }
        if (lane == 0) S->want[w] = req == REQ_BOUND ? G : (req == REQ_RATE ? 1 : 0);
        __syncthreads();
        if (tid == 0) {  // pack the columns: a rate request takes 1 column, a bound G residual + G Hessian columns.
            // A pass carries kLrMaxCols columns; a request that does not fit waits for the next pass (the chain's state
            // is untouched, it simply asks again).  The chain served first rotates so nobody starves.
            int col = 0;
            for (int j = 0; j < kLrChains; ++j) {
                const int cc = (j + (int)(round & (kLrChains - 1))) & (kLrChains - 1);
                const int wn = S->want[cc], wc = (wn == 1) ? 1 : 2 * wn;
                const bool fits = col + wc <= kLrMaxCols;
                S->col0[cc] = col;
                S->nt[cc] = fits ? wn : 0;
                col += fits ? wc : 0;
            }
            S->n_cols = col;
        }
        __syncthreads();
        if (S->nt[w] == 0) req = REQ_NONE;  // nothing asked, or deferred to the next pass
        if (req == REQ_BOUND) { ++n_builds; if (!half) ++steps; }
        if (req == REQ_RATE) ++n_rates;
        if (lane == 0 && req != REQ_NONE) {  // the times this chain wants evaluated
            S->tp[w] = tp;
            S->gc[w] = gc; S->grem[w] = grem; S->gh[w] = gh;
        }
        // the CTA goes on while any chain has work left (a request, or a pending fall-through transition)
        const int any = __syncthreads_or(live ? 1 : 0);
        if (!any) break;
        ++round;
        if (S->n_cols == 0) continue;  // only fall-through transitions this round

        // ---- 2. one pass over X for all requests (specialised on the number of 8-column tiles) ----
        switch ((S->n_cols + 7) >> 3) {
            case 1: lr_pass<1>(p, sm, L, phase); break;
            case 2: lr_pass<2>(p, sm, L, phase); break;
            case 3: lr_pass<3>(p, sm, L, phase); break;
            case 4: lr_pass<4>(p, sm, L, phase); break;
            case 5: lr_pass<5>(p, sm, L, phase); break;
            case 6: lr_pass<6>(p, sm, L, phase); break;
            case 7: lr_pass<7>(p, sm, L, phase); break;
            default: lr_pass<8>(p, sm, L, phase); break;
        }
        const double* A = sm + L.Xt;
        const int col0 = S->col0[w];

        // ---- 3. every chain consumes its columns ----
        if (req == REQ_BOUND) {
            // upper_bound_grid_vect, UpperBound.jl:203-247 (analytic derivative), cells of this lane's coordinates
            double bpart[kLrMaxG];
#pragma unroll
            for (int k = 0; k < kLrMaxG; ++k) bpart[k] = 0.0;
            for (int i = lane; i < d; i += 32) {
                const double xi = xr(i), vi = vr(i);
                double vl = 0, gl = 0, tl_ = 0;
#pragma unroll
                for (int k = 0; k < kLrMaxG; ++k)
                    if (k < G) {
                        const double tk = grid_t(k);
                        const double g = A[i * L.acc_stride + col0 + k] + (xi + tk * vi) * inv_s2;
                        const double hv = A[i * L.acc_stride + col0 + G + k] + vi * inv_s2;
                        double val = g * vi, dval = hv * vi;
                        if (!p.signed_bound) { dval = (0.0 > val) ? 0.0 : dval; val = (val > 0.0 ? val : 0.0); }
                        if (k > 0) {  // QUIRK-preserving tangent formula (UpperBound.jl:229-241)
                            double pos = (vl - val + dval * tk - gl * tl_) / (dval - gl);
                            if (pos != pos) pos = 0.0;
                            pos = fmin(fmax(pos, 0.0), step);
                            const double inter = vl + gl * pos;
                            bpart[k - 1] += fmax(fmax(fmax(vl, val), inter), 0.0);
                        }
                        vl = val; gl = dval; tl_ = tk;
                    }
            }
            double cs = 0.0;
            if (lane == 0) cum[0] = 0.0;
#pragma unroll
            for (int k = 0; k < kLrMaxG; ++k)
                if (k < G - 1) {
                    const double b = warp_sum(bpart[k]);
                    cs += b;
                    if (lane == 0) { box[k] = b; cum[k + 1] = cs * step; }
                }
            __syncwarp();
            const double e = rand_exp();
            next_event(e, tp, lambda_bar);
            exp_rv = e;
            if (exhausted) { status = PDMPFLUX_CHAIN_TAPE_EXHAUSTED; live = false; }
            else if (half) {  // erroneous_acceptance_rate!, SamplingLoopInplace.jl:131-151
                horizon = p.adaptive ? hbound : horizon;
                eb += 1;
                eva[eb % 5] = ar;
                half = false;
                need_build = false;
            } else if (tp > horizon) {  // move_to_horizon!, :87-101
                flow(horizon);
                ts += horizon; hh += 1;
                horizon = p.adaptive ? horizon * 1.01 : horizon;
            } else need_build = false;
        } else if (req == REQ_RATE) {
            // sampler.rate at tp (ZigZagSamplers.jl:83-86); lambda_i are also the categorical weights of the flip
            double part = 0.0;
            for (int i = lane; i < d; i += 32) {
                const double vi = vr(i);
                const double g = A[i * L.acc_stride + col0] + (xr(i) + tp * vi) * inv_s2;
                const double y = g * vi;
                const double l = (y > 0.0 ? y : 0.0);
                lam[i] = l;
                part += l;
            }
            const double lt = warp_sum(part);
            __syncwarp();
            ar = lt / lambda_bar;  // ac_step!, :113-129
            if (ar > 1.0) { need_build = true; half = true; }
            else if (rand_uniform() < ar) {  // if_accept!, :170-186
                if (p.use_t_stop && t + tp + ts > p.t_stop) {  // time-horizon variant, src/sample.jl:385-420
                    flow(p.t_stop - (t + ts));
                    t = p.t_stop;
                    ar = 0.0; eb = 0; rej = 0; hh = 0;
                    for (int k = 0; k < 5; ++k) eva[k] = 0.0;
                    record(p.col0 + ev);
                    ++ev;
                    status = PDMPFLUX_CHAIN_DONE;
                    live = false;
                } else {
                    const double uS = rand_uniform() * lt;  // categorical draw against the rates just computed
                    flow(tp);
                    if (lane == 0) {                         // first index with cumulative lambda > u S, else the last
                        double cp = 0.0;
                        int m = d - 1;
                        for (int i = 0; i < d; ++i) { cp += lam[i]; if (cp > uS) { m = i; break; } }
                        vr(m) = -vr(m);
                    }
                    __syncwarp();
                    t = t + tp + ts;
                    ts = 0.0; tp = 0.0;
                    if (exhausted) { status = PDMPFLUX_CHAIN_TAPE_EXHAUSTED; live = false; }
                    else {
                        record(p.col0 + ev);
                        ++ev;
                        steps = 0;
                        eb = 0; rej = 0; hh = 0;
                        for (int k = 0; k < 5; ++k) eva[k] = 0.0;
                        key.event = (uint32_t)(p.event0 + ev + 1);
                        sE = sU = 0;
                        need_build = true;
                        live = ev < p.n_events;
                        if (p.use_t_stop && !(t < p.t_stop)) { status = PDMPFLUX_CHAIN_DONE; live = false; }
                    }
                }
            } else {  // if_reject!, :188-203
                const double e3 = exp_rv + rand_exp();
                next_event(e3, tp, lambda_bar);
                horizon = p.adaptive ? horizon / 1.04 : horizon;
                exp_rv = e3;
                rej += 1;
                if (exhausted) { status = PDMPFLUX_CHAIN_TAPE_EXHAUSTED; live = false; }
                else if (tp > horizon) { flow(horizon); ts += horizon; hh += 1; need_build = true; }  // move_to_horizon2!
            }
        }
        __syncthreads();  // accumulators consumed before the next pass reuses the ring
    }

    if (!valid) return;
    for (int i = lane; i < d; i += 32) {
        p.sx[c * d + i] = xr(i);
        p.sv[c * d + i] = vr(i);
    }
    if (lane == 0) {
        p.st[c] = t; p.shorizon[c] = horizon; p.sar[c] = ar; p.status[c] = status;
        p.counters[2 * c] = n_builds; p.counters[2 * c + 1] = n_rates;
        p.tape_pos[3 * c] = pE; p.tape_pos[3 * c + 1] = pU;
        p.ncols[c] += ev;
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
int logreg_chains_per_block() { return kLrChains; }

}  // namespace pdmpflux
