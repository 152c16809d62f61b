// Zig-Zag on a Bayesian logistic-regression posterior (BASELINE.json config 4): four chains per CTA, FP64 DMMA pipeline.
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
// Execution: a CTA (384 threads, one per SM) owns 4 chains.  Between passes consumer warp c runs chain c's thinning
// state machine; whenever the chains need gradients they post a request (a bound: G times with Hessian columns, or a
// rate: 1 time) and the whole CTA makes ONE pass over X serving all of them (up to 64 packed residual columns; a
// request that does not fit waits a round).  A pass is a warp-specialised pipeline handed over with mbarriers only:
//   * feed: one producer thread per group issues TMA bulk copies (cp.async.bulk + mbarrier expect-tx, SASS UBLKCP.S.G)
//     of 32-row tiles of X (rows are contiguous in row-major X) and of y into a 5-deep ring;
//   * producers (warpgroups 1, 2, alternating tiles): z, w of all four chains by one DMMA product with
//     B = [x1 v1 x2 v2 x3 v3 x4 v4] -- exactly the 8 columns of mma.sync.m8n8k4.f64, no padding -- or, when every
//     asking chain has valid cached rows, by the incremental update z += dt w, w -= 2 v_m X[:, m]; then the residual
//     columns (two exps and one division per row of a bound) into a 3-deep ring;
//   * consumers (warpgroup 0): acc (d x columns) += Xtile^T . R with FP64 MMAs (SASS DMMA; tcgen05 has no FP64 kind),
//     accumulators in registers for the whole pass (setmaxnreg: 232 registers per consumer thread, 128 per producer).
// Four chains share every byte of X read from L2 and fill the MMA tile widths that a single chain would pad
// (z/w: 2 of 8 columns, a rate request: 1 of 8).  X (80 MB at C4 size) stays resident in the 126 MB L2.
// DESIGN.md 2b has the measurements behind each of these choices.
#include "common.cuh"
#include "philox.cuh"

namespace pdmpflux {

constexpr int kLrThreads = 384; // warpgroup 0: consumers (chain logic + main product); warpgroups 1, 2: producers
constexpr int kLrChains = 4;   // chains per CTA = consumer warps
constexpr int kLrXStages = 5;  // X-tile ring (TMA); 4 or 3 when d is so large that five tiles do not fit
constexpr int kLrRStages = 3;  // residual-tile ring (producers -> consumers)
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void group_barrier(int id, int nthreads) {  // named barrier of one warpgroup
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct LrShared {  // offsets (in doubles) into dynamic shared memory
    int bar, Xt, rs, ybuf, xv, zw, lam, box, cum, sched, tile, rs_tile, acc_stride, xstages, total;
};
// request schedule of one pass (shared memory, written by the chain warps, read by everybody)
struct LrSched {
    int n_cols;                    // packed residual columns
    int col0[kLrChains];           // first column of chain c
    int want[kLrChains];           // times chain c asks for (0: none, 1: rate, G: bound)
    int nt[kLrChains];             // times granted to chain c this pass (0 when it has to wait for the next one)
    double tp[kLrChains];          // rate request: the proposal time
    double gc[kLrChains], grem[kLrChains], gh[kLrChains];  // bound request: its time grid (see grid_t)
    // (z, w) cache in global memory (see lr_produce): mode 1 = the chain's cached z = X x0, w = X v0 are valid up to the
    // pending move x = x0 + pdt v0 followed by the flip of coordinate pm (whose velocity was pvm); mode 0 = recompute
    int mode[kLrChains], pm[kLrChains];
    double pdt[kLrChains], pvm[kLrChains];
    long long chain_id[kLrChains];  // the chain each consumer warp currently runs (chains are pulled from a work queue)
};
// mbarriers: full_x[s] (TMA landed), empty_x[s] (consumers done with the tile), full_rs[r] (residual tile written),
// empty_rs[r] (consumers done with it)
constexpr int kBarFullX = 0, kBarEmptyX = kLrXStages, kBarFullR = 2 * kLrXStages, kBarEmptyR = 2 * kLrXStages + kLrRStages;
constexpr int kLrBars = 2 * kLrXStages + 2 * kLrRStages;
__host__ __device__ inline LrShared lr_layout(int d, int xstages) {
    LrShared L;
    const int dp = (d + 7) / 8 * 8;
    L.xstages = xstages;
    L.tile = kLrRows * d + 32;                  // one X tile, rows contiguous (+ slack for fragment over-reads)
    L.rs_tile = kLrRows * kLrNcs + 32;          // one residual tile
    L.acc_stride = kLrMaxCols;
    int o = 0;
    L.bar = o; o += kLrBars;
    L.Xt = o; o += xstages * L.tile;            // X-tile ring (16-byte aligned: tile is even)
    L.rs = o; o += kLrRStages * L.rs_tile;      // residual-tile ring
    // after a pass the accumulators ((dp + 24) rows x kLrMaxCols) overlay [X ring | rs ring]: they fit by a wide margin
    L.ybuf = o; o += xstages * kLrRows;
    L.xv = o; o += (dp + 4) * 8;                // [k][8]: columns (2c, 2c+1) = (x, v) of chain c -- state AND B operand
    L.zw = o; o += 4 * kLrRows * 8;             // [producer group][parity][row][8]: (z, w) of chain c in columns (2c, 2c+1)
    L.lam = o; o += kLrChains * dp;
    L.box = o; o += kLrChains * 16;
    L.cum = o; o += kLrChains * 16;
    L.sched = o; o += (int)((sizeof(LrSched) + 7) / 8);
    L.total = o;
    return L;
}
__host__ __device__ inline LrShared lr_layout(int d) {  // the deepest X ring that fits the 227 KB of one SM
    int xs = kLrXStages;
    while (xs > 3 && sizeof(double) * (size_t)lr_layout(d, xs).total > 227u * 1024u) --xs;
    return lr_layout(d, xs);
}
size_t logreg_smem_bytes(int d, int G) { (void)G; return sizeof(double) * (size_t)lr_layout(d).total; }

// Ring bookkeeping is stateless: tile g (counted across passes, modulo 120 = lcm of the rings' parity periods for 3, 4
// or 5 X stages) lives in X slot g % xstages / residual slot g % 3, and the mbarrier phase parity of its use is
// (g / xstages) & 1 resp. (g / 3) & 1.
struct LrRing {
    int xs, rs;
    uint32_t xpar, rpar;
};
__device__ __forceinline__ LrRing lr_ring(uint32_t g, int xstages) {
    LrRing r;
    r.xs = (int)(g % (uint32_t)xstages); r.xpar = (g / (uint32_t)xstages) & 1u;
    r.rs = (int)(g % kLrRStages); r.rpar = (g / kLrRStages) & 1u;
    return r;
}

// ---- producer warpgroup pg (0 or 1): tiles pg, pg + 2, ...: TMA feed, (z, w) of the four chains, residual columns ----
__device__ void lr_produce(const KernelParams& p, double* sm, const LrShared& L, uint32_t gbase, int pg, int ptid) {
    const int d = p.d, lane = ptid & 31, pw = ptid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t n = p.pot.n;
    const LrSched* S = reinterpret_cast<const LrSched*>(sm + L.sched);
    const int kz = (d + 3) / 4;                      // k-steps of the z/w product
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
    const int64_t ntiles = (n + kLrRows - 1) / kLrRows;
    auto issue = [&](int64_t tile) {  // one thread: wait until the slot is free, then TMA the tile's rows of X and y
        const LrRing r = lr_ring(gbase + (uint32_t)tile, L.xstages);
        mbar_wait(&bar[kBarEmptyX + r.xs], r.xpar ^ 1u);
        const int64_t r0 = tile * kLrRows;
        const int rows = (int)min((int64_t)kLrRows, n - r0);
        // sizes rounded up to 16 bytes: the odd tail reads one double past the tile (zero padding behind X and y)
        const uint32_t xb = ((uint32_t)rows * d * 8u + 15u) & ~15u, yb = ((uint32_t)rows * 8u + 15u) & ~15u;
        mbar_expect_tx(&bar[kBarFullX + r.xs], xb + yb);
        bulk_load(sm + L.Xt + r.xs * L.tile, p.pot.vec + r0 * d, xb, &bar[kBarFullX + r.xs]);
        bulk_load(sm + L.ybuf + r.xs * kLrRows, p.pot.vec2 + r0, yb, &bar[kBarFullX + r.xs]);
    };
    if (ptid == 0 && pg < ntiles) issue(pg);
    // (z, w) = (X x, X v) of a chain change between two of its requests only by a straight move and at most one flip:
    // z += dt w, w -= 2 v_m X[:, m].  When every chain asking in this pass has valid cached rows (global memory,
    // [chain][tile][row][2], streamed with evict-first hints so that X keeps the L2) the 100-MMA z/w product of the
    // tile is skipped and each producer warp updates its own chain's rows in registers.  A chain without a valid cache
    // (its first request of a launch, the periodic exact refresh) makes the pass run the product, but only such chains
    // take its result: a chain's numbers never depend on which other chains happen to share its CTA.
    bool use_dmma = p.scratch == nullptr;
#pragma unroll
    for (int c = 0; c < kLrChains; ++c) use_dmma |= (S->nt[c] > 0 && S->mode[c] == 0);
    const int my_nt = S->nt[pw];
    double2* cache = nullptr;
    if (p.scratch != nullptr && my_nt > 0)
        cache = reinterpret_cast<double2*>(p.scratch) + ((int64_t)S->chain_id[pw] * ntiles) * kLrRows + lane;
    const double pdt = S->pdt[pw], pvm2 = 2.0 * S->pvm[pw];
    const int pm = S->pm[pw];
    const bool my_cached = cache != nullptr && S->mode[pw] == 1;  // this warp's chain updates its cached rows
    double2 zw_next = make_double2(0.0, 0.0);
    if (my_cached && pg < ntiles) zw_next = __ldcs(cache + (int64_t)pg * kLrRows);
    for (int64_t tile = pg; tile < ntiles; tile += 2) {
        const LrRing r = lr_ring(gbase + (uint32_t)tile, L.xstages);
        const int rows = (int)min((int64_t)kLrRows, n - tile * kLrRows);
        if (ptid == 0 && tile + 2 < ntiles) issue(tile + 2);  // this group's next tile: two consumer periods ahead
        mbar_wait(&bar[kBarFullX + r.xs], r.xpar);
        const double* Xs = sm + L.Xt + r.xs * L.tile;
        const double* ys = sm + L.ybuf + r.xs * kLrRows;
        double* zwb = sm + L.zw + (pg * 2 + (int)((tile >> 1) & 1)) * (kLrRows * 8);
        double z = 0.0, w = 0.0;  // this lane's row of this warp's chain
        if (my_cached) {
            z = zw_next.x; w = zw_next.y;
            if (tile + 2 < ntiles) zw_next = __ldcs(cache + (tile + 2) * kLrRows);  // next own tile, one iteration ahead
            z = fma(pdt, w, z);
            if (pm >= 0) w = fma(-pvm2, Xs[lane * d + pm], w);
            if (pdt != 0.0 || pm >= 0) __stcs(cache + tile * kLrRows, make_double2(z, w));
        }
        if (use_dmma) {
        // ---- (z, w) of the four chains for the tile: DMMA with B = [x1 v1 .. x4 v4] (k x 8); 8 rows per warp ----
        {
            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;  // two accumulator pairs: halves the dependent MMA chain
            const double* arow = Xs + (pw * 8 + gid) * d + tig;
            const double* bcol = sm + L.xv + tig * 8 + gid;
            int ks = 0;
            for (; ks + 1 < kz; ks += 2) {
                dmma(c0, c1, arow[4 * ks], bcol[32 * ks]);
                dmma(e0, e1, arow[4 * ks + 4], bcol[32 * ks + 32]);
            }
            if (ks < kz) dmma(c0, c1, arow[4 * ks], bcol[32 * ks]);
            double* zr = zwb + (pw * 8 + gid) * 8 + 2 * tig;  // C[row][2 tig], C[row][2 tig + 1] = (z, w) of chain tig
            zr[0] = c0 + e0;
            zr[1] = c1 + e1;
        }
        group_barrier(1 + pg, 128);
        if (!my_cached) {
            z = zwb[lane * 8 + 2 * pw]; w = zwb[lane * 8 + 2 * pw + 1];
            if (cache != nullptr) __stcs(cache + tile * kLrRows, make_double2(z, w));
        }
        }
        mbar_wait(&bar[kBarEmptyR + r.rs], r.rpar ^ 1u);  // consumers are done with the tile that used this slot
        // ---- residual columns: warp c evaluates chain c's requested times, lane = row of the tile ----
        {
            const int nt = my_nt;
            if (nt) {
                const bool live = lane < rows;
                const double yy = ys[lane];
                double* rp = sm + L.rs + r.rs * L.rs_tile + lane * kLrNcs + S->col0[pw];
                if (nt == 1) {  // rate at the proposal time
                    const double sg = 1.0 / (1.0 + exp(-(z + S->tp[pw] * w)));
                    rp[0] = live ? sg - yy : 0.0;
                } else {        // bound: residual and Hessian-weight columns at the G grid times
                    const double gc = S->gc[pw], grem = S->grem[pw], gh = S->gh[pw];
                    if (fabs(z) + fabs(gh * w) < 40.0) {
                        // eta_k = z + t_k w on a uniform grid: exp(-eta_k) is a geometric progression -- two exps per
                        // row instead of G -- and the G reciprocals 1 / (1 + q_k) come from ONE division (prefix
                        // products, invert the last, peel backwards).  Every factor is below e^40, so the product of
                        // up to 12 stays finite; the accumulated rounding (a few tens of ulp) is far inside the parity
                        // tolerance.
                        double q[kLrMaxG], pre[kLrMaxG];
                        const double rho = exp(-fma(1.0, gc, grem) * w);
                        q[0] = exp(-z);
                        pre[0] = 1.0 + q[0];
#pragma unroll
                        for (int k = 1; k < kLrMaxG; ++k) {
                            q[k] = q[k - 1] * rho;
                            pre[k] = (k < nt) ? pre[k - 1] * (1.0 + q[k]) : pre[k - 1];
                        }
                        double inv = 1.0 / pre[kLrMaxG - 1];
#pragma unroll
                        for (int k = kLrMaxG - 1; k >= 0; --k)
                            if (k < nt) {
                                const double sg = (k > 0) ? inv * pre[k - 1] : inv;
                                inv *= 1.0 + q[k];
                                rp[k] = live ? sg - yy : 0.0;
                                rp[nt + k] = live ? q[k] * sg * sg * w : 0.0;  // sigma (1 - sigma) w with 1 - sigma = q sigma
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[kBarFullR + r.rs]);  // release: this warp's columns of the residual tile are written
    }
}

// ---- consumer warpgroup: acc (d x columns) += Xtile^T . R over all tiles, then parked in shared memory ----
//   A[m][k] = Xt[row0 + k][i0 + m], B[k][n] = R[row0 + k][n0 + n]
// Work split (compile-time, so that no MMA is predicated -- a predicated mma.sync costs a WARPSYNC per group and ran the
// loop 1.5x slower): every warp takes NA whole m-tiles {cw, cw + 4, ..} (a tile past the last one computes garbage that
// is never stored); when n_mt % 4 == 1 (SPLIT) the odd last m-tile is split over the four warps by k-steps (two of the
// eight each) so that every warp issues the same number of MMAs; its four partial sums are parked in separate row
// blocks and added by the reader.
template <int NT, int NA, bool SPLIT>
__device__ __forceinline__ void lr_consume(const KernelParams& p, double* sm, const LrShared& L, uint32_t gbase, int ctid) {
    const int d = p.d, lane = ctid & 31, cw = ctid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t n = p.pot.n;
    const int n_mt = (d + 7) / 8;                    // m-tiles (coordinates), <= 16
    constexpr int kSlots = NA + (SPLIT ? 1 : 0);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
    double acc[kSlots][NT][2];
#pragma unroll
    for (int a = 0; a < kSlots; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int64_t ntiles = (n + kLrRows - 1) / kLrRows;
    for (int64_t tile = 0; tile < ntiles; ++tile) {
        const LrRing r = lr_ring(gbase + (uint32_t)tile, L.xstages);
        mbar_wait(&bar[kBarFullX + r.xs], r.xpar);
        mbar_wait(&bar[kBarFullR + r.rs], r.rpar);
        const double* ap = sm + L.Xt + r.xs * L.tile + tig * d + cw * 8 + gid;
        const double* bp = sm + L.rs + r.rs * L.rs_tile + tig * kLrNcs + gid;
        if constexpr (NA > 0) {
            // Operands of k-step ks + 1 are loaded while the MMAs of k-step ks issue (register double buffer).  Rows
            // past this warp's last real m-tile stay inside the rings, so the loads need no guard either.
            double av[2][NA], bv[2][NT];
            auto load = [&](int ks, int buf) {
#pragma unroll
                for (int b = 0; b < NT; ++b) bv[buf][b] = bp[4 * ks * kLrNcs + 8 * b];
#pragma unroll
                for (int a = 0; a < NA; ++a) av[buf][a] = ap[4 * ks * d + 32 * a];
            };
            load(0, 0);
#pragma unroll
            for (int ks = 0; ks < kLrRows / 4; ++ks) {
                const int cur = ks & 1;
                if (ks + 1 < kLrRows / 4) load(ks + 1, cur ^ 1);
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int b = 0; b < NT; ++b) dmma(acc[a][b][0], acc[a][b][1], av[cur][a], bv[cur][b]);
            }
        }
        if constexpr (SPLIT) {  // k-steps 2 cw, 2 cw + 1 of the last m-tile
            const double* as_ = sm + L.Xt + r.xs * L.tile + tig * d + (n_mt - 1) * 8 + gid;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const int ks = 2 * cw + kk;
                const double a_ = as_[4 * ks * d];
#pragma unroll
                for (int b = 0; b < NT; ++b) dmma(acc[NA][b][0], acc[NA][b][1], a_, bp[4 * ks * kLrNcs + 8 * b]);
            }
        }
        __syncwarp();
        if (lane == 0) {  // this warp is done with both tiles
            mbar_arrive(&bar[kBarEmptyR + r.rs]);
            mbar_arrive(&bar[kBarEmptyX + r.xs]);
        }
    }
    // ---- accumulator fragments -> shared memory acc[i][col], overlaying the (now idle) rings ----
    group_barrier(3, 128);  // every consumer warp has finished reading the rings
    double* A = sm + L.Xt;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
        const int mt = cw + 4 * a;
        if (mt < n_mt - (SPLIT ? 1 : 0)) {
#pragma unroll
            for (int b = 0; b < NT; ++b) {
                const int i = mt * 8 + gid, c = 8 * b + 2 * tig;
                A[i * L.acc_stride + c] = acc[a][b][0];
                A[i * L.acc_stride + c + 1] = acc[a][b][1];
            }
        }
    }
    if constexpr (SPLIT) {
        const int i = (cw == 0 ? (n_mt - 1) * 8 : n_mt * 8 + (cw - 1) * 8) + gid;
#pragma unroll
        for (int b = 0; b < NT; ++b) {
            const int c = 8 * b + 2 * tig;
            A[i * L.acc_stride + c] = acc[NA][b][0];
            A[i * L.acc_stride + c + 1] = acc[NA][b][1];
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NA, bool SPLIT>
__global__ void __launch_bounds__(kLrThreads, 1) logreg_zigzag_kernel(const __grid_constant__ KernelParams p) {
    extern __shared__ __align__(128) double sm[];
    const int d = p.d, G = p.G, tid = threadIdx.x, lane = tid & 31;
    const int wgroup = tid >> 7;                     // 0: consumers, 1 / 2: producers
    const int w = (tid >> 5) & (kLrChains - 1);      // consumer warp w runs chain w
    const LrShared L = lr_layout(d);
    const int dp = (d + 7) / 8 * 8;
    const double inv_s2 = p.pot.inv_s2;
    LrSched* S = reinterpret_cast<LrSched*>(sm + L.sched);
    // Consumer warp w starts with chain blockIdx.x * 4 + w and pulls further chains from a global work queue as soon as
    // its chain has finished this launch's events (persistent CTAs: no CTA waits for the slowest of four fixed chains).
    const int64_t c_raw = (int64_t)blockIdx.x * kLrChains + w;
    bool valid = wgroup == 0 && c_raw < p.n_chains;
    int64_t c = valid ? c_raw : 0;

    // ---- shared state: xv[k][8] holds (x, v) of chain w in columns (2w, 2w+1); it is also the z/w B operand ----
    for (int e = tid; e < (dp + 4) * 8; e += kLrThreads) sm[L.xv + e] = 0.0;
    for (int e = tid; e < L.xstages * L.tile + kLrRStages * L.rs_tile; e += kLrThreads) sm[L.Xt + e] = 0.0;  // rings, slack
    __syncthreads();
    if (valid)
        for (int i = lane; i < d; i += 32) {
            sm[L.xv + i * 8 + 2 * w] = p.sx[c * d + i];
            sm[L.xv + i * 8 + 2 * w + 1] = p.sv[c * d + i];
        }
    if (wgroup == 0 && lane == 0) S->chain_id[w] = valid ? (long long)c : -1;
    fence_async_smem();  // order these generic-proxy writes before the TMA (async-proxy) writes into the same buffers
    if (tid == 0) {
        uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
        for (int s = 0; s < kLrXStages; ++s) {
            mbar_init(&bar[kBarFullX + s], 1);             // the feeding thread's expect-tx arrival (+ the TMA bytes)
            mbar_init(&bar[kBarEmptyX + s], kLrChains);    // one arrival per consumer warp
        }
        for (int s = 0; s < kLrRStages; ++s) {
            mbar_init(&bar[kBarFullR + s], 4);             // one arrival per warp of the producer group
            mbar_init(&bar[kBarEmptyR + s], kLrChains);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    uint32_t gbase = 0;  // tiles streamed so far, modulo 120 (see lr_ring)
    const int64_t ntiles_pass = (p.pot.n + kLrRows - 1) / kLrRows;
    // ---- PDMPState scalars of chain w (warp-uniform registers) ----
    double t = p.st[c], horizon = p.shorizon[c], ar = p.sar[c];
    int status = valid ? p.status[c] : 0;
    int64_t n_builds = p.counters[2 * c], n_rates = p.counters[2 * c + 1];
    DrawKey key;
    uint64_t gchain = (uint64_t)(p.chain_offset + c);
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

    if (wgroup != 0) {
        // ---- producer warpgroups: mirror the consumers' CTA-wide barriers, work only inside the passes ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 128;");
        while (true) {
            __syncthreads();                       // requests posted
            __syncthreads();                       // schedule packed
            const int any = __syncthreads_or(0);
            if (!any) break;
            if (S->n_cols == 0) continue;
            fence_async_smem();
            __syncthreads();                       // pass start: the accumulator overlay of the last pass is dead
            lr_produce(p, sm, L, gbase, wgroup - 1, tid & 127);
            gbase = (gbase + (uint32_t)(ntiles_pass % 120)) % 120u;
            __syncthreads();                       // pass end: accumulators parked
            __syncthreads();                       // accumulators consumed
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");

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
    // (z, w) cache bookkeeping of this chain (see lr_produce): what happened to (x, v) since its rows were last written
    const bool cache_on = p.scratch != nullptr;
    bool cache_valid = false;
    int served = 0, pend_m = -1;
    double pend_dt = 0.0, pend_vm = 0.0;
    auto flow = [&](double tt) {  // ZigZagSamplers.jl:80
        for (int i = lane; i < d; i += 32) xr(i) += vr(i) * tt;
        pend_dt += tt;
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

    // ---- work queue: write a finished chain's PDMPState back, take over the next chain ----
    auto store_chain = [&]() {
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
    };
    auto load_chain = [&](int64_t cc) {
        valid = cc < p.n_chains;
        c = valid ? cc : 0;
        __syncwarp();
        for (int i = lane; i < d; i += 32) {
            xr(i) = valid ? p.sx[c * d + i] : 0.0;
            vr(i) = valid ? p.sv[c * d + i] : 0.0;
        }
        if (lane == 0) S->chain_id[w] = valid ? (long long)c : -1;
        __syncwarp();
        t = p.st[c]; horizon = p.shorizon[c]; ar = p.sar[c];
        status = valid ? p.status[c] : 0;
        n_builds = p.counters[2 * c]; n_rates = p.counters[2 * c + 1];
        gchain = (uint64_t)(p.chain_offset + c);
        key.chain_lo = (uint32_t)gchain; key.chain_hi8 = (uint32_t)(gchain >> 32) << 8;
        key.event = (uint32_t)(p.event0 + 1);
        sE = sU = 0;
        pE = p.tape_pos[3 * c]; pU = p.tape_pos[3 * c + 1];
        tE = p.tE + c * p.nE; tU = p.tU + c * p.nU;
        exhausted = false;
        eb = rej = hh = 0;
        for (int k = 0; k < 5; ++k) eva[k] = 0.0;
        ev = 0; steps = 0;
        ts = tp = lambda_bar = exp_rv = hbound = 0.0;
        need_build = true; half = false;
        cache_valid = false; served = 0; pend_m = -1; pend_dt = 0.0;
        if (p.use_t_stop && status == 0 && !(t < p.t_stop)) status = PDMPFLUX_CHAIN_DONE;
        live = valid && status == 0;
    };

    while (true) {
        // ---- 0. a chain that is done with this launch hands its warp to the next chain in the queue ----
        if (p.work_counter != nullptr)
            while (valid && !live) {
                store_chain();
                long long next = 0;
                if (lane == 0) next = p.work_start + (long long)atomicAdd(reinterpret_cast<unsigned long long*>(p.work_counter), 1ull);
                next = __shfl_sync(0xffffffffu, next, 0);
                load_chain(next);
            }

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
            S->mode[w] = (cache_valid && (served & 63) != 63) ? 1 : 0;  // every 64th request recomputes exactly
            S->pdt[w] = pend_dt; S->pm[w] = pend_m; S->pvm[w] = pend_vm;
        }
        if (req != REQ_NONE) { cache_valid = cache_on; pend_dt = 0.0; pend_m = -1; ++served; }  // rows current after this pass
        // the CTA goes on while any chain has work left (a request, or a pending fall-through transition)
        const int any = __syncthreads_or(live ? 1 : 0);
        if (!any) break;
        ++round;
        if (S->n_cols == 0) continue;  // only fall-through transitions this round

        // ---- 2. one pass over X for all requests (specialised on the number of 8-column tiles) ----
        // The accumulators of the previous pass overlaid the rings through the generic proxy: order that before the
        // TMA (async proxy) refills them.
        fence_async_smem();
        __syncthreads();
        switch ((S->n_cols + 7) >> 3) {
            case 1: lr_consume<1, NA, SPLIT>(p, sm, L, gbase, tid); break;
            case 2: lr_consume<2, NA, SPLIT>(p, sm, L, gbase, tid); break;
            case 3: lr_consume<3, NA, SPLIT>(p, sm, L, gbase, tid); break;
            case 4: lr_consume<4, NA, SPLIT>(p, sm, L, gbase, tid); break;
            case 5: lr_consume<5, NA, SPLIT>(p, sm, L, gbase, tid); break;
            case 6: lr_consume<6, NA, SPLIT>(p, sm, L, gbase, tid); break;
            case 7: lr_consume<7, NA, SPLIT>(p, sm, L, gbase, tid); break;
            default: lr_consume<8, NA, SPLIT>(p, sm, L, gbase, tid); break;
        }
        gbase = (gbase + (uint32_t)(ntiles_pass % 120)) % 120u;
        __syncthreads();  // pass end: every warp's accumulators are parked
        const double* Aov = sm + L.Xt;
        const int n_mt = (d + 7) / 8;
        constexpr bool split = SPLIT;
        auto A = [&](int i, int col) -> double {  // accumulator (i, col); the split last m-tile is the sum of four partials
            double v = Aov[i * L.acc_stride + col];
            if (split && i >= (n_mt - 1) * 8) {
                const double* q = Aov + (n_mt * 8 + (i - (n_mt - 1) * 8)) * L.acc_stride + col;
                v += q[0] + q[8 * L.acc_stride] + q[16 * L.acc_stride];
            }
            return v;
        };
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
                        const double g = A(i, col0 + k) + (xi + tk * vi) * inv_s2;
                        const double hv = A(i, col0 + G + k) + vi * inv_s2;
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
                const double g = A(i, col0) + (xr(i) + tp * vi) * inv_s2;
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
                    int m = d - 1;
                    if (lane == 0) {                         // first index with cumulative lambda > u S, else the last
                        double cp = 0.0;
                        for (int i = 0; i < d; ++i) { cp += lam[i]; if (cp > uS) { m = i; break; } }
                    }
                    m = __shfl_sync(0xffffffffu, m, 0);
                    pend_m = m; pend_vm = vr(m);
                    __syncwarp();
                    if (lane == 0) vr(m) = -pend_vm;
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
        __syncthreads();  // accumulators consumed before the next pass reuses the rings
    }

    if (valid) store_chain();  // (without a work queue: the chain this warp started with)
}

template <int NA, bool SPLIT>
static cudaError_t launch_variant(const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(logreg_zigzag_kernel<NA, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    logreg_zigzag_kernel<NA, SPLIT><<<grid, kLrThreads, smem, stream>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_logreg_zigzag(const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream) {
    const int n_mt = (p.d + 7) / 8;              // m-tiles of 8 coordinates (d <= 128: at most 16)
    const bool split = (n_mt & 3) == 1;          // see lr_consume
    const int na = split ? (n_mt - 1) / 4 : (n_mt + 3) / 4;
    switch (na * 2 + (split ? 1 : 0)) {
        case 1: return launch_variant<0, true>(p, grid, smem, stream);
        case 3: return launch_variant<1, true>(p, grid, smem, stream);
        case 5: return launch_variant<2, true>(p, grid, smem, stream);
        case 7: return launch_variant<3, true>(p, grid, smem, stream);
        case 2: return launch_variant<1, false>(p, grid, smem, stream);
        case 4: return launch_variant<2, false>(p, grid, smem, stream);
        case 6: return launch_variant<3, false>(p, grid, smem, stream);
        case 8: return launch_variant<4, false>(p, grid, smem, stream);
    }
    return cudaErrorInvalidValue;
}
int logreg_chains_per_block() { return kLrChains; }

}  // namespace pdmpflux
