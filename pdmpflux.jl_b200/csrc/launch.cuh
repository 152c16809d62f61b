// Host-side dispatch from (team width, potential kind) to a kernel instantiation.
#pragma once
#include <cstdlib>

#include "chain.cuh"

namespace pdmpflux {

// Zig-Zag + Brent on a team of lanes: which line model the kernel keeps (the NW template argument).
//   -1  transposed Brent: teams of 8 with at most 8 owned coordinates per lane (d <= 64) -- every lane keeps the chain's
//       compressed model and takes its own iterations of the search (chain.cuh, build_bound_brent)
//   8 / 16  register-resident (A_j, B_j) of the owned coordinates, team reduction per rate evaluation
//   0   the shared-memory A/B arrays
// Mirrors the instantiations below; api.cu sizes the shared memory with it.  PDMPFLUX_TSPEC=0 switches the transposed
// search off (the tests compare the two).
inline int brent_reg_nw(int sampler, int path, int team, int n_own) {
    if (sampler != PDMPFLUX_ZIGZAG || path != kPathFastBrent) return 0;
    if (team != 4 && team != 8) return 0;
    if (team == 8 && n_own <= 8) {
        const char* e = std::getenv("PDMPFLUX_TSPEC");
        if (!e || std::atoi(e) != 0) return -1;
    }
    return n_own <= 8 ? 8 : (n_own <= 16 ? 16 : 0);
}

template <int TEAM, int SAMPLER, int POT, int PATH, int NW = 0>
cudaError_t launch_one(const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream) {
    // combinations without a fast path fall back to the generic kernel (same results, more passes)
    constexpr bool ok = PATH == kPathGeneric ||
                        (SAMPLER != PDMPFLUX_STICKY_ZIGZAG && SAMPLER != PDMPFLUX_SPEEDUP_ZIGZAG && Pot<POT>::kAffine &&
                         !(SAMPLER == PDMPFLUX_BOOMERANG && Pot<POT>::kSpecial > 0));
    if constexpr (!ok) return cudaErrorInvalidValue;
    else {
        auto kern = skeleton_kernel<TEAM, SAMPLER, POT, PATH, NW>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, block_threads_rt(TEAM, SAMPLER, PATH), smem, stream>>>(p);
        return cudaGetLastError();
    }
}

template <int TEAM, int SAMPLER, int POT>
cudaError_t launch_for_pot(int path, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream) {
    if constexpr (TEAM == 4) {  // only built for the register-resident Zig-Zag x Brent kernels
        if constexpr (SAMPLER == PDMPFLUX_ZIGZAG && Pot<POT>::kAffine) {
            switch (p.brent_nw) {
            case 8: return launch_one<TEAM, SAMPLER, POT, kPathFastBrent, 8>(p, grid, smem, stream);
            case 16: return launch_one<TEAM, SAMPLER, POT, kPathFastBrent, 16>(p, grid, smem, stream);
            }
        }
        return cudaErrorInvalidValue;
    } else {
        switch (path) {
        case kPathGeneric: return launch_one<TEAM, SAMPLER, POT, kPathGeneric>(p, grid, smem, stream);
        case kPathFastBrent:
            if constexpr (TEAM == 8 && SAMPLER == PDMPFLUX_ZIGZAG && Pot<POT>::kAffine) {
                switch (p.brent_nw) {
                case -1: return launch_one<TEAM, SAMPLER, POT, kPathFastBrent, -1>(p, grid, smem, stream);
                case 8: return launch_one<TEAM, SAMPLER, POT, kPathFastBrent, 8>(p, grid, smem, stream);
                case 16: return launch_one<TEAM, SAMPLER, POT, kPathFastBrent, 16>(p, grid, smem, stream);
                }
            }
            return launch_one<TEAM, SAMPLER, POT, kPathFastBrent>(p, grid, smem, stream);
        case kPathFastGrid: return launch_one<TEAM, SAMPLER, POT, kPathFastGrid>(p, grid, smem, stream);
        default: return cudaErrorInvalidValue;
        }
    }
}

template <int TEAM, int SAMPLER>
cudaError_t launch_for_team(int pot, int path, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream) {
    switch (pot) {
    case PDMPFLUX_GAUSS_STD: return launch_for_pot<TEAM, SAMPLER, PDMPFLUX_GAUSS_STD>(path, p, grid, smem, stream);
    case PDMPFLUX_GAUSS_DIAG: return launch_for_pot<TEAM, SAMPLER, PDMPFLUX_GAUSS_DIAG>(path, p, grid, smem, stream);
    case PDMPFLUX_GAUSS_EQUICORR: return launch_for_pot<TEAM, SAMPLER, PDMPFLUX_GAUSS_EQUICORR>(path, p, grid, smem, stream);
    case PDMPFLUX_BANANA: return launch_for_pot<TEAM, SAMPLER, PDMPFLUX_BANANA>(path, p, grid, smem, stream);
    case PDMPFLUX_BANANA_README_SCALAR:
        return launch_for_pot<TEAM, SAMPLER, PDMPFLUX_BANANA_README_SCALAR>(path, p, grid, smem, stream);
    default: return cudaErrorInvalidValue;
    }
}

template <int SAMPLER>
cudaError_t launch_for_sampler(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem,
                               cudaStream_t stream) {
    switch (team) {
    case 1: return launch_for_team<1, SAMPLER>(pot, path, p, grid, smem, stream);
    case 4:
        if constexpr (SAMPLER == PDMPFLUX_ZIGZAG) return launch_for_team<4, SAMPLER>(pot, path, p, grid, smem, stream);
        else return cudaErrorInvalidValue;
#ifdef PDMPFLUX_EXTRA_TEAM
    case PDMPFLUX_EXTRA_TEAM: return launch_for_team<PDMPFLUX_EXTRA_TEAM, SAMPLER>(pot, path, p, grid, smem, stream);
#endif
    case 8: return launch_for_team<8, SAMPLER>(pot, path, p, grid, smem, stream);
    case 32: return launch_for_team<32, SAMPLER>(pot, path, p, grid, smem, stream);
    default: return cudaErrorInvalidValue;
    }
}

// which path a (sampler, potential, config) combination runs on; mirrors the `ok` condition of launch_one
inline int select_path(int sampler, int pot, int grid_size, int vectorized, int deriv_mode) {
    const bool affine = pot == PDMPFLUX_GAUSS_STD || pot == PDMPFLUX_GAUSS_DIAG || pot == PDMPFLUX_GAUSS_EQUICORR ||
                        pot == PDMPFLUX_BANANA;
    if (!affine || (sampler == PDMPFLUX_BOOMERANG && pot == PDMPFLUX_BANANA)) return kPathGeneric;
    if (sampler == PDMPFLUX_STICKY_ZIGZAG) return kPathGeneric;  // masked velocities: per-node evaluation only
    if (sampler == PDMPFLUX_SPEEDUP_ZIGZAG) return kPathGeneric; // nonlinear flow: no affine line model
    if (grid_size == 0) return kPathFastBrent;
    if (sampler == PDMPFLUX_ZIGZAG) {
        // vectorised bound with analytic derivatives only: with finite differences the reference's cell maximum
        // depends on the O(sqrt(eps)) derivative noise, which the affine shortcut does not reproduce
        if (!vectorized || deriv_mode != PDMPFLUX_DERIV_JVP) return kPathGeneric;
    }
    return kPathFastGrid;  // BPS / FECMC / Boomerang: closed-form nodes, analytic or finite-difference derivative
}

// logreg.cu: Zig-Zag x logistic regression, four chains per CTA sharing each pass over X, FP64 DMMA
cudaError_t launch_logreg_zigzag(const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
size_t logreg_smem_bytes(int d, int G);
int logreg_chains_per_block();

// diagnostic.cu: RV_diagnostic (src/diagnostic.jl:37-75) over a batch of skeletons; uval is a [C][ld_u >= B+1] scratch
cudaError_t launch_rv_diagnostic(int kind, const PotParams& pp, int flow_kind, int d, int64_t ld_sk, int64_t n_sk,
                                 int64_t n_chains, const int64_t* ncols, int64_t B, const double* X, const double* V,
                                 const double* t, double* uval, int64_t ld_u, double* rv, cudaStream_t stream);

cudaError_t launch_skeleton_zigzag(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t launch_skeleton_bps(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t launch_skeleton_fecmc(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t launch_skeleton_boomerang(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t launch_skeleton_sticky(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t launch_skeleton_speedup(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);

}  // namespace pdmpflux
