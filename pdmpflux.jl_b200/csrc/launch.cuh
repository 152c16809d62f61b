// Host-side dispatch from (team width, potential kind) to a kernel instantiation.
#pragma once
#include "chain.cuh"

namespace pdmpflux {

template <int TEAM, int SAMPLER, int POT>
cudaError_t launch_one(const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream) {
    auto kern = skeleton_kernel<TEAM, SAMPLER, POT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, kBlockThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int TEAM, int SAMPLER>
cudaError_t launch_for_team(int pot, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream) {
    switch (pot) {
    case PDMPFLUX_GAUSS_STD: return launch_one<TEAM, SAMPLER, PDMPFLUX_GAUSS_STD>(p, grid, smem, stream);
    case PDMPFLUX_GAUSS_DIAG: return launch_one<TEAM, SAMPLER, PDMPFLUX_GAUSS_DIAG>(p, grid, smem, stream);
    case PDMPFLUX_GAUSS_EQUICORR: return launch_one<TEAM, SAMPLER, PDMPFLUX_GAUSS_EQUICORR>(p, grid, smem, stream);
    case PDMPFLUX_BANANA: return launch_one<TEAM, SAMPLER, PDMPFLUX_BANANA>(p, grid, smem, stream);
    case PDMPFLUX_BANANA_README_SCALAR:
        return launch_one<TEAM, SAMPLER, PDMPFLUX_BANANA_README_SCALAR>(p, grid, smem, stream);
    default: return cudaErrorInvalidValue;
    }
}

template <int SAMPLER>
cudaError_t launch_for_sampler(int team, int pot, const KernelParams& p, unsigned grid, size_t smem,
                               cudaStream_t stream) {
    switch (team) {
    case 1: return launch_for_team<1, SAMPLER>(pot, p, grid, smem, stream);
    case 8: return launch_for_team<8, SAMPLER>(pot, p, grid, smem, stream);
    case 32: return launch_for_team<32, SAMPLER>(pot, p, grid, smem, stream);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_skeleton_zigzag(int team, int pot, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t launch_skeleton_bps(int team, int pot, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t launch_skeleton_fecmc(int team, int pot, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);
cudaError_t launch_skeleton_boomerang(int team, int pot, const KernelParams& p, unsigned grid, size_t smem, cudaStream_t stream);

}  // namespace pdmpflux
