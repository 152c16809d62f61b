// Instantiations of the thinning kernel for the Speed-Up Zig-Zag sampler (generic path only; one translation unit per
// sampler so the build parallelises).  See chain.cuh (flow_coef / speedup_rate_and_slope) for the device logic.
#include "chain.cuh"
#include "launch.cuh"

namespace pdmpflux {
cudaError_t launch_skeleton_speedup(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem,
                                    cudaStream_t stream) {
    if (path != kPathGeneric) return cudaErrorInvalidValue;
    return launch_for_sampler<PDMPFLUX_SPEEDUP_ZIGZAG>(team, pot, path, p, grid, smem, stream);
}
}  // namespace pdmpflux
