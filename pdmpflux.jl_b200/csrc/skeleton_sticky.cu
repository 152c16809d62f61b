// Instantiations of the thinning kernel for the Sticky Zig-Zag sampler (generic path only; one translation unit per
// sampler so the build parallelises).  See chain.cuh (run_events_sticky) for the device logic and the reference lines.
#include "chain.cuh"
#include "launch.cuh"

namespace pdmpflux {
cudaError_t launch_skeleton_sticky(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem,
                                   cudaStream_t stream) {
    if (path != kPathGeneric) return cudaErrorInvalidValue;
    return launch_for_sampler<PDMPFLUX_STICKY_ZIGZAG>(team, pot, path, p, grid, smem, stream);
}
}  // namespace pdmpflux
