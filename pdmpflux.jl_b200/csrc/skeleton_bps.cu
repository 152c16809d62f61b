// Instantiations of the thinning kernel for the bps sampler (one translation unit per sampler so the
// build parallelises).  See chain.cuh for the device logic and the reference lines it replaces.
#include "chain.cuh"
#include "launch.cuh"

namespace pdmpflux {
cudaError_t launch_skeleton_bps(int team, int pot, int path, const KernelParams& p, unsigned grid, size_t smem,
                                 cudaStream_t stream) {
    return launch_for_sampler<PDMPFLUX_BPS>(team, pot, path, p, grid, smem, stream);
}
}  // namespace pdmpflux
