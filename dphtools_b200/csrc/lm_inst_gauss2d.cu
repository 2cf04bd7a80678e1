// explicit instantiations: symmetric 2-D Gaussian
#include "lm_launch.cuh"
namespace dphlm {
int launch_gauss2d(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) { DPH_DISPATCH_G(Gauss2D, G, mle, a, pl, ctx, s); }
int launch_cov_gauss2d(const CovArgs& a, cudaStream_t s) { return launch_cov<Gauss2D>(a, s); }
}  // namespace dphlm
