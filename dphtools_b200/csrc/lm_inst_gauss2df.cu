// explicit instantiations: sampled symmetric 2-D Gaussian of fixed width
#include "lm_launch.cuh"
namespace dphlm {
int launch_gauss2d_fixed(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) { DPH_DISPATCH_G(Gauss2DFixed, G, mle, a, pl, ctx, s); }
int launch_cov_gauss2d_fixed(const CovArgs& a, cudaStream_t s) { return launch_cov<Gauss2DFixed>(a, s); }
}  // namespace dphlm
