// explicit instantiations: rotated elliptical 2-D Gaussian
#include "lm_launch.cuh"
namespace dphlm {
int launch_gauss2d_elliptical(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) { DPH_DISPATCH_G(Gauss2DElliptical, G, mle, a, pl, ctx, s); }
int launch_cov_gauss2d_elliptical(const CovArgs& a, cudaStream_t s) { return launch_cov<Gauss2DElliptical>(a, s); }
}  // namespace dphlm
