// explicit instantiations: sum of 3 exponential(s), with and without offset
#include "lm_launch.cuh"
namespace dphlm {
int launch_multi_exp3(bool offset, int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) {
    using M0 = MultiExp<3, false>;
    using M1 = MultiExp<3, true>;
    if (offset) { DPH_DISPATCH_G(M1, G, mle, a, pl, ctx, s); }
    DPH_DISPATCH_G(M0, G, mle, a, pl, ctx, s);
}
}  // namespace dphlm
