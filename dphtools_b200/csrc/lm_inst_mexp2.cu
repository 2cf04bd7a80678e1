// explicit instantiations: sum of 2 exponential(s), with and without offset
#include "lm_launch.cuh"
namespace dphlm {
int launch_multi_exp2(bool offset, int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) {
    using M0 = MultiExp<2, false>;
    using M1 = MultiExp<2, true>;
    if (offset) { DPH_DISPATCH_G(M1, G, mle, a, pl, ctx, s); }
    DPH_DISPATCH_G(M0, G, mle, a, pl, ctx, s);
}
}  // namespace dphlm
