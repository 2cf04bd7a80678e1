// explicit instantiations: sum of 1 exponential(s), with and without offset
#include "lm_launch.cuh"
namespace dphlm {
int launch_multi_exp1(bool offset, int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) {
    using M0 = MultiExp<1, false>;
    using M1 = MultiExp<1, true>;
    if (offset) { DPH_DISPATCH_G(M1, G, mle, a, pl, ctx, s); }
    DPH_DISPATCH_G(M0, G, mle, a, pl, ctx, s);
}
int launch_cov_multi_exp(int n_param, const CovArgs& a, cudaStream_t s) {
    switch (n_param) {
        case 2: return launch_cov<MultiExp<1, false>>(a, s);
        case 3: return launch_cov<MultiExp<1, true>>(a, s);
        case 4: return launch_cov<MultiExp<2, false>>(a, s);
        case 5: return launch_cov<MultiExp<2, true>>(a, s);
        case 6: return launch_cov<MultiExp<3, false>>(a, s);
        case 7: return launch_cov<MultiExp<3, true>>(a, s);
    }
    return ::dphlm::fail("multi_exp supports 2..7 parameters");
}
}  // namespace dphlm
