// explicit instantiations: one or two exponentials convolved with an instrument response function
#include "lm_launch.cuh"
namespace dphlm {
namespace {
using I1 = MultiExpIrf<1, false>;
using I1o = MultiExpIrf<1, true>;
using I2 = MultiExpIrf<2, false>;
using I2o = MultiExpIrf<2, true>;
int go1(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) { DPH_DISPATCH_G_WIDE(I1, G, mle, a, pl, ctx, s); }
int go1o(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) { DPH_DISPATCH_G_WIDE(I1o, G, mle, a, pl, ctx, s); }
int go2(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) { DPH_DISPATCH_G_WIDE(I2, G, mle, a, pl, ctx, s); }
int go2o(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) { DPH_DISPATCH_G_WIDE(I2o, G, mle, a, pl, ctx, s); }
}  // namespace
int launch_multi_exp_irf(int n_param, int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) {
    switch (n_param) {
        case 2: return go1(G, mle, a, pl, ctx, s);
        case 3: return go1o(G, mle, a, pl, ctx, s);
        case 4: return go2(G, mle, a, pl, ctx, s);
        case 5: return go2o(G, mle, a, pl, ctx, s);
    }
    return ::dphlm::fail("multi_exp_irf supports 2..5 parameters (1-2 exponentials, optional offset)");
}
int launch_cov_multi_exp_irf(int n_param, const CovArgs& a, cudaStream_t s) {
    switch (n_param) {
        case 2: return launch_cov<I1>(a, s);
        case 3: return launch_cov<I1o>(a, s);
        case 4: return launch_cov<I2>(a, s);
        case 5: return launch_cov<I2o>(a, s);
    }
    return ::dphlm::fail("multi_exp_irf supports 2..5 parameters (1-2 exponentials, optional offset)");
}
}  // namespace dphlm
