// Host build of the kernel algorithm (lm_core.cuh with one lane per fit).  TEST INFRASTRUCTURE ONLY:
// lets the CPU test-suite execute the fp32 fitter logic against the float64 oracle without a GPU.
// Nothing under dphtools_b200/ loads this library; the product path is the CUDA build.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../dphtools_b200/csrc/lm_core.cuh"

using namespace dphlm;

template <class Mo, bool MLE>
static void run(int h, int w, long n, const float* data, const float* xaxis, const float* p0, const Options& opt,
                float* popt, float* p_ls, int* info, int* nfev, float* chisq, int* counts) {
    const int M = h * w;
    std::vector<Coord> cxy(M);
    for (int i = 0; i < M; ++i) {
        if (xaxis) { cxy[i].x = xaxis[i]; cxy[i].y = 0.f; }
        else { cxy[i].x = float(i % w); cxy[i].y = float(i / w); }
    }
    std::vector<float> y(M), e(2 * Mo::NE * M);
    Lanes<1> ln{0};
    for (long f = 0; f < n; ++f) {
        for (int i = 0; i < M; ++i) {
            float v = data[f * M + i];
            y[i] = MLE ? clamp_nonneg(v) : v;
        }
        PixelState<Mo::NE> px{y.data(), e.data(), cxy.data(), M, 0};
        FitResult<Mo::P> r;
        fit_one<Mo, MLE, 1>(ln, px, p0 + f * Mo::P, opt, true, r);
        for (int k = 0; k < Mo::P; ++k) { popt[f * Mo::P + k] = r.x[k]; p_ls[f * Mo::P + k] = r.p_ls[k]; }
        info[f] = r.info; nfev[f] = r.nfev; chisq[f] = r.chisq;
        counts[4 * f + 0] = r.nfev_ls; counts[4 * f + 1] = r.n_acc; counts[4 * f + 2] = r.n_rej; counts[4 * f + 3] = 0;
    }
}

extern "C" int emu_fit_batch(int model, int method, int n_param, int h, int w, long n, const float* data,
                             const float* xaxis, const float* p0, float ftol, float xtol, float gtol, float factor,
                             int maxfev, float* popt, float* p_ls, int* info, int* nfev, float* chisq, int* counts) {
    Options opt{ftol, xtol, gtol, factor, maxfev};
#define RUN(...)                                                                                                 \
    do {                                                                                                         \
        if (method) run<__VA_ARGS__, true>(h, w, n, data, xaxis, p0, opt, popt, p_ls, info, nfev, chisq, counts);          \
        else run<__VA_ARGS__, false>(h, w, n, data, xaxis, p0, opt, popt, p_ls, info, nfev, chisq, counts);                \
        return 0;                                                                                                \
    } while (0)
    if (model == 0 && n_param == 5) RUN(Gauss2D);
    if (model == 1 && n_param == 7) RUN(Gauss2DElliptical);
    if (model == 2 && n_param == 2) RUN(MultiExp<1, false>);
    if (model == 2 && n_param == 3) RUN(MultiExp<1, true>);
    if (model == 2 && n_param == 4) RUN(MultiExp<2, false>);
    if (model == 2 && n_param == 5) RUN(MultiExp<2, true>);
    if (model == 2 && n_param == 6) RUN(MultiExp<3, false>);
    if (model == 2 && n_param == 7) RUN(MultiExp<3, true>);
    return -1;
}
