// Host build of the kernel algorithm (lm_core.cuh with one lane per fit).  TEST INFRASTRUCTURE ONLY:
// lets the CPU test-suite execute the fp32 fitter logic against the float64 oracle without a GPU.
// Nothing under dphtools_b200/ loads this library; the product path is the CUDA build.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../dphtools_b200/csrc/lm_core.cuh"

using namespace dphlm;

struct Parked {
    long fit;
    std::vector<float> rec;
};

struct Buffers {
    int M;
    long n;
    const float* data;
    const float* p0;
    float *popt, *p_ls, *chisq;
    int *info, *nfev, *counts;
    std::vector<int> ier;
    int budget;  // trips after which a fit is parked and later resumed (exercises save/resume); 0 = never
    std::vector<Parked> parked;
    std::vector<char> was_parked;
    const float* consts;  // model constants of the call, or null
    int skip_prefit = 0;  // lm() alone: stage B starts at p0 (DPHLM_FLAG_SKIP_PREFIT)
};

template <class Mo>
struct HostIoPrefit {
    Buffers& b;
    long next = 0;
    size_t resumed = 0;
    int budget() const { return b.budget > 0 ? b.budget : 1 << 30; }
    const float* consts() const { return b.consts; }
    bool park(const Lanes<1>&, const StageA<Mo>& st, long long fit) {
        if (b.was_parked[fit]) return false;  // park once
        b.was_parked[fit] = 1;
        Parked pk{fit, std::vector<float>(StageA<Mo>::kRecord)};
        st.save(pk.rec.data());
        b.parked.push_back(pk);
        return true;
    }
    bool fetch(const Lanes<1>&, PixelState<Mo::NE>& px, StageA<Mo>& st, long long& fit) {
        float pstart[Mo::P];
        for (;;) {
            if (next >= b.n) {
                if (resumed >= b.parked.size()) return false;
                const Parked& pk = b.parked[resumed++];
                for (int i = 0; i < b.M; ++i) px.y[i] = b.data[pk.fit * b.M + i];
                st.resume(pk.rec.data(), b.consts);
                fit = pk.fit;
                return true;
            }
            long idx = next++;
            bool bad = false;
            for (int i = 0; i < b.M; ++i) { float v = b.data[idx * b.M + i]; bad |= !std::isfinite(v); px.y[i] = v; }
            for (int k = 0; k < Mo::P; ++k) pstart[k] = b.p0[idx * Mo::P + k];
            if (!bad) { fit = idx; st.start(pstart, b.consts); return true; }
            b.ier[idx] = 100;
            for (int k = 0; k < Mo::P; ++k) b.p_ls[idx * Mo::P + k] = pstart[k];
            for (int k = 0; k < 4; ++k) b.counts[4 * idx + k] = 0;
        }
    }
    void store(const Lanes<1>&, const StageA<Mo>& st, long long fit) {
        for (int k = 0; k < Mo::P; ++k) b.p_ls[fit * Mo::P + k] = st.x[k];
        b.ier[fit] = st.ier;
        b.counts[4 * fit] = st.nfev;
    }
};

template <class Mo, bool MLE>
struct HostIoMain {
    Buffers& b;
    long next = 0;
    size_t resumed = 0;
    int budget() const { return b.budget > 0 ? b.budget : 1 << 30; }
    const float* consts() const { return b.consts; }
    bool park(const Lanes<1>&, const StageB<Mo, MLE>& st, long long fit) {
        if (b.was_parked[fit]) return false;
        b.was_parked[fit] = 1;
        Parked pk{fit, std::vector<float>(StageB<Mo, MLE>::kRecord)};
        st.save(pk.rec.data());
        b.parked.push_back(pk);
        return true;
    }
    bool fetch(const Lanes<1>&, PixelState<Mo::NE>& px, StageB<Mo, MLE>& st, long long& fit) {
        float pstart[Mo::P];
        for (;;) {
            if (next >= b.n) {
                if (resumed >= b.parked.size()) return false;
                const Parked& pk = b.parked[resumed++];
                for (int i = 0; i < b.M; ++i) { float v = b.data[pk.fit * b.M + i]; px.y[i] = MLE ? clamp_nonneg(v) : v; }
                st.resume(pk.rec.data(), b.consts);
                fit = pk.fit;
                return true;
            }
            long idx = next++;
            int ier = b.ier[idx];
            for (int k = 0; k < Mo::P; ++k) pstart[k] = b.p_ls[idx * Mo::P + k];
            if (ier >= 1 && ier <= 4) {
                for (int i = 0; i < b.M; ++i) { float v = b.data[idx * b.M + i]; px.y[i] = MLE ? clamp_nonneg(v) : v; }
                fit = idx;
                st.start(pstart, b.consts);
                return true;
            }
            for (int k = 0; k < Mo::P; ++k) b.popt[idx * Mo::P + k] = pstart[k];
            b.info[idx] = ier == 0 ? -99 : -ier;
            b.nfev[idx] = 0;
            b.chisq[idx] = 0.f;
        }
    }
    void store(const Lanes<1>&, const StageB<Mo, MLE>& st, long long fit) {
        for (int k = 0; k < Mo::P; ++k) b.popt[fit * Mo::P + k] = st.xret[k];
        b.info[fit] = st.info; b.nfev[fit] = st.ev; b.chisq[fit] = st.chisq_ret;
        b.counts[4 * fit + 1] = st.n_acc; b.counts[4 * fit + 2] = st.n_rej; b.counts[4 * fit + 3] = st.n_tie;
    }
};

template <class Mo, bool MLE>
static void run(int h, int w, const float* xaxis, const Options& opt, Buffers& b) {
    const int M = h * w;
    const int stride = (M + 1) & ~1;  // even: pairs of adjacent pixels are 8-byte aligned
    std::vector<double> backing(4 * stride);  // double: 8-byte aligned storage
    float* cx = reinterpret_cast<float*>(backing.data());
    float* cy = cx + stride;
    for (int i = 0; i < stride; ++i) fill_coords<Mo>(cx, cy, i, h, w, xaxis);
    std::vector<double> ybuf(stride), ebuf(Mo::NE * stride);
    float* y = reinterpret_cast<float*>(ybuf.data());
    float* e = reinterpret_cast<float*>(ebuf.data());
    Lanes<1> ln{0, 1u};
    PixelState<Mo::NE> px{y, e, cx, cy, M, stride, 0};
    std::vector<double> table(Mo::table_floats(h, w) / 2 + 2);  // double: 16-byte aligned by operator new
    px.tbl = reinterpret_cast<float*>(table.data());
    px.w = w;
    float stash[32];
    px.stash = stash;
    if (b.skip_prefit) {
        for (long i = 0; i < b.n; ++i) {
            b.ier[i] = 1;
            b.counts[4 * i] = 0;
            for (int k = 0; k < Mo::P; ++k) b.p_ls[i * Mo::P + k] = b.p0[i * Mo::P + k];
        }
    } else {
        HostIoPrefit<Mo> ioa{b};
        run_stage<StageA<Mo>, Mo, 1>(ln, px, opt, ioa);
    }
    b.parked.clear();
    b.was_parked.assign(b.n, 0);
    HostIoMain<Mo, MLE> iob{b};
    run_stage<StageB<Mo, MLE>, Mo, 1>(ln, px, opt, iob);
}

extern "C" int emu_fit_batch(int model, int method, int n_param, int h, int w, long n, const float* data,
                             const float* xaxis, const float* p0, float ftol, float xtol, float gtol, float factor,
                             int maxfev, int budget, float* popt, float* p_ls, int* info, int* nfev, float* chisq, int* counts,
                             const float* consts, int skip_prefit, int exact_ties) {
    Options opt{ftol, xtol, gtol, factor, maxfev, exact_ties};
    Buffers b{h * w, n, data, p0, popt, p_ls, chisq, info, nfev, counts, std::vector<int>(n, 0), budget, {}, std::vector<char>(n, 0), consts, skip_prefit};
#define RUN(...)                                                 \
    do {                                                         \
        if (method) run<__VA_ARGS__, true>(h, w, xaxis, opt, b);  \
        else run<__VA_ARGS__, false>(h, w, xaxis, opt, b);        \
        return 0;                                                \
    } while (0)
    if (model == 0 && n_param == 5) RUN(Gauss2D);
    if (model == 1 && n_param == 7) RUN(Gauss2DElliptical);
    if (model == 3 && n_param == 5) RUN(Gauss2DIntegrated);
    if (model == 4 && n_param == 4 && consts) RUN(Gauss2DIntegratedFixed);
    if (model == 6 && n_param == 4 && consts) RUN(Gauss2DFixed);
    if (model == 5 && n_param == 2 && consts) RUN(MultiExpIrf<1, false>);
    if (model == 5 && n_param == 3 && consts) RUN(MultiExpIrf<1, true>);
    if (model == 5 && n_param == 4 && consts) RUN(MultiExpIrf<2, false>);
    if (model == 5 && n_param == 5 && consts) RUN(MultiExpIrf<2, true>);
    if (model == 2 && n_param == 2) RUN(MultiExp<1, false>);
    if (model == 2 && n_param == 3) RUN(MultiExp<1, true>);
    if (model == 2 && n_param == 4) RUN(MultiExp<2, false>);
    if (model == 2 && n_param == 5) RUN(MultiExp<2, true>);
    if (model == 2 && n_param == 6) RUN(MultiExp<3, false>);
    if (model == 2 && n_param == 7) RUN(MultiExp<3, true>);
    return -1;
}
