// lm_kernels.cu -- sm_100a kernels and the C ABI (include/dphlm.h) around lm_core.cuh.
//
// Execution model.  One fit is owned by a group of G lanes of a warp (G = 4, 8, 16 or 32, so a warp
// carries 32/G fits).  Warps are persistent: each one pulls the next batch of 32/G consecutive fits from
// a global atomic counter, stages their ROIs into its private slice of shared memory with coalesced
// loads (clamping to >= 0 for the MLE estimator, lm.py:127, and checking finiteness, lm.py:651-652),
// runs both stages of the fit (lm_core.cuh: fit_one) and writes popt / info / nfev / chisq.  There is no
// CTA-level synchronisation after the coordinate table is built: warps never wait for each other, which
// is what absorbs the 6..30+ spread in iteration counts between fits.
//
// Shared memory per warp: (32/G) * M * (1 + 2*NE) floats -- the data y and the double-buffered stored
// exponentials of the accepted and the trial point (accepting a step flips a buffer index).
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/dphlm.h"
#include "lm_core.cuh"

using namespace dphlm;

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }
#define CUDA_TRY(x)                                                                                    \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_));           \
    } while (0)

struct KArgs {
    const float* data;
    const float* xaxis;
    const float* p0;
    float* popt;
    int* info;
    int* nfev;
    float* chisq;
    float* p_ls;
    int* counts;
    unsigned long long* counter;
    long long n_fits;
    int h, w;
    Options opt;
};

constexpr int kWarps = 8;  // warps per CTA

template <class Mo, bool MLE, int G>
__global__ void __launch_bounds__(kWarps * 32) dphlm_fit_kernel(const KArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int FPW = 32 / G;  // fits per warp
    constexpr int P = Mo::P;
    const int M = a.h * a.w;
    Coord* cxy = reinterpret_cast<Coord*>(smem);
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        Coord c;
        if (a.xaxis) { c.x = a.xaxis[i]; c.y = 0.0f; }
        else { c.x = float(i % a.w); c.y = float(i / a.w); }
        cxy[i] = c;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G;
    float* wy = smem + 2 * M + warp * (FPW * M * (1 + 2 * Mo::NE));
    float* we = wy + FPW * M;
    Lanes<G> ln{lane % G};

    for (;;) {
        unsigned long long batch = 0;
        if (lane == 0) batch = atomicAdd(a.counter, 1ull);
        batch = __shfl_sync(0xffffffffu, batch, 0);
        const long long f0 = (long long)batch * FPW;
        if (f0 >= a.n_fits) break;
        const long long left = a.n_fits - f0;
        const int nload = int(left < FPW ? left : FPW) * M;
        const float* src = a.data + f0 * M;
        unsigned badmask = 0;
        for (int i = lane; i < FPW * M; i += 32) {
            float v = i < nload ? __ldg(src + i) : 0.0f;
            if (!f_finite(v)) badmask |= 1u << (i / M);
            wy[i] = MLE ? clamp_nonneg(v) : v;
        }
        badmask = __reduce_or_sync(0xffffffffu, badmask);
        __syncwarp();

        const long long fit = f0 + grp;
        const bool active = fit < a.n_fits;
        const bool bad = (badmask >> grp) & 1u;
        float p0[P];
#pragma unroll
        for (int k = 0; k < P; ++k) p0[k] = active ? __ldg(a.p0 + fit * P + k) : 1.0f;
        PixelState<Mo::NE> px{wy + grp * M, we + grp * (2 * Mo::NE * M), cxy, M, 0};
        FitResult<P> r;
        fit_one<Mo, MLE, G>(ln, px, p0, a.opt, active && !bad, r);
        if (active && ln.lane == 0) {
            if (bad) {
                r.info = -100;
#pragma unroll
                for (int k = 0; k < P; ++k) r.x[k] = r.p_ls[k] = p0[k];
            }
#pragma unroll
            for (int k = 0; k < P; ++k) a.popt[fit * P + k] = r.x[k];
            a.info[fit] = r.info;
            a.nfev[fit] = r.nfev;
            a.chisq[fit] = r.chisq;
            if (a.p_ls) {
#pragma unroll
                for (int k = 0; k < P; ++k) a.p_ls[fit * P + k] = r.p_ls[k];
            }
            if (a.counts) {
                reinterpret_cast<int4*>(a.counts)[fit] = make_int4(r.nfev_ls, r.n_acc, r.n_rej, 0);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ launch
struct DeviceCtx {
    std::mutex mu;
    unsigned long long* counters = nullptr;
    unsigned next = 0;
    int sms = 0;
    // host-path staging buffers, one set per pipeline slot
    struct Slot {
        cudaStream_t stream = nullptr;
        char* buf = nullptr;
        size_t cap = 0;
    } slots[3];
};
constexpr int kCounterRing = 1024;
DeviceCtx g_ctx[64];

int get_ctx(int dev, DeviceCtx** out) {
    if (dev < 0 || dev >= 64) return fail("device index out of range");
    DeviceCtx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.counters) {
        CUDA_TRY(cudaMalloc(&c.counters, kCounterRing * sizeof(unsigned long long)));
        CUDA_TRY(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
    }
    *out = &c;
    return 0;
}

template <class Mo, bool MLE, int G>
int launch(const KArgs& a0, DeviceCtx* ctx, cudaStream_t stream) {
    KArgs a = a0;
    const int M = a.h * a.w;
    constexpr int FPW = 32 / G;
    const size_t smem = sizeof(float) * (2 * (size_t)M + (size_t)kWarps * FPW * M * (1 + 2 * Mo::NE));
    auto kern = dphlm_fit_kernel<Mo, MLE, G>;
    if (smem > 227 * 1024) return fail("ROI too large for the shared-memory layout of this group size");
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarps * 32, smem));
    if (occ < 1) return fail("kernel does not fit on an SM");
    unsigned slot;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        slot = ctx->next++ % kCounterRing;
    }
    a.counter = ctx->counters + slot;
    CUDA_TRY(cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), stream));
    const long long batches = (a.n_fits + FPW - 1) / FPW;
    long long grid = (long long)ctx->sms * occ;
    const long long need = (batches + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kWarps * 32, smem, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int auto_group(int M) {
    if (M <= 64) return 8;
    if (M <= 160) return 16;
    return 32;
}

template <class Mo>
int dispatch_g(const dphlm_desc* d, const KArgs& a, DeviceCtx* ctx, cudaStream_t s) {
    int G = d->group_lanes ? d->group_lanes : auto_group(a.h * a.w);
    const bool mle = d->method == DPHLM_MLE;
#define DPH_CASE(GG)                                                                  \
    case GG:                                                                          \
        return mle ? launch<Mo, true, GG>(a, ctx, s) : launch<Mo, false, GG>(a, ctx, s);
    switch (G) {
        DPH_CASE(4)
        DPH_CASE(8)
        DPH_CASE(16)
        DPH_CASE(32)
    }
#undef DPH_CASE
    return fail("group_lanes must be 0, 4, 8, 16 or 32");
}

int validate(const dphlm_desc* d) {
    if (!d) return fail("null descriptor");
    if (d->n_fits < 0) return fail("n_fits < 0");
    if (d->dim0 < 1 || d->dim1 < 1) return fail("dim0/dim1 must be >= 1");
    if (d->method != DPHLM_LS && d->method != DPHLM_MLE) return fail("Method not recognized");
    if (d->n_fits > 0 && (!d->data || !d->p0 || !d->popt || !d->info || !d->nfev || !d->chisq))
        return fail("null data/p0/popt/info/nfev/chisq pointer");
    switch (d->model) {
        case DPHLM_GAUSS2D:
            if (d->n_param != 5) return fail("gauss2d takes 5 parameters");
            break;
        case DPHLM_GAUSS2D_ELLIPTICAL:
            if (d->n_param != 7) return fail("gauss2d_elliptical takes 7 parameters");
            break;
        case DPHLM_MULTI_EXP:
            if (d->n_param < 2 || d->n_param > 7) return fail("multi_exp supports 2..7 parameters (1-3 exponentials, optional offset)");
            if (!d->xaxis && d->n_fits > 0) return fail("multi_exp needs xaxis");
            if (d->dim1 != 1) return fail("multi_exp is 1-D: dim1 must be 1");
            break;
        default:
            return fail("unknown model id");
    }
    if (d->n_param > d->dim0 * d->dim1) return fail("Improper input: N must not exceed M");
    return 0;
}

int fit_device(const dphlm_desc* d, DeviceCtx* ctx, cudaStream_t s) {
    if (d->n_fits == 0) return 0;
    KArgs a{};
    a.data = d->data; a.xaxis = d->xaxis; a.p0 = d->p0;
    a.popt = d->popt; a.info = d->info; a.nfev = d->nfev; a.chisq = d->chisq;
    a.p_ls = d->p_ls; a.counts = d->counts;
    a.n_fits = d->n_fits; a.h = d->dim0; a.w = d->dim1;
    a.opt = Options{d->ftol, d->xtol, d->gtol, d->factor, d->maxfev};
    switch (d->model) {
        case DPHLM_GAUSS2D: return dispatch_g<Gauss2D>(d, a, ctx, s);
        case DPHLM_GAUSS2D_ELLIPTICAL: return dispatch_g<Gauss2DElliptical>(d, a, ctx, s);
        case DPHLM_MULTI_EXP:
            switch (d->n_param) {
                case 2: return dispatch_g<MultiExp<1, false>>(d, a, ctx, s);
                case 3: return dispatch_g<MultiExp<1, true>>(d, a, ctx, s);
                case 4: return dispatch_g<MultiExp<2, false>>(d, a, ctx, s);
                case 5: return dispatch_g<MultiExp<2, true>>(d, a, ctx, s);
                case 6: return dispatch_g<MultiExp<3, false>>(d, a, ctx, s);
                case 7: return dispatch_g<MultiExp<3, true>>(d, a, ctx, s);
            }
    }
    return fail("unsupported model / n_param");
}

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace

extern "C" {

void dphlm_desc_init(dphlm_desc* d) {
    if (!d) return;
    *d = dphlm_desc{};
    d->ftol = 1.49012e-8f;
    d->xtol = 1.49012e-8f;
    d->gtol = 0.0f;
    d->factor = 100.0f;
    d->maxfev = 0;
}

int dphlm_fit_batch(const dphlm_desc* d, void* stream) {
    if (validate(d)) return -1;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DeviceCtx* ctx;
    if (get_ctx(dev, &ctx)) return -1;
    return fit_device(d, ctx, static_cast<cudaStream_t>(stream));
}

int dphlm_fit_batch_host(const dphlm_desc* d, int device) {
    if (validate(d)) return -1;
    if (d->n_fits == 0) return 0;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(device));
    DeviceCtx* ctx;
    if (get_ctx(device, &ctx)) return -1;
    const size_t M = (size_t)d->dim0 * d->dim1, P = d->n_param;
    // chunking: >= 4 chunks for overlap once the batch is large, each at most 256k fits
    long long chunk = d->n_fits / 8;
    if (chunk < 32768) chunk = 32768;
    if (chunk > 262144) chunk = 262144;
    const size_t b_data = align_up(chunk * M * 4), b_par = align_up(chunk * P * 4), b_i = align_up(chunk * 4),
                 b_cnt = align_up(chunk * 16), b_x = align_up(M * 4);
    const size_t need = b_data + 3 * b_par + 3 * b_i + b_cnt + b_x;
    int rc = 0;
    // one host call at a time per device: the staging slots are shared
    static std::mutex host_mu[64];
    std::lock_guard<std::mutex> lk(host_mu[device]);
    for (auto& s : ctx->slots) {
        if (!s.stream) CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        if (s.cap < need) {
            if (s.buf) CUDA_TRY(cudaFree(s.buf));
            s.buf = nullptr; s.cap = 0;
            CUDA_TRY(cudaMalloc(&s.buf, need));
            s.cap = need;
        }
    }
    int k = 0;
    for (long long lo = 0; lo < d->n_fits && rc == 0; lo += chunk, k = (k + 1) % 3) {
        const long long n = d->n_fits - lo < chunk ? d->n_fits - lo : chunk;
        auto& s = ctx->slots[k];
        char* b = s.buf;
        dphlm_desc c = *d;
        c.n_fits = n;
        float* ddata = (float*)b; b += b_data;
        float* dp0 = (float*)b; b += b_par;
        float* dpopt = (float*)b; b += b_par;
        float* dpls = (float*)b; b += b_par;
        int* dinfo = (int*)b; b += b_i;
        int* dnfev = (int*)b; b += b_i;
        float* dchi = (float*)b; b += b_i;
        int* dcnt = (int*)b; b += b_cnt;
        float* dx = (float*)b;
        CUDA_TRY(cudaMemcpyAsync(ddata, d->data + lo * M, n * M * 4, cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(dp0, d->p0 + lo * P, n * P * 4, cudaMemcpyHostToDevice, s.stream));
        if (d->xaxis) CUDA_TRY(cudaMemcpyAsync(dx, d->xaxis, M * 4, cudaMemcpyHostToDevice, s.stream));
        c.data = ddata; c.p0 = dp0; c.xaxis = d->xaxis ? dx : nullptr;
        c.popt = dpopt; c.info = dinfo; c.nfev = dnfev; c.chisq = dchi;
        c.p_ls = d->p_ls ? dpls : nullptr;
        c.counts = d->counts ? dcnt : nullptr;
        rc = fit_device(&c, ctx, s.stream);
        if (rc) break;
        CUDA_TRY(cudaMemcpyAsync(d->popt + lo * P, dpopt, n * P * 4, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaMemcpyAsync(d->info + lo, dinfo, n * 4, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaMemcpyAsync(d->nfev + lo, dnfev, n * 4, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaMemcpyAsync(d->chisq + lo, dchi, n * 4, cudaMemcpyDeviceToHost, s.stream));
        if (d->p_ls) CUDA_TRY(cudaMemcpyAsync(d->p_ls + lo * P, dpls, n * P * 4, cudaMemcpyDeviceToHost, s.stream));
        if (d->counts) CUDA_TRY(cudaMemcpyAsync(d->counts + lo * 4, dcnt, n * 16, cudaMemcpyDeviceToHost, s.stream));
    }
    for (auto& s : ctx->slots) {
        cudaError_t e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess && rc == 0) rc = fail(std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    return rc;
}

void* dphlm_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        g_err = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}

void dphlm_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int dphlm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { g_err = "cudaGetDeviceCount failed"; return -1; }
    return n;
}

int dphlm_version(void) { return DPHLM_VERSION; }

const char* dphlm_last_error(void) { return g_err.c_str(); }

}  // extern "C"
