// lm_launch.cuh -- kernel template and launcher shared by the per-model translation units.
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
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>

#include "lm_core.cuh"

#ifndef DPH_WARPS
#define DPH_WARPS 4  // warps per CTA
#endif
#ifndef DPH_MINB
#define DPH_MINB 3   // CTAs per SM the register allocation is bounded for (measured: spills cost more than occupancy gains)
#endif

namespace dphlm {

int fail(const std::string& m);  // sets the thread-local error text, returns -1
#define CUDA_TRY(x)                                                                                    \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) return ::dphlm::fail(std::string(#x) + ": " + cudaGetErrorString(e_));  \
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

constexpr int kWarps = DPH_WARPS;

template <class Mo, bool MLE, int G>
__global__ void __launch_bounds__(kWarps * 32, DPH_MINB) dphlm_fit_kernel(const KArgs a) {
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

template <class Mo, bool MLE, int G>
int launch(const KArgs& a0, DeviceCtx* ctx, cudaStream_t stream) {
    KArgs a = a0;
    const int M = a.h * a.w;
    constexpr int FPW = 32 / G;
    const size_t smem = sizeof(float) * (2 * (size_t)M + (size_t)kWarps * FPW * M * (1 + 2 * Mo::NE));
    auto kern = dphlm_fit_kernel<Mo, MLE, G>;
    if (smem > 227 * 1024) return fail("ROI too large for the shared-memory layout of this group size");
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
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


// one per model family, defined in lm_inst_*.cu
int launch_gauss2d(int G, bool mle, const KArgs& a, DeviceCtx* ctx, cudaStream_t s);
int launch_gauss2d_elliptical(int G, bool mle, const KArgs& a, DeviceCtx* ctx, cudaStream_t s);
int launch_multi_exp1(bool offset, int G, bool mle, const KArgs& a, DeviceCtx* ctx, cudaStream_t s);
int launch_multi_exp2(bool offset, int G, bool mle, const KArgs& a, DeviceCtx* ctx, cudaStream_t s);
int launch_multi_exp3(bool offset, int G, bool mle, const KArgs& a, DeviceCtx* ctx, cudaStream_t s);

#define DPH_DISPATCH_G(MO, G, mle, a, ctx, s)                                                      \
    switch (G) {                                                                                   \
        case 4: return (mle) ? launch<MO, true, 4>(a, ctx, s) : launch<MO, false, 4>(a, ctx, s);     \
        case 8: return (mle) ? launch<MO, true, 8>(a, ctx, s) : launch<MO, false, 8>(a, ctx, s);     \
        case 16: return (mle) ? launch<MO, true, 16>(a, ctx, s) : launch<MO, false, 16>(a, ctx, s);  \
        case 32: return (mle) ? launch<MO, true, 32>(a, ctx, s) : launch<MO, false, 32>(a, ctx, s);  \
    }                                                                                              \
    return ::dphlm::fail("group_lanes must be 0, 4, 8, 16 or 32")

}  // namespace dphlm
