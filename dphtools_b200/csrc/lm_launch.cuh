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
    float* p_ls;   // never null inside the kernels (user buffer or workspace)
    int* ier;      // workspace [n_fits]: exit code of the pre-fit (100 = non-finite input data)
    int* counts;   // optional
    unsigned long long* counter;
    long long n_fits;
    int h, w;
    Options opt;
};

constexpr int kWarps = DPH_WARPS;

// per-fit stride of the shared-memory arrays: >= M and congruent to G modulo 32, so that the G-lane
// groups of a warp touch disjoint banks
__host__ __device__ inline int padded_stride(int M, int G) {
    int s = M + ((G % 32) - (M % 32) + 32) % 32;
    return s;
}

#if defined(__CUDACC__)
// Fit supply / result sink of the pre-fit kernel.
template <class Mo, int G>
struct IoPrefit {
    const KArgs& a;
    __device__ bool fetch(const Lanes<G>& ln, PixelState<Mo::NE>& px, float* pstart, long long& fit) const {
        constexpr int P = Mo::P;
        const int src = __ffs(ln.gmask) - 1;
        for (;;) {
            unsigned long long idx = 0;
            if (ln.lane == 0) idx = atomicAdd(a.counter, 1ull);
            idx = __shfl_sync(ln.gmask, idx, src);
            if ((long long)idx >= a.n_fits) return false;
            const float* d = a.data + idx * px.M;
            bool bad = false;
            for (int i = ln.lane; i < px.M; i += G) {
                float v = __ldg(d + i);
                bad |= !f_finite(v);
                px.y[i] = v;
            }
#pragma unroll
            for (int k = 0; k < P; ++k) pstart[k] = __ldg(a.p0 + idx * P + k);
            bad = __any_sync(ln.gmask, bad);
            __syncwarp(ln.gmask);
            if (!bad) { fit = (long long)idx; return true; }
            if (ln.lane == 0) {  // the reference raises ValueError (lm.py:651-652)
                a.ier[idx] = 100;
#pragma unroll
                for (int k = 0; k < P; ++k) a.p_ls[idx * P + k] = pstart[k];
                if (a.counts) reinterpret_cast<int4*>(a.counts)[idx] = make_int4(0, 0, 0, 0);
            }
        }
    }
    __device__ void store(const Lanes<G>& ln, const StageA<Mo>& st, long long fit) const {
        if (ln.lane != 0) return;
#pragma unroll
        for (int k = 0; k < Mo::P; ++k) a.p_ls[fit * Mo::P + k] = st.x[k];
        a.ier[fit] = st.ier;
        if (a.counts) a.counts[4 * fit] = st.nfev;
    }
};

// Fit supply / result sink of the main (lm) kernel: starts from the pre-fit results.
template <class Mo, bool MLE, int G>
struct IoMain {
    const KArgs& a;
    __device__ bool fetch(const Lanes<G>& ln, PixelState<Mo::NE>& px, float* pstart, long long& fit) const {
        constexpr int P = Mo::P;
        const int src = __ffs(ln.gmask) - 1;
        for (;;) {
            unsigned long long idx = 0;
            if (ln.lane == 0) idx = atomicAdd(a.counter, 1ull);
            idx = __shfl_sync(ln.gmask, idx, src);
            if ((long long)idx >= a.n_fits) return false;
            const int ier = __ldg(a.ier + idx);
#pragma unroll
            for (int k = 0; k < P; ++k) pstart[k] = a.p_ls[idx * P + k];
            if (ier >= 1 && ier <= 4) {
                const float* d = a.data + idx * px.M;
                for (int i = ln.lane; i < px.M; i += G) {
                    float v = __ldg(d + i);
                    px.y[i] = MLE ? clamp_nonneg(v) : v;  // lm.py:127
                }
                __syncwarp(ln.gmask);
                fit = (long long)idx;
                return true;
            }
            if (ln.lane == 0) {  // the pre-fit failed: report it, do not fit (scipy raises, lm.py:642-644)
#pragma unroll
                for (int k = 0; k < P; ++k) a.popt[idx * P + k] = pstart[k];
                a.info[idx] = ier == 0 ? -99 : -ier;
                a.nfev[idx] = 0;
                a.chisq[idx] = 0.0f;
                if (a.counts) { a.counts[4 * idx + 1] = 0; a.counts[4 * idx + 2] = 0; a.counts[4 * idx + 3] = 0; }
            }
        }
    }
    __device__ void store(const Lanes<G>& ln, const StageB<Mo, MLE>& st, long long fit) const {
        if (ln.lane != 0) return;
#pragma unroll
        for (int k = 0; k < Mo::P; ++k) a.popt[fit * Mo::P + k] = st.xret[k];
        a.info[fit] = st.info;
        a.nfev[fit] = st.ev;
        a.chisq[fit] = st.chisq_ret;
        if (a.counts) { a.counts[4 * fit + 1] = st.n_acc; a.counts[4 * fit + 2] = st.n_rej; a.counts[4 * fit + 3] = 0; }
    }
};

template <class Stage, class Mo, int G, class IO>
__global__ void __launch_bounds__(kWarps * 32, DPH_MINB) dphlm_stage_kernel(const KArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int FPW = 32 / G;  // fits in flight per warp
    const int M = a.h * a.w;
    const int stride = padded_stride(M, G);
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
    const int region = stride * (1 + 2 * Mo::NE);
    float* base = smem + 2 * ((M + 1) & ~1) + (warp * FPW + grp) * region;
    Lanes<G> ln{lane % G, G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (grp * G))};
    PixelState<Mo::NE> px{base, base + stride, cxy, M, stride, 0};
    IO io{a};
    run_stage<Stage, Mo, G>(ln, px, a.opt, io);
}
#endif

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

template <class Kern>
int launch_stage(Kern kern, const KArgs& a0, int fpw, size_t smem, DeviceCtx* ctx, cudaStream_t stream) {
    KArgs a = a0;
    if (smem > 227 * 1024) return ::dphlm::fail("ROI too large for the shared-memory layout of this group size");
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarps * 32, smem));
    if (occ < 1) return ::dphlm::fail("kernel does not fit on an SM");
    unsigned slot;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        slot = ctx->next++ % kCounterRing;
    }
    a.counter = ctx->counters + slot;
    CUDA_TRY(cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), stream));
    long long grid = (long long)ctx->sms * occ;
    const long long need = (a.n_fits + (long long)kWarps * fpw - 1) / ((long long)kWarps * fpw);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kWarps * 32, smem, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// both stages of curve_fit(method='ls'|'mle') for one batch: pre-fit kernel, then the lm() kernel
template <class Mo, bool MLE, int G>
int launch(const KArgs& a, DeviceCtx* ctx, cudaStream_t stream) {
    const int M = a.h * a.w;
    constexpr int FPW = 32 / G;
    const size_t smem = sizeof(float) * (2 * (size_t)((M + 1) & ~1) + (size_t)kWarps * FPW * padded_stride(M, G) * (1 + 2 * Mo::NE));
    int rc = launch_stage(dphlm_stage_kernel<StageA<Mo>, Mo, G, IoPrefit<Mo, G>>, a, FPW, smem, ctx, stream);
    if (rc) return rc;
    return launch_stage(dphlm_stage_kernel<StageB<Mo, MLE>, Mo, G, IoMain<Mo, MLE, G>>, a, FPW, smem, ctx, stream);
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
