// lm_kernels.cu -- the C ABI (include/dphlm.h): validation, dispatch, the host-buffer pipeline.
// The kernels themselves are lm_launch.cuh (template) instantiated per model in lm_inst_*.cu.
#include <cuda_runtime.h>

#include <cstdlib>
#include <algorithm>
#include <mutex>
#include <vector>
#include <string>

#include "../../include/dphlm.h"
#include "lm_launch.cuh"

using namespace dphlm;

namespace dphlm {
thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }
}  // namespace dphlm

namespace {

DeviceCtx g_ctx[64];

struct TimingEvents {
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool valid = false;
};
thread_local TimingEvents g_timing;

int get_ctx(int dev, DeviceCtx** out) {
    if (dev < 0 || dev >= 64) return fail("device index out of range");
    DeviceCtx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.sms) {
        CUDA_TRY(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
        cudaMemPool_t pool;
        unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        CUDA_TRY(cudaMalloc(&c.err_word, sizeof(int)));
        CUDA_TRY(cudaMemset(c.err_word, 0, sizeof(int)));
    }
    *out = &c;
    return 0;
}

// number of floats `consts` holds for a model (host path: `consts` is a host pointer)
constexpr int kMaxConsts = 3 + 32;
int model_consts(int model, const float* consts) {
    if (model == DPHLM_GAUSS2D_INTEGRATED_FIXED || model == DPHLM_GAUSS2D_FIXED) return 1;
    if (model == DPHLM_MULTI_EXP_IRF && consts) {
        const int w = (int)consts[0];
        return 3 + (w < 1 ? 1 : w > 32 ? 32 : w);
    }
    return 0;
}

int auto_group(int M, int model, long long n_fits, bool tail_hidden) {
    // measured on B200 (profiles/): 7x7 -> 2 lanes; 11x11 -> 4, and 2 for the sampled Gaussian from 3e5 fits per call on (5 %
    // faster there: the per-trip scalar work is shared by 16 fits; equal at 1e5, where the longer per-fit latency shows in the
    // tails); 15x15 -> 4 for the sampled and the elliptical Gaussian, which keep no per-pixel state in shared memory, 8 otherwise
    if (model == DPHLM_MULTI_EXP_IRF) return M <= 512 ? 16 : 32;  // built for wide groups only (lm_inst_mexp_irf.cu)
    // multi-exponentials keep no per-bin state in shared memory: the narrowest group whose two staging buffers still leave room
    // for three CTAs per SM (measured on 256 bins, bi-exponential + offset: 5.65e7 fits/s at 4 lanes, 5.18e7 at 8, 4.27e7 at
    // 16, 3.4e7 at 2)
    if (model == DPHLM_MULTI_EXP) return M <= 64 ? 2 : M <= 276 ? 4 : M <= 576 ? 8 : M <= 1176 ? 16 : 32;
    if (M <= 64) return 2;
    // (pipelined calls hide the tail under the next call: 2 lanes win from 1e5 fits on, 1.12e8 against 1.04e8 fits/s)
    if (M <= 128) return (model == DPHLM_GAUSS2D || model == DPHLM_GAUSS2D_FIXED) && (n_fits >= 300000 || (tail_hidden && n_fits >= 50000)) ? 2 : 4;
    if (M <= 240) return (model == DPHLM_GAUSS2D || model == DPHLM_GAUSS2D_ELLIPTICAL || model == DPHLM_GAUSS2D_FIXED) ? 4 : 8;  // (elliptical, 15x15: 3.24e7 fits/s at 4 lanes, 2.57e7 at 8)
    return 16;
}

int validate_shape(const dphlm_desc* d) {
    if (!d) return fail("null descriptor");
    if (d->n_fits < 0) return fail("n_fits < 0");
    if (d->n_fits > 2000000000LL) return fail("n_fits above 2e9 per call: split the batch");  // straggler lists hold 32-bit fit indices
    if (d->dim0 < 1 || d->dim1 < 1) return fail("dim0/dim1 must be >= 1");
    switch (d->model) {
        case DPHLM_GAUSS2D:
            if (d->n_param != 5) return fail("gauss2d takes 5 parameters");
            break;
        case DPHLM_GAUSS2D_ELLIPTICAL:
            if (d->n_param != 7) return fail("gauss2d_elliptical takes 7 parameters");
            break;
        case DPHLM_GAUSS2D_INTEGRATED:
            if (d->n_param != 5) return fail("gauss2d_integrated takes 5 parameters");
            if (d->dim0 > 32 || d->dim1 > 32) return fail("gauss2d_integrated: ROI sides up to 32 pixels");
            break;
        case DPHLM_GAUSS2D_INTEGRATED_FIXED:
            if (d->n_param != 4) return fail("gauss2d_integrated_fixed takes 4 parameters");
            if (d->dim0 > 32 || d->dim1 > 32) return fail("gauss2d_integrated_fixed: ROI sides up to 32 pixels");
            if (!d->consts && d->n_fits > 0) return fail("gauss2d_integrated_fixed needs consts[0] = sigma");
            break;
        case DPHLM_GAUSS2D_FIXED:
            if (d->n_param != 4) return fail("gauss2d_fixed takes 4 parameters");
            if (!d->consts && d->n_fits > 0) return fail("gauss2d_fixed needs consts[0] = sigma");
            break;
        case DPHLM_MULTI_EXP_IRF:
            if (d->n_param < 2 || d->n_param > 5) return fail("multi_exp_irf supports 2..5 parameters (1-2 exponentials, optional offset)");
            if (!d->xaxis && d->n_fits > 0) return fail("multi_exp_irf needs xaxis");
            if (!d->consts && d->n_fits > 0) return fail("multi_exp_irf needs consts = [W, m0, dt, irf[0..W-1]]");
            if (d->dim1 != 1) return fail("multi_exp_irf is 1-D: dim1 must be 1");
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

int validate(const dphlm_desc* d) {
    if (!d) return fail("null descriptor");
    if (d->n_fits < 0) return fail("n_fits < 0");
    if (d->n_fits > 2000000000LL) return fail("n_fits above 2e9 per call: split the batch");  // straggler lists hold 32-bit fit indices
    if (d->dim0 < 1 || d->dim1 < 1) return fail("dim0/dim1 must be >= 1");
    if (d->method != DPHLM_LS && d->method != DPHLM_MLE) return fail("Method not recognized");
    if (d->n_fits > 0 && (!d->data || !d->p0 || !d->popt || !d->info || !d->nfev || !d->chisq))
        return fail("null data/p0/popt/info/nfev/chisq pointer");
    return validate_shape(d);
}

// `ws` (optional): caller-provided device workspace of at least workspace_bytes(d)
// capacity of each straggler list.  With DPHLM_FLAG_EXACT_TIES up to a third of the fits pass through the lm() list
// (lm_core.cuh: tie_break), hence n / 2 there.
int park_capacity(long long n, bool exact) {
    long long c = exact ? n / 2 : n / 16;
    if (c < 256) c = 256;
    if (c > (exact ? 8388608 : 65536)) c = exact ? 8388608 : 65536;
    return (int)c;
}
size_t park_ready_bytes(long long n, bool exact) { return ((3 * (size_t)park_capacity(n, exact) * sizeof(int)) + 255) & ~size_t(255); }
size_t park_bytes(long long n, bool exact) {  // header (sync words + work counters), three ready arrays, three record arrays
    return 256 + park_ready_bytes(n, exact) + 3 * (size_t)park_capacity(n, exact) * kParkRecord * sizeof(float);
}
size_t workspace_bytes(const dphlm_desc* d) {
    return (((size_t)d->n_fits * sizeof(int) + 255) & ~size_t(255)) + park_bytes(d->n_fits, d->flags & DPHLM_FLAG_EXACT_TIES) +
           (d->p_ls ? 0 : (size_t)d->n_fits * d->n_param * sizeof(float));
}

__global__ void widen_u16(const unsigned short* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = float(in[i]);
}

int widen_async(const void* in, float* out, long long n, cudaStream_t s) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    widen_u16<<<(unsigned)blocks, 256, 0, s>>>(static_cast<const unsigned short*>(in), out, n);
    CUDA_TRY(cudaGetLastError());
    return 0;
}


// One launch prepares a call: the pre-fit exit codes (0 = "not published yet" for an lm() kernel that starts early; 1 when the
// pre-fit is skipped), the header and ready flags of the straggler lists (0), the pre-fit parameters (NaN = not published)
// and every fit's `info` (DPHLM_INFO_UNFINISHED, so that a fit nobody finished cannot be mistaken for a result).
__global__ void dphlm_init_kernel(int* ier, int ier_value, long long n_fits, int* zero, long long n_zero, unsigned* p_ls, long long n_pls,
                                  int* info) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = t; i < n_fits; i += stride) { ier[i] = ier_value; info[i] = DPHLM_INFO_UNFINISHED; }
    for (long long i = t; i < n_zero; i += stride) zero[i] = 0;
    for (long long i = t; i < n_pls; i += stride) p_ls[i] = 0xffffffffu;
}

int launch_model(const dphlm_desc* d, int G, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s) {
    const bool mle = d->method == DPHLM_MLE;
    switch (d->model) {
        case DPHLM_GAUSS2D: return launch_gauss2d(G, mle, a, pl, ctx, s);
        case DPHLM_GAUSS2D_ELLIPTICAL: return launch_gauss2d_elliptical(G, mle, a, pl, ctx, s);
        case DPHLM_GAUSS2D_INTEGRATED: return launch_gauss2d_integrated(G, mle, a, pl, ctx, s);
        case DPHLM_GAUSS2D_INTEGRATED_FIXED: return launch_gauss2d_integrated_fixed(G, mle, a, pl, ctx, s);
        case DPHLM_MULTI_EXP_IRF: return launch_multi_exp_irf(d->n_param, G, mle, a, pl, ctx, s);
        case DPHLM_GAUSS2D_FIXED: return launch_gauss2d_fixed(G, mle, a, pl, ctx, s);
        case DPHLM_MULTI_EXP:
            switch (d->n_param / 2) {
                case 1: return launch_multi_exp1(d->n_param & 1, G, mle, a, pl, ctx, s);
                case 2: return launch_multi_exp2(d->n_param & 1, G, mle, a, pl, ctx, s);
                case 3: return launch_multi_exp3(d->n_param & 1, G, mle, a, pl, ctx, s);
            }
    }
    return fail("unsupported model / n_param");
}

// Enqueue the kernels for a batch whose buffers are on the device.
int fit_device(const dphlm_desc* d, DeviceCtx* ctx, cudaStream_t s, char* ws_in = nullptr) {
    if (d->n_fits == 0) return 0;
    KArgs a{};
    a.data = d->data; a.xaxis = d->xaxis; a.consts = d->consts; a.p0 = d->p0;
    a.popt = d->popt; a.info = d->info; a.nfev = d->nfev; a.chisq = d->chisq;
    a.p_ls = d->p_ls; a.counts = d->counts; a.x_acc = d->x_acc;
    a.n_fits = d->n_fits; a.h = d->dim0; a.w = d->dim1;
    a.opt = Options{d->ftol, d->xtol, d->gtol, d->factor, d->maxfev, (d->flags & DPHLM_FLAG_EXACT_TIES) ? 1 : 0};
    if (d->flags & DPHLM_FLAG_TIMING) {
        for (auto& e : g_timing.ev)
            if (!e) CUDA_TRY(cudaEventCreate(&e));
        a.ev = g_timing.ev;
        g_timing.valid = true;
    }
    const int G = d->group_lanes ? d->group_lanes : auto_group(a.h * a.w, d->model, d->n_fits, d->flags & DPHLM_FLAG_PIPELINED);
    // workspace: pre-fit exit codes, and the pre-fit parameters when the caller does not want them
    const size_t b_ier = ((size_t)d->n_fits * sizeof(int) + 255) & ~size_t(255);
    const bool exact = d->flags & DPHLM_FLAG_EXACT_TIES;
    const size_t b_park = park_bytes(d->n_fits, exact);
    const size_t b_pls = d->p_ls ? 0 : (size_t)d->n_fits * d->n_param * sizeof(float);
    const bool u16 = (d->flags & DPHLM_FLAG_DATA_U16) && !ws_in;  // the host path widens into its own slot buffer
    const size_t b_wide = u16 ? (size_t)d->n_fits * a.h * a.w * sizeof(float) : 0;
    char* ws = ws_in;
    if (!ws) CUDA_TRY(cudaMallocAsync(&ws, b_ier + b_park + ((b_pls + 255) & ~size_t(255)) + b_wide, s));
    if (u16) {
        float* wide = reinterpret_cast<float*>(ws + b_ier + b_park + ((b_pls + 255) & ~size_t(255)));
        if (widen_async(d->data, wide, d->n_fits * a.h * a.w, s)) return -1;
        a.data = wide;
    }
    a.ier = reinterpret_cast<int*>(ws);
    if (!d->p_ls) a.p_ls = reinterpret_cast<float*>(ws + b_ier + b_park);
    // 256-byte header (sync words, then the work counters of the two main kernels at byte 64) and the straggler
    // lists; DPHLM_FLAG_NO_HANDOFF keeps every fit in the group that started it
    ParkLists pl{};
    pl.cap = (d->flags & DPHLM_FLAG_NO_HANDOFF) ? 0 : park_capacity(d->n_fits, exact);
    pl.sync = reinterpret_cast<int*>(ws + b_ier);
    a.counter = reinterpret_cast<unsigned long long*>(ws + b_ier + 64);
    pl.ready = reinterpret_cast<int*>(ws + b_ier + 256);
    pl.rec = reinterpret_cast<float*>(ws + b_ier + 256 + park_ready_bytes(d->n_fits, exact));
    // the pre-fit results start as "not published yet" (exit code 0, parameters NaN) for an lm() kernel that starts early
    const bool skip = d->flags & DPHLM_FLAG_SKIP_PREFIT;
    a.skip_prefit = skip ? 1 : 0;
    if (skip) a.p_ls = const_cast<float*>(d->p0);  // lm() alone starts from p0 (nothing writes p_ls then)
    {
        const long long n_zero = (long long)((256 + (pl.cap ? park_ready_bytes(d->n_fits, exact) : 0)) / sizeof(int));
        const long long n_pls = skip ? 0 : d->n_fits * d->n_param;
        long long blocks = (std::max<long long>(d->n_fits, n_pls) + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        dphlm_init_kernel<<<(unsigned)blocks, 256, 0, s>>>(a.ier, skip ? 1 : 0, d->n_fits, reinterpret_cast<int*>(ws + b_ier), n_zero,
                                                          reinterpret_cast<unsigned*>(skip ? nullptr : a.p_ls), n_pls, d->info);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { if (!ws_in) cudaFreeAsync(ws, s); return fail(std::string("init kernel: ") + cudaGetErrorString(e)); }
    }
    int rc;
    {
        std::lock_guard<std::mutex> lk(ctx->fork_mu);  // the side stream and its events are shared per device
        if (!ctx->sides[0].stream_a) {
            for (auto& sd : ctx->sides) {
                CUDA_TRY(cudaStreamCreateWithFlags(&sd.stream_a, cudaStreamNonBlocking));
                CUDA_TRY(cudaStreamCreateWithFlags(&sd.stream_b, cudaStreamNonBlocking));
                for (cudaEvent_t* e : {&sd.fork_ev, &sd.mid_ev, &sd.join_a, &sd.join_b}) CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
            }
        }
        rc = launch_model(d, G, a, pl, ctx, s);
        if (rc) {
            // a launch failed part-way: straggler kernels may already be running on the side streams and use the workspace.
            // Let them end (their producers never started, so they leave at once) before the workspace goes back to the pool.
            for (auto& sd : ctx->sides) { cudaStreamSynchronize(sd.stream_a); cudaStreamSynchronize(sd.stream_b); }
        }
    }
    if (!ws_in) cudaFreeAsync(ws, s);
    return rc;
}

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

// FP32 pipe microbenchmark: independent scalar FFMA chains (two invariant operands), the fastest fp32 instruction stream
// this part sustains (scripts/microbench/pipes.cu: 0.97 FFMA per clock and scheduler; packed FFMA2 issues every 2.2 clocks).
__global__ void __launch_bounds__(256) dphlm_ffma_chain_kernel(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = threadIdx.x * 1e-3f + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = __fmaf_rn(acc[k], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


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

int dphlm_fit_batch(const dphlm_desc* d_in, void* stream) {
    if (validate(d_in)) return -1;
    dphlm_desc tuned = *d_in;  // DPHLM_FLAGS (environment) ORs extra flags in: tuning aid
    if (const char* e = getenv("DPHLM_FLAGS")) tuned.flags |= atoi(e);
    const dphlm_desc* d = &tuned;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DeviceCtx* ctx;
    if (get_ctx(dev, &ctx)) return -1;
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    if (!(d->flags & DPHLM_FLAG_PIPELINED) || d->n_fits == 0) return fit_device(d, ctx, user);
    // Pipelined: the whole call goes to one of the library's own streams, behind what the caller's stream holds now; the
    // caller's stream does not wait (dphlm_join).  Consecutive calls alternate between the streams, so the drain and the
    // stragglers of one call run under the main kernels of the next.
    tuned.flags &= ~DPHLM_FLAG_TIMING;  // the timing events assume one call at a time on the caller's stream
    std::lock_guard<std::mutex> lk(ctx->pipe_mu);
    int depth = 3;  // measured on 1e5-fit calls: 1.06e8 (2), 1.14e8 (3), 1.13e8 (4) fits/s when one call in four holds a 600-evaluation straggler
    if (const char* e = getenv("DPHLM_PIPE_DEPTH")) depth = std::min(std::max(atoi(e), 1), (int)DeviceCtx::kPipeMax);  // tuning aid
    DeviceCtx::Pipe& pp = ctx->pipes[ctx->next_pipe++ % depth];
    if (!pp.stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&pp.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&pp.in_ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&pp.done_ev, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(pp.in_ev, user));
    CUDA_TRY(cudaStreamWaitEvent(pp.stream, pp.in_ev, 0));
    const int rc = fit_device(d, ctx, pp.stream);
    CUDA_TRY(cudaEventRecord(pp.done_ev, pp.stream));
    pp.pending = true;
    return rc;
}

int dphlm_join(void* stream) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DeviceCtx* ctx;
    if (get_ctx(dev, &ctx)) return -1;
    std::lock_guard<std::mutex> lk(ctx->pipe_mu);
    for (auto& pp : ctx->pipes) {
        if (!pp.pending) continue;
        CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), pp.done_ev, 0));
        pp.pending = false;
    }
    return 0;
}

int dphlm_call_status(int* tripped) {
    if (!tripped) return fail("null pointer");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DeviceCtx* ctx;
    if (get_ctx(dev, &ctx)) return -1;
    CUDA_TRY(cudaDeviceSynchronize());
    int word = 0;
    CUDA_TRY(cudaMemcpy(&word, ctx->err_word, sizeof(int), cudaMemcpyDeviceToHost));
    if (word) CUDA_TRY(cudaMemset(ctx->err_word, 0, sizeof(int)));
    *tripped = word ? 1 : 0;
    return 0;
}

static int host_enqueue(const dphlm_desc* d, int device, bool async) {
    if (validate(d)) return -1;
    if (d->n_fits == 0) return 0;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(device));
    DeviceCtx* ctx;
    if (get_ctx(device, &ctx)) return -1;
    const size_t M = (size_t)d->dim0 * d->dim1, P = d->n_param;
    // Chunking: the copies of one chunk overlap the kernels of the others (one stream per chunk, kHostSlots in flight).  The
    // kernels cannot start before the first chunk has landed and need longer per fit than the copy does, so the chunks
    // grow: n/10, 2n/10, 3n/10, 4n/10 for a batch up to ~6e5 fits, then the cap.  DPHLM_CHUNK = fixed size (tuning).
    // A call that is enqueued without waiting overlaps with its neighbours anyway (the copy of the next call runs under the
    // kernels of this one), and small chunks cost kernel efficiency: two halves then.  Measured at 1e5 fits per call over 60
    // calls: 1.05e8 / 1.03e8 / 0.995e8 fits/s for one / two / three chunks (53 GB/s of host-to-device copies: the PCIe link),
    // 0.9e8 for the growing schedule; over 10 calls, where filling and draining the pipeline shows, three chunks lead by 2 %.
    const long long kChunkMin = 8192, kChunkMax = 262144;
    long long step = d->n_fits / 10;
    if (step < kChunkMin) step = kChunkMin;
    if (step > kChunkMax) step = kChunkMax;
    long long fixed = 0;
    if (async) fixed = std::min<long long>(std::max<long long>((d->n_fits + 1) / 2, kChunkMin), kChunkMax);
    if (const char* e = getenv("DPHLM_CHUNK")) { long long v = atoll(e); if (v > 0) fixed = v; }
    std::vector<long long> bounds{0};
    for (long long c = fixed ? fixed : step; bounds.back() < d->n_fits;) {
        long long hi = bounds.back() + c;
        if (hi > d->n_fits || d->n_fits - hi < kChunkMin / 2) hi = d->n_fits;  // no crumb at the end
        bounds.push_back(hi);
        if (!fixed && c + step <= kChunkMax) c += step;
    }
    // the largest chunk either schedule would make of this batch sizes the slot buffers (so that a blocking call followed by an
    // asynchronous one of the same size does not reallocate them)
    long long chunk = std::min<long long>(std::max<long long>(std::max<long long>(4 * step, (d->n_fits + 1) / 2), kChunkMin), kChunkMax);
    for (size_t i = 1; i < bounds.size(); ++i) chunk = std::max(chunk, bounds[i] - bounds[i - 1]);
    const size_t b_data = align_up(chunk * M * 4), b_par = align_up(chunk * P * 4), b_i = align_up(chunk * 4),
                 b_cnt = align_up(chunk * 16), b_x = align_up(M * 4) + 256;  // xaxis + model constants
    const size_t b_ws = align_up(chunk * 4) + align_up(park_bytes(chunk, d->flags & DPHLM_FLAG_EXACT_TIES)) + align_up(chunk * P * 4);
    const bool u16 = d->flags & DPHLM_FLAG_DATA_U16;
    const size_t b_raw = u16 ? align_up(chunk * M * 2) : 0;
    const size_t need = b_data + b_raw + 4 * b_par + 3 * b_i + b_cnt + b_x + b_ws;
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
    int n_slots = kHostSlots;  // 8 against 4: consecutive calls land on different slots (+4 % end to end at 1e5 fits per call)
    if (const char* e = getenv("DPHLM_HOST_SLOTS")) n_slots = std::min(std::max(atoi(e), 1), (int)kHostSlots);  // tuning aid
    static unsigned next_slot[64];  // round-robin across calls too (under host_mu)
    int k = next_slot[device] % n_slots;
    for (size_t ci = 0; ci + 1 < bounds.size() && rc == 0; ++ci, k = (k + 1) % n_slots) {
        const long long lo = bounds[ci], n = bounds[ci + 1] - lo;
        auto& s = ctx->slots[k];
        char* b = s.buf;
        dphlm_desc c = *d;
        c.n_fits = n;
        float* ddata = (float*)b; b += b_data;
        char* draw = b; b += b_raw;
        float* dp0 = (float*)b; b += b_par;
        float* dpopt = (float*)b; b += b_par;
        float* dpls = (float*)b; b += b_par;
        float* dxacc = (float*)b; b += b_par;
        int* dinfo = (int*)b; b += b_i;
        int* dnfev = (int*)b; b += b_i;
        float* dchi = (float*)b; b += b_i;
        int* dcnt = (int*)b; b += b_cnt;
        float* dx = (float*)b; b += b_x;
        char* dws = b;
        if (u16) {
            CUDA_TRY(cudaMemcpyAsync(draw, reinterpret_cast<const unsigned short*>(d->data) + lo * M, n * M * 2, cudaMemcpyHostToDevice, s.stream));
            if (widen_async(draw, ddata, n * (long long)M, s.stream)) { rc = -1; break; }
        } else {
            CUDA_TRY(cudaMemcpyAsync(ddata, d->data + lo * M, n * M * 4, cudaMemcpyHostToDevice, s.stream));
        }
        CUDA_TRY(cudaMemcpyAsync(dp0, d->p0 + lo * P, n * P * 4, cudaMemcpyHostToDevice, s.stream));
        if (d->xaxis) CUDA_TRY(cudaMemcpyAsync(dx, d->xaxis, M * 4, cudaMemcpyHostToDevice, s.stream));
        float* dconsts = reinterpret_cast<float*>(reinterpret_cast<char*>(dx) + align_up(M * 4));
        const int n_consts = model_consts(d->model, d->consts);
        if (n_consts) CUDA_TRY(cudaMemcpyAsync(dconsts, d->consts, n_consts * 4, cudaMemcpyHostToDevice, s.stream));
        c.data = ddata; c.p0 = dp0; c.xaxis = d->xaxis ? dx : nullptr; c.consts = n_consts ? dconsts : nullptr;
        c.popt = dpopt; c.info = dinfo; c.nfev = dnfev; c.chisq = dchi;
        c.p_ls = dpls;  // always materialised in the slot buffer: the lm() kernel starts from it
        c.counts = d->counts ? dcnt : nullptr;
        c.x_acc = d->x_acc ? dxacc : nullptr;
        rc = fit_device(&c, ctx, s.stream, dws);
        if (rc) break;
        // the pollers' watchdog word (never set in correct operation) comes back with the results
        if (!s.err_word) CUDA_TRY(cudaHostAlloc(&s.err_word, sizeof(int), cudaHostAllocDefault));
        CUDA_TRY(cudaMemcpyAsync(s.err_word, ctx->err_word, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaMemcpyAsync(d->popt + lo * P, dpopt, n * P * 4, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaMemcpyAsync(d->info + lo, dinfo, n * 4, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaMemcpyAsync(d->nfev + lo, dnfev, n * 4, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaMemcpyAsync(d->chisq + lo, dchi, n * 4, cudaMemcpyDeviceToHost, s.stream));
        if (d->p_ls) CUDA_TRY(cudaMemcpyAsync(d->p_ls + lo * P, dpls, n * P * 4, cudaMemcpyDeviceToHost, s.stream));
        if (d->counts) CUDA_TRY(cudaMemcpyAsync(d->counts + lo * 4, dcnt, n * 16, cudaMemcpyDeviceToHost, s.stream));
        if (d->x_acc) CUDA_TRY(cudaMemcpyAsync(d->x_acc + lo * P, dxacc, n * P * 4, cudaMemcpyDeviceToHost, s.stream));
    }
    next_slot[device] = (unsigned)k;
    cudaSetDevice(prev);
    return rc;
}

static int host_wait(int device) {
    if (device < 0 || device >= 64) return fail("device index out of range");
    int prev = 0, rc = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(device));
    DeviceCtx* ctx;
    if (get_ctx(device, &ctx)) return -1;
    for (auto& s : ctx->slots) {
        if (!s.stream) continue;
        cudaError_t e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess && rc == 0) rc = fail(std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
        if (s.err_word && *s.err_word && rc == 0) {
            rc = fail("a straggler poller kernel gave up waiting (watchdog); unfinished fits keep info = DPHLM_INFO_UNFINISHED");
            cudaMemset(ctx->err_word, 0, sizeof(int));
        }
        if (s.err_word) *s.err_word = 0;
    }
    cudaSetDevice(prev);
    return rc;
}

int dphlm_fit_batch_host(const dphlm_desc* d, int device) {
    const int rc = host_enqueue(d, device, false);
    const int rw = host_wait(device);  // also after a failed enqueue: nothing of this call may still be in flight on return
    return rc ? rc : rw;
}

int dphlm_fit_batch_host_async(const dphlm_desc* d, int device) { return host_enqueue(d, device, true); }

int dphlm_host_wait(int device) { return host_wait(device); }

int dphlm_covariance(const dphlm_desc* d, const float* popt, float* pcov, int kind, void* stream) {
    if (validate_shape(d)) return -1;
    if (kind != DPHLM_COV_JTJ && kind != DPHLM_COV_CRLB) return fail("kind must be DPHLM_COV_JTJ or DPHLM_COV_CRLB");
    if (d->n_fits > 0 && (!popt || !pcov)) return fail("null popt/pcov pointer");
    if (d->n_fits == 0) return 0;
    CovArgs a{popt, d->xaxis, d->consts, pcov, d->n_fits, d->dim0, d->dim1, kind};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (d->model) {
        case DPHLM_GAUSS2D: return launch_cov_gauss2d(a, s);
        case DPHLM_GAUSS2D_ELLIPTICAL: return launch_cov_gauss2d_elliptical(a, s);
        case DPHLM_GAUSS2D_INTEGRATED: return launch_cov_gauss2d_integrated(a, s);
        case DPHLM_GAUSS2D_INTEGRATED_FIXED: return launch_cov_gauss2d_integrated_fixed(a, s);
        case DPHLM_MULTI_EXP_IRF: return launch_cov_multi_exp_irf(d->n_param, a, s);
        case DPHLM_GAUSS2D_FIXED: return launch_cov_gauss2d_fixed(a, s);
        case DPHLM_MULTI_EXP: return launch_cov_multi_exp(d->n_param, a, s);
    }
    return fail("unknown model id");
}

void* dphlm_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        dphlm::g_err = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}

void dphlm_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int dphlm_last_kernel_ms(float ms[3]) {
    if (!ms || !g_timing.valid) return fail("no timed dphlm_fit_batch call on this thread");
    CUDA_TRY(cudaEventSynchronize(g_timing.ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&ms[0], g_timing.ev[0], g_timing.ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&ms[1], g_timing.ev[1], g_timing.ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&ms[2], g_timing.ev[2], g_timing.ev[3]));
    return 0;
}

int dphlm_measure_fp32_peak(double* tflops) {
    if (!tflops) return fail("null pointer");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256, iters = 8192;
    float* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, sizeof(float) * blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {  // the first repetitions also bring the clocks up
        cudaEventRecord(e0);
        dphlm_ffma_chain_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = 2.0 * 16 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return 0;
}

int dphlm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { dphlm::g_err = "cudaGetDeviceCount failed"; return -1; }
    return n;
}

int dphlm_version(void) { return DPHLM_VERSION; }

const char* dphlm_last_error(void) { return dphlm::g_err.c_str(); }

}  // extern "C"
