// lm_launch.cuh -- kernel template and launcher shared by the per-model translation units.
//
// Execution model.  One fit is owned by a group of G lanes of a warp (G = 2, 4, 8, 16 or 32, so a warp carries 32/G fits).
// Each stage of curve_fit(method='ls'|'mle') is one persistent kernel (dphlm_stage_kernel<StageA|StageB, Model, G, IO>):
// every lane group pulls its next fit from a global atomic counter as soon as its current one ends, stages the ROI into its
// own slice of shared memory with cp.async while the current fit iterates (clamping to >= 0 for the MLE estimator,
// lm.py:127, and checking finiteness, lm.py:651-652), runs lm_core.cuh: run_stage and writes its results.  There is no
// CTA-level synchronisation after the coordinate tables are built: groups never wait for each other, which is what absorbs
// the 6..30+ spread in iteration counts between fits.  Fits beyond the trip budget are parked in a
// queue and finished by a concurrently running poller kernel with a whole warp per fit; the lm() kernel is launched so that
// it starts in the CTA slots the pre-fit kernel frees while it drains (see launch()).
//
// Shared memory per lane group: two data buffers (the working ROI and the staged next one, each followed by the staged
// start parameters and pre-fit exit code), the double-buffered per-pixel values of the accepted and the trial point for
// models that keep them (accepting a step flips a buffer index), and the per-trial tables of models that have them.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "lm_core.cuh"

#ifndef DPH_WARPS
#define DPH_WARPS 4  // warps per CTA
#endif
#ifndef DPH_MINB
#define DPH_MINB 3   // CTAs per SM the register allocation is bounded for (measured: spills cost more than occupancy gains)
#endif

namespace dphlm {

constexpr int kParkRecord = 40;  // floats per continuation record (>= 4 P + 9 for P <= 7)

int fail(const std::string& m);  // sets the thread-local error text, returns -1
#define CUDA_TRY(x)                                                                                    \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) return ::dphlm::fail(std::string(#x) + ": " + cudaGetErrorString(e_));  \
    } while (0)

// One list of parked fits: producers append (atomic count, record, then the ready flag = fit + 1), consumers claim
// with a CAS on `cons` once count > cons and wait for the flag.  Slots beyond `cap` do not exist.
struct ParkQueue {
    int* count;
    int* cons;
    int* ready;    // [cap], zero-initialised per call
    float* rec;    // [cap][kParkRecord]
    int cap;
};

struct KArgs {
    const float* data;
    const float* xaxis;
    const float* consts;  // model constants of the call (Mo::set_consts), or null
    const float* p0;
    float* popt;
    int* info;
    int* nfev;
    float* chisq;
    float* p_ls;   // never null inside the kernels (user buffer or workspace)
    int* ier;      // workspace [n_fits]: exit code of the pre-fit (100 = non-finite input data)
    int* counts;   // optional
    float* x_acc;  // optional: last accepted point of lm()
    unsigned long long* counter;
    // straggler hand-off: a main kernel parks fits that used up `budget` trips in `produce`; a concurrently running
    // poller kernel (one warp per fit) takes them from `consume[]` as they appear and finishes them
    ParkQueue produce;      // where this kernel puts fits (cap == 0: nowhere)
    ParkQueue consume[2];   // poller kernels: the queues they serve; consume[1] holds fits that START this stage here
    int* wait_done[2];      // poller kernels: "producer has exited" flags of the kernels feeding consume[] (null: none)
    int* exit_counter;      // CTAs of this kernel that have exited; the last one raises *done_flag
    int* done_flag;
    int* error_flag;        // raised if a poller gave up waiting (never in correct operation)
    int budget;             // trips before parking; <= 0: never
    int max_backlog;        // parked fits waiting for a poller beyond which the main kernels stop parking
    int on_demand_tail;     // != 0: the last on_demand_tail / 4 waves of fits are claimed on demand, not ahead (IoDevice::stage)
    int skip_prefit;        // host-side only: DPHLM_FLAG_SKIP_PREFIT
    int no_vec16;           // tuning aid (DPHLM_NO_VEC16): stage with 4-byte copies even where 16-byte ones are possible
    long long n_fits;
    int h, w;
    Options opt;
    cudaEvent_t* ev;  // host-side only: 3 events around the two kernels when timing is requested, else null
};

constexpr int kWarps = DPH_WARPS;
constexpr int kIerParked = -1;   // pre-fit exit code of a fit whose pre-fit continues in the poller kernel
constexpr long long kPollTimeoutCycles = 120000000000ll;  // ~60 s without work or news: a poller gives up (and says so)

// per-fit stride of the shared-memory arrays: >= M and congruent to G modulo 32, so that the G-lane
// groups of a warp touch disjoint banks
__host__ __device__ inline int padded_stride(int M, int G) {
    int s = M + ((G % 32) - (M % 32) + 32) % 32;
    return s;
}

#if defined(__CUDACC__)
// ---- double-buffered staging of the next fit's inputs (cp.async, 4-byte granularity: ROIs of 121 floats
// are only 4-byte aligned in the batch).  While a group iterates on one fit, the ROI, the start
// parameters and the pre-fit exit code of the fit it will run next are already on their way into the
// group's second shared-memory buffer.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {  // both 16-byte aligned
    unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// release / acquire on a word another kernel of the same call polls
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int kStageExtra = 8;  // floats after the staged ROI: start parameters (<= 7) and one int
__device__ __forceinline__ bool staging_4byte_only(const KArgs& a) { return a.no_vec16 != 0; }

template <class Stage, class Mo, int G, bool MAIN, bool MLE, bool RESUME>
struct IoDevice {
    const KArgs& a;
    float* ynext;          // staging buffer [stride + kStageExtra] of this group
    long long next = -1;   // index of the staged fit, -1: nothing staged yet
    bool vec16 = false;    // 128-bit staging copies: ROI size a multiple of 16 bytes and both staging buffers 16-byte aligned

    __device__ int budget() const { return (RESUME || a.budget <= 0 || a.produce.cap <= 0) ? 0x7fffffff : a.budget; }
    __device__ const float* consts() const { return a.consts; }

    // append a record to a queue: count, record, fence, ready flag (in this order)
    __device__ static bool publish(const ParkQueue& q, const float* r, int nrec, long long fit) {
        const int slot = atomicAdd(q.count, 1);
        if (slot >= q.cap) return false;
        for (int k = 0; k < nrec; ++k) q.rec[(size_t)slot * kParkRecord + k] = r[k];
        __threadfence();
        *reinterpret_cast<volatile int*>(q.ready + slot) = (int)fit + 1;
        return true;
    }

    // hand a fit that has used up its trip budget to the poller kernel
    __device__ bool park(const Lanes<G>& ln, const Stage& st, long long fit) const {
        if (RESUME) return false;
        const int src = __ffs(ln.gmask) - 1;
        int ok = 0;
        if (ln.lane == 0) {
            // Park only while the pollers keep up: a backlog of more than a few fits per poller warp means the tail of this
            // workload is heavier than the pollers' capacity, and then a straggler is better off staying in its group
            // (which costs one group slot, as before the hand-off existed) than waiting in the list.
            const int count = *reinterpret_cast<volatile int*>(a.produce.count);
            const int backlog = count - *reinterpret_cast<volatile int*>(a.produce.cons);
            if (count < a.produce.cap && backlog <= a.max_backlog) {
                float r[kParkRecord];
                st.save(r);
                ok = publish(a.produce, r, Stage::kRecord, fit) ? 1 : 0;  // on failure the fit stays here
                // the lm() kernel skips a parked pre-fit (the lm() poller runs it once its pre-fit is done); written only after
                // the hand-off succeeded because the lm() kernel may already be reading this array (early start, see launch())
                if (!MAIN && ok) *reinterpret_cast<volatile int*>(a.ier + fit) = kIerParked;
            }
        }
        return __shfl_sync(ln.gmask, ok, src) != 0;
    }

    // hand a fit whose xtol test is too close to call in float32 to the poller kernel (Options::exact_ties); no backlog limit:
    // the pollers are sized for it (launch())
    template <class St> __device__ bool park_tie(const Lanes<G>& ln, const PixelState<Mo::NE>& px, const St& st, long long fit) const {
        if (RESUME || a.produce.cap <= 0) return false;
        const int src = __ffs(ln.gmask) - 1;
        int ok = 0;
        if (ln.lane == 0) {
            float r[kParkRecord];
            st.save(r);
#pragma unroll
            for (int k = 0; k <= Mo::P; ++k) r[St::kKeepAt + k] = px.y[px.stride + k];
            ok = publish(a.produce, r, St::kRecord, fit) ? 1 : 0;
        }
        return __shfl_sync(ln.gmask, ok, src) != 0;
    }

    // Claim the next fit index for this group and start copying its inputs.  `after` is the index of the fit the group is
    // about to work on (-1: none).  Near the end of the batch nothing is claimed ahead: a fit staged behind a long-running one
    // would be stuck there while other groups have already run out of work, so the last wave is handed out on demand
    // (fetch() then claims and loads synchronously).
    __device__ void stage(const Lanes<G>& ln, int M, int stride, long long after) {
        constexpr int P = Mo::P;
        const int src = __ffs(ln.gmask) - 1;
        const long long groups = (long long)gridDim.x * kWarps * (32 / G);
        if (a.on_demand_tail && after >= 0 && after + (3 + a.on_demand_tail / 4) * groups >= a.n_fits) {
            unsigned long long c = 0;
            if (ln.lane == 0) c = *reinterpret_cast<volatile unsigned long long*>(a.counter);
            c = __shfl_sync(ln.gmask, c, src);
            if ((long long)c + groups * a.on_demand_tail / 4 >= a.n_fits) { next = -1; return; }
        }
        unsigned long long idx = 0;
        if (ln.lane == 0) idx = atomicAdd(a.counter, 1ull);
        idx = __shfl_sync(ln.gmask, idx, src);
        next = (long long)idx;
        if (next >= a.n_fits) return;
        const float* d = a.data + idx * M;
        if (vec16) {
            for (int i = 4 * ln.lane; i < M; i += 4 * G) cp_async16(ynext + i, d + i);
        } else {
            for (int i = ln.lane; i < M; i += G) cp_async4(ynext + i, d + i);
        }
        const float* ps = (MAIN ? a.p_ls : a.p0) + idx * P;
        for (int k = ln.lane; k < P; k += G) cp_async4(ynext + stride + k, ps + k);
        if (MAIN && ln.lane == 0) cp_async4(ynext + stride + 7, a.ier + idx);
    }

    // poller kernel: wait for the next parked fit of the queues this kernel serves (one warp per fit).  Returns false
    // when the queues are empty and every producer has exited.
    __device__ bool fetch_parked(const Lanes<G>& ln, PixelState<Mo::NE>& px, Stage& st, long long& fit) const {
        constexpr int P = Mo::P;
        const long long t0 = clock64();
        int q = -1, slot = 0, r = 0;
        for (;;) {  // until a published record has been claimed
        for (;;) {
            // all lanes run the same loop; lane 0 decides
            int claim = -1, got = 0, finished = 0;
            if (ln.lane == 0) {
                for (int k = 1; k >= 0 && claim < 0; --k) {  // fits waiting to START this stage first: they have the longest way to go
                    const ParkQueue& pq = a.consume[k];
                    if (pq.cap <= 0) continue;
                    const int c = *reinterpret_cast<volatile int*>(pq.cons);
                    int n = *reinterpret_cast<volatile int*>(pq.count);
                    if (n > pq.cap) n = pq.cap;
                    // a ticket, not a compare-and-swap: with hundreds of warps on one list (DPHLM_FLAG_EXACT_TIES) the losers of a
                    // CAS round sleep and the list drains at one fit per microsecond.  A ticket beyond the current count (two warps
                    // saw the same last entry) is waited for below like any other slot, and given up when the producers have gone
                    if (c < n) {
                        const int t = atomicAdd(pq.cons, 1);
                        if (t < pq.cap) { claim = k; got = t; }
                    }
                }
                if (claim < 0) {
                    bool done = true;
                    for (int k = 0; k < 2; ++k)
                        if (a.wait_done[k] && *reinterpret_cast<volatile int*>(a.wait_done[k]) == 0) done = false;
                    if (done) {  // producers are gone: whatever they published is visible now; look once more
                        __threadfence();
                        bool empty = true;
                        for (int k = 0; k < 2; ++k) {
                            const ParkQueue& pq = a.consume[k];
                            if (pq.cap <= 0) continue;
                            int n = *reinterpret_cast<volatile int*>(pq.count);
                            if (n > pq.cap) n = pq.cap;
                            if (*reinterpret_cast<volatile int*>(pq.cons) < n) empty = false;
                        }
                        finished = empty ? 1 : 0;
                    }
                    if (!finished && clock64() - t0 > kPollTimeoutCycles) {
                        if (a.error_flag) *reinterpret_cast<volatile int*>(a.error_flag) = 1;
                        finished = 1;
                    }
                }
            }
            claim = __shfl_sync(ln.gmask, claim, 0);
            got = __shfl_sync(ln.gmask, got, 0);
            finished = __shfl_sync(ln.gmask, finished, 0);
            if (claim >= 0) { q = claim; slot = got; break; }
            if (finished) return false;
            __nanosleep(2000);
        }
        // wait for the record of the claimed slot (lane 0 polls, all lanes follow): 1 = there, 2 = it will never come (the
        // ticket is beyond what the producers published before they left), 3 = watchdog
        int state = 0;
        if (ln.lane == 0) {
            const ParkQueue& pq = a.consume[q];
            while (state == 0) {
                r = *reinterpret_cast<volatile int*>(pq.ready + slot);
                if (r != 0) { state = 1; break; }
                bool gone = true;
                for (int k = 0; k < 2; ++k)
                    if (a.wait_done[k] && *reinterpret_cast<volatile int*>(a.wait_done[k]) == 0) gone = false;
                if (gone) {
                    __threadfence();
                    int n = *reinterpret_cast<volatile int*>(pq.count);
                    if (n > pq.cap) n = pq.cap;
                    r = *reinterpret_cast<volatile int*>(pq.ready + slot);
                    if (r != 0) { state = 1; break; }
                    if (slot >= n) { state = 2; break; }
                }
                if (clock64() - t0 > kPollTimeoutCycles) {
                    if (a.error_flag) *reinterpret_cast<volatile int*>(a.error_flag) = 1;
                    state = 3;
                    break;
                }
                __nanosleep(200);
            }
        }
        state = __shfl_sync(ln.gmask, state, 0);
        r = __shfl_sync(ln.gmask, r, 0);
        if (state == 3) return false;
        if (state == 1) break;
        }
        const ParkQueue& pq = a.consume[q];
        __threadfence();
        const long long idx = r - 1;
        const float* d = a.data + idx * px.M;
        for (int i = ln.lane; i < px.M; i += G) {
            float v = __ldg(d + i);
            px.y[i] = (MAIN && MLE) ? clamp_nonneg(v) : v;
        }
        __syncwarp(ln.gmask);
        float rec[kParkRecord];
        const volatile float* src = pq.rec + (size_t)slot * kParkRecord;
        if (q == 1) {  // starts this stage here: the record is just the start parameters
#pragma unroll
            for (int k = 0; k < P; ++k) rec[k] = src[k];
            st.start(rec, a.consts);
        } else {
#pragma unroll
            for (int k = 0; k < Stage::kRecord; ++k) rec[k] = src[k];
            st.resume(rec, a.consts);
            if constexpr (Stage::kTie) {  // the previous accepted point goes back behind the working ROI (StageB::note_accepted)
                if (ln.lane <= Mo::P) px.y[px.stride + ln.lane] = rec[Stage::kKeepAt + ln.lane];
                __syncwarp(ln.gmask);
            }
        }
        fit = idx;
        return true;
    }

    __device__ bool fetch(const Lanes<G>& ln, PixelState<Mo::NE>& px, Stage& st, long long& fit) {
        constexpr int P = Mo::P;
        if (RESUME) return fetch_parked(ln, px, st, fit);
        float pstart[P];
        for (;;) {
            if (next < 0) stage(ln, px.M, px.stride, -1);  // nothing staged (first fit, or the last wave): claim now
            if (next >= a.n_fits) {
                // No new fits are left, only the ones in flight: from here on a kernel launched behind this one with the
                // programmatic-serialization attribute (the lm() kernel behind the pre-fit kernel) may take the CTA slots
                // this grid frees while it drains.  No effect on ordinary launches.
                if (!MAIN) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                return false;
            }
            const long long idx = next;
            cp_async_wait_all();
            __syncwarp(ln.gmask);
            float* t = px.y; px.y = ynext; ynext = t;  // the staged buffer becomes the working one
            stage(ln, px.M, px.stride, idx);            // and the following fit starts loading
#pragma unroll
            for (int k = 0; k < P; ++k) pstart[k] = px.y[px.stride + k];
            if (MAIN) {
                int ier = __float_as_int(px.y[px.stride + 7]);
                // This kernel is launched so that it may start while the pre-fit kernel is still draining (programmatic
                // dependent launch).  The pre-fit publishes the parameters of a fit, then its exit code (release); both arrays
                // start the call as "not yet" (exit code 0, parameters NaN), so a staged copy made too early shows.
                bool early = ier == 0;
#pragma unroll
                for (int k = 0; k < P; ++k) early |= pstart[k] != pstart[k];
                if (early) {
                    const long long t0 = clock64();
                    while ((ier = ld_acquire(a.ier + idx)) == 0) {
                        if (clock64() - t0 > kPollTimeoutCycles) {  // never in correct operation: report instead of hanging
                            if (a.error_flag) *reinterpret_cast<volatile int*>(a.error_flag) = 1;
                            ier = 99;
                            break;
                        }
                        __nanosleep(200);
                    }
#pragma unroll
                    for (int k = 0; k < P; ++k) pstart[k] = __ldcg(a.p_ls + idx * P + k);
                }
                if (ier <= kIerParked) continue;  // pre-fit parked: still running (or just finished) in the follow-up launch, which will also run lm()
                if (ier < 1 || ier > 4) {  // the pre-fit failed: report it, do not fit (scipy raises, lm.py:642-644)
                    if (ln.lane == 0) {
#pragma unroll
                        for (int k = 0; k < P; ++k) a.popt[idx * P + k] = pstart[k];
                        a.info[idx] = -ier;
                        a.nfev[idx] = 0;
                        a.chisq[idx] = 0.0f;
                        if (a.counts) { a.counts[4 * idx + 1] = 0; a.counts[4 * idx + 2] = 0; a.counts[4 * idx + 3] = 0; }
                        if (a.x_acc) {
#pragma unroll
                            for (int k = 0; k < P; ++k) a.x_acc[idx * P + k] = pstart[k];
                        }
                    }
                    continue;
                }
                if (MLE) {  // data <= 0 -> 0 (lm.py:127)
                    for (int i = ln.lane; i < px.M; i += G) px.y[i] = clamp_nonneg(px.y[i]);
                    __syncwarp(ln.gmask);
                }
                fit = idx;
                st.start(pstart, a.consts);
                return true;
            }
            bool bad = false;
            for (int i = ln.lane; i < px.M; i += G) bad |= !f_finite(px.y[i]);
            bad = __any_sync(ln.gmask, bad);
            if (!bad) { fit = idx; st.start(pstart, a.consts); return true; }
            if (ln.lane == 0) {  // the reference raises ValueError (lm.py:651-652)
#pragma unroll
                for (int k = 0; k < P; ++k) a.p_ls[idx * P + k] = pstart[k];
                if (a.counts) reinterpret_cast<int4*>(a.counts)[idx] = make_int4(0, 0, 0, 0);
                st_release(a.ier + idx, 100);
            }
        }
    }
    __device__ void store(const Lanes<G>& ln, const StageA<Mo>& st, long long fit) const {
        if (ln.lane != 0) return;
#pragma unroll
        for (int k = 0; k < Mo::P; ++k) a.p_ls[fit * Mo::P + k] = st.x[k];
        if (a.counts) a.counts[4 * fit] = st.nfev;
        const int ier = st.ier == 0 ? 99 : st.ier;  // 0 is "not published yet" to the lm() kernel
        if (!RESUME) { st_release(a.ier + fit, ier); return; }
        // a resumed pre-fit ends while the lm() kernel may be reading this array: keep the entry <= kIerParked ...
        *reinterpret_cast<volatile int*>(a.ier + fit) = -ier - 2;
        if (st.ier >= 1 && st.ier <= 4) {  // ... and pass the fit on to the lm() poller, which starts its lm() loop
            publish(a.produce, st.x, Mo::P, fit);
            return;
        }
#pragma unroll
        for (int k = 0; k < Mo::P; ++k) a.popt[fit * Mo::P + k] = st.x[k];  // the pre-fit failed: report it (scipy raises)
        a.info[fit] = -ier;
        a.nfev[fit] = 0;
        a.chisq[fit] = 0.0f;
        if (a.counts) { a.counts[4 * fit + 1] = 0; a.counts[4 * fit + 2] = 0; a.counts[4 * fit + 3] = 0; }
        if (a.x_acc) {
#pragma unroll
            for (int k = 0; k < Mo::P; ++k) a.x_acc[fit * Mo::P + k] = st.x[k];
        }
    }
    __device__ void store(const Lanes<G>& ln, const StageB<Mo, MLE>& st, long long fit) const {
        if (ln.lane != 0) return;
#pragma unroll
        for (int k = 0; k < Mo::P; ++k) a.popt[fit * Mo::P + k] = st.xret[k];
        a.info[fit] = st.info;
        a.nfev[fit] = st.ev;
        a.chisq[fit] = st.chisq_ret;
        if (a.counts) { a.counts[4 * fit + 1] = st.n_acc; a.counts[4 * fit + 2] = st.n_rej; a.counts[4 * fit + 3] = st.n_tie; }
        if (a.x_acc) {
#pragma unroll
            for (int k = 0; k < Mo::P; ++k) a.x_acc[fit * Mo::P + k] = st.x[k];
        }
    }
};

// shared-memory floats of one lane group: two data buffers (working + staged, each with the extra
// words) and the double-buffered exponentials
__host__ __device__ inline int group_region(int stride, int ne, int table, int G) { return 2 * (stride + kStageExtra) + 2 * ne * stride + table + stash_floats(G, ne > 0); }

// CTAs per SM the register allocation is bounded for: 3 (<= 168 registers) where that costs no spills (the symmetric
// Gaussian, single exponentials); models with more parameters or more per-pixel state declare kMinCtas = 2 (<= 255
// registers): measured +28 % (mle) / +37 % (ls) for the 7-parameter elliptical Gaussian
template <class Mo>
constexpr int min_ctas() { return (Mo::kMinCtas < DPH_MINB ? Mo::kMinCtas : DPH_MINB) * 4 / DPH_WARPS; }  // kMinCtas counts four-warp CTAs

template <class Stage, class Mo, int G, class IO>
__global__ void __launch_bounds__(kWarps * 32, min_ctas<Mo>()) dphlm_stage_kernel(const KArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int FPW = 32 / G;  // fits in flight per warp
    const int M = a.h * a.w;
    const int stride = padded_stride(M, G);
    const int Mp = (M + 1) & ~1;
    float* cx = smem;
    float* cy = smem + Mp;
    for (int i = threadIdx.x; i < Mp; i += blockDim.x) fill_coords<Mo>(cx, cy, i, a.h, a.w, a.xaxis);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G;
    const int table = Mo::table_floats(a.h, a.w);
    const int region = group_region(stride, Mo::kStoreE ? Mo::NE : 0, table, G);
    float* base = smem + 2 * ((M + 1) & ~1) + (warp * FPW + grp) * region;
    Lanes<G> ln{lane % G, G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (grp * G))};
    PixelState<Mo::NE> px{base, base + 2 * (stride + kStageExtra), cx, cy, M, stride, 0};
    px.tbl = base + region - table - stash_floats(G, Mo::kStoreE);
    px.stash = base + region - stash_floats(G, Mo::kStoreE);
    px.w = a.w;
    IO io{a, base + stride + kStageExtra};
    // ROIs whose size is a multiple of 16 bytes (8x8, 12x12, 256-bin histograms ...) are staged with 128-bit copies; an 11x11
    // ROI (484 B) is only 4-byte aligned in the batch ("Staging")
    io.vec16 = !staging_4byte_only(a) && (M % 4 == 0) && ((reinterpret_cast<size_t>(a.data) & 15) == 0) &&
               (((__cvta_generic_to_shared(base) | __cvta_generic_to_shared(base + stride + kStageExtra)) & 15) == 0);
    run_stage<Stage, Mo, G>(ln, px, a.opt, io);
    if (a.exit_counter) {  // tell the pollers fed by this kernel when its last CTA has gone
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(a.exit_counter, 1) == (int)gridDim.x - 1) {
                __threadfence();
                *reinterpret_cast<volatile int*>(a.done_flag) = 1;
            }
        }
    }
}
#endif

// ------------------------------------------------------------------------------------ covariance
// Per-fit parameter covariance from the Jacobian at a given point (second return value of curve_fit; the reference
// evaluates it at the last ACCEPTED point of lm(), lm.py:468-473, 489 -- `x_acc` of dphlm_fit_batch):
//   kind 0: pinv(J^T J) of the unweighted Jacobian, the reference's pcov (dphtools/utils/lm.py:672-677);
//   kind 1: Poisson Cramer-Rao bound (J^T diag(1/f) J)^-1, the uncertainty an MLE user actually wants
//           (precedent: notebooks/MLE_LevMarq_4Param_2013_11_12.m:145-168).
// Off the hot path, so it is done in float64: model and Jacobian from Mo::eval64, J^T J summed over eight lanes per fit,
// then a cyclic Jacobi eigen-decomposition V diag(l) V^T on one lane and pinv = V diag(1/l) V^T over the eigenvalues above
// eps max(M, P) l_max.  The reference cuts the SINGULAR VALUES of J at eps max(M, P) s_max (lm.py:675); from J^T J only
// eigenvalues down to ~1e-16 l_max exist, so an exactly rank-deficient Jacobian (a zero or repeated column) gives the
// reference's answer, while directions with 3e-14 < s / s_max < 2e-7 -- kept upstream, numerically meaningless -- are dropped.
struct CovArgs {
    const float* popt;
    const float* xaxis;
    const float* consts;
    float* pcov;
    long long n_fits;
    int h, w, kind;
};

template <int P>
__device__ void sym_pinv64(double (&a)[P][P], double cut_rel, double (&out)[P][P]) {
    double v[P][P];
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j < P; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            diag += a[i][i] * a[i][i];
#pragma unroll
            for (int j = i + 1; j < P; ++j) off += a[i][j] * a[i][j];
        }
        if (off <= 1e-60 || off <= 1e-34 * diag) break;
#pragma unroll
        for (int p = 0; p < P - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < P; ++q) {
                const double apq = a[p][q];
                if (apq == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
                for (int k = 0; k < P; ++k) {  // columns p, q of A and V
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq; a[k][q] = sn * akp + c * akq;
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq; v[k][q] = sn * vkp + c * vkq;
                }
#pragma unroll
                for (int k = 0; k < P; ++k) {  // rows p, q of A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk; a[q][k] = sn * apk + c * aqk;
                }
            }
        }
    }
    double lmax = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) lmax = fmax(lmax, a[i][i]);
    double inv[P];
#pragma unroll
    for (int i = 0; i < P; ++i) inv[i] = (a[i][i] > cut_rel * lmax && a[i][i] > 0.0) ? 1.0 / a[i][i] : 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j < P; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < P; ++k) s += v[i][k] * inv[k] * v[j][k];
            out[i][j] = s;
        }
}

template <class Mo>
__global__ void __launch_bounds__(128) dphlm_cov_kernel(const CovArgs a) {
    constexpr int P = Mo::P, G = 8;
    const int lane = threadIdx.x & 31;
    const long long fit = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool active = fit < a.n_fits;
    const int gl = lane % G;
    double p[P];
#pragma unroll
    for (int k = 0; k < P; ++k) p[k] = active ? double(a.popt[fit * P + k]) : 1.0;
    double A[P][P];
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j < P; ++j) A[i][j] = 0.0;
    const int M = a.h * a.w;
    for (int i = gl; i < M; i += G) {
        double f, J[P];
        const double cx = a.xaxis ? double(a.xaxis[i]) : double(i % a.w), cy = a.xaxis ? double(i) : double(i / a.w);
        Mo::eval64(p, a.consts, cx, cy, f, J);
        double wgt = 1.0;
        if (a.kind == 1) wgt = f > 0.0 ? 1.0 / f : 0.0;
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
            for (int c = r; c < P; ++c) A[r][c] += wgt * J[r] * J[c];
    }
#pragma unroll
    for (int r = 0; r < P; ++r)
#pragma unroll
        for (int c = r; c < P; ++c) {
            double v = A[r][c];
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            A[r][c] = A[c][r] = v;
        }
    if (!active || gl != 0) return;
    double out[P][P];
    sym_pinv64<P>(A, 2.220446049250313e-16 * double(M > P ? M : P), out);
#pragma unroll
    for (int r = 0; r < P; ++r)
#pragma unroll
        for (int c = 0; c < P; ++c) a.pcov[(fit * P + r) * P + c] = float(out[r][c]);
}

template <class Mo>
int launch_cov(const CovArgs& a, cudaStream_t stream) {
    const long long threads = a.n_fits * 8;
    const long long grid = (threads + 127) / 128;
    if (grid > 0) dphlm_cov_kernel<Mo><<<(unsigned)grid, 128, 0, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

constexpr int kHostSlots = 8;  // staging slots of the host path (DPHLM_HOST_SLOTS of them are used: chunks in flight)

struct DeviceCtx {
    std::mutex mu;
    int sms = 0;
    // host-path staging buffers, one set per pipeline slot
    struct Slot {
        cudaStream_t stream = nullptr;
        char* buf = nullptr;
        size_t cap = 0;
        int* err_word = nullptr;  // pinned
    } slots[kHostSlots];
    // side stream for the follow-up launch of the pre-fit stragglers (forked from / joined into the caller's)
    // (a few, used round-robin, so that concurrent calls on different streams do not serialise behind each other)
    std::mutex fork_mu;
    struct Side {
        cudaStream_t stream_a = nullptr, stream_b = nullptr;
        cudaEvent_t fork_ev = nullptr, mid_ev = nullptr, join_a = nullptr, join_b = nullptr;
    } sides[4];
    unsigned next_side = 0;
    // pipelined calls (DPHLM_FLAG_PIPELINED): the library's own streams, used round-robin; `in_ev` orders a call behind the
    // caller's stream, `done_ev` is what dphlm_join() makes the caller's stream wait for
    static constexpr int kPipeMax = 4;
    std::mutex pipe_mu;
    struct Pipe {
        cudaStream_t stream = nullptr;
        cudaEvent_t in_ev = nullptr, done_ev = nullptr;
        bool pending = false;
    } pipes[kPipeMax];
    unsigned next_pipe = 0;
    int* err_word = nullptr;  // device: raised by a straggler kernel that gave up waiting (sticky until dphlm_call_status)
};

template <class Kern>
int launch_stage(Kern kern, const KArgs& a0, int fpw, size_t smem, DeviceCtx* ctx, cudaStream_t stream, int reserve = 0, int fixed_grid = 0,
                 bool early_start = false) {
    KArgs a = a0;
    if (smem > 227 * 1024) return ::dphlm::fail("ROI too large for the shared-memory layout of this group size");
    // attribute setup once per (kernel instantiation, device) -- at the maximum, so that it never shrinks under a concurrent
    // caller with a larger ROI -- and one occupancy query per shared-memory size
    struct Cached { bool attr = false; std::map<size_t, int> occ; };
    static std::map<std::pair<const void*, int>, Cached> cache;  // Kern is one function-pointer type for all kernels
    static std::mutex cache_mu;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int occ;
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        Cached& c = cache[std::make_pair(reinterpret_cast<const void*>(kern), dev)];
        if (!c.attr) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            c.attr = true;
        }
        int& o = c.occ[smem];
        if (o == 0) {
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kWarps * 32, smem));
            if (o < 1) { c.occ.erase(smem); return ::dphlm::fail("kernel does not fit on an SM"); }
        }
        occ = o;
    }
    // a.counter points into the call's own workspace header (zeroed once per call by fit_device)
    // main kernels fill the machine minus the CTA slots reserved for the concurrently running pollers
    long long grid = (long long)ctx->sms * occ - reserve;
    const long long need = (a.n_fits + (long long)kWarps * fpw - 1) / ((long long)kWarps * fpw);
    if (grid > need) grid = need;
    if (fixed_grid > 0) grid = fixed_grid;
    if (grid < 1) grid = 1;
    if (early_start) {  // may begin while the previous kernel of the stream drains; it synchronises per fit, not per grid
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kWarps * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
        return 0;
    }
    kern<<<(unsigned)grid, kWarps * 32, smem, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// both stages of curve_fit(method='ls'|'mle') for one batch: pre-fit kernel, then the lm() kernel, each followed
// by a small warp-per-fit launch that finishes the stragglers the main kernel parked
constexpr int kBudgetPrefit = 16, kBudgetMain = 24;  // ~p99.9 of the trip counts (mean 7 and 9.7)
inline int budget_from_env(const char* name, int dflt) {  // tuning aid: DPHLM_BUDGET_A / DPHLM_BUDGET_B
    const char* e = getenv(name);
    return e && atoi(e) > 0 ? atoi(e) : dflt;
}

// header words of the per-call workspace (zeroed once per call)
enum { kSyncCountA = 0, kSyncCountB, kSyncCountL, kSyncConsA, kSyncConsB, kSyncConsL, kSyncDoneAMain, kSyncDoneAWide, kSyncDoneBMain,
       kSyncExitAMain, kSyncExitAWide, kSyncExitBMain, kSyncError, kSyncWords = 16 };
constexpr int kPollerCtasDefault = 8;  // CTA slots kept free of main-kernel CTAs for each poller kernel (32 warps = 32 stragglers at a time)

struct ParkLists {  // carved from the workspace by fit_device
    int* sync;      // [kSyncWords]
    int* ready;     // [3][cap]
    float* rec;     // [3][cap][kParkRecord]
    int cap;
    __host__ ParkQueue queue(int which) const {  // 0: parked pre-fits, 1: parked lm() loops, 2: late pre-fits waiting for lm()
        return ParkQueue{sync + kSyncCountA + which, sync + kSyncConsA + which, ready + (size_t)which * cap,
                         rec + (size_t)which * cap * kParkRecord, cap};
    }
};

// Both stages of curve_fit(method='ls'|'mle') for one batch.  Four kernels, two of them pollers:
//   caller's stream : pre-fit main --------> lm() main ------------------> (join)
//   side stream A   : pre-fit poller (parked pre-fits; passes each finished one on to the lm() poller)
//   side stream B   :                        lm() poller (parked lm() loops + late pre-fits)
// The pollers run concurrently with the main kernels, which leave kPollerCtas CTA slots free for each of them, so
// every kernel is resident whatever order the hardware starts them in; a poller only ever waits for work or for the
// "last CTA has exited" flag of the kernels that feed it, never the other way round.
template <class Mo, bool MLE, int G>
int launch(const KArgs& a0, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t stream) {
    const bool skip_prefit = a0.skip_prefit != 0;  // lm() alone: p_ls aliases p0, every pre-fit exit code starts as 1
    const int M = a0.h * a0.w;
    constexpr int FPW = 32 / G;
    const size_t smem = sizeof(float) * (2 * (size_t)((M + 1) & ~1) + (size_t)kWarps * FPW * group_region(padded_stride(M, G), Mo::kStoreE ? Mo::NE : 0, Mo::table_floats(a0.h, a0.w), G));
    const size_t smem_w = sizeof(float) * (2 * (size_t)((M + 1) & ~1) + (size_t)kWarps * group_region(padded_stride(M, 32), Mo::kStoreE ? Mo::NE : 0, Mo::table_floats(a0.h, a0.w), 32));
    const bool handoff = G < 32 && pl.cap > 0;
    const int kPollerCtas = budget_from_env("DPHLM_POLLER_CTAS", kPollerCtasDefault);  // tuning aid
    // With Options::exact_ties 3-20 % of the fits pass through the lm() poller (it re-decides the borderline xtol tests in
    // float64, lm_core.cuh: tie_break): a whole-machine launch BEHIND the lm() main kernel then, instead of eight CTAs beside it
    // (one warp per fit is latency-bound: it takes the width of the machine to get through tens of thousands of them).
    const bool exact = a0.opt.exact_ties != 0;
    const int kPollerCtasB = exact ? 0 : kPollerCtas;
    using SA = StageA<Mo>;
    using SB = StageB<Mo, MLE>;
    unsigned long long* const counters = a0.counter;  // work counters of the two main kernels
    const ParkQueue none{nullptr, nullptr, nullptr, nullptr, 0};
    KArgs a = a0;
    a.counter = counters;
    a.on_demand_tail = getenv("DPHLM_NO_ON_DEMAND_TAIL") ? 0 : 4;  // in quarters of a wave (one wave = one fit per group)
    if (const char* e = getenv("DPHLM_TAIL_QUARTERS")) { if (a.on_demand_tail && atoi(e) > 0) a.on_demand_tail = atoi(e); }  // tuning aid
    a.no_vec16 = getenv("DPHLM_NO_VEC16") ? 1 : 0;                  // tuning aid
    a.produce = a.consume[0] = a.consume[1] = none;
    a.wait_done[0] = a.wait_done[1] = nullptr;
    a.exit_counter = a.done_flag = nullptr;
    a.error_flag = ctx->err_word;
    KArgs b = a;
    b.counter = counters + 1;
    if (a.ev) CUDA_TRY(cudaEventRecord(a.ev[0], stream));
    if (!handoff) {
        a.budget = b.budget = 0;
        int rc = 0;
        if (!skip_prefit) rc = launch_stage(dphlm_stage_kernel<SA, Mo, G, IoDevice<SA, Mo, G, false, false, false>>, a, FPW, smem, ctx, stream);
        if (rc) return rc;
        if (a.ev) CUDA_TRY(cudaEventRecord(a.ev[1], stream));
        rc = launch_stage(dphlm_stage_kernel<SB, Mo, G, IoDevice<SB, Mo, G, true, MLE, false>>, b, FPW, smem, ctx, stream);
        if (rc) return rc;
        if (a.ev) { CUDA_TRY(cudaEventRecord(a.ev[2], stream)); CUDA_TRY(cudaEventRecord(a.ev[3], stream)); }
        return 0;
    }
    DeviceCtx::Side& sd = ctx->sides[ctx->next_side++ % 4];  // the caller holds ctx->fork_mu
    // pre-fit main + its poller
    a.budget = budget_from_env("DPHLM_BUDGET_A", kBudgetPrefit);
    a.max_backlog = b.max_backlog = budget_from_env("DPHLM_MAX_BACKLOG", 3 * kPollerCtas * kWarps);
    a.produce = pl.queue(0);
    a.exit_counter = pl.sync + kSyncExitAMain; a.done_flag = pl.sync + kSyncDoneAMain;
    KArgs aw = a;
    aw.budget = 0;
    aw.consume[0] = pl.queue(0); aw.produce = pl.queue(2);
    aw.wait_done[0] = pl.sync + kSyncDoneAMain;
    aw.exit_counter = pl.sync + kSyncExitAWide; aw.done_flag = pl.sync + kSyncDoneAWide;
    CUDA_TRY(cudaEventRecord(sd.fork_ev, stream));
    if (!skip_prefit) CUDA_TRY(cudaStreamWaitEvent(sd.stream_a, sd.fork_ev, 0));
    CUDA_TRY(cudaStreamWaitEvent(sd.stream_b, sd.fork_ev, 0));
    int rc = 0;
    if (!skip_prefit) {
        rc = launch_stage(dphlm_stage_kernel<SA, Mo, G, IoDevice<SA, Mo, G, false, false, false>>, a, FPW, smem, ctx, stream, kPollerCtas + kPollerCtasB);
        if (rc) return rc;
        rc = launch_stage(dphlm_stage_kernel<SA, Mo, 32, IoDevice<SA, Mo, 32, false, false, true>>, aw, 1, smem_w, ctx, sd.stream_a, 0, kPollerCtas);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(sd.join_a, sd.stream_a));
    }
    if (a.ev) CUDA_TRY(cudaEventRecord(a.ev[1], stream));
    // lm() main + its poller (both behind the pre-fit main kernel)
    b.budget = budget_from_env("DPHLM_BUDGET_B", kBudgetMain);
    b.produce = pl.queue(1);
    b.exit_counter = pl.sync + kSyncExitBMain; b.done_flag = pl.sync + kSyncDoneBMain;
    KArgs bw = b;
    bw.budget = 0;
    bw.produce = none;
    bw.consume[0] = pl.queue(1); bw.consume[1] = skip_prefit ? none : pl.queue(2);
    bw.wait_done[0] = pl.sync + kSyncDoneBMain; bw.wait_done[1] = skip_prefit ? nullptr : pl.sync + kSyncDoneAWide;
    bw.exit_counter = bw.done_flag = nullptr;
    if (exact) b.max_backlog = pl.cap;  // nobody takes fits off the list while the main kernel runs: no limit on what waits there
    // The lm() kernel directly follows the pre-fit kernel in the stream and may start in the CTA slots that one frees while its
    // last fits finish: it waits per fit for the pre-fit's exit code (IoDevice::fetch), not for the grid.  Timing events
    // between the two launches would serialise them, so a timed call keeps the plain stream order.
    const bool early = !a.ev && !skip_prefit && !getenv("DPHLM_NO_EARLY_START");
    rc = launch_stage(dphlm_stage_kernel<SB, Mo, G, IoDevice<SB, Mo, G, true, MLE, false>>, b, FPW, smem, ctx, stream, kPollerCtas + kPollerCtasB, 0, early);
    if (rc) return rc;
    if (a.ev) CUDA_TRY(cudaEventRecord(a.ev[2], stream));  // end of the lm() main kernel itself
    if (exact) {  // after the main kernel and the pre-fit poller (which feeds it the late pre-fits), on the caller's stream
        if (!skip_prefit) CUDA_TRY(cudaStreamWaitEvent(stream, sd.join_a, 0));
        rc = launch_stage(dphlm_stage_kernel<SB, Mo, 32, IoDevice<SB, Mo, 32, true, MLE, true>>, bw, 1, smem_w, ctx, stream);
        if (rc) return rc;
    } else {
        rc = launch_stage(dphlm_stage_kernel<SB, Mo, 32, IoDevice<SB, Mo, 32, true, MLE, true>>, bw, 1, smem_w, ctx, sd.stream_b, 0, kPollerCtasB);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(sd.join_b, sd.stream_b));
        if (!skip_prefit) CUDA_TRY(cudaStreamWaitEvent(stream, sd.join_a, 0));
        CUDA_TRY(cudaStreamWaitEvent(stream, sd.join_b, 0));
    }
    if (a.ev) CUDA_TRY(cudaEventRecord(a.ev[3], stream));  // the pollers have finished the stragglers
    return 0;
}

// models that are only built for wide groups (256-bin histograms)
#define DPH_DISPATCH_G_WIDE(MO, G, mle, a, pl, ctx, s)                                                    \
    switch (G) {                                                                                          \
        case 16: return (mle) ? launch<MO, true, 16>(a, pl, ctx, s) : launch<MO, false, 16>(a, pl, ctx, s); \
        case 32: return (mle) ? launch<MO, true, 32>(a, pl, ctx, s) : launch<MO, false, 32>(a, pl, ctx, s); \
    }                                                                                                     \
    return ::dphlm::fail("this model is built for group_lanes 0, 16 or 32")

// one per model family, defined in lm_inst_*.cu
int launch_gauss2d(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_gauss2d_elliptical(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_gauss2d_integrated(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_gauss2d_integrated_fixed(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_multi_exp1(bool offset, int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_multi_exp2(bool offset, int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_multi_exp3(bool offset, int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_cov_gauss2d(const CovArgs& a, cudaStream_t s);
int launch_cov_gauss2d_elliptical(const CovArgs& a, cudaStream_t s);
int launch_cov_gauss2d_integrated(const CovArgs& a, cudaStream_t s);
int launch_cov_gauss2d_integrated_fixed(const CovArgs& a, cudaStream_t s);
int launch_gauss2d_fixed(int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_cov_gauss2d_fixed(const CovArgs& a, cudaStream_t s);
int launch_multi_exp_irf(int n_param, int G, bool mle, const KArgs& a, const ParkLists& pl, DeviceCtx* ctx, cudaStream_t s);
int launch_cov_multi_exp_irf(int n_param, const CovArgs& a, cudaStream_t s);
int launch_cov_multi_exp(int n_param, const CovArgs& a, cudaStream_t s);

#define DPH_DISPATCH_G(MO, G, mle, a, pl, ctx, s)                                                     \
    switch (G) {                                                                                   \
        case 2: return (mle) ? launch<MO, true, 2>(a, pl, ctx, s) : launch<MO, false, 2>(a, pl, ctx, s);     \
        case 4: return (mle) ? launch<MO, true, 4>(a, pl, ctx, s) : launch<MO, false, 4>(a, pl, ctx, s);     \
        case 8: return (mle) ? launch<MO, true, 8>(a, pl, ctx, s) : launch<MO, false, 8>(a, pl, ctx, s);     \
        case 16: return (mle) ? launch<MO, true, 16>(a, pl, ctx, s) : launch<MO, false, 16>(a, pl, ctx, s);  \
        case 32: return (mle) ? launch<MO, true, 32>(a, pl, ctx, s) : launch<MO, false, 32>(a, pl, ctx, s);  \
    }                                                                                              \
    return ::dphlm::fail("group_lanes must be 0, 2, 4, 8, 16 or 32")

}  // namespace dphlm
