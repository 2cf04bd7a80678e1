/* dphlm.h -- C ABI of the B200 batched Levenberg-Marquardt / Poisson-MLE fitter.
 *
 * The reference (david-hoffman/dphtools) is pure Python and has no FFI layer; its boundary for this
 * path is the Python call surface of dphtools/utils/lm.py.  Each entry point below names the reference
 * interface it stands in for.  The Python host side (dphtools_b200/lm.py) binds these with ctypes; the
 * stub a dphtools maintainer would add is shown in INTEGRATION.md.
 *
 * All functions are thread-safe and re-entrant across host threads that use distinct streams/devices.
 * Return value: 0 on success, negative on error (text from dphlm_last_error(), thread-local).
 */
#ifndef DPHLM_H
#define DPHLM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPHLM_VERSION 200

/* model ids: callables a caller would pass as `f` to curve_fit (dphtools/utils/lm.py:510-512).
 * MULTI_EXP is the reference's own model (dphtools/utils/fitfuncs.py:20-52); the Gaussians are new:
 * GAUSS2D (amp, x0, y0, sigma, offset), GAUSS2D_ELLIPTICAL (amp, x0, y0, sigma_x, sigma_y, theta, offset) sampled at
 * the pixel centres; GAUSS2D_INTEGRATED (photons, x0, y0, sigma, offset) integrated over each pixel (erf model);
 * GAUSS2D_INTEGRATED_FIXED (photons, x0, y0, offset) the same with the width held at consts[0] (the 4-parameter model of
 * notebooks/MLE_LevMarq_4Param_2013_11_12.m:44-54, where the PSF width is an argument of the call);
 * GAUSS2D_FIXED (amp, x0, y0, offset): GAUSS2D with the width held at consts[0];
 * MULTI_EXP_IRF: MULTI_EXP (1-2 exponentials, optional offset) convolved with an instrument response function given on the
 * histogram's own uniform grid: f_j = offset + sum_i a_i sum_q irf[q] exp(-k_i (j - m0 - q) dt) over 0 <= q < W, q <= j - m0;
 * consts = [W (<= 32), m0, dt, irf[0..W-1]].  Built for group_lanes 0, 16, 32. */
enum dphlm_model {
    DPHLM_GAUSS2D = 0, DPHLM_GAUSS2D_ELLIPTICAL = 1, DPHLM_MULTI_EXP = 2, DPHLM_GAUSS2D_INTEGRATED = 3,
    DPHLM_GAUSS2D_INTEGRATED_FIXED = 4, DPHLM_MULTI_EXP_IRF = 5, DPHLM_GAUSS2D_FIXED = 6
};

/* `method` of curve_fit / lm (dphtools/utils/lm.py:235, 519, 621-626) */
enum dphlm_method { DPHLM_LS = 0, DPHLM_MLE = 1 };

/* Per-fit `info` written by the fitter:
 *    1, 2, 4   converged, same meaning as lm()'s info (dphtools/utils/lm.py:310-333, 409-432, 468-470)
 *    5         stage 2 reached maxfev (lm.py:479-481); curve_fit raises RuntimeError (lm.py:679-680)
 *   -k         the least-squares pre-fit (scipy curve_fit -> MINPACK lmder, lm.py:642-644) ended with
 *              ier = k not in 1..4 (k = 5: maxfev); scipy raises RuntimeError there
 *   -9         non-finite model value at p0
 *   -100       non-finite input data; the reference raises ValueError (lm.py:651-662)
 *   DPHLM_INFO_UNFINISHED   the fit was never finished: every call pre-fills `info` with this value, so a fit dropped by a
 *              failed launch or by a straggler kernel that gave up waiting (see dphlm_call_status) cannot be mistaken
 *              for a result                                                                         */
#define DPHLM_INFO_UNFINISHED (-2139062144) /* 0x80808080 */

typedef struct dphlm_desc {
    int32_t model;       /* enum dphlm_model */
    int32_t method;      /* enum dphlm_method */
    int32_t n_param;     /* 5 (GAUSS2D, GAUSS2D_INTEGRATED), 4 (their _FIXED forms), 7 (ELLIPTICAL), 2..7 (MULTI_EXP: 2 per exponential, +1 = offset) */
    int32_t dim0, dim1;  /* ROI height, width; (n_bins, 1) for 1-D models */
    int32_t group_lanes; /* lanes of a warp that own one fit: 0 = automatic, else 2, 4, 8, 16 or 32 */
    int64_t n_fits;
    const float* data;   /* [n_fits, dim0, dim1] ydata, row-major (uint16 with DPHLM_FLAG_DATA_U16)  (lm.py:512) */
    const float* xaxis;  /* MULTI_EXP(_IRF): [dim0] shared xdata; NULL for 2-D models (pixel indices)   */
    const float* p0;     /* [n_fits, n_param] initial guesses                               (lm.py:514) */
    float* popt;         /* [n_fits, n_param] fitted parameters                             (lm.py:682) */
    int32_t* info;       /* [n_fits]                                                        (lm.py:669) */
    int32_t* nfev;       /* [n_fits] infodict['nfev'] of stage 2                            (lm.py:489) */
    float* chisq;        /* [n_fits] chi-square at popt (0.5 sum r^2, or Poisson deviance / 2)          */
    float* p_ls;         /* optional [n_fits, n_param]: result of the least-squares pre-fit (lm.py:648) */
    int32_t* counts;     /* optional [n_fits, 4]: pre-fit evaluations, accepted, rejected steps, float64 tie-breaks */
    float* x_acc;        /* optional [n_fits, n_param]: last ACCEPTED point of lm(); infodict['fjac'] and pcov of the
                            reference are evaluated there, one step before popt                 (lm.py:468-473, 489) */
    float ftol, xtol, gtol, factor; /* lm()/leastsq keywords, forwarded to both stages (lm.py:643, 668) */
    int32_t maxfev;      /* <= 0: 100 * (n_param + 1)                                   (lm.py:306-307) */
    int32_t flags;       /* DPHLM_FLAG_* */
    const float* consts; /* constants of the model, shared by all fits: GAUSS2D_FIXED, GAUSS2D_INTEGRATED_FIXED: [1] sigma;
                            MULTI_EXP_IRF: [3 + W] = W, m0, dt, irf[0..W-1]; else NULL */
} dphlm_desc;

/* flags: record CUDA events around the two kernels of dphlm_fit_batch (read with dphlm_last_kernel_ms) */
#define DPHLM_FLAG_TIMING 1
/* flags: do not hand stragglers (fits beyond ~p99.9 of the trip counts) to the warp-per-fit follow-up launches */
#define DPHLM_FLAG_NO_HANDOFF 2
/* flags: `data` points to uint16 camera counts instead of float32 (half the PCIe traffic on the host entry); they are
 * widened on the device before the fit */
#define DPHLM_FLAG_DATA_U16 4
/* flags (dphlm_fit_batch only): pipelined call.  The kernels are enqueued on one of the library's own streams behind everything
 * that is in `stream` at the time of the call, and `stream` does NOT wait for them: consecutive pipelined calls overlap
 * (three in flight; the tail of one call -- its last fits, its stragglers -- runs under the next call's kernels), which is
 * worth ~25 % at 1e5 fits per call.  The caller must treat the outputs as pending, keep all buffers alive, and not feed
 * them to later work until dphlm_join(stream) has made `stream` wait for every pending pipelined call. */
#define DPHLM_FLAG_PIPELINED 8
/* flags: skip the least-squares pre-fit: `p0` is the start of the lm() loop itself.  This is the reference's second public
 * function, lm(func, x0, ..., method=) (dphtools/utils/lm.py:221-236), as opposed to curve_fit (lm.py:510-685). */
#define DPHLM_FLAG_SKIP_PREFIT 16
/* flags: re-decide in float64 the lm() trips whose step is within a factor 16 of the xtol threshold (lm.py:365-367, 430-432).
 * xtol = 1.49e-8 is four times below float32 resolution of x, so the float32 path splits the fits that end that close --
 * 0.1 % (5-parameter Gaussian) to 2 % (4-parameter models) of a batch -- between info 1 and info 2 by noise; both mean
 * "converged" and the parameters agree to 1e-7 either way.  With this flag those fits get the reference's flag (info equality
 * with the reference 99.8-100 % instead of 98-99.9 %), at the price of two float64 passes for the ~5 % of fits that come close. */
#define DPHLM_FLAG_EXACT_TIES 32

/* Fill the tolerances with the reference defaults (ftol = xtol = 1.49012e-8, gtol = 0, factor = 100,
 * maxfev = 0 -> 100 (P + 1); dphtools/utils/lm.py:228-233) and zero everything else. */
void dphlm_desc_init(dphlm_desc* d);

/* Batched replacement for per-ROI curve_fit(f, xdata, ydata, p0, jac=..., method='ls'|'mle')
 * (dphtools/utils/lm.py:510-685).  All pointers are DEVICE pointers on the current device; the work is
 * enqueued on `stream` (a cudaStream_t, NULL = default stream) and the call returns without waiting. */
int dphlm_fit_batch(const dphlm_desc* d, void* stream);

/* Makes `stream` wait for all pending DPHLM_FLAG_PIPELINED calls of the current device (asynchronous: enqueues event waits). */
int dphlm_join(void* stream);

/* Watchdog of the current device: the straggler kernels wait for work from the main kernels and give up after ~60 s without
 * news (never in correct operation; reachable under a debugger or on a GPU shared with a long-running foreign kernel).
 * Synchronises the device, writes 1 to *tripped if that has happened since the last call of this function, else 0, and
 * clears the record.  Fits that were dropped keep info = DPHLM_INFO_UNFINISHED.  (dphlm_fit_batch_host checks by itself.) */
int dphlm_call_status(int* tripped);

/* Same contract with HOST pointers: copies the inputs to `device`, fits, copies the results back and
 * returns when they are in place.  Large batches are cut into chunks whose copies overlap the kernels. */
int dphlm_fit_batch_host(const dphlm_desc* d, int device);

/* The same without the final wait: returns once the copies and kernels of the call are enqueued (from page-locked buffers
 * nothing blocks), so that the caller can prepare and enqueue the next batch -- whose host-to-device copies then overlap
 * this batch's kernels and result copies.  The buffers must stay untouched until dphlm_host_wait(device), which waits for
 * every enqueued call of the device and reports their errors. */
int dphlm_fit_batch_host_async(const dphlm_desc* d, int device);
int dphlm_host_wait(int device);

/* Parameter covariance of every fit from the model Jacobian at the given points `popt` (DEVICE pointers, asynchronous on `stream`;
 * uses model, n_param, dim0, dim1, n_fits, xaxis of the descriptor).  pcov is [n_fits, n_param, n_param].
 *   DPHLM_COV_JTJ   pinv(J^T J) of the unweighted Jacobian: the `pcov` curve_fit returns (lm.py:672-677)
 *   DPHLM_COV_CRLB  (J^T diag(1/f) J)^-1: the Poisson Cramer-Rao bound of the fitted parameters
 * Computed in float64 from the eigen-decomposition of J^T J: eigenvalues below eps max(M, P) l_max are dropped, so a
 * rank-deficient Jacobian gives the pseudo-inverse like the reference's singular-value cut (lm.py:675).  For the reference's
 * pcov pass `x_acc` of dphlm_fit_batch (the last ACCEPTED point, where lm() evaluates fjac), not popt. */
enum dphlm_cov_kind { DPHLM_COV_JTJ = 0, DPHLM_COV_CRLB = 1 };
int dphlm_covariance(const dphlm_desc* d, const float* popt, float* pcov, int kind, void* stream);

/* The step upstream of the fit: cut roi x roi windows around peak positions out of a stack of frames and derive
 * the Gaussian initial guesses on the device (DEVICE pointers, asynchronous on `stream`).
 * Window = slice_maker's (dphtools/utils/__init__.py:244-295): rows [row - roi/2, row - roi/2 + roi), same for columns.
 * p0 = (max - min, x0, y0, sigma0, max(min, 0.5)) with (x0, y0) the ROI centre or, with refine != 0, the vertex of the
 * parabolas through the centre row / column (localize_peak, dphtools/utils/__init__.py:590-617), in ROI coordinates.
 * A window crossing the frame border is filled by clamping and reported with valid = 0. */
enum dphlm_dtype { DPHLM_F32 = 0, DPHLM_U16 = 1 };
typedef struct dphlm_extract_desc {
    const void* frames;    /* [n_frames, height, width], float32 or uint16 (camera counts) */
    int32_t frame_dtype;   /* enum dphlm_dtype */
    int32_t n_frames, height, width;
    const int32_t* peaks;  /* [n_peaks, 3]: frame, row, column of the peak pixel */
    int64_t n_peaks;
    int32_t roi;           /* window size, 1..31 */
    int32_t refine;        /* 0: centre of the window; 1: parabola vertex */
    float sigma0;          /* initial width */
    float* rois;           /* out [n_peaks, roi, roi] float32: `data` of dphlm_fit_batch */
    float* p0;             /* out [n_peaks, 5]:            `p0`   of dphlm_fit_batch (DPHLM_GAUSS2D) */
    int32_t* valid;        /* out [n_peaks], optional */
} dphlm_extract_desc;
int dphlm_extract_rois(const dphlm_extract_desc* d, void* stream);

/* Initial guesses of multi_exp_fit (dphtools/utils/fitfuncs.py:63-80, 131-152) for n_fits histograms [n_fits, n_bins] on the
 * shared x axis, on the device: per component one segment of the x range -- bins [seg_bounds[k], seg_bounds[k+1]), the
 * reference's log-spaced split, computed once by the caller (DEVICE pointer to components + 1 ints) -- and in each segment
 * the log-linear estimate (amplitude, rate, offset) of _estimate_exponent_params; p0 is [n_fits, 2 components (+ 1)], the
 * layout dphlm_fit_batch takes for DPHLM_MULTI_EXP.  The data must be finite.  Asynchronous on `stream`. */
int dphlm_guess_multi_exp(const float* data, const float* xaxis, const int32_t* seg_bounds, int32_t components, int32_t with_offset,
                          int32_t n_bins, int64_t n_fits, float* p0, void* stream);

/* Page-locked host buffers for dphlm_fit_batch_host (plain malloc'ed memory works too, slower). */
void* dphlm_host_alloc(size_t bytes);
void dphlm_host_free(void* p);

/* Durations (ms, CUDA events on the launch stream) of this thread's last dphlm_fit_batch call that had DPHLM_FLAG_TIMING set:
 * ms[0] the pre-fit main kernel, ms[1] the lm() main kernel, ms[2] from the end of the lm() main kernel until the poller
 * kernels have finished the last straggler (the pollers run concurrently with the main kernels; a pre-fit that needs
 * hundreds of evaluations ends here).  Waits for that call to finish. */
int dphlm_last_kernel_ms(float ms[3]);

/* Measured FP32 throughput of the current device in TFLOP/s (independent FFMA chains, best of a few launches, ~10 ms):
 * the denominator bench.py reports beside the nominal SMs x 128 x 2 x clock. */
int dphlm_measure_fp32_peak(double* tflops);

int dphlm_device_count(void);
int dphlm_version(void);
const char* dphlm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* DPHLM_H */
