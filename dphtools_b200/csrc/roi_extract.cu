// roi_extract.cu -- the step upstream of the fitter: cut fixed-size ROIs around peak
// positions out of a stack of frames and derive the initial guesses, on the device.
//
// Reference pointers: slice_maker (dphtools/utils/__init__.py:244-295) defines the window -- centre rounded to
// the nearest integer, [c - w//2, c - w//2 + w) -- and localize_peak (:590-617) the parabola refinement of the
// centre through the row and the column that pass through the middle of the ROI.  slice_maker coerces a window
// that crosses the frame border into a smaller slice; a batch needs fixed-size ROIs, so such peaks are marked
// valid = 0 (their ROI is filled by clamping to the border) and left to the caller.
//
// One warp per ROI: coalesced row segments in, contiguous ROI out, max / min / parabola sums by warp shuffles.
// p0 = (max - min, x0, y0, sigma0, max(min, 0.5)) -- the recipe of BASELINE config 1.
#include <cuda_runtime.h>

#include <string>

#include "../../include/dphlm.h"

namespace dphlm {
int fail(const std::string& m);
}

namespace {

template <class T>
__global__ void __launch_bounds__(128) extract_kernel(const dphlm_extract_desc d) {
    const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= d.n_peaks) return;
    const int roi = d.roi, half = roi / 2, M = roi * roi;
    const int f = d.peaks[3 * n], r0 = d.peaks[3 * n + 1] - half, c0 = d.peaks[3 * n + 2] - half;
    const bool inside = f >= 0 && f < d.n_frames && r0 >= 0 && c0 >= 0 && r0 + roi <= d.height && c0 + roi <= d.width;
    const int fc = min(max(f, 0), d.n_frames - 1);
    const T* frame = static_cast<const T*>(d.frames) + (size_t)fc * d.height * d.width;
    float* out = d.rois + (size_t)n * M;
    float vmax = -3.4e38f, vmin = 3.4e38f;
    for (int i = lane; i < M; i += 32) {
        const int rr = min(max(r0 + i / roi, 0), d.height - 1), cc = min(max(c0 + i % roi, 0), d.width - 1);
        const float v = float(frame[(size_t)rr * d.width + cc]);
        out[i] = v;
        vmax = fmaxf(vmax, v);
        vmin = fminf(vmin, v);
    }
    // parabola through the centre row (x) and centre column (y): with symmetric abscissae t = -half..half,
    // y = a t^2 + b t + c has b = S(t y) / S(t^2), a = (n S(t^2 y) - S(t^2) S(y)) / (n S(t^4) - S(t^2)^2)
    float sx1 = 0.f, sx2 = 0.f, sx0 = 0.f, sy1 = 0.f, sy2 = 0.f, sy0 = 0.f;
    if (d.refine && lane < roi) {
        const float t = float(lane - half);
        const int rm = min(max(r0 + half, 0), d.height - 1), cm = min(max(c0 + half, 0), d.width - 1);
        const int rl = min(max(r0 + lane, 0), d.height - 1), cl = min(max(c0 + lane, 0), d.width - 1);
        const float vx = float(frame[(size_t)rm * d.width + cl]);  // along x (centre row)
        const float vy = float(frame[(size_t)rl * d.width + cm]);  // along y (centre column)
        sx0 = vx; sx1 = t * vx; sx2 = t * t * vx;
        sy0 = vy; sy1 = t * vy; sy2 = t * t * vy;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        sx0 += __shfl_xor_sync(0xffffffffu, sx0, o); sx1 += __shfl_xor_sync(0xffffffffu, sx1, o); sx2 += __shfl_xor_sync(0xffffffffu, sx2, o);
        sy0 += __shfl_xor_sync(0xffffffffu, sy0, o); sy1 += __shfl_xor_sync(0xffffffffu, sy1, o); sy2 += __shfl_xor_sync(0xffffffffu, sy2, o);
    }
    if (lane == 0) {
        float x0 = float(half), y0 = float(half);
        if (d.refine) {
            float t2 = 0.f, t4 = 0.f;
            for (int k = -half; k < roi - half; ++k) { t2 += float(k * k); t4 += float(k * k) * float(k * k); }
            const float den = roi * t4 - t2 * t2;
            const float ax = (roi * sx2 - t2 * sx0) / den, bx = sx1 / t2;
            const float ay = (roi * sy2 - t2 * sy0) / den, by = sy1 / t2;
            // a peak has negative curvature; otherwise keep the centre
            if (ax < 0.f) x0 = fminf(fmaxf(half - bx / (2.f * ax), 0.f), float(roi - 1));
            if (ay < 0.f) y0 = fminf(fmaxf(half - by / (2.f * ay), 0.f), float(roi - 1));
        }
        float* p = d.p0 + 5 * n;
        p[0] = vmax - vmin; p[1] = x0; p[2] = y0; p[3] = d.sigma0; p[4] = fmaxf(vmin, 0.5f);
        if (d.valid) d.valid[n] = inside ? 1 : 0;
    }
}

// Initial guess of multi_exp_fit (dphtools/utils/fitfuncs.py:63-80, 131-152) for a batch of histograms on one x axis: per
// segment of the x range (the reference's log-spaced split, the same for every histogram and passed in as bin bounds) a
// log-linear estimate of (amplitude, rate, offset): offset = min, straight line through log(data - offset) over the bins
// where that is finite, mirrored for a rising segment; the offsets of all but the last segment are dropped.  One warp per
// histogram, float64 sums (np.polyfit of degree 1 = the closed-form regression).
__global__ void __launch_bounds__(128) guess_multi_exp_kernel(const float* __restrict__ data, const float* __restrict__ xaxis,
                                                              const int* __restrict__ bounds, int n_seg, int with_offset, int M,
                                                              long long n, float* __restrict__ p0) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* y = data + row * M;
    const int P = 2 * n_seg + (with_offset ? 1 : 0);
    for (int sgm = 0; sgm < n_seg; ++sgm) {
        const int lo = bounds[sgm], hi = bounds[sgm + 1];
        double amp = 0.0, rate = 0.0, off = 0.0;
        if (hi > lo) {
            const float sign = y[lo] >= y[hi - 1] ? 1.0f : -1.0f;  // rising: estimate the mirrored decay
            float vmin = 3.4e38f;
            for (int i = lo + lane; i < hi; i += 32) vmin = fminf(vmin, sign * y[i]);
            for (int o = 16; o > 0; o >>= 1) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            double s0 = 0.0, sx = 0.0, sl = 0.0, sxx = 0.0, sxl = 0.0;
            for (int i = lo + lane; i < hi; i += 32) {
                const double d = double(sign * y[i]) - double(vmin);
                if (d > 0.0) {
                    const double l = log(d), x = double(xaxis[i]);
                    s0 += 1.0; sx += x; sl += l; sxx += x * x; sxl += x * l;
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o); sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sl += __shfl_xor_sync(0xffffffffu, sl, o); sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
                sxl += __shfl_xor_sync(0xffffffffu, sxl, o);
            }
            const double den = s0 * sxx - sx * sx;
            double m = 0.0, b = 0.0;
            if (den != 0.0) { m = (s0 * sxl - sx * sl) / den; b = (sl - m * sx) / s0; }
            amp = double(sign) * exp(b); rate = -m; off = double(sign) * double(vmin);
            if (!(amp == amp)) amp = 0.0;  // nan_to_num
            if (!(rate == rate)) rate = 0.0;
            amp = fmin(fmax(amp, -3.0e38), 3.0e38); rate = fmin(fmax(rate, -3.0e38), 3.0e38);
        }
        if (lane == 0) {
            p0[row * P + 2 * sgm] = float(amp);
            p0[row * P + 2 * sgm + 1] = float(rate);
            if (with_offset && sgm == n_seg - 1) p0[row * P + 2 * n_seg] = float(off);
        }
    }
}

}  // namespace

extern "C" int dphlm_guess_multi_exp(const float* data, const float* xaxis, const int32_t* seg_bounds, int32_t components, int32_t with_offset,
                                     int32_t n_bins, int64_t n_fits, float* p0, void* stream) {
    using dphlm::fail;
    if (n_fits < 0 || n_bins < 1 || components < 1 || components > 3) return fail("bad shape: 1..3 components, n_bins >= 1");
    if (n_fits == 0) return 0;
    if (!data || !xaxis || !seg_bounds || !p0) return fail("null data/xaxis/seg_bounds/p0 pointer");
    const long long grid = (n_fits * 32 + 127) / 128;
    guess_multi_exp_kernel<<<(unsigned)grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(data, xaxis, seg_bounds, components, with_offset,
                                                                                        n_bins, n_fits, p0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("guess_multi_exp_kernel launch: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int dphlm_extract_rois(const dphlm_extract_desc* d, void* stream) {
    using dphlm::fail;
    if (!d) return fail("null descriptor");
    if (d->n_peaks < 0 || d->n_frames < 1 || d->height < 1 || d->width < 1) return fail("bad frame stack shape");
    if (d->roi < 1 || d->roi > 31) return fail("roi must be 1..31 pixels");
    if (d->roi > d->height || d->roi > d->width) return fail("roi larger than the frame");
    // the parabola vertex is derived for abscissae symmetric about the window centre (localize_peak asserts odd shapes too)
    if (d->refine && d->roi % 2 == 0) return fail("refine needs an odd roi");
    if (d->frame_dtype != DPHLM_F32 && d->frame_dtype != DPHLM_U16) return fail("frame_dtype must be DPHLM_F32 or DPHLM_U16");
    if (d->n_peaks == 0) return 0;
    if (!d->frames || !d->peaks || !d->rois || !d->p0) return fail("null frames/peaks/rois/p0 pointer");
    const long long grid = (d->n_peaks * 32 + 127) / 128;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (d->frame_dtype == DPHLM_U16) extract_kernel<unsigned short><<<(unsigned)grid, 128, 0, s>>>(*d);
    else extract_kernel<float><<<(unsigned)grid, 128, 0, s>>>(*d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("extract_kernel launch: ") + cudaGetErrorString(e));
    return 0;
}
