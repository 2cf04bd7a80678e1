// lm_core.cuh -- the batched Levenberg-Marquardt / Poisson-MLE fitter, one fit per lane-group.
//
// This header is the whole numerical algorithm.  It is written once and compiled twice:
//   * by nvcc for sm_100a (lm_kernels.cu): a fit is owned by a group of G lanes of a warp, the
//     per-pixel state lives in shared memory, partial sums are combined with warp shuffles;
//   * by g++ with G = 1 (oracle/host_emu/host_emulation.cpp) so that the CPU test-suite can execute the very
//     same fp32 logic against the float64 oracle without a GPU.  That host build is test
//     infrastructure; the product path is the CUDA build only.
//
// What it replaces (reference file:line, upstream dphtools):
//   stage A  scipy.optimize.curve_fit -> MINPACK lmder/lmpar pre-fit   dphtools/utils/lm.py:642-648
//            (algorithm text: notebooks/lmder.f:186-452, notebooks/lmpar.f:100-264), restated over the
//            normal equations A = J^T J, g = J^T r instead of a QR factorisation;
//   stage B  lm()                                                      dphtools/utils/lm.py:221-507
//            with _chi2_ls/_update_ls (lm.py:30-52), _chi2_mle/_update_mle (lm.py:55-103) and the
//            non-negativity clamps (lm.py:106-132).
//
// fp32 and the 1.49e-8 tolerances.  The reference decides accept/reject and convergence from
// differences of chi-square values that are ~1e-8 of chi-square itself -- below fp32 resolution
// (SURVEY.md 7.3).  Here chi-square differences are never formed from two sums.  Every model provides
// delta(): f(x+d) - f(x) per pixel computed from the STEP d (relative error ~1e-6, not absolute
// error ~1e-7 f), and the reductions are accumulated per pixel as
//      LS : -delta (r + delta/2)              MLE: -delta + y log1p(delta / f_old)
// so that "actual" and "predicted" keep ~1e-6 relative accuracy down to 1e-10.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DPH_HD __host__ __device__ __forceinline__
#define DPH_HD_COLD __host__ __device__ __noinline__  // rarely executed code that must not weigh on the callers' registers
#else
#define DPH_HD inline
#define DPH_HD_COLD inline
#endif

#ifndef DPH_GAUSS2D_STORE
#define DPH_GAUSS2D_STORE false
#endif
#ifndef DPH_PAIR_UNROLL
#define DPH_PAIR_UNROLL 2  // unroll factor of the pixel-pair loop: two pairs in flight per lane (fits the register bound since the stash)
#endif

// per model (measured at 1e5 fits per call, one pair against two): elliptical Gaussian +11 %, IRF-convolved decay +14 %,
// pixel-integrated Gaussians +2..6 % with ONE pair in flight (their loop bodies are large: with two, instruction fetch
// -- `no_instruction` -- becomes the leading stall); multi-exponentials equal; sampled Gaussians +3-5 % with two
#ifndef DPH_UNROLL_ELLIP
#define DPH_UNROLL_ELLIP 1
#endif
#ifndef DPH_UNROLL_GINT
#define DPH_UNROLL_GINT 1
#endif
#ifndef DPH_UNROLL_MEXP
#define DPH_UNROLL_MEXP DPH_PAIR_UNROLL
#endif
#ifndef DPH_UNROLL_IRF
#define DPH_UNROLL_IRF 1
#endif

namespace dphlm {

constexpr int kPairUnroll = DPH_PAIR_UNROLL;

// ------------------------------------------------------------------------------------------ math
#if defined(__CUDA_ARCH__)
DPH_HD float f_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DPH_HD float f_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DPH_HD float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DPH_HD float f_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DPH_HD float f_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DPH_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
DPH_HD bool f_finite(float x) { return fabsf(x) < __int_as_float(0x7f800000); }
#else
DPH_HD float f_ex2(float x) { return exp2f(x); }
DPH_HD float f_lg2(float x) { return log2f(x); }
DPH_HD float f_rcp(float x) { return 1.0f / x; }
DPH_HD float f_rsqrt(float x) { return 1.0f / sqrtf(x); }
DPH_HD float f_sqrt(float x) { return sqrtf(x); }
DPH_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
DPH_HD bool f_finite(float x) { return std::isfinite(x); }
#endif

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kInf = __builtin_huge_valf();

// ------------------------------------------------------------------------- packed pairs (f32x2)
// The pixel passes run on PAIRS of adjacent pixels held in the two halves of a 64-bit register pair and
// use Blackwell's packed fp32 instructions (SASS FFMA2 / FMUL2 / FADD2, PTX fma.rn.f32x2, sm_100+): one
// issue slot per two FMAs.  Measured on B200 (scripts/microbench/fp32_peak.cu): scalar FFMA chains reach
// 42 TFLOP/s, FFMA2 chains 66 TFLOP/s of the nominal 74.  A per-fit scalar used as an operand is encoded
// as a broadcast source, so splatting costs nothing.  The same templates instantiated with T = float
// serve the odd tail pixel, the careful (irregular) pass and the host build.
struct alignas(8) f2 {
    float x, y;
    DPH_HD f2() {}
    DPH_HD f2(float s) : x(s), y(s) {}
    DPH_HD f2(float a, float b) : x(a), y(b) {}
};
#if defined(__CUDA_ARCH__)
DPH_HD f2 f_fma(f2 a, f2 b, f2 c) {
    float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    return f2(r.x, r.y);
}
DPH_HD f2 operator*(f2 a, f2 b) { float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return f2(r.x, r.y); }
DPH_HD f2 operator+(f2 a, f2 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return f2(r.x, r.y); }
DPH_HD f2 operator-(f2 a, f2 b) { return f_fma(b, f2(-1.0f), a); }
#else
DPH_HD f2 f_fma(f2 a, f2 b, f2 c) { return f2(f_fma(a.x, b.x, c.x), f_fma(a.y, b.y, c.y)); }
DPH_HD f2 operator*(f2 a, f2 b) { return f2(a.x * b.x, a.y * b.y); }
DPH_HD f2 operator+(f2 a, f2 b) { return f2(a.x + b.x, a.y + b.y); }
DPH_HD f2 operator-(f2 a, f2 b) { return f2(a.x - b.x, a.y - b.y); }
#endif
// component-wise helpers, defined for float and f2 alike
DPH_HD float v_ex2(float a) { return f_ex2(a); }
DPH_HD f2 v_ex2(f2 a) { return f2(f_ex2(a.x), f_ex2(a.y)); }
DPH_HD float v_rcp(float a) { return f_rcp(a); }
DPH_HD f2 v_rcp(f2 a) { return f2(f_rcp(a.x), f_rcp(a.y)); }
DPH_HD float v_lg2(float a) { return f_lg2(a); }
DPH_HD f2 v_lg2(f2 a) { return f2(f_lg2(a.x), f_lg2(a.y)); }
DPH_HD float v_erf(float a) { return erff(a); }
DPH_HD f2 v_erf(f2 a) { return f2(erff(a.x), erff(a.y)); }
DPH_HD float v_max(float a, float b) { return fmaxf(a, b); }
DPH_HD f2 v_max(f2 a, float b) { return f2(fmaxf(a.x, b), fmaxf(a.y, b)); }
DPH_HD float v_min2(float a, float b) { return fminf(a, b); }
DPH_HD f2 v_min2(f2 a, f2 b) { return f2(fminf(a.x, b.x), fminf(a.y, b.y)); }
// |a| <= thr ? x : y
DPH_HD float v_sel_small(float a, float thr, float x, float y) { return fabsf(a) <= thr ? x : y; }
DPH_HD f2 v_sel_small(f2 a, float thr, f2 x, f2 y) { return f2(fabsf(a.x) <= thr ? x.x : y.x, fabsf(a.y) <= thr ? x.y : y.y); }
DPH_HD float v_pick(bool c, float x, float y) { return c ? x : y; }
DPH_HD f2 v_pick(bool c, f2 x, f2 y) { return f2(c ? x.x : y.x, c ? x.y : y.y); }
DPH_HD float v_gather(const float* t, float i) { return t[int(i)]; }  // table lookup by (integer-valued) index
DPH_HD f2 v_gather(const float* t, f2 i) { return f2(t[int(i.x)], t[int(i.y)]); }
struct alignas(16) quad { float x, y, z, w; };
struct alignas(8) duo { float x, y; };
template <class T> struct Q4 { T a, b, c, d; };
template <class T> struct Q2 { T a, b; };
DPH_HD Q4<float> v_gather4(const float* t, float i) { const quad v = reinterpret_cast<const quad*>(t)[int(i)]; return {v.x, v.y, v.z, v.w}; }
DPH_HD Q4<f2> v_gather4(const float* t, f2 i) {
    const quad u = reinterpret_cast<const quad*>(t)[int(i.x)], v = reinterpret_cast<const quad*>(t)[int(i.y)];
    return {f2(u.x, v.x), f2(u.y, v.y), f2(u.z, v.z), f2(u.w, v.w)};
}
DPH_HD Q2<float> v_gather2(const float* t, float i) { const duo v = reinterpret_cast<const duo*>(t)[int(i)]; return {v.x, v.y}; }
DPH_HD Q2<f2> v_gather2(const float* t, f2 i) {
    const duo u = reinterpret_cast<const duo*>(t)[int(i.x)], v = reinterpret_cast<const duo*>(t)[int(i.y)];
    return {f2(u.x, v.x), f2(u.y, v.y)};
}
DPH_HD float v_clamp(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
DPH_HD f2 v_clamp(f2 a, float lo, float hi) { return f2(fminf(fmaxf(a.x, lo), hi), fminf(fmaxf(a.y, lo), hi)); }
DPH_HD float v_nonneg01(float a) { return a >= 0.0f ? 1.0f : 0.0f; }  // 1 where a >= 0, else 0
DPH_HD f2 v_nonneg01(f2 a) { return f2(a.x >= 0.0f ? 1.0f : 0.0f, a.y >= 0.0f ? 1.0f : 0.0f); }
DPH_HD float v_hsum(float a) { return a; }
DPH_HD float v_hsum(f2 a) { return a.x + a.y; }
DPH_HD float v_hmin(float a) { return a; }
DPH_HD float v_hmin(f2 a) { return fminf(a.x, a.y); }

// log1p(u): Taylor series of degree 6 for |u| <= 1/32 (truncation < 2e-10 relative), lg2-based beyond
// (absolute error ~2e-7 there, on steps that change the model by more than 3%).
constexpr float kLog1pSmall = 0.03125f;
template <class T>
DPH_HD T log1p_acc(T u) {
    T p = T(-1.0f / 6.0f);
    p = f_fma(p, u, T(0.2f));
    p = f_fma(p, u, T(-0.25f));
    p = f_fma(p, u, T(1.0f / 3.0f));
    p = f_fma(p, u, T(-0.5f));
    p = f_fma(p * u, u, u);
    T big = v_lg2(v_max(u + T(1.0f), 1e-37f)) * T(kLn2);
    return v_sel_small(u, kLog1pSmall, p, big);
}
// the same for callers that use the result only when 1 + u > 0 is known (the branch-free pass): no clamp
template <class T>
DPH_HD T log1p_pos(T u) {
    T p = T(-1.0f / 6.0f);
    p = f_fma(p, u, T(0.2f));
    p = f_fma(p, u, T(-0.25f));
    p = f_fma(p, u, T(1.0f / 3.0f));
    p = f_fma(p, u, T(-0.5f));
    p = f_fma(p * u, u, u);
    T big = v_lg2(u + T(1.0f)) * T(kLn2);
    return v_sel_small(u, kLog1pSmall, p, big);
}

// expm1(a) for |a| <= 1/16 (Taylor, degree 4, truncation a^4 / 120 < 1.3e-7 relative); callers handle larger |a| by
// direct differences of the stored exponentials.
template <class T>
DPH_HD T expm1_small(T a) {
    T p = T(1.0f / 24.0f);
    p = f_fma(p, a, T(1.0f / 6.0f));
    p = f_fma(p, a, T(0.5f));
    return f_fma(p * a, a, a);
}
constexpr float kSmallArg = 0.0625f;

// Chi-square differences of the branch-free pass, expanded about the ACCEPTED point.  With f0 the model value there,
// u = (f1 - f0) / f0 and the Poisson term t(f) = f - y ln f:
//   t(f0) - t(f1) = (y/f0 - 1) (f1 - f0) - y u^2 S(u),     S(u) = 1/2 - u/3 + u^2/4 - ...   (u - ln(1 + u) = u^2 S(u))
// The first-order term uses the very residual 1 - y/f0 the gradient at the accepted point was summed from (same inputs, same
// instructions), so first-order term and step stay consistent down to rounding: that, not the accuracy of the logarithm,
// is what keeps the signs of reductions ~1e-9 of chi-square right.  The series is short (see neg_series_ln); beyond the switch
// point |u| = 1/8 -- a step that changes a pixel's model value by more than 12 % -- a rarely taken branch substitutes the logarithm.
constexpr float kSeriesMax = 0.125f;
// -S(u), degree 2: the sign is folded into the coefficients.  Truncation u^3/5 of S: 8e-4 of the second-order term at the switch
// point, 1e-9 at the |u| ~ 1e-3 of the last trips, where the tolerances are decided (degrees 3 and 4 give the same flags and
// the same nfev on every golden fit).
template <class T>
DPH_HD T neg_series_ln(T u) {
    T p = T(-0.25f);
    p = f_fma(p, u, T(1.0f / 3.0f));
    return f_fma(p, u, T(-0.5f));
}
DPH_HD float v_absmax(float a, float b) { return fmaxf(fabsf(a), fabsf(b)); }
DPH_HD float v_absmax(f2 a, f2 b) { return fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y))); }
DPH_HD float v_absmax(float a) { return fabsf(a); }
DPH_HD float v_absmax(f2 a) { return fmaxf(fabsf(a.x), fabsf(a.y)); }

// e(new) - e(old) of an exponential whose argument changed by darg: e(old) expm1(darg), or the direct difference of the two
// values for a large change.  Both are computed and one is selected: putting the rare one behind a warp-uniform branch was
// measured slower (pre-fit kernel 0.304 -> 0.323 ms per 1e5 fits)
template <class T>
DPH_HD T exp_step(T darg, T ea, T eb) {
    return v_sel_small(darg, kSmallArg, ea * expm1_small(darg), eb - ea);
}

// ------------------------------------------------------------------------------- small dense algebra
// packed upper triangle, row-major: (i,j), i <= j
template <int P>
DPH_HD constexpr int tri(int i, int j) { return i * P - (i * (i - 1)) / 2 + (j - i); }
template <int P>
struct NT { static constexpr int value = P * (P + 1) / 2; };

// Cholesky factor of (A + lam * diag(D)): U^T U, stored packed with the RECIPROCAL diagonal.
// Returns false on a non-positive or non-finite pivot (the reference's LinAlgError branch, lm.py:426).
template <int P>
DPH_HD bool chol_factor(const float* A, const float* D, float lam, float* U) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        float d = f_fma(lam, D[i], A[tri<P>(i, i)]);
#pragma unroll
        for (int k = 0; k < i; ++k) d = f_fma(-U[tri<P>(k, i)], U[tri<P>(k, i)], d);
        ok = ok && (d > 0.0f) && f_finite(d);
        float rinv = f_rsqrt(d);
        U[tri<P>(i, i)] = rinv;
#pragma unroll
        for (int j = i + 1; j < P; ++j) {
            float s = A[tri<P>(i, j)];
#pragma unroll
            for (int k = 0; k < i; ++k) s = f_fma(-U[tri<P>(k, i)], U[tri<P>(k, j)], s);
            U[tri<P>(i, j)] = s * rinv;
        }
    }
    return ok;
}

// z = U^-T b (forward substitution), returns |z|^2
template <int P>
DPH_HD float chol_forward(const float* U, const float* b, float* z) {
    float n2 = 0.0f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        float s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = f_fma(-U[tri<P>(k, i)], z[k], s);
        z[i] = s * U[tri<P>(i, i)];
        n2 = f_fma(z[i], z[i], n2);
    }
    return n2;
}

// x = -(U^T U)^-1 g
template <int P>
DPH_HD void chol_solve_neg(const float* U, const float* g, float* x) {
    float z[P];
    chol_forward<P>(U, g, z);
#pragma unroll
    for (int i = P - 1; i >= 0; --i) {
        float s = z[i];
#pragma unroll
        for (int k = i + 1; k < P; ++k) s = f_fma(-U[tri<P>(i, k)], x[k], s);
        x[i] = s * U[tri<P>(i, i)];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) x[i] = -x[i];
}

template <int P>
DPH_HD float quad_form(const float* A, const float* p) {  // p^T A p
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        float t = 0.5f * A[tri<P>(i, i)] * p[i];
#pragma unroll
        for (int j = i + 1; j < P; ++j) t = f_fma(A[tri<P>(i, j)], p[j], t);
        s = f_fma(2.0f * p[i], t, s);
    }
    return s;
}

// --------------------------------------------------------------------------------------- models
// A model describes one pixel: prepared parameters (Par), the stored exponentials (NE per pixel),
// the model value, its Jacobian row, the accurate step difference delta() and J(x_old).d .
// Pixel coordinates arrive as (cx, cy); 1-D models use cx only.

#ifndef DPH_GAUSS2D_MINCTAS
#define DPH_GAUSS2D_MINCTAS 3
#endif
struct Gauss2D {  // amp, x0, y0, sigma, offset
    static constexpr int P = 5, NE = 1, kMinCtas = DPH_GAUSS2D_MINCTAS, kUnroll = kPairUnroll;
    struct Par { float amp, x0, y0, sig, off, s2, s3, nh, as2, as3, anh; };  // anh: exponent scale of the accepted point
    // ad1, ad2: the REALISED centre step (fill_tables); k1, k2: the exponent change is k1 dot + k2 + nds r2_old
    struct Step { float d0, d1, d2, d4, nds, c2, c3, k0, k1, k2, ad1, ad2; };
    template <class T> struct Geo { T dx, dy, r2; };
    template <class T> struct Diff { T dot, r20; };
    static constexpr int kTable = 0;
    static constexpr bool kStoreE = DPH_GAUSS2D_STORE;  // false: the accepted point's exponential is recomputed, nothing per pixel is kept
    DPH_HD static int table_floats(int, int) { return 0; }
    template <class L> DPH_HD static void fill_tables(const L&, const Par& q0, Par& q1, Step& st, float*, int, int, int) {
        st.ad1 = q1.x0 - q0.x0; st.ad2 = q1.y0 - q0.y0; q1.anh = q0.nh;
    }
    // The accepted point's exponential, from the r2 there that diff() forms out of the REALISED step (ad1 = x0_new - x0_old is
    // exact in float32, so dx + ad1 is cx - x0_old to one rounding, like in the pass that evaluated that point).  Forming r2_old
    // from the nominal step instead -- x0 + d rounds, the two differ by up to half an ulp of x0 -- makes this value inconsistent
    // with the one the gradient of the accepted point was summed from, which is enough to lose the last trips (tried: nfev
    // equality 95 % -> 77 %, 0.2 % of the fits no longer converge).
    template <class T> DPH_HD static void expo_old(const Par& q, const Geo<T>&, const Diff<T>& df, T* e) { e[0] = v_ex2(df.r20 * T(q.anh)); }
    DPH_HD static void set_consts(Par&, const float*) {}
    DPH_HD static void prepare(const float* p, Par& q) {
        q.amp = p[0]; q.x0 = p[1]; q.y0 = p[2]; q.sig = p[3]; q.off = p[4];
        float is = f_rcp(q.sig);
        q.s2 = is * is;
        q.s3 = q.s2 * is;
        q.nh = -0.5f * kLog2e * q.s2;
        q.as2 = q.amp * q.s2; q.as3 = q.amp * q.s3;
    }
    template <class T> DPH_HD static void geometry(const Par& q, T cx, T cy, Geo<T>& g) {
        g.dx = cx - T(q.x0); g.dy = cy - T(q.y0); g.r2 = f_fma(g.dx, g.dx, g.dy * g.dy);
    }
    template <class T> DPH_HD static void expo(const Par& q, const Geo<T>& g, T* e) { e[0] = v_ex2(g.r2 * T(q.nh)); }
    template <class T> DPH_HD static T value(const Par& q, const T* e) { return f_fma(T(q.amp), e[0], T(q.off)); }
    // Jacobian row WITHOUT the per-fit factors of its columns (amp/sigma^2, amp/sigma^2, amp/sigma^3): they are the same for
    // every pixel, so the sums A = J^T W J and g = J^T r are formed from (e, e dx, e dy, e r2, 1) and scaled once per pass
    // (post_scale), which takes two multiplications out of every pixel
    template <class T> DPH_HD static void jac(const Par&, const Geo<T>& g, const T* e, T* J) {
        J[0] = e[0]; J[1] = e[0] * g.dx; J[2] = e[0] * g.dy; J[3] = e[0] * g.r2; J[4] = T(1.0f);
    }
    DPH_HD static void post_scale(const Par& q, float* A, float* g) {
        const float s[P] = {1.0f, q.as2, q.as2, q.as3, 1.0f};
#pragma unroll
        for (int i = 0; i < P; ++i) {
#pragma unroll
            for (int j = i; j < P; ++j) A[i * P - (i * (i - 1)) / 2 + (j - i)] *= s[i] * s[j];
            g[i] *= s[i];
        }
    }
    // float64 model value and Jacobian row at one pixel (the tie-break pass of StageB, see precise_normal; the covariance kernel)
    DPH_HD static void eval64(const double* p, const float*, double cx, double cy, double& f, double* J) {
        const double dx = cx - p[1], dy = cy - p[2], r2 = dx * dx + dy * dy, s2 = p[3] * p[3];
        const double e = exp(-r2 / (2.0 * s2)), ae = p[0] * e;
        f = ae + p[4];
        J[0] = e; J[1] = ae * dx / s2; J[2] = ae * dy / s2; J[3] = ae * r2 / (s2 * p[3]); J[4] = 1.0;
    }
    DPH_HD static void prepare_step(const Par& a, const Par& b, const float* d, Step& s) {
        s.d0 = d[0]; s.d1 = d[1]; s.d2 = d[2]; s.d4 = d[4];
        const float nh1 = -0.5f * b.s2;                                   // -1/(2 sigma_new^2)
        s.nds = d[3] * (2.0f * a.sig + d[3]) * 0.5f * a.s2 * b.s2;        // -(1/(2 s_n^2) - 1/(2 s_o^2))
        const float ndd = -f_fma(d[1], d[1], d[2] * d[2]);
        s.c2 = a.amp * a.s2; s.c3 = a.amp * a.s3 * d[3];
        s.k0 = f_fma(-s.c2, ndd, d[0]);                                   // d0 + c2 (d1^2 + d2^2)
        // exponent change = nh1 (r2_new - r2_old) + nds r2_old, with r2_new - r2_old = -2 dot - (d1^2 + d2^2)
        s.k1 = -2.0f * nh1; s.k2 = ndd * nh1;
    }
    // per-pixel step quantities, from the geometry at the NEW point (dx_old = dx_new + step, ...)
    template <class T> DPH_HD static void diff(const Par&, const Step& s, const Geo<T>& g1, Diff<T>& df) {
        df.dot = f_fma(g1.dx, T(s.d1), g1.dy * T(s.d2));
        const T dxa = g1.dx + T(s.ad1), dya = g1.dy + T(s.ad2);
        df.r20 = f_fma(dxa, dxa, dya * dya);         // r2_old
    }
    template <class T> DPH_HD static T delta(const Par& a, const Step& s, const Diff<T>& df, const T* ea, const T* eb) {
        T darg = f_fma(df.dot, T(s.k1), f_fma(df.r20, T(s.nds), T(s.k2)));
        T de = exp_step(darg, ea[0], eb[0]);
        return f_fma(T(s.d0), eb[0], f_fma(T(a.amp), de, T(s.d4)));
    }
    template <class T> DPH_HD static T jdx(const Par&, const Step& s, const Diff<T>& df, const T* ea) {
        T lin = f_fma(T(s.c2), df.dot, f_fma(T(s.c3), df.r20, T(s.k0)));
        return f_fma(ea[0], lin, T(s.d4));
    }
};

// The sampled symmetric Gaussian with the width held at consts[0]: amp, x0, y0, offset (the
// fixed-width variant for the pixel-centre model; the erf form is Gauss2DIntegratedFixed).  Same step algebra as Gauss2D
// with d_sigma = 0.
struct Gauss2DFixed {
    static constexpr int P = 4, NE = 1, kMinCtas = 3, kUnroll = kPairUnroll;
    struct Par { float amp, x0, y0, off, sig = 1.0f, s2 = 1.0f, nh = -0.5f * 1.4426950408889634f, as2; };
    struct Step { float d0, d1, d2, d3, nh1, ndd, c2, k0, ad1, ad2; };  // ad1, ad2: the REALISED centre step (fill_tables)
    template <class T> struct Geo { T dx, dy, r2; };
    template <class T> struct Diff { T dot, dr2, r20; };
    static constexpr int kTable = 0;
    static constexpr bool kStoreE = false;  // the accepted point's exponential is recomputed, see Gauss2D::expo_old
    DPH_HD static int table_floats(int, int) { return 0; }
    template <class L> DPH_HD static void fill_tables(const L&, const Par& q0, Par& q1, Step& st, float*, int, int, int) {
        st.ad1 = q1.x0 - q0.x0; st.ad2 = q1.y0 - q0.y0;
    }
    template <class T> DPH_HD static void expo_old(const Par& q, const Geo<T>&, const Diff<T>& df, T* e) { e[0] = v_ex2(df.r20 * T(q.nh)); }
    DPH_HD static void set_consts(Par& q, const float* consts) {
        if (!consts) return;
        q.sig = consts[0];
        const float is = f_rcp(q.sig);
        q.s2 = is * is;
        q.nh = -0.5f * kLog2e * q.s2;
    }
    DPH_HD static void prepare(const float* p, Par& q) {
        q.amp = p[0]; q.x0 = p[1]; q.y0 = p[2]; q.off = p[3];
        q.as2 = q.amp * q.s2;
    }
    template <class T> DPH_HD static void geometry(const Par& q, T cx, T cy, Geo<T>& g) {
        g.dx = cx - T(q.x0); g.dy = cy - T(q.y0); g.r2 = f_fma(g.dx, g.dx, g.dy * g.dy);
    }
    template <class T> DPH_HD static void expo(const Par& q, const Geo<T>& g, T* e) { e[0] = v_ex2(g.r2 * T(q.nh)); }
    template <class T> DPH_HD static T value(const Par& q, const T* e) { return f_fma(T(q.amp), e[0], T(q.off)); }
    template <class T> DPH_HD static void jac(const Par&, const Geo<T>& g, const T* e, T* J) {  // unscaled columns, see Gauss2D::jac
        J[0] = e[0]; J[1] = e[0] * g.dx; J[2] = e[0] * g.dy; J[3] = T(1.0f);
    }
    DPH_HD static void post_scale(const Par& q, float* A, float* g) {
        const float s[P] = {1.0f, q.as2, q.as2, 1.0f};
#pragma unroll
        for (int i = 0; i < P; ++i) {
#pragma unroll
            for (int j = i; j < P; ++j) A[i * P - (i * (i - 1)) / 2 + (j - i)] *= s[i] * s[j];
            g[i] *= s[i];
        }
    }
    DPH_HD static void eval64(const double* p, const float* consts, double cx, double cy, double& f, double* J) {
        const double sig = consts ? double(consts[0]) : 1.0;
        const double dx = cx - p[1], dy = cy - p[2], r2 = dx * dx + dy * dy, s2 = sig * sig;
        const double e = exp(-r2 / (2.0 * s2)), ae = p[0] * e;
        f = ae + p[3];
        J[0] = e; J[1] = ae * dx / s2; J[2] = ae * dy / s2; J[3] = 1.0;
    }
    DPH_HD static void prepare_step(const Par& a, const Par&, const float* d, Step& s) {
        s.d0 = d[0]; s.d1 = d[1]; s.d2 = d[2]; s.d3 = d[3];
        s.nh1 = -0.5f * a.s2;
        s.ndd = -f_fma(d[1], d[1], d[2] * d[2]);
        s.c2 = a.amp * a.s2;
        s.k0 = f_fma(-s.c2, s.ndd, d[0]);  // d0 + c2 (d1^2 + d2^2)
    }
    template <class T> DPH_HD static void diff(const Par&, const Step& s, const Geo<T>& g1, Diff<T>& df) {
        df.dot = f_fma(g1.dx, T(s.d1), g1.dy * T(s.d2));
        df.dr2 = f_fma(T(-2.0f), df.dot, T(s.ndd));  // r2_new - r2_old
        const T dxa = g1.dx + T(s.ad1), dya = g1.dy + T(s.ad2);
        df.r20 = f_fma(dxa, dxa, dya * dya);         // r2_old, from the realised step (the width is a constant: same exponent scale)
    }
    template <class T> DPH_HD static T delta(const Par& a, const Step& s, const Diff<T>& df, const T* ea, const T* eb) {
        T darg = df.dr2 * T(s.nh1);
        T de = exp_step(darg, ea[0], eb[0]);
        return f_fma(T(s.d0), eb[0], f_fma(T(a.amp), de, T(s.d3)));
    }
    template <class T> DPH_HD static T jdx(const Par&, const Step& s, const Diff<T>& df, const T* ea) {
        return f_fma(ea[0], f_fma(T(s.c2), df.dot, T(s.k0)), T(s.d3));
    }
};

struct Gauss2DElliptical {  // amp, x0, y0, sigma_x, sigma_y, theta, offset
#ifndef DPH_ELLIP_MINCTAS
#define DPH_ELLIP_MINCTAS 2
#endif
    // (one pixel pair in flight per lane: with 35 packed accumulators the two-pair loop body outgrows the instruction cache --
    // `no_instruction` was the leading stall -- and costs 11 %)
    static constexpr int P = 7, NE = 1, kMinCtas = DPH_ELLIP_MINCTAS, kUnroll = DPH_UNROLL_ELLIP;
    // ax = 1/sx^2, ay = 1/sy^2; hax, hay: the same times -log2(e)/2 (exponent of ex2); axc .. ayc: products with cos / sin theta
    // o_*, ad1, ad2 (fill_tables): cos, sin and exponent scales of the accepted point and the REALISED centre step to it
    struct Par { float amp, x0, y0, sx, sy, th, off, c, s, ax, ay, isx, isy, hax, hay, axc, axs, nays, ayc, o_c, o_s, o_hax, o_hay, ad1, ad2; };
    // K1..K5: amp_old times the coefficients of u, v, u^2, v^2, u v in J(x_old) . d (see jdx)
    struct Step { float d0, d1, d2, d6, dc, dsn, haxb, hayb, hdax, hday, t1, t2, K1, K2, K3, K4, K5; };
    template <class T> struct Geo { T dx, dy, u, v, u2, v2; };
    template <class T> struct Diff { T dx, dy, u, v; };
    static constexpr int kTable = 0;
#ifndef DPH_ELLIP_STORE
#define DPH_ELLIP_STORE false
#endif
    // false: the accepted point's exponential is recomputed -- its rotated offsets from the REALISED centre step with the very
    // instruction sequence of geometry(), so bit for bit the value of the pass that evaluated that point (see Gauss2D::expo_old)
    // -- instead of kept per pixel in shared memory; that halves the shared memory of a fit and makes four-lane groups possible
    static constexpr bool kStoreE = DPH_ELLIP_STORE;
    DPH_HD static int table_floats(int, int) { return 0; }
    template <class L> DPH_HD static void fill_tables(const L&, const Par& q0, Par& q1, Step&, float*, int, int, int) {
        q1.o_c = q0.c; q1.o_s = q0.s; q1.o_hax = q0.hax; q1.o_hay = q0.hay;
        q1.ad1 = q1.x0 - q0.x0; q1.ad2 = q1.y0 - q0.y0;
    }
    template <class T> DPH_HD static void expo_old(const Par& q, const Geo<T>& g, const Diff<T>&, T* e) {
        const T dxa = g.dx + T(q.ad1), dya = g.dy + T(q.ad2);
        const T ua = f_fma(T(q.o_c), dxa, dya * T(q.o_s));
        const T va = f_fma(T(q.o_c), dya, dxa * T(-q.o_s));
        e[0] = v_ex2(f_fma(ua * ua, T(q.o_hax), (va * va) * T(q.o_hay)));
    }
    DPH_HD static void set_consts(Par&, const float*) {}
    DPH_HD static void prepare(const float* p, Par& q) {
        q.amp = p[0]; q.x0 = p[1]; q.y0 = p[2]; q.sx = p[3]; q.sy = p[4]; q.th = p[5]; q.off = p[6];
#if defined(__CUDA_ARCH__)
        sincosf(q.th, &q.s, &q.c);
#else
        q.s = sinf(q.th); q.c = cosf(q.th);
#endif
        q.isx = f_rcp(q.sx); q.isy = f_rcp(q.sy);
        q.ax = q.isx * q.isx; q.ay = q.isy * q.isy;
        q.hax = -0.5f * kLog2e * q.ax; q.hay = -0.5f * kLog2e * q.ay;
        q.axc = q.ax * q.c; q.axs = q.ax * q.s; q.nays = -q.ay * q.s; q.ayc = q.ay * q.c;
    }
    template <class T> DPH_HD static void geometry(const Par& q, T cx, T cy, Geo<T>& g) {
        g.dx = cx - T(q.x0); g.dy = cy - T(q.y0);
        g.u = f_fma(T(q.c), g.dx, g.dy * T(q.s));
        g.v = f_fma(T(q.c), g.dy, g.dx * T(-q.s));
        g.u2 = g.u * g.u; g.v2 = g.v * g.v;
    }
    template <class T> DPH_HD static void expo(const Par& q, const Geo<T>& g, T* e) {
        e[0] = v_ex2(f_fma(g.u2, T(q.hax), g.v2 * T(q.hay)));
    }
    template <class T> DPH_HD static T value(const Par& q, const T* e) { return f_fma(T(q.amp), e[0], T(q.off)); }
    // Jacobian row without the per-fit factors of its columns (see Gauss2D::jac): (e, e (u ax c - v ay s), e (u ax s + v ay c),
    // e u^2, e v^2, e u v, 1); post_scale applies (1, amp, amp, amp ax / sx, amp ay / sy, amp (ay - ax), 1) to the sums
    template <class T> DPH_HD static void jac(const Par& q, const Geo<T>& g, const T* e, T* J) {
        J[0] = e[0];
        J[1] = e[0] * f_fma(g.u, T(q.axc), g.v * T(q.nays));
        J[2] = e[0] * f_fma(g.u, T(q.axs), g.v * T(q.ayc));
        J[3] = e[0] * g.u2;
        J[4] = e[0] * g.v2;
        J[5] = e[0] * (g.u * g.v);
        J[6] = T(1.0f);
    }
    DPH_HD static void post_scale(const Par& q, float* A, float* g) {
        const float s[P] = {1.0f, q.amp, q.amp, q.amp * q.ax * q.isx, q.amp * q.ay * q.isy, q.amp * (q.ay - q.ax), 1.0f};
#pragma unroll
        for (int i = 0; i < P; ++i) {
#pragma unroll
            for (int j = i; j < P; ++j) A[i * P - (i * (i - 1)) / 2 + (j - i)] *= s[i] * s[j];
            g[i] *= s[i];
        }
    }
    DPH_HD static void eval64(const double* p, const float*, double cx, double cy, double& f, double* J) {
        const double c = cos(p[5]), s = sin(p[5]), dx = cx - p[1], dy = cy - p[2];
        const double u = c * dx + s * dy, v = -s * dx + c * dy, ax = 1.0 / (p[3] * p[3]), ay = 1.0 / (p[4] * p[4]);
        const double e = exp(-0.5 * (u * u * ax + v * v * ay)), ae = p[0] * e;
        f = ae + p[6];
        J[0] = e; J[1] = ae * (u * ax * c - v * ay * s); J[2] = ae * (u * ax * s + v * ay * c);
        J[3] = ae * u * u * ax / p[3]; J[4] = ae * v * v * ay / p[4]; J[5] = ae * u * v * (ay - ax); J[6] = 1.0;
    }
    DPH_HD static void prepare_step(const Par& a, const Par& b, const float* d, Step& s) {
        s.d0 = d[0]; s.d1 = d[1]; s.d2 = d[2]; s.d6 = d[6];
        // cos/sin increments through the half-angle identities (no cancellation)
        float hs, hc;
        float thm = a.th + 0.5f * d[5];
#if defined(__CUDA_ARCH__)
        sincosf(thm, &hs, &hc);
#else
        hs = sinf(thm); hc = cosf(thm);
#endif
        float sh = sinf(0.5f * d[5]);
        s.dc = -2.0f * hs * sh;
        s.dsn = 2.0f * hc * sh;
        // exponent change = -1/2 (d(u^2) ax_new + u_old^2 d(ax) + the same in v); the factor -1/2 is folded in here
        s.haxb = -0.5f * b.ax; s.hayb = -0.5f * b.ay;
        s.hdax = 0.5f * d[3] * (2.0f * a.sx + d[3]) * a.ax * b.ax;
        s.hday = 0.5f * d[4] * (2.0f * a.sy + d[4]) * a.ay * b.ay;
        s.t1 = -f_fma(b.c, d[1], b.s * d[2]);   // rotation of the centre shift, new frame
        s.t2 = -f_fma(b.c, d[2], -b.s * d[1]);
        // J(x_old) . d = e_old (d0 + amp (u K1' + v K2' + u^2 K3' + v^2 K4' + u v K5')) + d6 with the old point's u, v
        s.K1 = a.amp * f_fma(a.axc, d[1], a.axs * d[2]);
        s.K2 = a.amp * f_fma(a.nays, d[1], a.ayc * d[2]);
        s.K3 = a.amp * a.ax * a.isx * d[3];
        s.K4 = a.amp * a.ay * a.isy * d[4];
        s.K5 = a.amp * (a.ay - a.ax) * d[5];
    }
    template <class T> DPH_HD static void diff(const Par& a, const Step& s, const Geo<T>& g1, Diff<T>& df) {
        df.dx = g1.dx + T(s.d1); df.dy = g1.dy + T(s.d2);  // offsets from the OLD centre
        df.u = f_fma(T(a.c), df.dx, df.dy * T(a.s));
        df.v = f_fma(T(a.c), df.dy, df.dx * T(-a.s));
    }
    template <class T> DPH_HD static T delta(const Par& a, const Step& s, const Diff<T>& ga, const T* ea, const T* eb) {
        T du = f_fma(T(s.dc), ga.dx, f_fma(T(s.dsn), ga.dy, T(s.t1)));
        T dv = f_fma(T(s.dc), ga.dy, f_fma(T(-s.dsn), ga.dx, T(s.t2)));
        T du2 = du * f_fma(T(2.0f), ga.u, du);
        T dv2 = dv * f_fma(T(2.0f), ga.v, dv);
        T darg = f_fma(du2, T(s.haxb), (ga.u * ga.u) * T(s.hdax)) + f_fma(dv2, T(s.hayb), (ga.v * ga.v) * T(s.hday));
        T de = exp_step(darg, ea[0], eb[0]);
        return f_fma(T(s.d0), eb[0], f_fma(T(a.amp), de, T(s.d6)));
    }
    template <class T> DPH_HD static T jdx(const Par&, const Step& s, const Diff<T>& ga, const T* ea) {
        T t1 = f_fma(ga.v, T(s.K5), f_fma(ga.u, T(s.K3), T(s.K1)));
        T t2 = f_fma(ga.v, T(s.K4), T(s.K2));
        T lin = f_fma(ga.u, t1, f_fma(ga.v, t2, T(s.d0)));
        return f_fma(ea[0], lin, T(s.d6));
    }
};

// Pixel-integrated symmetric Gaussian (erf model of a pixelated camera; SURVEY.md 8(f) rank 4, cf. upstream
// notebooks/MLE_LevMarq_4Param_2013_11_12.m:44-46): amp * Ex(x) * Ey(y) + offset with
// Ex = (erf((x - x0 + 1/2) a) - erf((x - x0 - 1/2) a)) / 2, a = 1 / (sqrt(2) sigma).  amp, x0, y0, sigma, offset.
// The model is SEPARABLE, so every transcendental is evaluated per column and per row of the ROI (w + h axis positions per
// trial point instead of w h pixels): fill_tables() computes, cooperatively over the lanes of the group and into the group's
// shared memory, for every axis position
//   A = (E_b, dE_b/dc, dE_b/dsigma, E_b - E_a)   at the trial point b,     B = (E_a, J_axis(a) . d)   at the accepted point a,
// and the pixel loop only gathers and multiplies.  The axis difference E_b - E_a uses the step difference any function
// with a derivative can use: for small steps E(x + d) - E(x) = (E'(x) + E'(x + d)) d / 2 + O(d^3) (trapezoid rule on the
// derivatives both points need anyway; relative error ~ (d/scale)^2 / 12), for large steps the direct difference; the
// switch sits where both have the same absolute error in fp32 (~1e-5 of the difference).
// Then  f_b - f_a = d_amp E_bx E_by + amp_a (dEx E_by + E_ax dEy) + d_off  without cancellation.
template <bool FIXED>
struct Gauss2DIntegratedT {
    static constexpr int P = FIXED ? 4 : 5, NE = 1, kMinCtas = 3, kOff = P - 1, kUnroll = DPH_UNROLL_GINT;
    static constexpr int kAxisMax = 32;                // ROI side limit (validated by the entry points)
    static constexpr int kTable = 10 * 2 * kAxisMax;   // static bound, used by the covariance kernel
    static constexpr bool kStoreE = false;             // the accepted point's pixel value is E_ax E_ay from the tables
    DPH_HD static int table_floats(int h, int w) { return (10 * (h + w) + 3) & ~3; }
    struct Par {
        float amp, x0, y0, sig = 1.0f, off, a = 0.70710678118654752f, c = 0.39894228040143268f, cs = 0.39894228040143268f;
        const float* tbl = nullptr;   // axis table of the trial point this Par describes (fill_tables)
        const float* tblb = nullptr;  // (E_accepted, J_axis . d) pairs
        float w = 0.0f;               // ROI width: rows follow the columns in the tables
    };
    struct Step { float d[P]; bool direct; };
    template <class T> struct Geo { T ex, ey, dex, dey, sex, sey, dEx, dEy, exa, eya, ux, uy; };
    template <class T> struct Diff { T dEx, dEy, exa, eya, eyb, ux, uy; };
    DPH_HD static void set_sigma(Par& q, float sig) {
        q.sig = sig;
        const float is = f_rcp(sig);
        q.a = 0.70710678118654752f * is;
        q.c = 0.39894228040143268f * is;
        q.cs = q.c * is;
    }
    // FIXED: the width is a constant of the call (consts[0]); prepare() leaves it alone
    DPH_HD static void set_consts(Par& q, const float* consts) {
        if (FIXED && consts) set_sigma(q, consts[0]);
    }
    DPH_HD static void prepare(const float* p, Par& q) {
        q.amp = p[0]; q.x0 = p[1]; q.y0 = p[2]; q.off = p[kOff];
        if (!FIXED) set_sigma(q, p[3]);
    }
    // pixel integral of the unit-area Gaussian along one axis and its derivatives w.r.t. centre and sigma
    DPH_HD static void axis(const Par& q, float t, float& e, float& de, float& se) {
        const float tp = t + 0.5f, tm = t - 0.5f;
        const float zp = tp * q.a, zm = tm * q.a;
        e = (erff(zp) - erff(zm)) * 0.5f;
        const float gp = f_ex2(zp * zp * -kLog2e), gm = f_ex2(zm * zm * -kLog2e);
        de = (gm - gp) * q.c;
        se = (tm * gm - tp * gp) * q.cs;
    }
    // Tables: two buffers of (E, dE/dc, dE/dsigma, E - E_accepted) per axis position -- the accepted point and the trial point;
    // accepting a step flips which is which (phase bit 0, the same index as the per-pixel buffers of the other models), so
    // every trip evaluates the axis integrals at ONE point -- and a (E_accepted, J_axis(accepted) . d) pair per position.
    template <class L> DPH_HD static void fill_tables(const L& ln, const Par& q0, Par& q1, Step& st, float* tbl, int h, int w, int phase) {
        const int n = h + w, cur = phase & 1;
        const bool init = phase & 2;
        quad* acc = reinterpret_cast<quad*>(tbl) + (cur ? n : 0);
        quad* trial = reinterpret_cast<quad*>(tbl) + (cur ? 0 : n);
        duo* B = reinterpret_cast<duo*>(tbl + 8 * n);
        q1.tbl = reinterpret_cast<const float*>(trial); q1.tblb = reinterpret_cast<const float*>(B); q1.w = float(w);
        ln.sync();  // the previous pass has finished reading the tables
        for (int t = ln.lane; t < n; t += L::kLanes) {
            const bool row = t >= w;
            const float pos = float(row ? t - w : t);
            const float dc = row ? st.d[2] : st.d[1], ds = FIXED ? 0.0f : st.d[FIXED ? 0 : 3];
            float ea, dea, sea, eb, deb, seb;
            axis(q1, pos - (row ? q1.y0 : q1.x0), eb, deb, seb);
            if (init) {  // first pass of a fit (fresh or resumed): accepted point = trial point, zero step
                ea = eb; dea = deb; sea = seb;
                acc[t] = quad{ea, dea, sea, 0.0f};
            } else {
                const quad a = acc[t];
                ea = a.x; dea = a.y; sea = a.z;
            }
            // small step: trapezoid rule on the derivatives at both ends, which are needed anyway
            const float dE = st.direct ? eb - ea : 0.5f * f_fma(dea + deb, dc, (sea + seb) * ds);
            trial[t] = quad{eb, deb, seb, dE};
            B[t] = duo{ea, f_fma(dea, dc, sea * ds)};
        }
        ln.sync();
    }
    template <class T> DPH_HD static void geometry(const Par& q, T cx, T cy, Geo<T>& g) {
        const T r = cy + T(q.w);
        const Q4<T> ca = v_gather4(q.tbl, cx), ra = v_gather4(q.tbl, r);
        const Q2<T> cb = v_gather2(q.tblb, cx), rb = v_gather2(q.tblb, r);
        g.ex = ca.a; g.dex = ca.b; g.sex = ca.c; g.dEx = ca.d;
        g.ey = ra.a; g.dey = ra.b; g.sey = ra.c; g.dEy = ra.d;
        g.exa = cb.a; g.ux = cb.b; g.eya = rb.a; g.uy = rb.b;
    }
    template <class T> DPH_HD static void expo(const Par&, const Geo<T>& g, T* e) { e[0] = g.ex * g.ey; }
    template <class T> DPH_HD static void expo_old(const Par&, const Geo<T>& g, const Diff<T>&, T* e) { e[0] = g.exa * g.eya; }  // at the accepted point
    template <class T> DPH_HD static T value(const Par& q, const T* e) { return f_fma(T(q.amp), e[0], T(q.off)); }
    template <class T> DPH_HD static void jac(const Par& q, const Geo<T>& g, const T* e, T* J) {
        J[0] = e[0];
        J[1] = g.dex * g.ey * T(q.amp);
        J[2] = g.ex * g.dey * T(q.amp);
        if (!FIXED) J[3] = f_fma(g.sex, g.ey, g.ex * g.sey) * T(q.amp);
        J[kOff] = T(1.0f);
    }
    DPH_HD static void post_scale(const Par&, float*, float*) {}
    DPH_HD static void axis64(double t, double sig, double& e, double& de, double& se) {
        const double a = 0.70710678118654752440 / sig, c = 0.39894228040143267794 / sig;
        const double tp = t + 0.5, tm = t - 0.5, zp = tp * a, zm = tm * a;
        e = 0.5 * (erf(zp) - erf(zm));
        const double gp = exp(-zp * zp), gm = exp(-zm * zm);
        de = c * (gm - gp);
        se = c / sig * (tm * gm - tp * gp);
    }
    DPH_HD static void eval64(const double* p, const float* consts, double cx, double cy, double& f, double* J) {
        const double sig = FIXED ? (consts ? double(consts[0]) : 1.0) : p[3];
        double ex, dex, sex, ey, dey, sey;
        axis64(cx - p[1], sig, ex, dex, sex);
        axis64(cy - p[2], sig, ey, dey, sey);
        f = p[0] * ex * ey + p[kOff];
        J[0] = ex * ey; J[1] = p[0] * dex * ey; J[2] = p[0] * ex * dey;
        if (!FIXED) J[3] = p[0] * (sex * ey + ex * sey);
        J[kOff] = 1.0;
    }
    DPH_HD static void prepare_step(const Par& a, const Par&, const float* d, Step& s) {
#pragma unroll
        for (int k = 0; k < P; ++k) s.d[k] = d[k];
        const float is = f_rcp(fabsf(a.sig));
        float rel = fmaxf(fabsf(d[1]), fabsf(d[2])) * is;
        if (!FIXED) rel = fmaxf(rel, fabsf(d[3]) * is);
        s.direct = !(rel <= 0.011f);
    }
    template <class T> DPH_HD static void diff(const Par&, const Step&, const Geo<T>& g1, Diff<T>& df) {
        df.dEx = g1.dEx; df.dEy = g1.dEy; df.exa = g1.exa; df.eya = g1.eya; df.eyb = g1.ey; df.ux = g1.ux; df.uy = g1.uy;
    }
    template <class T> DPH_HD static T delta(const Par& a, const Step& s, const Diff<T>& df, const T*, const T* eb) {
        const T de = f_fma(df.dEx, df.eyb, df.exa * df.dEy);  // E_bx E_by - E_ax E_ay
        return f_fma(T(s.d[0]), eb[0], f_fma(T(a.amp), de, T(s.d[kOff])));
    }
    template <class T> DPH_HD static T jdx(const Par& a, const Step& s, const Diff<T>& df, const T* ea) {
        const T u = f_fma(df.ux, df.eya, df.exa * df.uy);     // J_(x0, y0, sigma)(a) . d / amp
        return f_fma(T(s.d[0]), ea[0], f_fma(T(a.amp), u, T(s.d[kOff])));
    }
};
using Gauss2DIntegrated = Gauss2DIntegratedT<false>;       // (photons, x0, y0, sigma, offset)
using Gauss2DIntegratedFixed = Gauss2DIntegratedT<true>;   // (photons, x0, y0, offset), sigma = consts[0]

// sum of NEXP exponentials (+ offset): a0, k0, a1, k1, ..., [offset]   (fitfuncs.py:20-52)
#ifndef DPH_MEXP_STORE
#define DPH_MEXP_STORE false
#endif
template <int NEXP, bool OFFSET>
struct MultiExp {
    static constexpr int P = 2 * NEXP + (OFFSET ? 1 : 0), NE = NEXP, kMinCtas = NEXP >= 3 ? 2 : 3, kUnroll = DPH_UNROLL_MEXP;
    struct Par { float a[NEXP], k[NEXP], off, kl[NEXP], kla[NEXP]; };  // kl = -log2(e) k; kla: the same of the accepted point
    struct Step { float da[NEXP], dk[NEXP], doff; };
    template <class T> struct Geo { T x; };
    template <class T> struct Diff { T x; };
    static constexpr int kTable = 0;
    // false: the accepted point's exponentials are recomputed from its rates (the same instruction on the same inputs as in the
    // pass that evaluated that point, so bit for bit the values the gradient was summed from) instead of kept per bin in shared
    // memory: two more MUFU per bin and exponential against a 64-bit load and store
    static constexpr bool kStoreE = DPH_MEXP_STORE;
    DPH_HD static int table_floats(int, int) { return 0; }
    template <class L> DPH_HD static void fill_tables(const L&, const Par& q0, Par& q1, Step&, float*, int, int, int) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) q1.kla[i] = q0.kl[i];
    }
    DPH_HD static void set_consts(Par&, const float*) {}
    DPH_HD static void prepare(const float* p, Par& q) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) { q.a[i] = p[2 * i]; q.k[i] = p[2 * i + 1]; q.kl[i] = -kLog2e * q.k[i]; }
        q.off = OFFSET ? p[2 * NEXP] : 0.0f;
    }
    template <class T> DPH_HD static void geometry(const Par&, T cx, T, Geo<T>& g) { g.x = cx; }
    template <class T> DPH_HD static void expo(const Par& q, const Geo<T>& g, T* e) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) e[i] = v_ex2(g.x * T(q.kl[i]));
    }
    template <class T> DPH_HD static void expo_old(const Par& q, const Geo<T>& g, const Diff<T>&, T* e) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) e[i] = v_ex2(g.x * T(q.kla[i]));
    }
    template <class T> DPH_HD static T value(const Par& q, const T* e) {
        T f = T(q.off);
#pragma unroll
        for (int i = 0; i < NEXP; ++i) f = f_fma(T(q.a[i]), e[i], f);
        return f;
    }
    // Jacobian row without the per-fit factor -a_i of the rate columns (see Gauss2D::jac): (e_i, x e_i, ..., 1)
    template <class T> DPH_HD static void jac(const Par&, const Geo<T>& g, const T* e, T* J) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) { J[2 * i] = e[i]; J[2 * i + 1] = g.x * e[i]; }
        if (OFFSET) J[2 * NEXP] = T(1.0f);
    }
    DPH_HD static void post_scale(const Par& q, float* A, float* g) {
        float s[P];
#pragma unroll
        for (int i = 0; i < NEXP; ++i) { s[2 * i] = 1.0f; s[2 * i + 1] = -q.a[i]; }
        if (OFFSET) s[2 * NEXP] = 1.0f;
#pragma unroll
        for (int i = 0; i < P; ++i) {
#pragma unroll
            for (int j = i; j < P; ++j) A[i * P - (i * (i - 1)) / 2 + (j - i)] *= s[i] * s[j];
            g[i] *= s[i];
        }
    }
    DPH_HD static void eval64(const double* p, const float*, double cx, double, double& f, double* J) {
        f = OFFSET ? p[2 * NEXP] : 0.0;
#pragma unroll
        for (int i = 0; i < NEXP; ++i) {
            const double e = exp(-p[2 * i + 1] * cx);
            f += p[2 * i] * e;
            J[2 * i] = e; J[2 * i + 1] = -p[2 * i] * cx * e;
        }
        if (OFFSET) J[2 * NEXP] = 1.0;
    }
    DPH_HD static void prepare_step(const Par&, const Par&, const float* d, Step& s) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) { s.da[i] = d[2 * i]; s.dk[i] = d[2 * i + 1]; }
        s.doff = OFFSET ? d[2 * NEXP] : 0.0f;
    }
    template <class T> DPH_HD static void diff(const Par&, const Step&, const Geo<T>& g1, Diff<T>& df) { df.x = g1.x; }
    template <class T> DPH_HD static T delta(const Par& a, const Step& s, const Diff<T>& ga, const T* ea, const T* eb) {
        T r = T(s.doff);
#pragma unroll
        for (int i = 0; i < NEXP; ++i) {
            T darg = ga.x * T(-s.dk[i]);
            T de = exp_step(darg, ea[i], eb[i]);
            r = f_fma(T(s.da[i]), eb[i], f_fma(T(a.a[i]), de, r));
        }
        return r;
    }
    template <class T> DPH_HD static T jdx(const Par& a, const Step& s, const Diff<T>& ga, const T* ea) {
        T r = T(s.doff);
#pragma unroll
        for (int i = 0; i < NEXP; ++i) r = f_fma(ea[i], f_fma(ga.x, T(-a.a[i] * s.dk[i]), T(s.da[i])), r);
        return r;
    }
};

// Sum of NEXP exponentials convolved with an instrument response function (IRF) sampled on the histogram's own uniform
// grid (own formulation, not derived from the GPL supplement in the upstream notebooks):
//   f_j = offset + sum_i a_i C_i[j],   C_i[j] = sum_{q=0}^{min(j-m0, W-1)} irf[q] exp(-k_i (j - m0 - q) dt)   (0 for j < m0)
// with the W IRF taps starting at bin m0.  Parameters as MultiExp: a0, k0, a1, k1, ..., [offset].
// consts = [W, m0, dt, irf[0..W-1]].  With u = (j - m0) dt and e = exp(-k u):
//   C = e S[w],  D = -dC/dk = u C - e T[w],   w = min(j - m0, W-1),
//   S[w] = sum_{q<=w} irf[q] exp(k q dt),  T[w] = sum_{q<=w} irf[q] q dt exp(k q dt)
// so one exponential per bin and two prefix-sum tables per exponential and trial point (fill_tables: a shuffle scan over
// the lanes of the group, in the group's shared memory; the tables of the accepted point are kept, accepting a step flips the
// two buffers).  C and D of both points are rebuilt per bin from the tables, so the model keeps no per-bin state, J(x) d at the
// accepted point comes from its C and D, and f(x+d) - f(x) uses dC = -dk (D_a + D_b)/2 (trapezoid rule on dC/dk,
// relative error (dk u)^2 / 12) for small dk u and the direct difference otherwise.
constexpr int kIrfTaps = 32;
template <int NEXP, bool OFFSET>
struct MultiExpIrf {
    static constexpr int P = 2 * NEXP + (OFFSET ? 1 : 0), NE = 2 * NEXP, kMinCtas = 3, kUnroll = DPH_UNROLL_IRF;
    static constexpr int kHalf = 2 * NEXP * kIrfTaps;  // one buffer: S and T per exponential
    static constexpr int kTable = 2 * kHalf;           // accepted point and trial point; accepting a step flips which is which
    static constexpr bool kStoreE = false;             // C and D of the accepted point are rebuilt from its tables
    DPH_HD static int table_floats(int, int) { return kTable; }
    static constexpr float kTrapz = 0.009f;  // |dk u| below which the trapezoid rule beats the direct difference in fp32
    struct Par {
        float a[NEXP], k[NEXP], off;
        const float* consts = nullptr;  // [W, m0, dt, irf...]
        const float* tbl = nullptr;     // [NEXP][2][kIrfTaps] of this trial point
        const float* tbla = nullptr;    // the same of the accepted point
        float ka[NEXP];                 // rates of the accepted point
        float m0 = 0.0f, dt = 1.0f, wmax = 0.0f;
    };
    struct Step { float da[NEXP], dk[NEXP], doff; };
    template <class T> struct Geo { T u, w, in; };
    template <class T> struct Diff { T u; };
    DPH_HD static void set_consts(Par& q, const float* c) {
        q.consts = c;
        if (!c) return;
        const float w = fminf(fmaxf(c[0], 1.0f), float(kIrfTaps));
        q.wmax = w - 1.0f; q.m0 = c[1]; q.dt = c[2];
    }
    DPH_HD static void prepare(const float* p, Par& q) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) { q.a[i] = p[2 * i]; q.k[i] = p[2 * i + 1]; }
        q.off = OFFSET ? p[2 * NEXP] : 0.0f;
    }
    // prefix sums S, T over the IRF taps for every exponential of the point q, by the lanes of the group
    template <class L> DPH_HD static void fill_tables(const L& ln, const Par& q0, Par& q, Step&, float* tbl, int, int, int phase) {
        const int cur = phase & 1;
        const bool init = phase & 2;
        float* trial = tbl + (cur ? 0 : kHalf);
        float* acc = tbl + (cur ? kHalf : 0);
        q.tbl = trial; q.tbla = acc;
#pragma unroll
        for (int i = 0; i < NEXP; ++i) q.ka[i] = q0.k[i];
        if (!q.consts) return;  // idle group
        constexpr int G = L::kLanes;
        ln.sync();  // the previous pass has finished reading the tables
#pragma unroll
        for (int i = 0; i < NEXP; ++i) {
            float carry_s = 0.0f, carry_t = 0.0f;
            const float kd = q.k[i] * q.dt * kLog2e;
            for (int base = 0; base < kIrfTaps; base += G) {
                const int w = base + ln.lane;
                const float tap = float(w) <= q.wmax ? q.consts[3 + w] : 0.0f;
                float s = tap * f_ex2(kd * float(w));
                float t = s * (float(w) * q.dt);
#pragma unroll
                for (int o = 1; o < G; o <<= 1) {
                    const float us = ln.up(s, o), ut = ln.up(t, o);
                    if (ln.lane >= o) { s += us; t += ut; }
                }
                s += carry_s; t += carry_t;
                trial[(2 * i) * kIrfTaps + w] = s;
                trial[(2 * i + 1) * kIrfTaps + w] = t;
                if (init) {  // first pass of a fit: the accepted point is the trial point
                    acc[(2 * i) * kIrfTaps + w] = s;
                    acc[(2 * i + 1) * kIrfTaps + w] = t;
                }
                carry_s = ln.bcast(s, G - 1);
                carry_t = ln.bcast(t, G - 1);
            }
        }
        ln.sync();
    }
    // cy carries the bin index of 1-D data
    template <class T> DPH_HD static void geometry(const Par& q, T, T cy, Geo<T>& g) {
        const T jj = cy - T(q.m0);
        g.u = jj * T(q.dt);
        g.in = v_nonneg01(jj);
        g.w = v_clamp(jj, 0.0f, q.wmax);
    }
    template <class T> DPH_HD static void conv(const float* tbl, const float* k, const Geo<T>& g, T* e) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) {
            const T ex = v_ex2(g.u * T(-kLog2e * k[i])) * g.in;
            const T c = ex * v_gather(tbl + (2 * i) * kIrfTaps, g.w);
            e[2 * i] = c;
            e[2 * i + 1] = g.u * c - ex * v_gather(tbl + (2 * i + 1) * kIrfTaps, g.w);
        }
    }
    template <class T> DPH_HD static void expo(const Par& q, const Geo<T>& g, T* e) { conv(q.tbl, q.k, g, e); }
    template <class T> DPH_HD static void expo_old(const Par& q, const Geo<T>& g, const Diff<T>&, T* e) { conv(q.tbla, q.ka, g, e); }  // accepted point
    template <class T> DPH_HD static T value(const Par& q, const T* e) {
        T f = T(q.off);
#pragma unroll
        for (int i = 0; i < NEXP; ++i) f = f_fma(T(q.a[i]), e[2 * i], f);
        return f;
    }
    template <class T> DPH_HD static void jac(const Par& q, const Geo<T>&, const T* e, T* J) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) { J[2 * i] = e[2 * i]; J[2 * i + 1] = e[2 * i + 1] * T(-q.a[i]); }
        if (OFFSET) J[2 * NEXP] = T(1.0f);
    }
    DPH_HD static void post_scale(const Par&, float*, float*) {}
    // cy carries the bin index; consts = [W, m0, dt, irf...]
    DPH_HD static void eval64(const double* p, const float* consts, double, double cy, double& f, double* J) {
        f = OFFSET ? p[2 * NEXP] : 0.0;
        if (OFFSET) J[2 * NEXP] = 1.0;
        const int W = consts ? int(fminf(fmaxf(consts[0], 1.0f), float(kIrfTaps))) : 1;
        const double m0 = consts ? double(consts[1]) : 0.0, dt = consts ? double(consts[2]) : 1.0;
#pragma unroll
        for (int i = 0; i < NEXP; ++i) {
            double c = 0.0, d = 0.0;
            for (int q = 0; q < W; ++q) {
                const double lag = (cy - m0 - double(q)) * dt;
                if (lag >= 0.0) {
                    const double term = double(consts[3 + q]) * exp(-p[2 * i + 1] * lag);
                    c += term; d += term * lag;
                }
            }
            f += p[2 * i] * c;
            J[2 * i] = c; J[2 * i + 1] = -p[2 * i] * d;
        }
    }
    DPH_HD static void prepare_step(const Par&, const Par&, const float* d, Step& s) {
#pragma unroll
        for (int i = 0; i < NEXP; ++i) { s.da[i] = d[2 * i]; s.dk[i] = d[2 * i + 1]; }
        s.doff = OFFSET ? d[2 * NEXP] : 0.0f;
    }
    template <class T> DPH_HD static void diff(const Par&, const Step&, const Geo<T>& g1, Diff<T>& df) { df.u = g1.u; }
    template <class T> DPH_HD static T delta(const Par& a, const Step& s, const Diff<T>& df, const T* ea, const T* eb) {
        T r = T(s.doff);
#pragma unroll
        for (int i = 0; i < NEXP; ++i) {
            const T trap = (ea[2 * i + 1] + eb[2 * i + 1]) * T(-0.5f * s.dk[i]);
            const T dc = v_sel_small(df.u * T(s.dk[i]), kTrapz, trap, eb[2 * i] - ea[2 * i]);
            r = f_fma(T(s.da[i]), eb[2 * i], f_fma(T(a.a[i]), dc, r));
        }
        return r;
    }
    template <class T> DPH_HD static T jdx(const Par& a, const Step& s, const Diff<T>&, const T* ea) {
        T r = T(s.doff);
#pragma unroll
        for (int i = 0; i < NEXP; ++i) r = f_fma(ea[2 * i], T(s.da[i]), f_fma(ea[2 * i + 1], T(-a.a[i] * s.dk[i]), r));
        return r;
    }
};

// ------------------------------------------------------------------------------ lane group / storage
// Per-fit pixel state, pointer based so that the device (shared memory) and host builds share code.
//   y[M]               data (MLE: already clamped to >= 0)
//   e[2][NE][stride]   stored exponentials, double buffered: `cur` is the accepted point, 1-cur the trial
//   cx[M], cy[M]       pixel coordinates, shared by all fits (cy unused by 1-D models)
// All arrays are 8-byte aligned and indexed by pixel, so that a pair of adjacent pixels is one 64-bit access.

template <int NE>
struct PixelState {
    float* y;
    float* e;  // [2][NE][stride]
    const float* cx;
    const float* cy;
    int M, stride;
    int cur;
    float* tbl = nullptr;  // Mo::table_floats(h, w): per-trial tables of models that need them (fill_tables), else unused
    int w = 1;             // row length of a 2-D ROI (M / w rows)
    float* stash = nullptr;  // kStashFloats per group: where the per-fit scalars the pixel loop does not need wait during a pass
    DPH_HD float* ebuf(int which, int j) const { return e + (which * NE + j) * stride; }
};

#if defined(__CUDA_ARCH__)
DPH_HD int float_as_int(float f) { return __float_as_int(f); }
DPH_HD float int_as_float(int i) { return __int_as_float(i); }
#else
DPH_HD int float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
DPH_HD float int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
#endif

// Entry k of the coordinate tables shared by all fits of a CTA: the pixel's coordinates (cy = bin index for 1-D data).
template <class Mo>
DPH_HD void fill_coords(float* cx, float* cy, int k, int h, int w, const float* xaxis) {
    const int M = h * w, i = k < M ? k : M - 1;
    (void)h;
    cx[k] = xaxis ? xaxis[i] : float(i % w);
    cy[k] = xaxis ? float(i) : float(i / w);
}

// Per-fit scalars that the pixel loop does not touch (the normal equations of the accepted point, its gradient, the scaling
// and the step: 30 floats for 5 parameters) are parked in the group's shared memory for the duration of a pass and read back
// after it.  The pass runs at the register bound of three CTAs per SM; without this the compiler recomputes loop-invariant
// step constants inside the loop instead of keeping them (DPH_STASH=0 builds the old form).
#ifndef DPH_STASH
#define DPH_STASH 1
#endif
// floats per group: everything (30 for 5 parameters) where shared memory allows it; the normal equations alone for two-lane
// groups (64 fits per CTA leave 19 floats per group before a third CTA no longer fits on the SM) and for models that keep
// per-pixel values of the accepted point in shared memory (measured: the full size costs the fixed-width Gaussian its third CTA)
DPH_HD constexpr int stash_floats(int G, bool store_e) { return G >= 4 && !store_e ? 32 : 16; }
DPH_HD void stash_store(float* s, const float* v, int n) {
#pragma unroll
    for (int i = 0; i < n; ++i) s[i] = v[i];
}
DPH_HD void stash_load(const float* s, float* v, int n) {
#pragma unroll
    for (int i = 0; i < n; ++i) v[i] = *reinterpret_cast<const volatile float*>(s + i);
}

// Group of G lanes.  On the host G == 1 and the reductions are the identity.
template <int G>
struct Lanes {
    static constexpr int kLanes = G;
    int lane;        // index inside the group
    unsigned gmask;  // the warp lanes of this group
#if defined(__CUDA_ARCH__)
    DPH_HD float sum(float v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    DPH_HD int any(int v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    DPH_HD static bool warp_any(bool v) { return __any_sync(0xffffffffu, v); }
    DPH_HD double sum_group(double v) const {  // only the lanes of this group take part (group-divergent code)
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        return v;
    }
    DPH_HD float up(float v, int o) const { return __shfl_up_sync(gmask, v, o, G); }   // value of lane - o (own value below o)
    DPH_HD float bcast(float v, int src) const { return __shfl_sync(gmask, v, src, G); }  // value of group lane src
    DPH_HD void sync() const { __syncwarp(gmask); }
#else
    DPH_HD double sum_group(double v) const { return v; }
    DPH_HD float up(float v, int) const { return v; }
    DPH_HD float bcast(float v, int) const { return v; }
    DPH_HD void sync() const {}
    DPH_HD float sum(float v) const { return v; }
    DPH_HD int any(int v) const { return v; }
    DPH_HD static bool warp_any(bool v) { return v; }
#endif
    template <int N>
    DPH_HD void sum_all(float* v) const {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = sum(v[i]);
    }
};

// ------------------------------------------------------------------------------------ pixel passes
constexpr int kFlagNonFinite = 1, kFlagPredNeg = 2, kFlagIrregular = 4;

template <int P>
struct Normal {  // per-lane partial sums of one pass, group totals after the reduction
    float A[NT<P>::value];
    float g[P];
    float actual, predicted;  // chi2(old) - chi2(trial), chi2(old) - chi2(trial + J_old d)
    float chi1;               // chi2(trial) itself (LS: always; MLE: only when asked for)
    int flags;
    DPH_HD void clear() {
#pragma unroll
        for (int i = 0; i < NT<P>::value; ++i) A[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < P; ++i) g[i] = 0.0f;
        actual = predicted = chi1 = 0.0f;
        flags = 0;
    }
    template <int G>
    DPH_HD void reduce(const Lanes<G>& ln) {
        ln.template sum_all<NT<P>::value>(A);
        ln.template sum_all<P>(g);
        actual = ln.sum(actual);
        predicted = ln.sum(predicted);
        chi1 = ln.sum(chi1);
        flags = ln.any(flags);
    }
};

// partial sums of one lane while it walks its pixels; T = f2 for pixel pairs, float for single pixels
template <int P, class T>
struct Acc {
    T A[NT<P>::value];
    T g[P];
    T actual, predicted, chi1;
    T lo;  // smallest model value met (regular case: stays > 0)
    DPH_HD void clear() {
#pragma unroll
        for (int i = 0; i < NT<P>::value; ++i) A[i] = T(0.0f);
#pragma unroll
        for (int i = 0; i < P; ++i) g[i] = T(0.0f);
        actual = predicted = chi1 = T(0.0f);
        lo = T(1.0f);
    }
    DPH_HD void add(const T* J, T w, T res) {
#pragma unroll
        for (int i = 0; i < P; ++i) {
            T wj = w * J[i];
#pragma unroll
            for (int j = i; j < P; ++j) A[tri<P>(i, j)] = f_fma(wj, J[j], A[tri<P>(i, j)]);
            g[i] = f_fma(J[i], res, g[i]);
        }
    }
};

// Poisson term  t(f) = f - y ln(f / y)  with the reference's zeroing rule (lm.py:70-78); slow path.
DPH_HD float mle_term_direct(float f, float y) {
    float l = (y > 0.0f && f > 0.0f) ? y * kLn2 * (f_lg2(f) - f_lg2(y)) : 0.0f;
    return f - l;
}
DPH_HD f2 mle_term_direct(f2 f, f2 y) { return f2(mle_term_direct(f.x, y.x), mle_term_direct(f.y, y.y)); }
DPH_HD float clamp_nonneg(float f) { return f <= 0.0f ? 0.0f : f; }  // lm.py:117
// y ln(q) for q = y / f > 0, and 0 where y = 0 (the reference's zeroing rule)
DPH_HD float v_ylog(float y, float q) { return y > 0.0f ? y * kLn2 * f_lg2(q) : 0.0f; }
DPH_HD f2 v_ylog(f2 y, f2 q) { return f2(v_ylog(y.x, q.x), v_ylog(y.y, q.y)); }

// MLE weights at a model value f (already clamped): w = y/f^2, res = 1 - y/f; invalid -> (0, 1)  (lm.py:86-102)
DPH_HD void mle_weights(float f, float y, float& w, float& res) {
    float rf = f_rcp(f);
    float yf = y * rf;
    float ww = yf * rf;
    bool ok = (f > 0.0f) && f_finite(ww);
    w = ok ? ww : 0.0f;
    res = ok ? 1.0f - yf : 1.0f;
}

// One pixel (T = float) or one pair of adjacent pixels (T = f2) of the regular, branch-free trial
// evaluation: model at the trial point q1 = q0 + d, the chi-square reductions from per-pixel differences
// (see neg_series_ln) and, speculatively, the normal equations at the trial point.  QUIRK (lm.py:446-449): the
// predicted model is f(x_new) + J(x_old) d.
template <class Mo, bool MLE, bool PRED, class T>
DPH_HD void pixel_regular(const typename Mo::Par& q0, const typename Mo::Par& q1, const typename Mo::Step& st, bool init,
                          bool want_chi, const typename Mo::template Geo<T>& g1, T y, const T* e0in, T* e1, Acc<Mo::P, T>& n) {
    constexpr int P = Mo::P, NE = Mo::NE;
    typename Mo::template Diff<T> df;
    T e0[NE], J[P];
    Mo::expo(q1, g1, e1);
    Mo::diff(q0, st, g1, df);
    if constexpr (Mo::kStoreE) {
#pragma unroll
        for (int j = 0; j < NE; ++j) e0[j] = v_pick(init, e1[j], e0in[j]);
    } else {
        Mo::expo_old(q1, g1, df, e0);  // on the first pass the accepted point is the trial point (zero step, equal tables)
    }
    const T f0 = Mo::value(q0, e0);
    const T f1 = Mo::value(q1, e1);
    const T dl = Mo::delta(q0, st, df, e0, e1);
    Mo::jac(q1, g1, e1, J);
    if (MLE) {
        // The sums of this pass are used only if every model value involved is positive: the old one (the driver repeats the
        // pass in the literal form whenever the accepted point had a non-positive pixel), the trial one and the predicted one
        // (tracked in n.lo; the clamp turns a NaN into a 0 that the minimum sees).  So 1 + u = f_new / f_old > 0 where it counts.
        const T f1c = v_max(f1, 0.0f);
        const T rf0 = v_rcp(f0);
        const T yf0 = y * rf0;
        const T rn0 = yf0 - T(1.0f);  // -(1 - y/f0): minus the residual the gradient at the accepted point was summed from 
        // t(f0) - t(f1) = rn0 dl - (y/f0) dl u S(u)   (see neg_series_ln above).  The first-order term goes into the sum on its own,
        // with the very rn0 (an exact product inside the multiply-add): folding it into one factor q = rn0 - (y/f0) u S(u) rounds
        // q and costs 0.3 % of nfev equality on the golden sets; the second-order term follows as (y u dl) (-S(u)).
        const T u = dl * rf0;
        const T ma = (yf0 * u) * dl;                       // y u^2:  -y (u - ln(1 + u)) = ma (-S(u))
        n.actual = f_fma(rn0, dl, n.actual) + ma * neg_series_ln(u);
        if (PRED) {
            const T jd = Mo::jdx(q0, st, df, e0);
            const T dp = dl + jd;
            const T up = dp * rf0;
            const T mp = (yf0 * up) * dp;
            n.predicted = f_fma(rn0, dp, n.predicted) + mp * neg_series_ln(up);
            if (v_absmax(u, up) > kSeriesMax) {  // a step that changes this pixel by more than 12 %: rare -- the logarithm replaces the series
                const T la = y * f_fma(v_lg2(u + T(1.0f)), T(kLn2), T(0.0f) - u) - ma * neg_series_ln(u);
                const T lp = y * f_fma(v_lg2(up + T(1.0f)), T(kLn2), T(0.0f) - up) - mp * neg_series_ln(up);
                n.actual = n.actual + v_sel_small(u, kSeriesMax, T(0.0f), la);
                n.predicted = n.predicted + v_sel_small(up, kSeriesMax, T(0.0f), lp);
                n.lo = v_min2(n.lo, f1c + jd);  // the predicted value f0 (1 + up) can only reach 0 on such a step
            }
        } else if (v_absmax(u) > kSeriesMax) {
            const T la = y * f_fma(v_lg2(u + T(1.0f)), T(kLn2), T(0.0f) - u) - ma * neg_series_ln(u);
            n.actual = n.actual + v_sel_small(u, kSeriesMax, T(0.0f), la);
        }
        n.lo = v_min2(n.lo, f1c);
        const T rf1 = v_rcp(f1c);
        const T yf = y * rf1;
        n.add(J, yf * rf1, T(1.0f) - yf);
        // chi-square itself (first pass of a fit only): f - y - y ln(f / y) = (f - y) + y ln(y / f), with the y / f the weights
        // are made of -- one logarithm per pixel; y = 0 contributes f (lm.py:70-78)
        if (want_chi) n.chi1 = n.chi1 + ((f1c - y) + v_ylog(y, yf));
    } else {
        const T r0 = f0 - y, r1 = f1 - y;
        n.actual = n.actual - dl * f_fma(T(0.5f), dl, r0);
        if (PRED) {
            const T dp = dl + Mo::jdx(q0, st, df, e0);
            n.predicted = n.predicted - dp * f_fma(T(0.5f), dp, r0);
        }
        n.chi1 = f_fma(r1 * T(0.5f), r1, n.chi1);
        n.add(J, T(1.0f), r1);
    }
}

// Evaluation at the trial point for all pixels of the fit owned by this lane group.  Reads the exponentials
// of the accepted point (buffer cur), writes the trial ones (buffer 1-cur).  With `init` the accepted point
// does not exist yet: the pass is run with d = 0 and the old exponentials are taken equal to the new ones
// (actual = predicted = 0).
//
// FAST walks pairs of adjacent pixels with packed f32x2 arithmetic (pixel_regular<f2>) and treats only the
// regular case (all model values > 0, finite); a group that meets anything else raises kFlagIrregular and
// the caller repeats the pass with FAST = false, which implements the reference's clamps and zeroing rules
// literally, pixel by pixel.
template <class Mo, bool MLE, bool PRED, bool FAST, int G>
DPH_HD void trial_pass(const Lanes<G>& ln, const PixelState<Mo::NE>& px, const typename Mo::Par& q0,
                       const typename Mo::Par& q1, const typename Mo::Step& st, bool init, bool want_chi,
                       Normal<Mo::P>& n) {
    constexpr int P = Mo::P, NE = Mo::NE;
    const float* e0b = px.ebuf(px.cur, 0);
    float* e1b = px.ebuf(1 - px.cur, 0);
    const int stride = px.stride;
    n.clear();
    if (FAST) {
        Acc<P, f2> a2;
        a2.clear();
        Acc<P, float> a1;
        const int npair = px.M >> 1;
#if defined(__CUDA_ARCH__)
#pragma unroll Mo::kUnroll
#endif
        for (int q = ln.lane; q < npair; q += G) {
            const int i = 2 * q;
            f2 e0[NE], e1[NE];
            if constexpr (Mo::kStoreE) {
#pragma unroll
                for (int j = 0; j < NE; ++j) e0[j] = *reinterpret_cast<const f2*>(e0b + j * stride + i);
            }
            const f2 cx = *reinterpret_cast<const f2*>(px.cx + i);
            const f2 cy = *reinterpret_cast<const f2*>(px.cy + i);
            const f2 y = *reinterpret_cast<const f2*>(px.y + i);
            typename Mo::template Geo<f2> g1;
            Mo::geometry(q1, cx, cy, g1);
            pixel_regular<Mo, MLE, PRED, f2>(q0, q1, st, init, want_chi, g1, y, e0, e1, a2);
            if constexpr (Mo::kStoreE) {
#pragma unroll
                for (int j = 0; j < NE; ++j) *reinterpret_cast<f2*>(e1b + j * stride + i) = e1[j];
            }
        }
#pragma unroll
        for (int k = 0; k < NT<P>::value; ++k) a1.A[k] = v_hsum(a2.A[k]);
#pragma unroll
        for (int k = 0; k < P; ++k) a1.g[k] = v_hsum(a2.g[k]);
        a1.actual = v_hsum(a2.actual); a1.predicted = v_hsum(a2.predicted); a1.chi1 = v_hsum(a2.chi1);
        a1.lo = v_hmin(a2.lo);
        if ((px.M & 1) && ln.lane == npair % G) {  // odd pixel count: the last pixel alone
            const int i = px.M - 1;
            float e0[NE], e1[NE];
            if constexpr (Mo::kStoreE) {
#pragma unroll
                for (int j = 0; j < NE; ++j) e0[j] = e0b[j * stride + i];
            }
            typename Mo::template Geo<float> g1;
            Mo::geometry(q1, px.cx[i], px.cy[i], g1);
            pixel_regular<Mo, MLE, PRED, float>(q0, q1, st, init, want_chi, g1, px.y[i], e0, e1, a1);
            if constexpr (Mo::kStoreE) {
#pragma unroll
                for (int j = 0; j < NE; ++j) e1b[j * stride + i] = e1[j];
            }
        }
#pragma unroll
        for (int k = 0; k < NT<P>::value; ++k) n.A[k] = a1.A[k];
#pragma unroll
        for (int k = 0; k < P; ++k) n.g[k] = a1.g[k];
        n.actual = a1.actual; n.predicted = a1.predicted; n.chi1 = a1.chi1;
        if (MLE && !(a1.lo > 0.0f)) n.flags |= kFlagIrregular;  // also catches NaN
    } else {
        Acc<P, float> a1;
        a1.clear();
        for (int i = ln.lane; i < px.M; i += G) {
            typename Mo::template Geo<float> g1;
            typename Mo::template Diff<float> df;
            float e0[NE], e1[NE], J[P];
            Mo::geometry(q1, px.cx[i], px.cy[i], g1);
            Mo::expo(q1, g1, e1);
            if constexpr (Mo::kStoreE) {
#pragma unroll
                for (int j = 0; j < NE; ++j) {
                    float t = e0b[j * stride + i];
                    e0[j] = init ? e1[j] : t;
                    e1b[j * stride + i] = e1[j];
                }
            }
            const float y = px.y[i];
            Mo::diff(q0, st, g1, df);
            if constexpr (!Mo::kStoreE) Mo::expo_old(q1, g1, df, e0);
            const float f0 = Mo::value(q0, e0);
            const float f1 = Mo::value(q1, e1);
            const float dl = Mo::delta(q0, st, df, e0, e1);
            Mo::jac(q1, g1, e1, J);
            if (!f_finite(f1)) n.flags |= kFlagNonFinite;
            if (MLE) {
                const float f0c = clamp_nonneg(f0), f1c = clamp_nonneg(f1);
                const float rf0 = f_rcp(f0c);
                const bool reg = (f0c > 0.0f) && (f1c > 0.0f);
                if (reg) a1.actual += f_fma(y, log1p_acc(dl * rf0), -dl);
                else a1.actual += mle_term_direct(f0c, y) - mle_term_direct(f1c, y);
                if (PRED) {
                    const float jd = Mo::jdx(q0, st, df, e0);
                    const float fp = f1c + jd;
                    if (fp < 0.0f) n.flags |= kFlagPredNeg;
                    if (reg && fp > 0.0f) {
                        const float dp = dl + jd;
                        a1.predicted += f_fma(y, log1p_acc(dp * rf0), -dp);
                    } else {
                        a1.predicted += mle_term_direct(f0c, y) - mle_term_direct(fp, y);
                    }
                }
                float w, res;
                mle_weights(f1c, y, w, res);
                a1.add(J, w, res);
                if (want_chi) a1.chi1 += mle_term_direct(f1c, y) - y;
            } else {
                const float r0 = f0 - y, r1 = f1 - y;
                a1.actual = f_fma(-dl, f_fma(0.5f, dl, r0), a1.actual);
                if (PRED) {
                    const float dp = dl + Mo::jdx(q0, st, df, e0);
                    a1.predicted = f_fma(-dp, f_fma(0.5f, dp, r0), a1.predicted);
                }
                a1.chi1 = f_fma(0.5f * r1, r1, a1.chi1);
                a1.add(J, 1.0f, r1);
            }
        }
#pragma unroll
        for (int k = 0; k < NT<P>::value; ++k) n.A[k] = a1.A[k];
#pragma unroll
        for (int k = 0; k < P; ++k) n.g[k] = a1.g[k];
        n.actual = a1.actual; n.predicted = a1.predicted; n.chi1 = a1.chi1;
    }
    n.reduce(ln);
    Mo::post_scale(q1, n.A, n.g);  // per-fit column factors the model kept out of its Jacobian rows (no-op for most models)
    if (!MLE && !f_finite(n.chi1)) n.flags |= kFlagNonFinite;
}

// ---------------------------------------------------------------------------------- tie-break pass
// The reference ends lm() with info = 2 when |dx| <= xtol (|x| + xtol), xtol = 1.49e-8 (lm.py:365-367, 430-432): a step four
// times smaller than float32 resolves in x itself.  Whether that test fires is decided by the Newton step at the accepted
// point to ~1e-9 |x|, and the float32 trajectory carries ~3e-8 |x| of noise into that point (the gradient of the previous
// trip, summed from float32 model values, moved it by that much).  Left alone, the fits whose last step is within a small
// factor of the threshold -- up to 2 % of a batch for the 4-parameter models -- get a coin flip between info 1 and info 2.
// With Options::exact_ties a proposal within kTieBand of the threshold is re-decided in float64 (the one place float64 is
// used): the previous accepted point (kept in the working buffer's spare words) is taken as exact, its normal equations and
// gradient are summed in float64 and the damped step from there is repeated -- that removes the noise the float32 gradient
// put into the accepted point -- and the test is made with the float64 step at the point so reached.  Measured on the
// golden sets: info equality 98.0 -> 99.8 % (4-parameter models), 99.85 -> 99.95 % (5 parameters); cost: two float64 passes
// for the ~5 % of the fits that come that close.
template <class Mo, bool MLE, int G>
DPH_HD void precise_normal(const Lanes<G>& ln, const PixelState<Mo::NE>& px, const double* p, const float* consts, double* A, double* g) {
    constexpr int P = Mo::P;
#pragma unroll
    for (int k = 0; k < NT<P>::value; ++k) A[k] = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) g[k] = 0.0;
    for (int i = ln.lane; i < px.M; i += G) {
        double f, J[P];
        Mo::eval64(p, consts, double(px.cx[i]), double(px.cy[i]), f, J);
        const double y = double(px.y[i]);  // MLE: clamped when the fit was staged (lm.py:127)
        double w = 1.0, res;
        if (MLE) {
            if (f <= 0.0) f = 0.0;  // lm.py:132
            const double yf = y / f, yf2 = yf / f;
            const bool ok = (yf - yf == 0.0) && (yf2 - yf2 == 0.0);  // lm.py:86-102: invalid pixel -> weight 0, residual 1
            w = ok ? yf2 : 0.0;
            res = ok ? 1.0 - yf : 1.0;
        } else {
            res = f - y;
        }
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const double wj = w * J[r];
#pragma unroll
            for (int c = r; c < P; ++c) A[tri<P>(r, c)] += wj * J[c];
            g[r] += J[r] * res;
        }
    }
#pragma unroll
    for (int k = 0; k < NT<P>::value; ++k) A[k] = ln.sum_group(A[k]);
#pragma unroll
    for (int k = 0; k < P; ++k) g[k] = ln.sum_group(g[k]);
}

// d = -(A + lam diag(D))^-1 g in float64 (Cholesky); false if the matrix is not positive definite
template <int P>
DPH_HD bool solve_damped64(const double* A, const float* D, double lam, const double* g, double* d) {
    double U[NT<P>::value];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        double piv = A[tri<P>(i, i)] + lam * double(D[i]);
#pragma unroll
        for (int k = 0; k < i; ++k) piv -= U[tri<P>(k, i)] * U[tri<P>(k, i)];
        if (!(piv > 0.0)) return false;
        const double r = 1.0 / sqrt(piv);
        U[tri<P>(i, i)] = r;
#pragma unroll
        for (int j = i + 1; j < P; ++j) {
            double t = A[tri<P>(i, j)];
    #pragma unroll
        for (int k = 0; k < i; ++k) t -= U[tri<P>(k, i)] * U[tri<P>(k, j)];
            U[tri<P>(i, j)] = t * r;
        }
    }
    double z[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        double t = -g[i];
#pragma unroll
        for (int k = 0; k < i; ++k) t -= U[tri<P>(k, i)] * z[k];
        z[i] = t * U[tri<P>(i, i)];
    }
#pragma unroll
    for (int i = P - 1; i >= 0; --i) {
        double t = z[i];
#pragma unroll
        for (int k = i + 1; k < P; ++k) t -= U[tri<P>(i, k)] * d[k];
        d[i] = t * U[tri<P>(i, i)];
    }
    return true;
}

// ------------------------------------------------------------------------------------- options/result
struct Options {
    float ftol, xtol, gtol, factor;
    int maxfev;      // <= 0: 100 * (P + 1)
    int exact_ties;  // != 0: re-decide borderline xtol tests in float64 (precise_normal)
};

// ------------------------------------------------------------------------- stage A: MINPACK lmder
// lmpar (notebooks/lmpar.f:100-264) over the normal equations: find par >= 0 such that the step
// p = -(A + par D)^-1 g has |sqrt(D) p| within 10% of delta.  D holds the SQUARED scale factors.
template <int P>
DPH_HD void lmpar_normal(const float* A, const float* g, const float* D, float delta, float& par, float* p, float& pnorm) {
    constexpr float dwarf = 1.17549435e-38f;
    float U[NT<P>::value], w[P], z[P];
    bool full_rank = chol_factor<P>(A, D, 0.0f, U);
    float dxnorm, fp;
    if (full_rank) {
        chol_solve_neg<P>(U, g, p);
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < P; ++i) s = f_fma(D[i] * p[i], p[i], s);
        dxnorm = f_sqrt(s);
        fp = dxnorm - delta;
        if (fp <= 0.1f * delta) { par = 0.0f; pnorm = dxnorm; return; }
    } else {
        dxnorm = kInf; fp = kInf;
    }
    float parl = 0.0f;
    if (full_rank) {
        float rdx = f_rcp(dxnorm);
#pragma unroll
        for (int i = 0; i < P; ++i) w[i] = D[i] * p[i] * rdx;
        float t2 = chol_forward<P>(U, w, z);
        parl = fp * f_rcp(delta * t2);
    }
    float gn = 0.0f;
#pragma unroll
    for (int i = 0; i < P; ++i) gn = f_fma(g[i] * g[i], f_rcp(D[i]), gn);
    float gnorm = f_sqrt(gn);
    float paru = gnorm * f_rcp(delta);
    if (paru == 0.0f) paru = dwarf / fminf(delta, 0.1f);
    par = fmaxf(par, parl);
    par = fminf(par, paru);
    if (par == 0.0f) par = full_rank ? gnorm * f_rcp(dxnorm) : 0.0f;
    for (int it = 1;; ++it) {
        if (par == 0.0f) par = fmaxf(dwarf, 0.001f * paru);
        bool ok = chol_factor<P>(A, D, par, U);
        chol_solve_neg<P>(U, g, p);
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < P; ++i) s = f_fma(D[i] * p[i], p[i], s);
        dxnorm = f_sqrt(s);
        float temp = fp;
        fp = dxnorm - delta;
        if (!ok || !f_finite(fp)) {  // not factorizable in fp32: push par up and retry (bounded by it == 10)
            if (it == 10) break;
            parl = fmaxf(parl, par);
            par = fmaxf(2.0f * par, 0.001f * paru);
            fp = temp;
            continue;
        }
        if (fabsf(fp) <= 0.1f * delta || (parl == 0.0f && fp <= temp && temp < 0.0f) || it == 10) break;
        float rdx = f_rcp(dxnorm);
#pragma unroll
        for (int i = 0; i < P; ++i) w[i] = D[i] * p[i] * rdx;
        float t2 = chol_forward<P>(U, w, z);
        float parc = fp * f_rcp(delta * t2);
        if (fp > 0.0f) parl = fmaxf(parl, par);
        if (fp < 0.0f) paru = fminf(paru, par);
        par = fmaxf(parl, par + parc);
    }
    pnorm = dxnorm;
}


// Per-fit state of the least-squares pre-fit: MINPACK lmder (notebooks/lmder.f:186-452) over the normal
// equations.  One trip of the driver = propose() a step, one trial pass, consume() its sums.
template <class Mo>
struct StageA {
    static constexpr int P = Mo::P;
    static constexpr int NTP = NT<P>::value;
    static constexpr bool kMle = false, kPred = false, kTie = false;
    float x[P], A[NTP], g[P], D[P];
    typename Mo::Par q0;
    float chi;  // 0.5 * fnorm^2
    float par, delta, xnorm, pnorm;
    int ier, nfev;
    bool init, first, resumed;

    // `consts`: constants of the model for this call (Mo::set_consts), may be null
    DPH_HD void start(const float* p0, const float* consts = nullptr) {
#pragma unroll
        for (int i = 0; i < P; ++i) x[i] = p0[i];
        Mo::set_consts(q0, consts);
        Mo::prepare(x, q0);
        par = delta = xnorm = pnorm = chi = 0.0f;
        ier = 0; nfev = 0;
        init = true; first = true; resumed = false;
    }
    // Continuation record: everything a different lane group needs to carry this fit on from the last
    // accepted point (A and g are recomputed there by a zero-step pass).
    static constexpr int kRecord = 2 * P + 6;
    DPH_HD int trips() const { return nfev; }
    DPH_HD void save(float* r) const {
#pragma unroll
        for (int i = 0; i < P; ++i) { r[i] = x[i]; r[P + i] = D[i]; }
        r[2 * P] = chi; r[2 * P + 1] = par; r[2 * P + 2] = delta; r[2 * P + 3] = xnorm;
        r[2 * P + 4] = float(nfev); r[2 * P + 5] = first ? 1.0f : 0.0f;
    }
    DPH_HD void resume(const float* r, const float* consts = nullptr) {
#pragma unroll
        for (int i = 0; i < P; ++i) { x[i] = r[i]; D[i] = r[P + i]; }
        chi = r[2 * P]; par = r[2 * P + 1]; delta = r[2 * P + 2]; xnorm = r[2 * P + 3];
        nfev = int(r[2 * P + 4]); first = r[2 * P + 5] != 0.0f;
        Mo::set_consts(q0, consts);
        Mo::prepare(x, q0);
        pnorm = 0.0f; ier = 0;
        init = true; resumed = true;
    }
    DPH_HD bool wants_chi() const { return false; }
    // 0: the fit ended without needing the pass, 1: evaluate x + p (2, stage B only: tie-break wanted)
    DPH_HD int propose(const Options&, float* p) {
        if (init) {
#pragma unroll
            for (int i = 0; i < P; ++i) p[i] = 0.0f;
            return 1;
        }
        lmpar_normal<P>(A, g, D, delta, par, p, pnorm);
        if (first) delta = fminf(delta, pnorm);
        return 1;
    }
    DPH_HD void note_accepted(const float*, float*) {}
    DPH_HD float lam_of_step() const { return 0.0f; }
    DPH_HD int tie_fallback() { return 1; }
    template <int G> DPH_HD int tie_break(const Lanes<G>&, const PixelState<Mo::NE>&, const Options&, const float*) { return 1; }
    // what waits in the group's shared memory during a pass: A always, g, D and the step if `cap` floats hold them too
    static constexpr bool kStash = DPH_STASH && NTP <= 16;
    template <int CAP> DPH_HD void stash_put(float* s, const float* p) const {
        stash_store(s, A, NTP);
        if (NTP + 3 * P <= CAP) { stash_store(s + NTP, g, P); stash_store(s + NTP + P, D, P); stash_store(s + NTP + 2 * P, p, P); }
    }
    template <int CAP> DPH_HD void stash_get(const float* s, float* p) {
        stash_load(s, A, NTP);
        if (NTP + 3 * P <= CAP) { stash_load(s + NTP, g, P); stash_load(s + NTP + P, D, P); stash_load(s + NTP + 2 * P, p, P); }
    }
    // new Jacobian at an accepted point: gradient test and rescaling (lmder.f: gnorm, diag = max(diag, wa2))
    DPH_HD void fresh(const Options& opt) {
        if (opt.gtol > 0.0f && chi > 0.0f) {
            float gnorm = 0.0f, rf = f_rsqrt(2.0f * chi);
#pragma unroll
            for (int i = 0; i < P; ++i)
                if (A[tri<P>(i, i)] != 0.0f) gnorm = fmaxf(gnorm, fabsf(g[i] * rf * f_rsqrt(A[tri<P>(i, i)])));
            if (gnorm <= opt.gtol) ier = 4;
        }
#pragma unroll
        for (int i = 0; i < P; ++i) D[i] = fmaxf(D[i], A[tri<P>(i, i)]);
    }
    // returns true when the fit is finished; `accepted` tells the driver to flip the exponential buffers
    DPH_HD bool consume(const Options& opt, const Normal<P>& n, const float* p, const typename Mo::Par& q1, bool& accepted) {
        const int maxfev = opt.maxfev > 0 ? opt.maxfev : 100 * (P + 1);
        accepted = false;
        if (init && resumed) {  // carried over from another group: only the normal equations are rebuilt
            init = resumed = false;
#pragma unroll
            for (int i = 0; i < NTP; ++i) A[i] = n.A[i];
#pragma unroll
            for (int i = 0; i < P; ++i) g[i] = n.g[i];
            accepted = true;
            return false;
        }
        if (init) {
            init = false;
            nfev = 1;
            if (n.flags & kFlagNonFinite) { ier = 9; return true; }  // scipy raises ValueError on a non-finite start
#pragma unroll
            for (int i = 0; i < NTP; ++i) A[i] = n.A[i];
#pragma unroll
            for (int i = 0; i < P; ++i) g[i] = n.g[i];
            chi = n.chi1;
            accepted = true;
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                D[i] = A[tri<P>(i, i)] == 0.0f ? 1.0f : A[tri<P>(i, i)];
                s = f_fma(D[i] * x[i], x[i], s);
            }
            xnorm = f_sqrt(s);
            delta = opt.factor * xnorm;
            if (delta == 0.0f) delta = opt.factor;
            fresh(opt);
            return ier != 0;
        }
        ++nfev;
        const float fn2 = 2.0f * chi;  // fnorm^2
        const float actual = n.actual;
        const bool bad = (n.flags & kFlagNonFinite) || !f_finite(actual);
        const bool blown = bad || (fn2 - 2.0f * actual) >= 100.0f * fn2;  // 0.1 fnorm1 >= fnorm
        const float rfn2 = f_rcp(fn2);
        const float actred = blown ? -1.0f : 2.0f * actual * rfn2;
        const float t1 = quad_form<P>(A, p) * rfn2;
        const float t2 = par * pnorm * pnorm * rfn2;
        const float prered = t1 + 2.0f * t2;
        const float dirder = -(t1 + t2);
        const float ratio = prered != 0.0f ? actred * f_rcp(prered) : 0.0f;
        if (ratio <= 0.25f) {
            float temp = actred >= 0.0f ? 0.5f : 0.5f * dirder * f_rcp(dirder + 0.5f * actred);
            if (blown || temp < 0.1f) temp = 0.1f;
            delta = temp * fminf(delta, pnorm * 10.0f);
            par = par * f_rcp(temp);
        } else if (par == 0.0f || ratio >= 0.75f) {
            delta = 2.0f * pnorm;
            par = 0.5f * par;
        }
        if (ratio >= 1e-4f) {  // successful iteration
#pragma unroll
            for (int i = 0; i < P; ++i) x[i] += p[i];
            q0 = q1;
            accepted = true;
#pragma unroll
            for (int i = 0; i < NTP; ++i) A[i] = n.A[i];
#pragma unroll
            for (int i = 0; i < P; ++i) g[i] = n.g[i];
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < P; ++i) s = f_fma(D[i] * x[i], x[i], s);
            xnorm = f_sqrt(s);
            chi -= actual;
            first = false;
        }
        const bool fconv = fabsf(actred) <= opt.ftol && prered <= opt.ftol && 0.5f * ratio <= 1.0f;
        if (fconv) ier = 1;
        if (delta <= opt.xtol * xnorm) ier = fconv ? 3 : 2;
        if (ier == 0 && nfev >= maxfev) ier = 5;
        if (ier == 0 && !(delta > 0.0f)) ier = 7;
        if (ier == 0 && accepted) fresh(opt);
        return ier != 0;
    }
};

// Per-fit state of the reference's lm() loop (lm.py:389-481).
template <class Mo, bool MLE>
struct StageB {
    static constexpr int P = Mo::P;
    static constexpr int NTP = NT<P>::value;
    static constexpr bool kMle = MLE, kPred = true;
    static constexpr bool kTie = MLE;  // least squares ends at trip 0 on the pre-fit's own optimum (Q16): nothing to break
    static constexpr float kTieBand = 16.0f;  // |dx| within this factor of the xtol threshold: re-decided in float64 (Options::exact_ties)
    float x[P], A[NTP], g[P], dtd[P], xret[P];
    bool hist;  // the previous accepted point and its lambda are in the working buffer's spare words (note_accepted)
    bool tie_fires;  // what float32 says about a proposal that is too close to call
    typename Mo::Par q0;
    float chisq_old, chisq_ret, lam;
    int info, ev, n_acc, n_rej, n_tie;
    bool init, live, resumed;

    DPH_HD void start(const float* p_ls, const float* consts = nullptr) {
#pragma unroll
        for (int i = 0; i < P; ++i) x[i] = xret[i] = p_ls[i];
        hist = false;
        Mo::set_consts(q0, consts);
        Mo::prepare(x, q0);
        chisq_old = chisq_ret = lam = 0.0f;
        info = 5; ev = 0; n_acc = n_rej = n_tie = 0;
        init = true; live = false; resumed = false;
    }
    // continuation record: the state, then (words kKeepAt ...) whether a previous accepted point is known, that point and its lambda
    static constexpr int kKeepAt = 3 * P + 8, kRecord = 4 * P + 9;
    DPH_HD int trips() const { return ev; }
    DPH_HD void save(float* r) const {
#pragma unroll
        for (int i = 0; i < P; ++i) { r[i] = x[i]; r[P + i] = dtd[i]; r[2 * P + i] = xret[i]; }
        r[3 * P] = chisq_old; r[3 * P + 1] = chisq_ret; r[3 * P + 2] = lam;
        r[3 * P + 3] = float(ev); r[3 * P + 4] = float(n_acc); r[3 * P + 5] = float(n_rej);
        r[3 * P + 6] = hist ? 1.0f : 0.0f; r[3 * P + 7] = float(n_tie);
    }
    DPH_HD void resume(const float* r, const float* consts = nullptr) {
#pragma unroll
        for (int i = 0; i < P; ++i) { x[i] = r[i]; dtd[i] = r[P + i]; xret[i] = r[2 * P + i]; }
        chisq_old = r[3 * P]; chisq_ret = r[3 * P + 1]; lam = r[3 * P + 2];
        ev = int(r[3 * P + 3]); n_acc = int(r[3 * P + 4]); n_rej = int(r[3 * P + 5]); n_tie = int(r[3 * P + 7]);
        hist = r[3 * P + 6] != 0.0f;  // (the caller puts the kept point back behind its working ROI)
        Mo::set_consts(q0, consts);
        Mo::prepare(x, q0);
        info = 5;
        init = true; live = false; resumed = true;
    }
    DPH_HD bool wants_chi() const { return init && !resumed; }
    // 0: the fit ended before the evaluation (xtest / gtest), 1: evaluate x + p, 2: |dx| is within kTieBand of the xtol
    // threshold and Options::exact_ties is set -- the caller calls tie_break()
    DPH_HD int propose(const Options& opt, float* p) {
        live = false;
#pragma unroll
        for (int i = 0; i < P; ++i) p[i] = 0.0f;
        if (init) return 1;
        if (opt.gtol > 0.0f) {  // lm.py:358-363, 409-411
            float gm = 0.0f;
#pragma unroll
            for (int i = 0; i < P; ++i) gm = fmaxf(gm, fabsf(g[i]));
            if (gm <= opt.gtol) { info = 4; return 0; }
        }
        return solve_step(opt, g, p);
    }
    DPH_HD float lam_of_step() const { return lam; }
    // the float32 decision, when nobody can break the tie (the straggler list is full)
    DPH_HD int tie_fallback() {
        if (!tie_fires) return 1;
        info = 2;
        live = false;
        return 0;
    }
    // the accepted point a step was just taken from and the lambda of that step: what tie_break() repeats in float64.
    // `keep`: the P + 1 spare words behind the working ROI (free once the fit has been fetched)
    DPH_HD void note_accepted(const float* xold_lam, float* keep) {
#pragma unroll
        for (int i = 0; i <= P; ++i) keep[i] = xold_lam[i];
        hist = true;
    }
    // The proposal is within kTieBand of the xtol threshold: decide in float64.  1: evaluate the step, 0: xtest fires (info 2).
    template <int G> DPH_HD_COLD int tie_break(const Lanes<G>& ln, const PixelState<Mo::NE>& px, const Options& opt, const float* consts) {
        double A64[NTP], g64[P], xp[P], d[P];
        const volatile float* keep = px.y + px.stride;
        ++n_tie;
        if (hist) {  // repeat the last accepted step from the previous accepted point, exactly
#pragma unroll
            for (int i = 0; i < P; ++i) xp[i] = double(keep[i]);
            precise_normal<Mo, MLE, G>(ln, px, xp, consts, A64, g64);
            if (solve_damped64<P>(A64, dtd, double(keep[P]), g64, d)) {
#pragma unroll
                for (int i = 0; i < P; ++i) xp[i] += d[i];
            } else {
#pragma unroll
                for (int i = 0; i < P; ++i) xp[i] = double(x[i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < P; ++i) xp[i] = double(x[i]);
        }
        precise_normal<Mo, MLE, G>(ln, px, xp, consts, A64, g64);
        if (!solve_damped64<P>(A64, dtd, double(lam), g64, d)) return 1;  // keep the float32 decision (not firing)
        double dn = 0.0, xn = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i) { dn += d[i] * d[i]; xn += xp[i] * xp[i]; }
        if (sqrt(dn) <= double(opt.xtol) * (sqrt(xn) + double(opt.xtol))) {  // lm.py:365-367, 430-432
            info = 2;
            live = false;
            return 0;
        }
        return 1;
    }
    DPH_HD int solve_step(const Options& opt, const float* gg, float* p) {
        float U[NTP], d[P];
        bool ok = chol_factor<P>(A, dtd, lam, U);
        if (ok) {
            chol_solve_neg<P>(U, gg, d);
#pragma unroll
            for (int i = 0; i < P; ++i) ok = ok && f_finite(d[i]);
        }
        if (!ok) {
            lam *= opt.factor;  // lm.py:426-428: the trip is consumed
            return 1;
        }
        float dn = 0.0f, xn2 = 0.0f;
#pragma unroll
        for (int i = 0; i < P; ++i) { dn = f_fma(d[i], d[i], dn); xn2 = f_fma(x[i], x[i], xn2); }
        const float thr = opt.xtol * (sqrtf(xn2) + opt.xtol), dnorm = sqrtf(dn);
        const bool fires = dnorm <= thr;  // lm.py:365-367, 430-432
        if (kTie && opt.exact_ties && dnorm <= kTieBand * thr) {  // too close to call in float32 (also when it seems to fire)
            tie_fires = fires;
            live = true;
#pragma unroll
            for (int i = 0; i < P; ++i) p[i] = d[i];
            return 2;
        }
        if (fires) {
            info = 2;
            return 0;
        }
        live = true;
#pragma unroll
        for (int i = 0; i < P; ++i) p[i] = d[i];
        return 1;
    }
    static constexpr bool kStash = DPH_STASH && NTP <= 16;
    template <int CAP> DPH_HD void stash_put(float* s, const float* p) const {
        stash_store(s, A, NTP);
        if (NTP + 3 * P <= CAP) { stash_store(s + NTP, g, P); stash_store(s + NTP + P, dtd, P); stash_store(s + NTP + 2 * P, p, P); }
    }
    template <int CAP> DPH_HD void stash_get(const float* s, float* p) {
        stash_load(s, A, NTP);
        if (NTP + 3 * P <= CAP) { stash_load(s + NTP, g, P); stash_load(s + NTP + P, dtd, P); stash_load(s + NTP + 2 * P, p, P); }
    }
    DPH_HD int solve_step(const Options& opt, const float* gg, float* p, bool tie) {
        float U[NTP], d[P];
        bool ok = chol_factor<P>(A, dtd, lam, U);
        if (ok) {
            chol_solve_neg<P>(U, gg, d);
#pragma unroll
            for (int i = 0; i < P; ++i) ok = ok && f_finite(d[i]);
        }
        if (!ok) {
            lam *= opt.factor;  // lm.py:426-428: the trip is consumed
            return 1;
        }
        float dn = 0.0f, xn2 = 0.0f;
#pragma unroll
        for (int i = 0; i < P; ++i) { dn = f_fma(d[i], d[i], dn); xn2 = f_fma(x[i], x[i], xn2); }
        const float thr = opt.xtol * (sqrtf(xn2) + opt.xtol), dnorm = sqrtf(dn);
        if (tie && dnorm <= kTieBand * thr && dnorm * kTieBand >= thr) return 2;
        if (dnorm <= thr) {  // lm.py:365-367, 430-432
            info = 2;
            return 0;
        }
        live = true;
#pragma unroll
        for (int i = 0; i < P; ++i) p[i] = d[i];
        return 1;
    }
    DPH_HD bool consume(const Options& opt, const Normal<P>& n, const float* p, const typename Mo::Par& q1, bool& accepted) {
        const int maxfev = opt.maxfev > 0 ? opt.maxfev : 100 * (P + 1);
        accepted = false;
        if (init && resumed) {
            init = resumed = false;
#pragma unroll
            for (int i = 0; i < NTP; ++i) A[i] = n.A[i];
#pragma unroll
            for (int i = 0; i < P; ++i) g[i] = n.g[i];
            accepted = true;
            return false;
        }
        if (init) {
            init = false;
#pragma unroll
            for (int i = 0; i < NTP; ++i) A[i] = n.A[i];
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                g[i] = n.g[i];
                dtd[i] = A[tri<P>(i, i)];                     // lm.py:394
                s = f_fma(dtd[i] * x[i], x[i], s);
            }
            chisq_old = chisq_ret = n.chi1;
            lam = sqrtf(s);                                    // lm.py:398
            if (!(lam > 0.0f) || !f_finite(lam)) lam = 100.0f;  // lm.py:399-400
            accepted = true;
            return false;
        }
        if (live) {
#pragma unroll
            for (int i = 0; i < P; ++i) xret[i] = x[i] + p[i];
            const float actual = n.actual;
            // chi2 = +inf when the predicted model dips below 0 (lm.py:62-64)
            const float predicted = (n.flags & kFlagPredNeg) ? -kInf : n.predicted;
            chisq_ret = chisq_old - actual;
            float rho = actual / predicted;
            if (actual < 0.0f || predicted < 0.0f || !f_finite(rho)) rho = 0.0f;  // lm.py:454-463
            if (rho > 1e-2f) {
                if (actual <= opt.ftol * chisq_old) {  // lm.py:468-470
                    info = 1;
                    return true;
                }
#pragma unroll
                for (int i = 0; i < P; ++i) x[i] = xret[i];
                q0 = q1;
                accepted = true;
                chisq_old -= actual;
#pragma unroll
                for (int i = 0; i < NTP; ++i) A[i] = n.A[i];
#pragma unroll
                for (int i = 0; i < P; ++i) { g[i] = n.g[i]; dtd[i] = fmaxf(dtd[i], A[tri<P>(i, i)]); }  // lm.py:474
                lam = fmaxf(lam * 0.2f, 1e-7f);  // lm.py:475
                ++n_acc;
            } else {
                lam = fminf(lam * 1.5f, 1e7f);  // lm.py:477
                ++n_rej;
            }
        }
        ++ev;
        if (ev >= maxfev) { ev = maxfev - 1; return true; }  // info stays 5 (lm.py:479-481)
        return false;
    }
};

// ------------------------------------------------------------------------------------- stage driver
// Runs one stage for a stream of fits.  Every lane group owns one fit at a time and fetches the next one
// from `io` as soon as its own has finished, so the groups of a warp never wait for each other's
// iteration counts; the only warp-wide operations (the shuffles of the pass reductions) sit at
// warp-uniform points of the loop.  `io` supplies fits (fetch: start or resume `st`), takes results (store)
// and may take over a fit that has used up its trip budget (park): stragglers -- the reference's own
// trajectories reach 100-360 trips on degenerate ROIs -- are handed to a follow-up launch where a whole
// warp owns each fit, instead of holding a narrow group (and the tail of the kernel) for milliseconds.
#if defined(__CUDA_ARCH__)
template <int G> constexpr bool kTieInline = G == 32;
#else
template <int G> constexpr bool kTieInline = true;
#endif

template <class Stage, class Mo, int G, class IO>
DPH_HD void run_stage(const Lanes<G>& ln, PixelState<Mo::NE>& px, const Options& opt, IO& io) {
    constexpr int P = Mo::P;
    Stage st;
    Normal<P> nrm;
    float p[P];
    typename Mo::Par q1;
    typename Mo::Step step;
    bool has = false, exhausted = false;
    bool old_irregular = false;  // the accepted point of this group's fit has a non-positive (or non-finite) model value
    long long fit = -1;
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = 1.0f;
    st.start(p);  // defined state for idle groups (values irrelevant)
    for (;;) {
        if (Lanes<G>::warp_any(!has && !exhausted)) {
            if (!has && !exhausted) {
                if (io.fetch(ln, px, st, fit)) {
                    px.cur = 0;
                    has = true;
                } else {
                    exhausted = true;
                }
            }
        }
        if (!Lanes<G>::warp_any(has)) break;
        bool pass = has;
        int proposed = has ? st.propose(opt, p) : 1;
        if constexpr (Stage::kTie) {
          if (Lanes<G>::warp_any(proposed == 2)) {  // xtol tie-break: rare (see precise_normal)
            if (proposed == 2) {
                // The float64 passes run where a whole warp owns the fit (the poller kernels; the host build): a narrow group
                // hands the fit over like a straggler.  Inside a narrow group they would stall the other fits of the warp for
                // the length of ~20 ordinary trips (measured: 3x slower calls), and their code would sit in the hot kernels.
                if constexpr (kTieInline<G>) {
                    proposed = st.template tie_break<G>(ln, px, opt, io.consts());
                } else if (io.park_tie(ln, px, st, fit)) {
                    has = false;
                    pass = false;
                    proposed = 1;
                } else {
                    proposed = st.tie_fallback();
                }
            }
          }
        }
        if (has && proposed == 0) {  // ended before the evaluation (xtest / gtest)
            io.store(ln, st, fit);
            has = false;
            pass = false;
        }
        if (!pass) {
#pragma unroll
            for (int i = 0; i < P; ++i) p[i] = 0.0f;
        }
        float xn[P];
#pragma unroll
        for (int i = 0; i < P; ++i) xn[i] = st.x[i] + p[i];
        q1 = st.q0;  // carries the constants of the call (fixed-width model)
        Mo::prepare(xn, q1);
        Mo::prepare_step(st.q0, q1, p, step);
        // models with per-trial tables; no-op otherwise.  Last argument: bit 0 = which buffer holds the accepted point, bit 1 = first pass
        Mo::fill_tables(ln, st.q0, q1, step, px.tbl, px.M / px.w, px.w, px.cur | (st.init ? 2 : 0));
        const bool init = st.init, want_chi = Lanes<G>::warp_any(pass && st.wants_chi());
        if (Stage::kStash) st.template stash_put<stash_floats(G, Mo::kStoreE)>(px.stash, p);
        trial_pass<Mo, Stage::kMle, Stage::kPred, true, G>(ln, px, st.q0, q1, step, init, want_chi, nrm);
        // the branch-free pass is valid only between two regular points; otherwise repeat it in the literal form
        const bool trial_irregular = Stage::kMle && (nrm.flags & kFlagIrregular);
        const bool literal = Stage::kMle && (trial_irregular || (old_irregular && !init));
        if (Stage::kMle && Lanes<G>::warp_any(pass && literal)) {
            Normal<P> slow;
            trial_pass<Mo, Stage::kMle, Stage::kPred, false, G>(ln, px, st.q0, q1, step, init, want_chi, slow);
            if (literal) nrm = slow;
        }
        if (Stage::kStash) st.template stash_get<stash_floats(G, Mo::kStoreE)>(px.stash, p);
        if (pass) {
            bool accepted;
            float xold_lam[P + 1];
            if (Stage::kTie && opt.exact_ties) {
#pragma unroll
                for (int i = 0; i < P; ++i) xold_lam[i] = st.x[i];
                xold_lam[P] = st.lam_of_step();
            }
            const bool done = st.consume(opt, nrm, p, q1, accepted);
            if (Stage::kTie && opt.exact_ties && accepted && !init) st.note_accepted(xold_lam, px.y + px.stride);
            if (accepted) px.cur = 1 - px.cur;
            if (accepted || init) old_irregular = trial_irregular;  // the first pass evaluates the accepted point itself
            if (done) {
                io.store(ln, st, fit);
                has = false;
            } else if (st.trips() >= io.budget() && io.park(ln, st, fit)) {
                has = false;
            }
        }
    }
}

}  // namespace dphlm
