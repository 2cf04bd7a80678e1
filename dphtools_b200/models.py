"""Built-in model functions (float64 numpy) and their analytic Jacobians.

The reference ships no Gaussian PSF model; these callables are the
new API surface of this package.  Each one has the scipy ``f(xdata, *params)`` signature, so it can be
handed to ``scipy.optimize.curve_fit`` or to the reference's ``dphtools.utils.lm.curve_fit`` unchanged,
and carries ``model_id`` / ``n_param`` / ``jac`` attributes that :func:`dphtools_b200.curve_fit` uses as
the dispatch key for the CUDA kernels.  ``multi_exp`` follows the reference's convention
(``dphtools/utils/fitfuncs.py:20-52``): parameters ``(a0, k0, a1, k1, ..., [offset])``, the offset is
present iff the number of parameters is odd.

Conventions:

* 2-D models take ``xdata = np.stack([xx.ravel(), yy.ravel()])`` with ``xx`` the column index and
  ``yy`` the row index of a row-major ROI, integer pixel coordinates ``0..n-1``.
* ``gauss2d``:  ``amp * exp(-((x-x0)^2 + (y-y0)^2) / (2 sigma^2)) + offset``
* ``gauss2d_integrated``: ``amp * Ex(x) * Ey(y) + offset`` with ``Ex`` the integral of a unit-area Gaussian over the pixel
  (erf model of a pixelated camera); ``amp`` is the photon count of the spot.
* ``gauss2d_fixed(sigma)``: factory for ``gauss2d`` with the width held at ``sigma`` (4 parameters).
* ``gauss2d_integrated_fixed(sigma)``: factory for the 4-parameter ``(amp, x0, y0, offset)`` version with the width held
  at ``sigma`` (the model of ``notebooks/MLE_LevMarq_4Param_2013_11_12.m`` upstream, whose PSF width is a call argument).
* ``multi_exp_convolved(irf, dt, m0)``: factory for ``multi_exp`` convolved with an instrument response function sampled on
  the histogram's own grid (lifetime histograms): ``f_j = offset + sum_i a_i sum_q irf[q] exp(-k_i (j - m0 - q) dt)``.
* ``gauss2d_elliptical``: ``u =  c (x-x0) + s (y-y0)``, ``v = -s (x-x0) + c (y-y0)``,
  ``amp * exp(-(u^2/sigma_x^2 + v^2/sigma_y^2)/2) + offset`` with ``c, s = cos, sin(theta)``.
"""

import numpy as np

# numeric ids shared with include/dphlm.h (dphlm_model)
MODEL_GAUSS2D = 0
MODEL_GAUSS2D_ELLIPTICAL = 1
MODEL_MULTI_EXP = 2
MODEL_GAUSS2D_INTEGRATED = 3
MODEL_GAUSS2D_INTEGRATED_FIXED = 4
MODEL_MULTI_EXP_IRF = 5
MODEL_GAUSS2D_FIXED = 6
ONE_D = (MODEL_MULTI_EXP, MODEL_MULTI_EXP_IRF)  # models of a 1-D histogram: need the shared x axis
IRF_MAX_TAPS = 32


def gauss2d(xy, amp, x0, y0, sigma, offset):
    """Symmetric 2-D Gaussian, 5 parameters ``(amp, x0, y0, sigma, offset)``."""
    dx = xy[0] - x0
    dy = xy[1] - y0
    return amp * np.exp(-(dx * dx + dy * dy) / (2.0 * sigma * sigma)) + offset


def gauss2d_jac(xy, amp, x0, y0, sigma, offset):
    """Jacobian of :func:`gauss2d`, shape ``(M, 5)`` (row = pixel, column = parameter)."""
    dx = xy[0] - x0
    dy = xy[1] - y0
    r2 = dx * dx + dy * dy
    s2 = sigma * sigma
    e = np.exp(-r2 / (2.0 * s2))
    ae = amp * e
    return np.stack([e, ae * dx / s2, ae * dy / s2, ae * r2 / (s2 * sigma), np.ones_like(e)], axis=1)


def gauss2d_elliptical(xy, amp, x0, y0, sigma_x, sigma_y, theta, offset):
    """Rotated elliptical 2-D Gaussian, 7 parameters."""
    c, s = np.cos(theta), np.sin(theta)
    dx = xy[0] - x0
    dy = xy[1] - y0
    u = c * dx + s * dy
    v = -s * dx + c * dy
    return amp * np.exp(-0.5 * (u * u / sigma_x**2 + v * v / sigma_y**2)) + offset


def gauss2d_elliptical_jac(xy, amp, x0, y0, sigma_x, sigma_y, theta, offset):
    """Jacobian of :func:`gauss2d_elliptical`, shape ``(M, 7)``."""
    c, s = np.cos(theta), np.sin(theta)
    dx = xy[0] - x0
    dy = xy[1] - y0
    u = c * dx + s * dy
    v = -s * dx + c * dy
    ax = 1.0 / sigma_x**2
    ay = 1.0 / sigma_y**2
    e = np.exp(-0.5 * (u * u * ax + v * v * ay))
    ae = amp * e
    return np.stack(
        [
            e,
            ae * (u * ax * c - v * ay * s),
            ae * (u * ax * s + v * ay * c),
            ae * u * u * ax / sigma_x,
            ae * v * v * ay / sigma_y,
            ae * u * v * (ay - ax),
            np.ones_like(e),
        ],
        axis=1,
    )


def _pixel_integral(t, sigma):
    """Integral of a unit-area 1-D Gaussian over the pixel [t - 1/2, t + 1/2] and its derivatives w.r.t. centre and sigma."""
    from scipy.special import erf

    a = 1.0 / (np.sqrt(2.0) * sigma)
    tp, tm = t + 0.5, t - 0.5
    e = 0.5 * (erf(tp * a) - erf(tm * a))
    gp, gm = np.exp(-(tp * a) ** 2), np.exp(-(tm * a) ** 2)
    c = 1.0 / (np.sqrt(2.0 * np.pi) * sigma)
    return e, c * (gm - gp), c / sigma * (tm * gm - tp * gp)


def gauss2d_integrated(xy, amp, x0, y0, sigma, offset):
    """Pixel-integrated symmetric Gaussian: ``amp`` photons spread by a Gaussian of width ``sigma`` and integrated over
    each unit pixel (erf model, cf. ``notebooks/MLE_LevMarq_4Param_2013_11_12.m:44-46`` upstream), plus ``offset``."""
    ex, _, _ = _pixel_integral(xy[0] - x0, sigma)
    ey, _, _ = _pixel_integral(xy[1] - y0, sigma)
    return amp * ex * ey + offset


def gauss2d_integrated_jac(xy, amp, x0, y0, sigma, offset):
    """Jacobian of :func:`gauss2d_integrated`, shape ``(M, 5)``."""
    ex, dex, sex = _pixel_integral(xy[0] - x0, sigma)
    ey, dey, sey = _pixel_integral(xy[1] - y0, sigma)
    return np.stack([ex * ey, amp * dex * ey, amp * ex * dey, amp * (sex * ey + ex * sey), np.ones_like(ex)], axis=1)


def gauss2d_fixed(sigma):
    """:func:`gauss2d` with the width held at ``sigma``: returns the model ``f(xy, amp, x0, y0, offset)`` (with ``f.jac``);
    ``f.consts = (sigma,)`` travels to the kernels as ``dphlm_desc.consts``."""
    sigma = float(sigma)
    if not sigma > 0:
        raise ValueError("sigma must be positive")

    def model(xy, amp, x0, y0, offset):
        return gauss2d(xy, amp, x0, y0, sigma, offset)

    def jac(xy, amp, x0, y0, offset):
        return gauss2d_jac(xy, amp, x0, y0, sigma, offset)[:, [0, 1, 2, 4]]

    _tag(model, jac, MODEL_GAUSS2D_FIXED, 4, "gauss2d_fixed")
    model.consts = (sigma,)
    return model


def gauss2d_integrated_fixed(sigma):
    """Pixel-integrated Gaussian of FIXED width: returns the model ``f(xy, amp, x0, y0, offset)`` (with ``f.jac``), the
    4-parameter fit of ``notebooks/MLE_LevMarq_4Param_2013_11_12.m:44-54`` upstream.  ``f.consts = (sigma,)`` travels to
    the kernels as ``dphlm_desc.consts``."""
    sigma = float(sigma)
    if not sigma > 0:
        raise ValueError("sigma must be positive")

    def model(xy, amp, x0, y0, offset):
        return gauss2d_integrated(xy, amp, x0, y0, sigma, offset)

    def jac(xy, amp, x0, y0, offset):
        return gauss2d_integrated_jac(xy, amp, x0, y0, sigma, offset)[:, [0, 1, 2, 4]]

    _tag(model, jac, MODEL_GAUSS2D_INTEGRATED_FIXED, 4, "gauss2d_integrated_fixed")
    model.consts = (sigma,)
    return model


def multi_exp(xdata, *args):
    r"""Sum of exponentials ``offset + sum_i a_i exp(-k_i x)`` (reference ``fitfuncs.py:20-34``)."""
    xdata = np.asarray(xdata, dtype=float)
    n_exp = len(args) // 2
    res = np.full_like(xdata, args[-1] if len(args) % 2 else 0.0)
    for i in range(n_exp):
        res = res + args[2 * i] * np.exp(-args[2 * i + 1] * xdata)
    return res


def multi_exp_jac(xdata, *args):
    """Jacobian of :func:`multi_exp` (reference ``fitfuncs.py:37-52``), shape ``(M, len(args))``."""
    xdata = np.asarray(xdata, dtype=float)
    cols = []
    for i in range(len(args) // 2):
        e = np.exp(-args[2 * i + 1] * xdata)
        cols += [e, -args[2 * i] * xdata * e]
    if len(args) % 2:
        cols.append(np.ones_like(xdata))
    return np.stack(cols, axis=1)


def multi_exp_convolved(irf, dt, m0=0):
    """Sum of exponentials convolved with an instrument response function: returns the model ``f(x, a0, k0, ..., [offset])``
    (parameters as :func:`multi_exp`) with ``f.jac``.  ``irf`` holds the ``W <= 32`` response taps on the histogram's own
    uniform grid of spacing ``dt``, the first one at bin ``m0``:

        ``f_j = offset + sum_i a_i C_i[j]``,  ``C_i[j] = sum_{q=0}^{min(j-m0, W-1)} irf[q] exp(-k_i (j - m0 - q) dt)``

    (bins before ``m0`` see only the offset).  ``x`` must be that grid (any origin); only its length is used.
    ``f.consts = (W, m0, dt, *irf)`` travels to the kernels as ``dphlm_desc.consts``.  (the model
    is our own discrete formulation.)"""
    irf = np.asarray(irf, dtype=np.float32).astype(float).ravel()  # the kernels see float32 taps
    dt, m0 = float(np.float32(dt)), int(m0)
    if not 1 <= irf.size <= IRF_MAX_TAPS:
        raise ValueError("irf must have 1..%d taps" % IRF_MAX_TAPS)
    if not dt > 0 or m0 < 0:
        raise ValueError("dt must be positive and m0 non-negative")

    def _lags(xdata):
        j = np.arange(np.asarray(xdata).size)
        lag = (j[:, None] - m0 - np.arange(irf.size)[None, :]) * dt  # (M, W)
        return lag, lag >= 0

    def _cd(xdata, k):
        lag, ok = _lags(xdata)
        term = np.where(ok, irf[None, :] * np.exp(-k * np.where(ok, lag, 0.0)), 0.0)
        return term.sum(1), (term * lag).sum(1)

    def model(xdata, *args):
        res = np.full(np.asarray(xdata).size, args[-1] if len(args) % 2 else 0.0, dtype=float)
        for i in range(len(args) // 2):
            res = res + args[2 * i] * _cd(xdata, args[2 * i + 1])[0]
        return res

    def jac(xdata, *args):
        cols = []
        for i in range(len(args) // 2):
            c, d = _cd(xdata, args[2 * i + 1])
            cols += [c, -args[2 * i] * d]
        if len(args) % 2:
            cols.append(np.ones(np.asarray(xdata).size))
        return np.stack(cols, axis=1)

    _tag(model, jac, MODEL_MULTI_EXP_IRF, None, "multi_exp_convolved")
    model.consts = (float(irf.size), float(m0), dt) + tuple(float(v) for v in irf)
    return model


def _tag(func, jac, model_id, n_param, name):
    func.jac = jac
    func.model_id = model_id
    func.n_param = n_param  # None: taken from len(p0)
    func.model_name = name
    return func


_tag(gauss2d, gauss2d_jac, MODEL_GAUSS2D, 5, "gauss2d")
_tag(gauss2d_elliptical, gauss2d_elliptical_jac, MODEL_GAUSS2D_ELLIPTICAL, 7, "gauss2d_elliptical")
_tag(multi_exp, multi_exp_jac, MODEL_MULTI_EXP, None, "multi_exp")
_tag(gauss2d_integrated, gauss2d_integrated_jac, MODEL_GAUSS2D_INTEGRATED, 5, "gauss2d_integrated")

BUILTIN = {m.model_name: m for m in (gauss2d, gauss2d_elliptical, multi_exp, gauss2d_integrated)}


def resolve(name, consts=None):
    """Model by name; the factories are called with the constants they stored in ``f.consts``."""
    if name == "gauss2d_integrated_fixed":
        return gauss2d_integrated_fixed(consts[0])
    if name == "gauss2d_fixed":
        return gauss2d_fixed(consts[0])
    if name == "multi_exp_convolved":
        w = int(consts[0])
        return multi_exp_convolved(consts[3:3 + w], consts[2], int(consts[1]))
    return BUILTIN[name]


def roi_grid(height, width):
    """``xdata`` for a row-major ``height x width`` ROI: ``(2, M)`` array of (column, row) indices."""
    yy, xx = np.mgrid[0:height, 0:width]
    return np.stack([xx.ravel(), yy.ravel()]).astype(float)
