"""Float64 numpy restatement of the reference fitter ``dphtools/utils/lm.py``.  TEST INFRASTRUCTURE.

This is the checker the CUDA path is compared with; it is never on the product path.

What it restates (reference file:line, relative to the upstream repo root):

* ``chi2_ls`` / ``update_ls``        -> ``dphtools/utils/lm.py:30-52``
* ``chi2_mle`` / ``update_mle``      -> ``dphtools/utils/lm.py:55-103``
* ``clamp_nonneg``                   -> ``dphtools/utils/lm.py:106-118``
* ``lm``                             -> ``dphtools/utils/lm.py:221-507``
* ``curve_fit`` (``ls``/``mle``)     -> ``dphtools/utils/lm.py:602-685``

Stage 1 of ``curve_fit`` is third-party code that is not in the reference tree:
``scipy.optimize.curve_fit`` -> ``leastsq`` -> compiled MINPACK ``lmder`` (call site
``lm.py:642-644``; scipy is unpinned upstream, 1.18.1 in this image).  The oracle calls the very same
scipy routine, so stage 1 is *the* reference implementation, not a restatement.

Pinning: the reference's own tests never execute ``method='ls'|'mle'``, so the
oracle is pinned against outputs of the live reference instead: ``oracle/make_golden.py`` imports
``/root/reference`` in the build container and stores inputs + reference outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against them bit-for-bit
(identical ``info`` / ``nfev``, ``popt`` equal to 1e-12 relative).
"""

import numpy as np
import scipy.optimize

FTOL = XTOL = 1.49012e-8

MESSAGES = {
    1: "Both actual and predicted relative reductions in the sum of squares are at most {ftol}",
    2: "The relative error between two consecutive iterates is at most {xtol}",
    4: "The cosine of the angle between func(x) and any column of the\n  Jacobian is at most {gtol:f} in absolute value",
    5: "Number of calls to function has reached maxfev = {maxfev:d}.",
}


def clamp_nonneg(a):
    """Copy with every entry <= 0 set to 0 (``lm.py:106-118``)."""
    return np.where(a <= 0, 0.0, a)


def chi2_ls(r):
    return 0.5 * np.dot(r, r)


def update_ls(jmat, r):
    return jmat.T @ jmat, jmat.T @ r


def chi2_mle(f, y):
    """Poisson deviance / 2 with the reference's zeroing rule (``lm.py:55-78``)."""
    if f.min() < 0:
        return np.inf
    with np.errstate(invalid="ignore", divide="ignore"):
        logterm = y * np.log(f / y)
    logterm[~np.isfinite(logterm)] = 0.0
    return (f - y).sum() - logterm.sum()


def update_mle(jmat, f, y):
    """``A = J^T diag(y/f^2) J``, ``g = J^T (1 - y/f)``; invalid pixels: weight 0, residual 1 (``lm.py:81-103``)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        y_f = y / f
        y_f2 = y_f / f
    bad = ~(np.isfinite(y_f) & np.isfinite(y_f2))
    y_f[bad] = 0.0
    y_f2[bad] = 0.0
    return (jmat.T * y_f2) @ jmat, jmat.T @ (1.0 - y_f)


def lm(model, jac, y, x0, method, ftol=FTOL, xtol=XTOL, gtol=0.0, maxfev=None, factor=100.0, trace=None):
    """The damped normal-equation loop of ``lm.py:221-507``.

    ``model(p)`` returns the raw model vector, ``jac(p)`` the (M, P) Jacobian, ``y`` the data.
    Returns ``(x, info, nfev, fvec, fjac, chisq)``.  ``trace`` (a list) receives one tuple per trip:
    ``(ev, lambda, actual, predicted, rho, accepted)``.
    """
    mle = method == "mle"
    x0 = np.array(x0, dtype=float).ravel()
    if maxfev is None:
        maxfev = 100 * (x0.size + 1)  # lm.py:306-307
    if mle:
        y = clamp_nonneg(np.asarray(y, dtype=float))  # lm.py:127

        def evaluate(p):
            return clamp_nonneg(model(p))  # lm.py:132

        def chi2(f):
            return chi2_mle(f, y)

        def update(p, f):
            j = jac(p)
            return (j,) + update_mle(j, f, y)

    else:
        y = np.asarray(y, dtype=float)

        def evaluate(p):
            return model(p) - y  # lm.py:167

        chi2 = chi2_ls

        def update(p, f):
            j = jac(p)
            return (j,) + update_ls(j, f)

    f = evaluate(x0)
    j, a, g = update(x0, f)
    dtd = np.diag(a).copy()  # lm.py:394 (kept as a vector here)
    chisq_old = chi2(f)
    lam = np.sqrt(np.dot(x0 * dtd, x0))  # lm.py:398
    if lam <= 0 or not np.isfinite(lam):
        lam = 100.0
    x = x0
    info = 5
    ev = -1
    for ev in range(maxfev):
        if gtol and np.abs(g).max() <= gtol:  # lm.py:358-363, 409-411
            info = 4
            break
        try:
            dx = np.linalg.solve(a + lam * np.diag(dtd), -g)  # lm.py:420-425
        except np.linalg.LinAlgError:
            lam *= factor  # lm.py:426-428, the trip is consumed
            continue
        if np.linalg.norm(dx) <= xtol * (np.linalg.norm(x0) + xtol):  # lm.py:365-367, 430
            info = 2
            break
        x = x0 + dx
        f = evaluate(x)
        chisq_new = chi2(f)
        chisq_pred = chi2(f + j @ dx)  # lm.py:446-449: f at the NEW point, J at the old one
        actual = chisq_old - chisq_new
        with np.errstate(invalid="ignore", divide="ignore"):
            predicted = chisq_old - chisq_pred
            rho = np.float64(actual) / np.float64(predicted)
        if actual < 0 or predicted < 0 or not np.isfinite(rho):  # lm.py:454-463
            rho = 0.0
        accepted = rho > 1e-2
        if trace is not None:
            trace.append((ev, float(lam), float(actual), float(predicted), float(rho), bool(accepted)))
        if accepted:
            if actual <= ftol * chisq_old:  # lm.py:468-470 (breaks before x0 <- x)
                info = 1
                break
            x0, chisq_old = x, chisq_new
            j, a, g = update(x0, f)
            dtd = np.fmax(dtd, np.diag(a))  # lm.py:474
            lam = max(lam / 5, 1e-7)  # lm.py:475
        else:
            lam = min(lam * 1.5, 1e7)  # lm.py:477
    nfev = max(ev, 0) if maxfev > 0 else 0  # lm.py:489 (zero-based index of the last trip)
    return x, info, nfev, f, j, chi2(f)


def pcov_from_jac(fjac):
    """Moore-Penrose ``pinv(J^T J)`` from the SVD of the last Jacobian (``lm.py:672-677``)."""
    _, s, vt = np.linalg.svd(fjac, full_matrices=False)
    keep = s > np.finfo(float).eps * max(fjac.shape) * s[0]
    s, vt = s[keep], vt[: keep.sum()]
    return np.dot(vt.T / s**2, vt)


def curve_fit(f, xdata, ydata, p0, jac, method, full_output=False, return_stage1=False, **kwargs):
    """``curve_fit(..., method='ls'|'mle')`` of ``lm.py:602-685``: scipy LS pre-fit, then :func:`lm`.

    ``kwargs`` go to both stages, as upstream (``lm.py:643, 668``).
    """
    if method not in ("ls", "mle"):
        raise TypeError("Method {} not recognized".format(method))
    if jac is None:
        raise NotImplementedError("You need a Jacobian")
    # stage 1: the reference's own third-party call (lm.py:642-644)
    p_ls = scipy.optimize.curve_fit(f, xdata, ydata, p0, None, False, True, (-np.inf, np.inf), None, jac, **kwargs)[0]
    ydata = np.asarray_chkfinite(ydata)
    xdata = np.asarray_chkfinite(xdata)
    x, info, nfev, fvec, fjac, chisq = lm(
        lambda p: f(xdata, *p), lambda p: jac(xdata, *p), ydata, p_ls, method, **kwargs
    )
    pcov = pcov_from_jac(fjac)
    if info not in (1, 2, 3, 4):
        raise RuntimeError("Optimal parameters not found: " + MESSAGES[info].format(maxfev=kwargs.get("maxfev") or 100 * (len(x) + 1)))
    out = (x, pcov)
    if full_output:
        out += (dict(fvec=fvec, fjac=fjac, nfev=nfev, chisq=chisq), info)
    if return_stage1:
        out += (p_ls,)
    return out


def fit_batch(model, data, p0, method, xdata, **kwargs):
    """Per-ROI :func:`curve_fit` over a batch.  Failed fits get ``info`` 5 (stage 2) or -1 (stage 1).

    Returns ``dict(popt[N,P], info[N], nfev[N], chisq[N], p_ls[N,P])``.
    """
    n = data.shape[0]
    p0 = np.asarray(p0, dtype=float)
    out = dict(
        popt=np.full(p0.shape, np.nan), info=np.zeros(n, np.int32), nfev=np.zeros(n, np.int32),
        chisq=np.full(n, np.nan), p_ls=np.full(p0.shape, np.nan),
    )
    flat = np.asarray(data, dtype=float).reshape(n, -1)
    for i in range(n):
        try:
            p_ls = scipy.optimize.curve_fit(model, xdata, flat[i], p0[i], jac=model.jac, **kwargs)[0]
        except RuntimeError:
            out["info"][i] = -1
            continue
        out["p_ls"][i] = p_ls
        x, info, nfev, _, _, chisq = lm(
            lambda p: model(xdata, *p), lambda p: model.jac(xdata, *p), flat[i], p_ls, method, **kwargs
        )
        out["popt"][i], out["info"][i], out["nfev"][i], out["chisq"][i] = x, info, nfev, chisq
    return out
