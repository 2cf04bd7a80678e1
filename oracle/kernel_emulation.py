"""Batch-vectorised numpy emulation of the CUDA kernel's ALGORITHM at a chosen precision.  TEST INFRASTRUCTURE.

PROTOTYPE of the numerics study (the production algorithm is dphtools_b200/csrc/lm_core.cuh, whose host build is
oracle/host_emu).  There is no GPU in the build container, so the numerics of the fp32 kernel (differential chi-square
evaluation, tight normal-equation LS pre-fit, the reference's lambda schedule and accept/reject quirks)
are developed and regression-tested here against the float64 oracle.  The CUDA kernels in
``dphtools_b200/csrc/`` implement the same steps; this file is their executable specification, never a
product path.

``dtype=np.float32`` emulates the kernel (``approx=True`` additionally injects ~2^-22 relative errors in
exp/rcp to stand in for MUFU approximations); ``dtype=np.float64`` checks the algorithm itself.
"""

import numpy as np


class _Approx:
    def __init__(self, dtype, approx, seed=0):
        self.dt = dtype
        self.approx = approx and dtype == np.float32
        self.rng = np.random.default_rng(seed)

    def _noise(self, a):
        if not self.approx:
            return a
        return (a * (1 + self.rng.uniform(-1, 1, a.shape) * 2.0**-22).astype(self.dt)).astype(self.dt)

    def exp(self, a):
        return self._noise(np.exp(a))

    def rcp(self, a):
        with np.errstate(divide="ignore"):
            return self._noise((1 / a).astype(self.dt))


class Gauss2D:
    """amp, x0, y0, sigma, offset."""

    P = 5

    def __init__(self, h, w, dtype, ap):
        yy, xx = np.mgrid[0:h, 0:w]
        self.X = xx.ravel().astype(dtype)[None, :]
        self.Y = yy.ravel().astype(dtype)[None, :]
        self.dt = dtype
        self.ap = ap

    def eval(self, p):
        dt = self.dt
        dx = self.X - p[:, 1:2]
        dy = self.Y - p[:, 2:3]
        r2 = dx * dx + dy * dy
        s = (dt(0.5) * self.ap.rcp(p[:, 3:4] * p[:, 3:4])).astype(dt)
        e = self.ap.exp(-r2 * s)
        f = p[:, 0:1] * e + p[:, 4:5]
        return f, dict(dx=dx, dy=dy, r2=r2, s=s, e=e)

    def jac(self, p, aux):
        ae = p[:, 0:1] * aux["e"]
        is2 = 2 * aux["s"]
        isg = self.ap.rcp(p[:, 3:4])
        return np.stack([aux["e"], ae * aux["dx"] * is2, ae * aux["dy"] * is2, ae * aux["r2"] * is2 * isg, np.ones_like(ae)], axis=2)

    def delta(self, p, d, aux, aux_new):
        """f(p+d) - f(p), computed from the step so that its RELATIVE error is ~1e-6."""
        dt = self.dt
        d1, d2, d3 = d[:, 1:2], d[:, 2:3], d[:, 3:4]
        dr2 = -d1 * (2 * aux["dx"] - d1) - d2 * (2 * aux["dy"] - d2)
        s, sn = aux["s"], aux_new["s"]
        ds = -d3 * (2 * p[:, 3:4] + d3) * 2 * s * sn
        darg = -(dr2 * sn + aux["r2"] * ds)
        em1 = np.expm1(darg).astype(dt)
        return d[:, 0:1] * aux_new["e"] + p[:, 0:1] * aux["e"] * em1 + d[:, 4:5]


class Gauss2DElliptical:
    """amp, x0, y0, sigma_x, sigma_y, theta, offset."""

    P = 7

    def __init__(self, h, w, dtype, ap):
        yy, xx = np.mgrid[0:h, 0:w]
        self.X = xx.ravel().astype(dtype)[None, :]
        self.Y = yy.ravel().astype(dtype)[None, :]
        self.dt = dtype
        self.ap = ap

    def eval(self, p):
        dt = self.dt
        c, s = np.cos(p[:, 5:6]).astype(dt), np.sin(p[:, 5:6]).astype(dt)
        dx = self.X - p[:, 1:2]
        dy = self.Y - p[:, 2:3]
        u = c * dx + s * dy
        v = c * dy - s * dx
        ax = self.ap.rcp(p[:, 3:4] * p[:, 3:4])
        ay = self.ap.rcp(p[:, 4:5] * p[:, 4:5])
        e = self.ap.exp(dt(-0.5) * (u * u * ax + v * v * ay))
        f = p[:, 0:1] * e + p[:, 6:7]
        return f, dict(dx=dx, dy=dy, u=u, v=v, ax=ax, ay=ay, c=c, s=s, e=e)

    def jac(self, p, aux):
        a = aux
        ae = p[:, 0:1] * a["e"]
        uax, vay = a["u"] * a["ax"], a["v"] * a["ay"]
        return np.stack(
            [a["e"], ae * (uax * a["c"] - vay * a["s"]), ae * (uax * a["s"] + vay * a["c"]),
             ae * a["u"] * uax * self.ap.rcp(p[:, 3:4]), ae * a["v"] * vay * self.ap.rcp(p[:, 4:5]),
             ae * a["u"] * a["v"] * (a["ay"] - a["ax"]), np.ones_like(ae)], axis=2)

    def delta(self, p, d, aux, aux_new):
        dt = self.dt
        a, b = aux, aux_new
        d1, d2, dth = d[:, 1:2], d[:, 2:3], d[:, 5:6]
        # cos/sin increments from the half-angle identities (no cancellation)
        sh = np.sin(dt(0.5) * dth).astype(dt)
        thm = p[:, 5:6] + dt(0.5) * dth
        dc = -2 * np.sin(thm).astype(dt) * sh
        dsn = 2 * np.cos(thm).astype(dt) * sh
        du = -b["c"] * d1 - b["s"] * d2 + dc * a["dx"] + dsn * a["dy"]
        dv = -b["c"] * d2 + b["s"] * d1 + dc * a["dy"] - dsn * a["dx"]
        du2 = du * (2 * a["u"] + du)
        dv2 = dv * (2 * a["v"] + dv)
        dax = -d[:, 3:4] * (2 * p[:, 3:4] + d[:, 3:4]) * a["ax"] * b["ax"]
        day = -d[:, 4:5] * (2 * p[:, 4:5] + d[:, 4:5]) * a["ay"] * b["ay"]
        darg = dt(-0.5) * (du2 * b["ax"] + a["u"] * a["u"] * dax + dv2 * b["ay"] + a["v"] * a["v"] * day)
        em1 = np.expm1(darg).astype(dt)
        return d[:, 0:1] * b["e"] + p[:, 0:1] * a["e"] * em1 + d[:, 6:7]


class MultiExp:
    """a0, k0, a1, k1, ..., [offset]."""

    def __init__(self, xaxis, n_param, dtype, ap):
        self.x = np.asarray(xaxis, dtype)[None, :]
        self.P = n_param
        self.ne = n_param // 2
        self.off = n_param % 2 == 1
        self.dt = dtype
        self.ap = ap

    def eval(self, p):
        es = [self.ap.exp(-p[:, 2 * i + 1 : 2 * i + 2] * self.x) for i in range(self.ne)]
        f = sum(p[:, 2 * i : 2 * i + 1] * es[i] for i in range(self.ne))
        if self.off:
            f = f + p[:, -1:]
        return f.astype(self.dt), dict(es=es)

    def jac(self, p, aux):
        cols = []
        for i in range(self.ne):
            cols += [aux["es"][i], -p[:, 2 * i : 2 * i + 1] * self.x * aux["es"][i]]
        if self.off:
            cols.append(np.ones_like(cols[0]))
        return np.stack(cols, axis=2)

    def delta(self, p, d, aux, aux_new):
        out = 0
        for i in range(self.ne):
            em1 = np.expm1(-d[:, 2 * i + 1 : 2 * i + 2] * self.x).astype(self.dt)
            out = out + d[:, 2 * i : 2 * i + 1] * aux_new["es"][i] + p[:, 2 * i : 2 * i + 1] * aux["es"][i] * em1
        if self.off:
            out = out + d[:, -1:]
        return out.astype(self.dt)


def _solve_spd(a, g, dt):
    """Solve a dx = -g per fit; returns (dx, ok).  ``ok`` False mimics LinAlgError (non-positive pivot)."""
    n, p, _ = a.shape
    a = a.astype(dt).copy()
    b = (-g).astype(dt).copy()
    ok = np.ones(n, bool)
    L = np.zeros_like(a)
    with np.errstate(invalid="ignore", divide="ignore"):
        for j in range(p):
            s = a[:, j, j] - (L[:, j, :j] ** 2).sum(1)
            ok &= s > 0
            ljj = np.sqrt(np.where(s > 0, s, 1)).astype(dt)
            L[:, j, j] = ljj
            for i in range(j + 1, p):
                L[:, i, j] = (a[:, i, j] - (L[:, i, :j] * L[:, j, :j]).sum(1)) / ljj
        z = np.zeros_like(b)
        for i in range(p):
            z[:, i] = (b[:, i] - (L[:, i, :i] * z[:, :i]).sum(1)) / L[:, i, i]
        x = np.zeros_like(b)
        for i in reversed(range(p)):
            x[:, i] = (z[:, i] - (L[:, i + 1 :, i] * x[:, i + 1 :]).sum(1)) / L[:, i, i]
    ok &= np.isfinite(x).all(1)
    return x.astype(dt), ok


def _norm(v):
    return np.sqrt((v * v).sum(1))


def _chi2_mle_direct(f, y, dt):
    with np.errstate(invalid="ignore", divide="ignore"):
        t = y * np.log(f / y)
    t[~np.isfinite(t)] = 0
    return ((f - y).sum(1) - t.sum(1)).astype(dt)


def _dchi_mle(fa, fb, delta, y, dt, rfa):
    """chi2(fa) - chi2(fb) from per-pixel differences; fb = fa + delta in the regular case."""
    with np.errstate(invalid="ignore", divide="ignore"):
        regular = (fa > 0) & (fb > 0)
        t = -delta + y * np.log1p(np.where(regular, delta * rfa, 0)).astype(dt)
        # irregular pixels (a clamped model value): evaluate both terms directly, reference zeroing rule
        la = y * np.log(fa / y)
        lb = y * np.log(fb / y)
        la[~np.isfinite(la)] = 0
        lb[~np.isfinite(lb)] = 0
        t_irr = (fa - fb) - la + lb
    return np.where(regular, t, t_irr).astype(dt).sum(1, dtype=dt)


def fit(model, data, p0, method, dtype=np.float32, approx=True, ftol=1.49012e-8, xtol=1.49012e-8,
        maxfev=None, factor=100.0, ls_tol=1e-7, ls_ftol=1.49012e-8, ls_maxit=100, differential=True, trace=False, ls_trace=None):
    """Emulate the kernel on a batch.  Returns dict(popt, info, nfev, chisq, p_ls, n_ls)."""
    dt = dtype
    n = data.shape[0]
    y_raw = data.reshape(n, -1).astype(dt)
    P = model.P
    if maxfev is None:
        maxfev = 100 * (P + 1)
    eyeP = np.eye(P, dtype=dt)[None]

    # ------------------------------------------------------------------ stage A: tight LS from p0
    x = p0.astype(dt).copy()
    f, aux = model.eval(x)
    r = (f - y_raw).astype(dt)
    J = model.jac(x, aux).astype(dt)
    A = np.einsum("nmi,nmj->nij", J, J).astype(dt)
    g = np.einsum("nmi,nm->ni", J, r).astype(dt)
    D = np.einsum("nii->ni", A).copy()
    lam = np.zeros(n, dt)
    active = np.ones(n, bool)
    ls_fail = np.zeros(n, bool)
    n_ls = np.zeros(n, np.int32)
    chi_ls = dt(0.5) * (r * r).sum(1, dtype=dt)
    for it in range(ls_maxit):
        if not active.any():
            break
        dx, ok = _solve_spd(A + lam[:, None, None] * D[:, :, None] * eyeP, g, dt)
        dx = np.where(ok[:, None], dx, 0).astype(dt)
        xn = (x + dx).astype(dt)
        fn, auxn = model.eval(xn)
        with np.errstate(invalid="ignore", over="ignore"):
            if differential:
                delta = model.delta(x, dx, aux, auxn).astype(dt)
            else:
                delta = (fn - f).astype(dt)
            actual = -(delta * (r + dt(0.5) * delta)).sum(1, dtype=dt)
            pred = dt(0.5) * (dx * (lam[:, None] * D * dx - g)).sum(1, dtype=dt)
            finite = ok & np.isfinite(actual) & np.isfinite(fn).all(1)
            good = finite & (actual > 1e-4 * pred) & (pred > 0)
            # MINPACK's ftol test (lmder.f: |actred| <= ftol, prered <= ftol, ratio <= 2), in half-sum-of-squares units
            fconv = finite & (np.abs(actual) <= dt(ls_ftol) * chi_ls) & (pred + dt(0.5) * lam * (D * dx * dx).sum(1) <= dt(ls_ftol) * chi_ls) & (actual <= 2 * pred)
            small = finite & (_norm(np.sqrt(D) * dx) <= ls_tol * _norm(np.sqrt(D) * x))
        acc = active & good
        if ls_trace is not None:
            ls_trace.append(dict(lam=lam.copy(), actual=actual.copy(), pred=pred.copy(), chi=chi_ls.copy(), ok=ok.copy(), good=good.copy(), fconv=fconv.copy(), small=small.copy(), dx=dx.copy(), x=x.copy()))
        rn = (fn - y_raw).astype(dt)
        Jn = model.jac(xn, auxn).astype(dt)
        An = np.einsum("nmi,nmj->nij", Jn, Jn).astype(dt)
        gn = np.einsum("nmi,nm->ni", Jn, rn).astype(dt)
        m = acc
        x[m], f[m], r[m], J[m], A[m], g[m] = xn[m], fn[m], rn[m], Jn[m], An[m], gn[m]
        chi_ls = np.where(m, chi_ls - actual, chi_ls).astype(dt)
        for k in aux:
            if isinstance(aux[k], list):
                for q in range(len(aux[k])):
                    aux[k][q] = np.where(m[:, None], auxn[k][q], aux[k][q])
            else:
                aux[k] = np.where(m[:, None], auxn[k], aux[k])
        D[m] = np.maximum(D[m], np.einsum("nii->ni", An)[m])
        lam_dn = np.where(lam * dt(0.1) < 1e-6, dt(0), lam * dt(0.1))
        lam_up = np.maximum(lam * dt(10), dt(1e-3))
        lam = np.where(acc, lam_dn, np.where(active, lam_up, lam)).astype(dt)
        n_ls += active
        done = active & (fconv | small)
        blown = active & (lam > 1e8)
        ls_fail |= blown & ~done
        active &= ~(done | blown)
    ls_fail |= active  # ran out of iterations
    p_ls = x.copy()

    # ------------------------------------------------------------------ stage B: the reference loop
    mle = method == "mle"
    x0 = x
    if mle:
        y = np.where(y_raw <= 0, dt(0), y_raw)
        f0 = np.where(f <= 0, dt(0), f)
        aux0 = aux
    else:
        y = y_raw
        f0 = f
        aux0 = aux

    def update(Jm, fm):
        if mle:
            with np.errstate(invalid="ignore", divide="ignore"):
                rf = model.ap.rcp(fm)
                yf = y * rf
                yf2 = yf * rf
            bad = ~(np.isfinite(yf) & np.isfinite(yf2))
            yf = np.where(bad, 0, yf).astype(dt)
            yf2 = np.where(bad, 0, yf2).astype(dt)
            Am = np.einsum("nmi,nm,nmj->nij", Jm, yf2, Jm).astype(dt)
            gm = np.einsum("nmi,nm->ni", Jm, (1 - yf)).astype(dt)
            return Am, gm, np.where(bad, 0, rf).astype(dt)
        Am = np.einsum("nmi,nmj->nij", Jm, Jm).astype(dt)
        gm = np.einsum("nmi,nm->ni", Jm, fm - y).astype(dt)
        return Am, gm, None

    def chi2(fm):
        if mle:
            return _chi2_mle_direct(fm, y, dt)
        return dt(0.5) * ((fm - y) ** 2).sum(1, dtype=dt)

    J0 = J
    A, g, rf0 = update(J0, f0)
    dtd = np.einsum("nii->ni", A).copy()
    chisq_old = chi2(f0)
    with np.errstate(invalid="ignore"):
        lam = np.sqrt((x0 * dtd * x0).sum(1, dtype=dt)).astype(dt)
    lam = np.where((lam <= 0) | ~np.isfinite(lam), dt(100), lam).astype(dt)
    xret = x0.copy()
    fret = f0.copy()
    info = np.full(n, 5, np.int32)
    nfev = np.zeros(n, np.int32)
    active = ~ls_fail
    traces = [] if trace else None
    for ev in range(maxfev):
        if not active.any():
            break
        nfev[active] = ev
        dx, ok = _solve_spd(A + lam[:, None, None] * dtd[:, :, None] * eyeP, g, dt)
        sing = active & ~ok
        lam = np.where(sing, lam * dt(factor), lam).astype(dt)
        live = active & ok
        xt = live & (_norm(dx) <= dt(xtol) * (_norm(x0) + dt(xtol)))
        info[xt] = 2
        active &= ~xt
        live &= ~xt
        dxl = np.where(live[:, None], dx, 0).astype(dt)
        xn = (x0 + dxl).astype(dt)
        fn_raw, auxn = model.eval(xn)
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            jdx = np.einsum("nmi,ni->nm", J0, dxl).astype(dt)
            if mle:
                fn = np.where(fn_raw <= 0, dt(0), fn_raw).astype(dt)
                if differential:
                    delta = model.delta(x0, dxl, aux0, auxn).astype(dt)
                    delta = np.where((fn_raw <= 0) | (f0 <= 0), fn - f0, delta).astype(dt)
                    actual = _dchi_mle(f0, fn, delta, y, dt, rf0)
                    fp = (fn + jdx).astype(dt)
                    predicted = _dchi_mle(f0, fp, (delta + jdx).astype(dt), y, dt, rf0)
                    predicted = np.where(fp.min(1) < 0, -np.inf, predicted)
                    chisq_new = None
                else:
                    chisq_new = chi2(fn)
                    fp = (fn + jdx).astype(dt)
                    chisq_pred = np.where(fp.min(1) < 0, np.inf, chi2(fp))
                    actual = chisq_old - chisq_new
                    predicted = chisq_old - chisq_pred
            else:
                fn = fn_raw
                if differential:
                    delta = model.delta(x0, dxl, aux0, auxn).astype(dt)
                    r0 = f0 - y
                    actual = -(delta * (r0 + dt(0.5) * delta)).sum(1, dtype=dt)
                    dp = (delta + jdx).astype(dt)
                    predicted = -(dp * (r0 + dt(0.5) * dp)).sum(1, dtype=dt)
                    chisq_new = None
                else:
                    chisq_new = chi2(fn)
                    actual = chisq_old - chisq_new
                    predicted = chisq_old - chi2(fn + jdx)
            rho = actual / predicted
            rho = np.where((actual < 0) | (predicted < 0) | ~np.isfinite(rho), 0, rho)
        xret = np.where(live[:, None], xn, xret)
        fret = np.where(live[:, None], fn, fret)
        accept = live & (rho > 1e-2)
        reject = live & ~accept
        fin = accept & (actual <= dt(ftol) * chisq_old)
        info[fin] = 1
        active &= ~fin
        upd = accept & ~fin
        if traces is not None:
            traces.append((lam.copy(), actual.copy(), predicted.copy(), rho.copy(), accept.copy(), live.copy()))
        Jn = model.jac(xn, auxn).astype(dt)
        An, gn, rfn = update(Jn, fn)
        m = upd
        x0 = np.where(m[:, None], xn, x0)
        if chisq_new is None:
            chisq_old = np.where(m, chisq_old - actual, chisq_old).astype(dt)
        else:
            chisq_old = np.where(m, chisq_new, chisq_old).astype(dt)
        f0 = np.where(m[:, None], fn, f0)
        if mle:
            rf0 = np.where(m[:, None], rfn, rf0)
        J0 = np.where(m[:, None, None], Jn, J0)
        A = np.where(m[:, None, None], An, A)
        g = np.where(m[:, None], gn, g)
        for k in aux0:
            if isinstance(aux0[k], list):
                aux0[k] = [np.where(m[:, None], auxn[k][q], aux0[k][q]) for q in range(len(aux0[k]))]
            else:
                aux0[k] = np.where(m[:, None], auxn[k], aux0[k])
        dtd = np.where(m[:, None], np.fmax(dtd, np.einsum("nii->ni", An)), dtd)
        lam = np.where(m, np.maximum(lam / dt(5), dt(1e-7)), lam)
        lam = np.where(reject, np.minimum(lam * dt(1.5), dt(1e7)), lam).astype(dt)
    nfev[active] = maxfev - 1
    info[ls_fail] = -1
    out = dict(popt=xret, info=info, nfev=nfev, chisq=chi2(fret), p_ls=p_ls, n_ls=n_ls)
    if trace:
        out["trace"] = traces
    return out
