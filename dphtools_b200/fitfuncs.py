"""Multi-exponential fit drivers: host-side mirror of ``dphtools/utils/fitfuncs.py:63-159``.

``multi_exp_fit`` keeps the reference's signature and guess logic (finite mask, log-spaced splitting of the x range into
one segment per component, log-linear estimate per segment, inner offsets dropped) and hands the guess to
:func:`dphtools_b200.curve_fit`; with ``method='ls'|'mle'`` in ``kwargs`` the fit runs on the GPU, otherwise scipy does it,
exactly as upstream (``fitfuncs.py:155-157`` -> ``lm.py:606-619``).  ``multi_exp_fit_batch`` is the batched form for
``[N, M]`` histograms sharing one x axis.
"""

import numpy as np

from .lm import curve_fit, curve_fit_batch
from .models import multi_exp, multi_exp_jac


def exponent(xdata, amp, rate, offset):
    """Single exponential ``amp * exp(-rate * x) + offset`` (``fitfuncs.py:55-60``)."""
    return multi_exp(xdata, amp, rate, offset)


def estimate_exponent_params(data, xdata):
    """``(amp, rate, offset)`` of a single decaying or rising exponential from a log-linear fit (``fitfuncs.py:63-80``)."""
    data, xdata = np.asarray(data, dtype=float), np.asarray(xdata, dtype=float)
    if not (np.isfinite(data).all() and np.isfinite(xdata).all()):
        raise AssertionError("data is not finite")
    if len(data) != len(xdata) or len(data) == 0:
        raise AssertionError("Lengths don't match")
    if data[0] < data[-1]:  # rising: estimate the mirrored decay
        amp, rate, offset = estimate_exponent_params(-data, xdata)
        return np.array((-amp, rate, -offset))
    offset = np.nanmin(data)
    with np.errstate(divide="ignore"):
        logged = np.log(data - offset)
    ok = np.isfinite(logged)
    slope, intercept = np.polyfit(xdata[ok], logged[ok], 1)
    return np.nan_to_num((np.exp(intercept), -slope, offset))


def _segments(xdata, components):
    """Index ranges that split the x axis into ``components`` log-spaced pieces (``fitfuncs.py:131-144``)."""
    if components <= 1:
        return [slice(None)]
    edges = np.logspace(np.log(xdata[xdata > 0].min()), np.log(xdata.max()), components + 1, base=np.e)
    cuts = [None] + list(np.searchsorted(xdata, edges)[1:-1]) + [None]
    return [slice(a, b) for a, b in zip(cuts[:-1], cuts[1:])]


def guess_multi_exp(data, xdata, components, offset=True):
    """Initial guess ``(a0, k0, ..., [offset])`` for :func:`multi_exp` (``fitfuncs.py:145-152``)."""
    parts = [estimate_exponent_params(data[s], xdata[s]) for s in _segments(xdata, components)]
    guess = np.concatenate([p[:-1] for p in parts[:-1]] + parts[-1:])
    return guess if offset else guess[:-1]


def multi_exp_fit(data, xdata=None, components=None, offset=True, **kwargs):
    """Fit a sum of exponentials; signature and behaviour of the reference's ``multi_exp_fit`` (``fitfuncs.py:93-159``)."""
    data = np.asarray(data)
    xdata = np.arange(len(data)) if xdata is None else np.asarray(xdata)
    if components is None:
        raise NotImplementedError  # the reference's _estimate_components is not implemented either (fitfuncs.py:83-85)
    good = np.isfinite(data)
    y, x = data[good], xdata[good]
    if len(y) <= 3:
        raise RuntimeError("Not enough good points to fit.")
    return curve_fit(multi_exp, x, y, p0=guess_multi_exp(y, x, components, offset), jac=multi_exp_jac, **kwargs)


def exponent_fit(data, xdata=None, offset=True):
    """Single-exponential fit (``fitfuncs.py:88-90``)."""
    return multi_exp_fit(data, xdata, components=1, offset=offset)


def guess_multi_exp_batch(data, xdata, components, offset=True):
    """:func:`guess_multi_exp` for ``[N, M]`` finite histograms on a shared x axis, on the device (``dphlm_guess_multi_exp``):
    the segment split is computed once on the host (it depends on the x axis only), the per-segment log-linear estimates run
    one warp per histogram.  ``data``: CUDA tensor (stays on the device) or numpy array; returns a CUDA tensor ``[N, P]``."""
    import ctypes

    import torch

    from . import _lib

    xdata = np.asarray(xdata, dtype=np.float64)
    d = torch.as_tensor(data)
    d = (d if d.is_cuda else d.cuda()).float().contiguous()
    n, m = d.shape
    if xdata.shape != (m,):
        raise ValueError("xdata must have shape (M,)")
    cuts = [0 if s.start is None else int(s.start) for s in _segments(xdata, components)] + [m]
    bounds = torch.tensor(cuts, dtype=torch.int32, device=d.device)
    xa = torch.as_tensor(xdata.astype(np.float32), device=d.device)
    p0 = torch.empty((n, 2 * components + (1 if offset else 0)), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        rc = _lib.lib().dphlm_guess_multi_exp(d.data_ptr(), xa.data_ptr(), bounds.data_ptr(), components, int(bool(offset)), m, n,
                                              p0.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream(d.device).cuda_stream))
    if rc != 0:
        raise RuntimeError("dphlm: " + _lib.last_error())
    return p0


def multi_exp_fit_batch(data, xdata, components, offset=True, method="mle", **kwargs):
    """Batched :func:`multi_exp_fit` for ``[N, M]`` histograms on a shared x axis: the reference's guess recipe on the device
    (:func:`guess_multi_exp_batch`), then one :func:`curve_fit_batch` call, everything on the GPU (a numpy ``data`` is uploaded
    once; the results come back as numpy arrays then).  Histograms with non-finite bins -- the reference masks those bins per
    histogram (``fitfuncs.py:124-126``), which changes that histogram's length -- are fitted one by one with
    :func:`multi_exp_fit` afterwards (``info = -100`` marks them in the batch result until then; failures leave NaN).
    Returns ``(BatchResult, p0)``."""
    import torch

    as_numpy = not (type(data).__module__.split(".")[0] == "torch")
    d = torch.as_tensor(data)
    d = (d if d.is_cuda else d.cuda()).float().contiguous()
    xdata = np.asarray(xdata, dtype=np.float64)
    finite = torch.isfinite(d).all(dim=1)
    clean = torch.where(finite[:, None], d, torch.ones_like(d)) if not bool(finite.all()) else d
    p0 = guess_multi_exp_batch(clean, xdata, components, offset)
    res = curve_fit_batch(multi_exp, clean, p0, method, xdata.astype(np.float32), **kwargs)
    bad = torch.nonzero(~finite).flatten().tolist()
    if bad:
        res.info[~finite] = -100
        res.popt[~finite] = float("nan")
        host = d[~finite].cpu().numpy().astype(float)
        single = {k: v for k, v in kwargs.items() if k in ("ftol", "xtol", "gtol", "maxfev", "factor")}
        for row, i in zip(host, bad):
            try:
                popt, _, infodict, _, ier = multi_exp_fit(row, xdata, components, offset, method=method, full_output=True, **single)
            except (RuntimeError, ValueError, TypeError):
                continue
            res.popt[i] = torch.as_tensor(popt, dtype=torch.float32)
            res.info[i], res.nfev[i] = int(ier), int(infodict["nfev"])
            res.chisq[i] = float(infodict.get("chisq", float("nan")))
    if as_numpy:
        torch.cuda.synchronize(d.device)
        res = type(res)(*[None if v is None else v.cpu().numpy() for v in res])
        p0 = p0.cpu().numpy()
    return res, p0
