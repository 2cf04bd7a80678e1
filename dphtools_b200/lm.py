"""Host-side mirror of ``dphtools.utils.lm`` for the B200 fitter.

``curve_fit`` keeps the reference's signature, guards, return values and exceptions
(``dphtools/utils/lm.py:510-685``); for ``method='ls'|'mle'`` the fit itself runs in the CUDA library
through the C ABI of ``include/dphlm.h`` (no CPU fallback), which restricts ``f`` to the built-in models
of :mod:`dphtools_b200.models`.  ``curve_fit_batch`` is the batched entry point the reference lacks:
one call fits N independent ROIs.

Per-fit ``info`` codes of the batched call are listed in ``include/dphlm.h``; the single-fit wrapper
turns them back into the reference's exceptions.
"""

import ctypes
import logging
import threading
from collections import namedtuple

import numpy as np

from . import _lib, models

logger = logging.getLogger(__name__)  # the reference logs through a module logger too (lm.py:27)

FTOL = XTOL = 1.49012e-8

BatchResult = namedtuple("BatchResult", "popt info nfev chisq p_ls counts pcov x_acc", defaults=(None, None))

_MESSAGES = {  # dphtools/utils/lm.py:310-353 (taken over from scipy.optimize.leastsq)
    1: "Both actual and predicted relative reductions in the sum of squares are at most {ftol}",
    2: "The relative error between two consecutive iterates is at most {xtol}",
    3: "Both actual and predicted relative reductions in the sum of squares\n  are at most {ftol:f} and the "
       "relative error between two consecutive iterates is at \n  most {xtol:f}",
    4: "The cosine of the angle between func(x) and any column of the\n  Jacobian is at most {gtol:f} in absolute value",
    5: "Number of calls to function has reached maxfev = {maxfev:d}.",
}


def _resolve_model(model):
    if isinstance(model, str):
        if model not in models.BUILTIN:
            raise ValueError("unknown built-in model %r (have %s)" % (model, sorted(models.BUILTIN)))
        return models.BUILTIN[model]
    if getattr(model, "model_id", None) is None:
        raise NotImplementedError(
            "method='ls'|'mle' runs on the GPU and supports only the built-in models of dphtools_b200.models "
            "(gauss2d, gauss2d_elliptical, gauss2d_integrated, gauss2d_integrated_fixed(sigma), multi_exp); there is no CPU "
            "fallback for arbitrary callables"
        )
    return model


def _is_torch(a):
    return type(a).__module__.split(".")[0] == "torch"


def _fill_desc(model, method, n_param, dim0, dim1, n_fits, ftol, xtol, gtol, maxfev, factor, group_lanes):
    if method not in ("ls", "mle"):
        raise TypeError("Method {} not recognized".format(method))
    d = _lib.Desc()
    _lib.lib().dphlm_desc_init(ctypes.byref(d))
    d.model, d.method, d.n_param = model.model_id, 1 if method == "mle" else 0, n_param
    d.dim0, d.dim1, d.n_fits, d.group_lanes = dim0, dim1, n_fits, group_lanes
    d.ftol, d.xtol, d.gtol, d.factor = ftol, xtol, gtol, factor
    d.maxfev = int(maxfev) if maxfev else 0
    return d


def _consts_host(model):
    """float32 array of the model's call constants (``dphlm_desc.consts``), or None."""
    c = getattr(model, "consts", None)
    return np.asarray(c, dtype=np.float32) if c else None


def _consts_device(model, device):
    """The same on ``device``, cached on the model so that repeated calls do not copy again."""
    c = _consts_host(model)
    if c is None:
        return None
    import torch

    cache = model.__dict__.setdefault("_consts_dev", {})
    key = str(device)
    if key not in cache:
        cache[key] = torch.as_tensor(c, device=device)
    return cache[key]


def shard_bounds(n, parts):
    """Contiguous split of ``n`` fits over ``parts`` devices / ranks: ``[(lo, hi), ...]``."""
    edges = [(n * k) // parts for k in range(parts + 1)]
    return list(zip(edges[:-1], edges[1:]))


def _check(rc):
    if rc != 0:
        raise RuntimeError("dphlm: " + _lib.last_error())


def curve_fit_batch(model, data, p0, method="mle", xdata=None, *, ftol=FTOL, xtol=XTOL, gtol=0.0, maxfev=None,
                    factor=100.0, return_stage1=False, return_counts=False, group_lanes=0, device=None, timing=False, out=None,
                    return_pcov=None, return_x_acc=False, pipelined=False, skip_prefit=False, wait=True, exact_ties=False):
    """Fit ``N`` independent ROIs / histograms: the batched form of per-ROI
    ``curve_fit(model, xdata, roi.ravel(), p0, jac=model.jac, method=method)`` (``lm.py:510-685``).

    Parameters
    ----------
    model : built-in model callable or its name (``"gauss2d"``, ``"gauss2d_elliptical"``, ``"multi_exp"``)
    data : ``[N, H, W]`` (2-D models) or ``[N, M]`` (``multi_exp``) float32; a uint16 numpy array (raw camera counts) is
        copied as it is -- half the PCIe traffic -- and widened on the device.  A CUDA ``torch.Tensor`` is fitted
        in place on its device and stream (asynchronously; results are CUDA tensors); a numpy array or CPU
        tensor goes through the host entry point (copies overlap the kernels; results are numpy arrays).
    p0 : ``[N, P]`` float32 initial guesses.
    xdata : ``multi_exp`` only: the shared ``[M]`` x axis.
    timing : device path only: record CUDA events around the two kernels (``_lib.last_kernel_ms()``).
    return_x_acc : also return the last ACCEPTED point of ``lm()`` (``[N, P]``): the reference evaluates ``infodict['fjac']``
        and ``pcov`` there, one converged step before ``popt`` (``lm.py:468-473, 489``).
    return_pcov : ``"jtj"`` or ``"crlb"``: also return the ``[N, P, P]`` covariances (see :func:`covariance`).  ``"jtj"`` is the
        reference's ``pcov``: evaluated at the last accepted point (``x_acc``) with its singular-value cut-off.
    pipelined : device path only: enqueue the call on one of the library's own streams and do NOT make the current stream
        wait for it (``DPHLM_FLAG_PIPELINED``): consecutive calls overlap, the drain and the stragglers of one run under the
        next.  The results are pending until :func:`join` has been called; the inputs are kept alive until then.
    wait : host path only: ``False`` returns as soon as the copies and kernels are enqueued (``dphlm_fit_batch_host_async``; use
        :func:`pinned_empty` buffers and leave them alone until :func:`host_wait`), so that the next batch's upload overlaps
        this batch's kernels.
    exact_ties : re-decide in float64 the ``lm()`` trips whose step is within a factor 64 of the ``xtol`` threshold
        (``DPHLM_FLAG_EXACT_TIES``): the fits that end that close -- 0.1 to 2 % of a batch -- then get the reference's
        ``info`` (1 or 2) instead of a coin flip; a few per cent slower.
    skip_prefit : start the ``lm()`` loop directly at ``p0`` (``DPHLM_FLAG_SKIP_PREFIT``): the reference's ``lm()`` rather
        than its ``curve_fit`` (``lm.py:221-236``).
    out : host path only: a ``BatchResult`` of preallocated C-contiguous arrays to write into (use
        :func:`pinned_empty` buffers, for the inputs too, so that the copies run asynchronously).
    device : host path only: CUDA device index, or a list of indices to shard the batch over several
        GPUs (contiguous shards, one host thread per device, no collective).

    Returns ``BatchResult(popt[N,P], info[N], nfev[N], chisq[N], p_ls, counts)``; ``p_ls`` (result of the
    least-squares pre-fit, ``lm.py:648``) and ``counts`` (``[N,4]``: pre-fit evaluations, accepted steps,
    rejected steps, 0) are ``None`` unless requested.  The call never raises for an individual fit.
    """
    model = _resolve_model(model)
    one_d = model.model_id in models.ONE_D
    if _is_torch(data) and data.is_cuda:
        want_acc = return_x_acc or return_pcov == "jtj"
        res = _fit_device(model, data, p0, method, xdata, one_d, ftol, xtol, gtol, maxfev, factor, return_stage1,
                          return_counts, group_lanes, timing, want_acc, pipelined, skip_prefit, exact_ties)
        if return_pcov:
            if pipelined:
                raise ValueError("return_pcov needs the results: call covariance() after join()")
            at = res.x_acc if return_pcov == "jtj" else res.popt
            res = res._replace(pcov=covariance(model, at, data.shape[1:], return_pcov, xdata))
        return res
    if pipelined:
        raise ValueError("pipelined=True needs CUDA tensors (the host entry point overlaps its own chunks; see wait=False)")
    if return_pcov and not wait:
        raise ValueError("return_pcov needs the results: call covariance() after host_wait()")
    if _is_torch(data):
        data = data.numpy()
    if _is_torch(p0):
        p0 = p0.cpu().numpy()
    u16 = isinstance(data, np.ndarray) and data.dtype == np.uint16  # camera counts: copied as they are, widened on the device
    data = np.ascontiguousarray(data, dtype=np.uint16 if u16 else np.float32)
    p0 = np.ascontiguousarray(p0, dtype=np.float32)
    n = data.shape[0]
    if data.ndim != (2 if one_d else 3) or p0.ndim != 2 or p0.shape[0] != n:
        raise ValueError("data must be [N,%s] and p0 [N,P]" % ("M" if one_d else "H,W"))
    n_param = p0.shape[1]
    dim0, dim1 = (data.shape[1], 1) if one_d else data.shape[1:]
    xa = None
    if one_d:
        if xdata is None:
            raise ValueError("multi_exp needs xdata")
        xa = np.ascontiguousarray(xdata, dtype=np.float32)
        if xa.shape != (dim0,):
            raise ValueError("xdata must have shape (M,)")
    if out is None:
        out = BatchResult(
            np.empty((n, n_param), np.float32), np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float32),
            np.empty((n, n_param), np.float32) if return_stage1 else None, np.empty((n, 4), np.int32) if return_counts else None, None,
            np.empty((n, n_param), np.float32) if (return_x_acc or return_pcov == "jtj") else None,
        )
    else:
        want = [((n, n_param), np.float32), ((n,), np.int32), ((n,), np.int32), ((n,), np.float32),
                ((n, n_param), np.float32) if return_stage1 else None, ((n, 4), np.int32) if return_counts else None]
        out = BatchResult(*(tuple(out) + (None,) * (8 - len(out))))
        if (return_x_acc or return_pcov == "jtj") and out.x_acc is None:
            out = out._replace(x_acc=np.empty((n, n_param), np.float32))
        for a, w in zip(out, want):
            if w is not None and (a is None or a.shape != w[0] or a.dtype != w[1] or not a.flags.c_contiguous):
                raise ValueError("out: expected a C-contiguous %s array of shape %s" % (np.dtype(w[1]).name, w[0]))
    devices = [0] if device is None else ([device] if isinstance(device, int) else list(device))

    consts = _consts_host(model)

    def run(dev, lo, hi):
        d = _fill_desc(model, method, n_param, dim0, dim1, hi - lo, ftol, xtol, gtol, maxfev, factor, group_lanes)
        d.data, d.p0 = data[lo:hi].ctypes.data, p0[lo:hi].ctypes.data
        d.xaxis = xa.ctypes.data if xa is not None else None
        d.consts = consts.ctypes.data if consts is not None else None
        d.popt, d.info, d.nfev, d.chisq = (a[lo:hi].ctypes.data for a in tuple(out)[:4])
        d.p_ls = out.p_ls[lo:hi].ctypes.data if return_stage1 else None
        d.counts = out.counts[lo:hi].ctypes.data if return_counts else None
        d.x_acc = out.x_acc[lo:hi].ctypes.data if out.x_acc is not None else None
        d.flags = (_lib.FLAG_DATA_U16 if u16 else 0) | (_lib.FLAG_SKIP_PREFIT if skip_prefit else 0) | (_lib.FLAG_EXACT_TIES if exact_ties else 0)
        if not wait:
            _host_pending.setdefault(dev, []).append((data, p0, xa, consts, out))  # keep the buffers alive until host_wait()
            return _lib.lib().dphlm_fit_batch_host_async(ctypes.byref(d), dev)
        return _lib.lib().dphlm_fit_batch_host(ctypes.byref(d), dev)

    if len(devices) == 1:
        _check(run(devices[0], 0, n))
    else:  # contiguous shards, one host thread per device; ctypes releases the GIL
        bounds = shard_bounds(n, len(devices))
        errs = []

        def worker(i):
            rc = run(devices[i], *bounds[i])
            if rc:
                errs.append(_lib.last_error())

        threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise RuntimeError("dphlm: " + errs[0])
    if return_pcov:
        at = out.x_acc if return_pcov == "jtj" else out.popt
        out = out._replace(pcov=covariance(model, at, data.shape[1:], return_pcov, xa))
    return out


_host_pending = {}


def host_wait(device=0):
    """Wait for every ``wait=False`` host call of ``device`` (an index or a list of indices) and raise their errors."""
    for dev in ([device] if isinstance(device, int) else list(device)):
        rc = _lib.lib().dphlm_host_wait(dev)
        _host_pending.pop(dev, None)
        _check(rc)


_pending = {}  # device index -> buffers of pipelined calls that must stay alive until join()


def join(device=None):
    """Make the current stream of ``device`` (default: the current device) wait for every pending ``pipelined=True`` call
    (``dphlm_join``).  Asynchronous; afterwards the results may be used on that stream like any other tensor."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    with torch.cuda.device(dev):
        _check(_lib.lib().dphlm_join(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    _pending.pop(dev.index, None)


def call_status(device=None):
    """True if a straggler kernel of ``device`` gave up waiting since the last call of this function (``dphlm_call_status``;
    never in correct operation).  Synchronises the device.  Fits that were dropped keep ``info == INFO_UNFINISHED``."""
    import torch

    word = ctypes.c_int(0)
    with torch.cuda.device(torch.cuda.current_device() if device is None else device):
        _check(_lib.lib().dphlm_call_status(ctypes.byref(word)))
    if word.value:
        logger.warning("dphlm: a straggler kernel gave up waiting; unfinished fits keep info = INFO_UNFINISHED")
    return bool(word.value)


class _PinnedBlock:
    """Page-locked host memory from ``dphlm_host_alloc``, released when the last array viewing it dies."""

    def __init__(self, nbytes):
        self.ptr = _lib.lib().dphlm_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("dphlm_host_alloc(%d) failed: %s" % (nbytes, _lib.last_error()))
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                _lib.lib().dphlm_host_free(ptr)
            except Exception:  # interpreter shutdown: the process is going away anyway
                pass


def pinned_empty(shape, dtype=np.float32):
    """``numpy.empty`` in page-locked host memory: copies to and from it overlap the kernels."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
    return np.asarray(_PinnedBlock(nbytes))[: int(np.prod(shape)) * np.dtype(dtype).itemsize].view(dtype).reshape(shape)


def _fit_device(model, data, p0, method, xdata, one_d, ftol, xtol, gtol, maxfev, factor, return_stage1, return_counts, group_lanes,
                timing=False, return_x_acc=False, pipelined=False, skip_prefit=False, exact_ties=False):
    import torch

    if data.dtype != torch.float32 or not data.is_contiguous():
        data = data.contiguous().float()
    p0 = torch.as_tensor(p0, dtype=torch.float32, device=data.device).contiguous()
    n = data.shape[0]
    if data.dim() != (2 if one_d else 3) or p0.dim() != 2 or p0.shape[0] != n:
        raise ValueError("data must be [N,%s] and p0 [N,P]" % ("M" if one_d else "H,W"))
    n_param = p0.shape[1]
    dim0, dim1 = (data.shape[1], 1) if one_d else data.shape[1:]
    xa = None
    if one_d:
        if xdata is None:
            raise ValueError("multi_exp needs xdata")
        xa = torch.as_tensor(xdata, dtype=torch.float32, device=data.device).contiguous()
        if tuple(xa.shape) != (dim0,):
            raise ValueError("xdata must have shape (M,)")
    dev = data.device
    popt = torch.empty((n, n_param), dtype=torch.float32, device=dev)
    info = torch.empty(n, dtype=torch.int32, device=dev)
    nfev = torch.empty(n, dtype=torch.int32, device=dev)
    chisq = torch.empty(n, dtype=torch.float32, device=dev)
    p_ls = torch.empty((n, n_param), dtype=torch.float32, device=dev) if return_stage1 else None
    counts = torch.empty((n, 4), dtype=torch.int32, device=dev) if return_counts else None
    x_acc = torch.empty((n, n_param), dtype=torch.float32, device=dev) if return_x_acc else None
    d = _fill_desc(model, method, n_param, dim0, dim1, n, ftol, xtol, gtol, maxfev, factor, group_lanes)
    d.data, d.p0, d.xaxis = data.data_ptr(), p0.data_ptr(), xa.data_ptr() if xa is not None else None
    consts = _consts_device(model, dev)
    d.consts = consts.data_ptr() if consts is not None else None
    d.popt, d.info, d.nfev, d.chisq = popt.data_ptr(), info.data_ptr(), nfev.data_ptr(), chisq.data_ptr()
    d.p_ls = p_ls.data_ptr() if return_stage1 else None
    d.counts = counts.data_ptr() if return_counts else None
    d.x_acc = x_acc.data_ptr() if return_x_acc else None
    d.flags = ((_lib.FLAG_TIMING if timing else 0) | (_lib.FLAG_PIPELINED if pipelined else 0) | (_lib.FLAG_SKIP_PREFIT if skip_prefit else 0)
               | (_lib.FLAG_EXACT_TIES if exact_ties else 0))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(_lib.lib().dphlm_fit_batch(ctypes.byref(d), ctypes.c_void_p(stream)))
    res = BatchResult(popt, info, nfev, chisq, p_ls, counts, None, x_acc)
    if pipelined:  # the kernels run on the library's streams: nothing they touch may go back to torch's allocator before join()
        _pending.setdefault(dev.index, []).append((data, p0, xa, consts, res))
    return res


def covariance(model, popt, shape, kind="jtj", xdata=None):
    """Batched parameter covariance from the model Jacobian at ``popt`` (``[N, P]``): ``[N, P, P]``.

    ``kind="jtj"``: ``pinv(J^T J)`` of the unweighted Jacobian -- the ``pcov`` the reference's ``curve_fit`` returns
    (``lm.py:672-677``).  ``kind="crlb"``: ``(J^T diag(1/f) J)^-1``, the Poisson Cramer-Rao bound.  ``shape`` is the ROI
    ``(H, W)`` or ``(M,)`` for ``multi_exp`` (which also needs ``xdata``).  CUDA tensors stay on the device (asynchronous on
    the current stream); numpy input is staged through torch and returned as numpy.
    """
    import torch

    model = _resolve_model(model)
    kinds = {"jtj": 0, "crlb": 1}
    if kind not in kinds:
        raise ValueError("kind must be 'jtj' or 'crlb'")
    as_numpy = not _is_torch(popt)
    pt = torch.as_tensor(popt, dtype=torch.float32)
    if not pt.is_cuda:
        pt = pt.cuda()
    pt = pt.contiguous()
    n, n_param = pt.shape
    shape = tuple(shape)
    dim0, dim1 = (shape[0], 1) if len(shape) == 1 else shape
    d = _fill_desc(model, "mle", n_param, dim0, dim1, n, FTOL, XTOL, 0.0, None, 100.0, 0)
    xa = None
    if model.model_id in models.ONE_D:
        if xdata is None:
            raise ValueError("multi_exp needs xdata")
        xa = torch.as_tensor(np.asarray(xdata, dtype=np.float32), device=pt.device).contiguous()
        d.xaxis = xa.data_ptr()
    consts = _consts_device(model, pt.device)
    d.consts = consts.data_ptr() if consts is not None else None
    pcov = torch.empty((n, n_param, n_param), dtype=torch.float32, device=pt.device)
    with torch.cuda.device(pt.device):
        stream = torch.cuda.current_stream(pt.device).cuda_stream
        _check(_lib.lib().dphlm_covariance(ctypes.byref(d), pt.data_ptr(), pcov.data_ptr(), kinds[kind], ctypes.c_void_p(stream)))
    return pcov.cpu().numpy() if as_numpy else pcov


def extract_rois(frames, peaks, roi=11, sigma0=1.3, refine=False):
    """Cut ``roi x roi`` windows around ``peaks`` (``[N, 3]`` int32: frame, row, column) out of ``frames``
    (``[F, H, W]`` float32 or uint16/int16 camera counts) and derive the ``gauss2d`` initial guesses, on the device.

    Window and guess follow ``slice_maker`` / ``localize_peak`` of the reference (``dphtools/utils/__init__.py:244-295,
    590-617``); see ``include/dphlm.h``.  Returns CUDA tensors ``(rois[N,roi,roi] float32, p0[N,5], valid[N] bool)`` ready
    for :func:`curve_fit_batch`; numpy inputs are staged through torch.
    """
    import torch

    fr = torch.as_tensor(frames)
    if fr.dtype == torch.uint16:
        fr, code = fr.view(torch.int16), 1  # raw camera counts: the kernel reads them as unsigned short
    else:
        fr, code = fr.float(), 0  # includes int16 (background-subtracted, signed) frames
    fr = fr.cuda().contiguous()
    pk = torch.as_tensor(peaks, dtype=torch.int32).to(fr.device).contiguous()
    if fr.dim() != 3 or pk.dim() != 2 or pk.shape[1] != 3:
        raise ValueError("frames must be [F,H,W] and peaks [N,3] (frame, row, column)")
    if refine and roi % 2 == 0:
        raise ValueError("refine needs an odd roi (the parabola vertex assumes a window symmetric about its centre)")
    n = pk.shape[0]
    rois = torch.empty((n, roi, roi), dtype=torch.float32, device=fr.device)
    p0 = torch.empty((n, 5), dtype=torch.float32, device=fr.device)
    valid = torch.empty(n, dtype=torch.int32, device=fr.device)
    d = _lib.ExtractDesc()
    d.frames, d.frame_dtype, d.n_frames, d.height, d.width = fr.data_ptr(), code, fr.shape[0], fr.shape[1], fr.shape[2]
    d.peaks, d.n_peaks, d.roi, d.refine, d.sigma0 = pk.data_ptr(), n, roi, int(bool(refine)), sigma0
    d.rois, d.p0, d.valid = rois.data_ptr(), p0.data_ptr(), valid.data_ptr()
    with torch.cuda.device(fr.device):
        stream = torch.cuda.current_stream(fr.device).cuda_stream
        _check(_lib.lib().dphlm_extract_rois(ctypes.byref(d), ctypes.c_void_p(stream)))
    return rois, p0, valid.bool()


def _grid_shape(xdata):
    """(h, w) if ``xdata`` is the (2, M) pixel grid of :func:`models.roi_grid`, else None."""
    if xdata.ndim != 2 or xdata.shape[0] != 2:
        return None
    w, h = int(xdata[0].max()) + 1, int(xdata[1].max()) + 1
    if h * w != xdata.shape[1] or not np.array_equal(xdata, models.roi_grid(h, w)):
        return None
    return h, w


def curve_fit(f, xdata, ydata, p0=None, sigma=None, absolute_sigma=False, check_finite=True,
              bounds=(-np.inf, np.inf), method=None, jac=None, **kwargs):
    """Drop-in for ``dphtools.utils.lm.curve_fit`` (``lm.py:510-685``).

    ``method`` in ``{None, 'lm', 'trf', 'dogbox'}`` delegates to :func:`scipy.optimize.curve_fit` exactly
    as the reference does (``lm.py:606-619``).  ``method='ls'|'mle'`` runs the two-stage fit (least-squares
    pre-fit, then the damped normal-equation loop with the Gaussian / Poisson estimator) on the GPU for the
    built-in models; guards and exceptions follow ``lm.py:621-639, 651-662, 679-685``.
    """
    return_full = kwargs.pop("full_output", False)
    if method in {"lm", "trf", "dogbox", None}:
        import scipy.optimize

        can_full_output = method not in {"trf", "dogbox"} and np.array_equal(bounds, (-np.inf, np.inf))
        if can_full_output:
            kwargs["full_output"] = return_full
        res = scipy.optimize.curve_fit(f, xdata, ydata, p0, sigma, absolute_sigma, check_finite, bounds, method, jac, **kwargs)
        if return_full and not can_full_output:
            return res[0], res[1], None, "No error", 1
        return res
    if method not in ("ls", "mle"):
        raise TypeError("Method {} not recognized".format(method))
    if bounds != (-np.inf, np.inf):
        raise NotImplementedError("Bounds has not been implemented")
    if sigma is not None:
        raise NotImplementedError("Weighting has not been implemented")
    if jac is None:
        raise NotImplementedError("You need a Jacobian")
    model = _resolve_model(f)
    if jac is not model.jac:
        raise NotImplementedError("method='ls'|'mle' uses the analytic Jacobian built into the kernels: pass jac=%s.jac" % model.model_name)
    unknown = set(kwargs) - {"ftol", "xtol", "gtol", "maxfev", "factor", "epsfcn", "diag", "col_deriv"}
    if unknown:
        raise TypeError("unexpected keyword arguments %s" % sorted(unknown))

    ydata = np.asarray_chkfinite(ydata, dtype=float) if check_finite else np.asarray(ydata, dtype=float)
    xdata = np.asarray_chkfinite(xdata, dtype=float) if check_finite else np.asarray(xdata, dtype=float)
    if p0 is None:
        if model.n_param is None:
            raise ValueError("Unable to determine number of fit parameters.")
        p0 = np.ones(model.n_param)
    p0 = np.atleast_1d(np.asarray(p0, dtype=float))
    if p0.size > ydata.size:
        raise TypeError("Improper input: func input vector length N=%d must not exceed func output vector length M=%d" % (p0.size, ydata.size))
    if model.model_id in models.ONE_D:
        data, xa = ydata.reshape(1, -1), xdata.ravel()
    else:
        shape = _grid_shape(xdata)
        if shape is None or shape[0] * shape[1] != ydata.size:
            raise NotImplementedError("2-D models need xdata = models.roi_grid(h, w), the pixel grid of a row-major ROI")
        data, xa = ydata.reshape((1,) + shape), None
    ftol, xtol, gtol = kwargs.get("ftol", FTOL), kwargs.get("xtol", XTOL), kwargs.get("gtol", 0.0)
    maxfev = kwargs.get("maxfev") or 100 * (p0.size + 1)
    res = curve_fit_batch(model, data, p0[None], method, xa, ftol=ftol, xtol=xtol, gtol=gtol, maxfev=maxfev,
                          factor=kwargs.get("factor", 100.0), return_x_acc=True)
    info, popt = int(res.info[0]), res.popt[0].astype(float)
    logger.debug("curve_fit(method=%r): info %d after %d lm() evaluations", method, info, int(res.nfev[0]) + 1)
    fmt = dict(ftol=ftol, xtol=xtol, gtol=gtol, maxfev=maxfev)
    if info == -100:
        raise ValueError("array must not contain infs or NaNs")
    if info < 0:  # the scipy pre-fit of the reference raises here (lm.py:642-644)
        raise RuntimeError("Optimal parameters not found: " + _MESSAGES.get(-info, "pre-fit failed (ier=%d)" % -info).format(**fmt))
    errmsg = _MESSAGES[info].format(**fmt)
    # pcov = pinv(J^T J) from the SVD of the unweighted Jacobian at the last ACCEPTED point, as in the reference
    # (lm.py:672-677 with fjac from lm.py:468-473, 489)
    fjac = model.jac(xdata, *res.x_acc[0].astype(float))
    _, s, vt = np.linalg.svd(fjac, full_matrices=False)
    keep = s > np.finfo(float).eps * max(fjac.shape) * s[0]
    s, vt = s[keep], vt[: keep.sum()]
    pcov = np.dot(vt.T / s**2, vt)
    if info not in (1, 2, 3, 4):
        raise RuntimeError("Optimal parameters not found: " + errmsg)
    if return_full:
        fvec = model(xdata, *popt)
        fvec = np.where(fvec <= 0, 0.0, fvec) if method == "mle" else fvec - ydata
        return popt, pcov, dict(fvec=fvec, fjac=fjac, nfev=int(res.nfev[0]), chisq=float(res.chisq[0])), errmsg, info
    return popt, pcov


def _fit_arrays(model, xdata, ydata, x0, what):
    """(data [1, ...], shared x axis or None) of one fit for :func:`curve_fit_batch`."""
    if model.model_id in models.ONE_D:
        return ydata.reshape(1, -1), xdata.ravel()
    shape = _grid_shape(xdata)
    if shape is None or shape[0] * shape[1] != ydata.size:
        raise NotImplementedError("2-D models need xdata = models.roi_grid(h, w), the pixel grid of a row-major ROI (%s)" % what)
    return ydata.reshape((1,) + shape), None


def lm(func, x0, args=(), Dfun=None, full_output=False, col_deriv=True, ftol=FTOL, xtol=XTOL, gtol=0.0, maxfev=None,
       epsfcn=None, factor=100, diag=None, method="ls"):
    """Counterpart of ``dphtools.utils.lm.lm`` (``lm.py:221-507``): the damped normal-equation loop ALONE, from ``x0``,
    without the least-squares pre-fit that ``curve_fit`` puts in front of it (``DPHLM_FLAG_SKIP_PREFIT``).

    Upstream ``func`` is a closure over the data -- ``func(params)`` returns the residuals (``'ls'``) or the pair
    ``(model, data)`` (``'mle'``, ``lm.py:121-146``) -- which a GPU kernel cannot look into.  Here ``func`` is a built-in
    model of :mod:`dphtools_b200.models`, ``Dfun`` its ``.jac`` and ``args = (xdata, ydata)``; everything else (keywords,
    guards in upstream order, return values, ``infodict`` keys, warnings and exceptions for ``info`` outside 1..4 when
    ``full_output`` is false, ``lm.py:491-507``) is as upstream.
    """
    x0 = np.asarray(x0, dtype=float).flatten()
    if not isinstance(args, tuple):
        args = (args,)
    if Dfun is None:
        raise NotImplementedError
    if not col_deriv:
        raise NotImplementedError("Column derivatives required")
    if maxfev is None:
        maxfev = 100 * (len(x0) + 1)
    if method not in ("ls", "mle"):
        raise TypeError("Method {} not recognized".format(method))
    model = _resolve_model(func)
    if Dfun is not model.jac:
        raise NotImplementedError("lm() uses the analytic Jacobian built into the kernels: pass Dfun=%s.jac" % model.model_name)
    if len(args) != 2:
        raise TypeError("lm() of dphtools_b200 takes args=(xdata, ydata)")
    xdata, ydata = np.asarray(args[0], dtype=float), np.asarray(args[1], dtype=float)
    data, xa = _fit_arrays(model, xdata, ydata, x0, "args[0]")
    res = curve_fit_batch(model, data, x0[None], method, xa, ftol=ftol, xtol=xtol, gtol=gtol, maxfev=maxfev, factor=float(factor),
                          return_x_acc=True, skip_prefit=True)
    info, popt = int(res.info[0]), res.popt[0].astype(float)
    fmt = dict(ftol=ftol, xtol=xtol, gtol=gtol, maxfev=maxfev)
    errmsg = _MESSAGES.get(info, "Unknown error.")
    errmsg = errmsg.format(**fmt) if info in _MESSAGES else errmsg
    logger.debug("Ended with %d function evaluations", int(res.nfev[0]) + 1)
    if info not in (1, 2, 3, 4) and not full_output:
        if info in (5, 6, 7, 8):
            logger.warning(errmsg)  # lm.py:491-493
        else:
            raise TypeError(errmsg)  # lm.py:494-498 (errors["unknown"])
    logger.debug(errmsg)
    if not full_output:
        return popt, None
    fvec = model(xdata, *popt)
    fvec = np.where(fvec <= 0, 0.0, fvec) if method == "mle" else fvec - ydata
    fjac = model.jac(xdata, *res.x_acc[0].astype(float))  # J at the last accepted point (lm.py:468-473, 489)
    return popt, None, dict(fvec=fvec, fjac=fjac, nfev=int(res.nfev[0])), errmsg, info
