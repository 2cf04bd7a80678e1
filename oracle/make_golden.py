"""Generate the golden vectors in ``tests/golden/`` by running the LIVE reference.  TEST INFRASTRUCTURE.

Run in the build container only (``/root/reference`` does not exist on the GPU box)::

    python oracle/make_golden.py            # all sets
    python oracle/make_golden.py c2_mle     # one set

For every fit it calls the unmodified ``dphtools.utils.lm.curve_fit(model, xdata, y, p0, jac=...,
method=..., full_output=True)`` imported from ``$DPHTOOLS_REF`` (default ``/root/reference``) and
records ``popt``, ``info``, ``nfev``, and the stage-1 result ``p_ls`` (the same scipy call the reference
makes at ``lm.py:642-644``, repeated here because the reference does not return it).  A fit on which the
reference raises ``RuntimeError`` is stored with ``info = 5`` if stage 2 hit ``maxfev`` and ``info = -1``
if the scipy pre-fit failed.
"""

import os
import sys
import time
from multiprocessing import Pool

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.environ.get("DPHTOOLS_REF", "/root/reference"))

import scipy.optimize  # noqa: E402
from dphtools.utils import lm as ref_lm  # noqa: E402  (the live reference)

from dphtools_b200 import models, synth  # noqa: E402

# name -> (workload, n_fits, seed, method)
SETS = {
    "c1_mle": ("c1_gauss2d_15", 256, 0, "mle"),
    "c2_mle": ("c2_gauss2d_11", 10000, 2, "mle"),
    "c2_ls": ("c2_gauss2d_11", 2000, 12, "ls"),
    "c3_mle": ("c3_gauss2d_7", 10000, 3, "mle"),
    "c4_mle": ("c4_ellip_15", 2000, 4, "mle"),
    "c4_ls": ("c4_ellip_15", 2000, 14, "ls"),
    "c5_mle": ("c5_multiexp_256", 2000, 5, "mle"),
    "c5_ls": ("c5_multiexp_256", 2000, 15, "ls"),
    "c6_mle": ("c6_gauss2d_int_11", 2000, 6, "mle"),
    "c7_mle": ("c7_gauss2d_intf_11", 2000, 7, "mle"),
    "c8_mle": ("c8_multiexp_irf_256", 1000, 8, "mle"),
    "c6_ls": ("c6_gauss2d_int_11", 1000, 16, "ls"),
    "c7_ls": ("c7_gauss2d_intf_11", 1000, 17, "ls"),
    "c8_ls": ("c8_multiexp_irf_256", 500, 18, "ls"),
    "c9_mle": ("c9_gauss2d_fixed_11", 1000, 9, "mle"),
}
MODEL_OF = {"c1": models.gauss2d, "c2": models.gauss2d, "c3": models.gauss2d, "c4": models.gauss2d_elliptical, "c5": models.multi_exp, "c6": models.gauss2d_integrated,
            "c7": models.gauss2d_integrated_fixed(1.3), "c8": models.multi_exp_convolved([1.0], 1.0),
            "c9": models.gauss2d_fixed(1.3)}

_G = {}


def _init(model_name, method, xdata, consts):
    _G.update(model=models.resolve(model_name, consts), method=method, xdata=xdata)


def _one(args):
    y, p0 = args
    model, method, xdata = _G["model"], _G["method"], _G["xdata"]
    npar = p0.size
    try:
        p_ls = scipy.optimize.curve_fit(model, xdata, y, p0, jac=model.jac)[0]
    except RuntimeError:
        return np.full(npar, np.nan), np.full(npar, np.nan), -1, 0
    try:
        popt, _, infodict, _, info = ref_lm.curve_fit(model, xdata, y, p0, jac=model.jac, method=method, full_output=True)
        return p_ls, popt, info, infodict["nfev"]
    except RuntimeError:
        # stage 2 ran out of iterations: recover popt the same way the reference computed it
        popt, _, infodict, _, info = ref_lm.lm(
            ref_lm._wrap_func_mle(model, xdata, y, None) if method == "mle" else ref_lm._wrap_func_ls(model, xdata, y, None),
            p_ls, Dfun=lambda p: model.jac(xdata, *p), full_output=1, method=method,
        )
        return p_ls, popt, info, infodict["nfev"]


def make(name):
    workload, n, seed, method = SETS[name]
    w = synth.WORKLOADS[workload](n, seed)
    consts = w.get("consts")
    model = models.resolve(MODEL_OF[name[:2]].model_name, consts)
    data, p0 = w["data"], w["p0"]
    if w["xaxis"] is not None:
        xdata = w["xaxis"].astype(np.float64)
    else:
        xdata = models.roi_grid(*data.shape[1:])
    flat = data.reshape(n, -1).astype(np.float64)
    t0 = time.time()
    with Pool(os.cpu_count(), initializer=_init, initargs=(model.model_name, method, xdata, consts)) as pool:
        res = pool.map(_one, zip(flat, p0.astype(np.float64)), chunksize=64)
    dt = time.time() - t0
    p_ls = np.stack([r[0] for r in res])
    popt = np.stack([r[1] for r in res])
    info = np.array([r[2] for r in res], np.int32)
    nfev = np.array([r[3] for r in res], np.int32)
    assert data.max() < 65536 and (data == np.floor(data)).all()
    out = os.path.join(REPO, "tests", "golden", name + ".npz")
    np.savez_compressed(
        out, data=data.astype(np.uint16), p0=p0, p_ls=p_ls, popt=popt, info=info, nfev=nfev,
        truth=w["truth"], xaxis=np.zeros(0, np.float32) if w["xaxis"] is None else w["xaxis"],
        meta=np.array([workload, str(n), str(seed), method, model.model_name]),
        consts=np.asarray(consts if consts else [], np.float32),
    )
    vals, cnt = np.unique(info, return_counts=True)
    print(f"{name}: {n} fits in {dt:.1f}s ({n / dt:.0f} fits/s on {os.cpu_count()} procs)  info={dict(zip(vals.tolist(), cnt.tolist()))}"
          f"  nfev mean {nfev.mean():.2f} max {nfev.max()}  -> {out} ({os.path.getsize(out) / 1e6:.2f} MB)")


if __name__ == "__main__":
    for nm in sys.argv[1:] or SETS:
        make(nm)
