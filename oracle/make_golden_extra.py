"""Extra golden vectors from the LIVE reference (run in the build container only).  TEST INFRASTRUCTURE.

    python oracle/make_golden_extra.py

* ``tests/golden/x_lm.npz``: the reference's second public function, ``dphtools.utils.lm.lm(func, x0, Dfun=..., method=)``
  (``lm.py:221-507``), called DIRECTLY from the raw initial guesses of the c2 / c5 workloads (no least-squares pre-fit),
  ``method='mle'`` and ``'ls'``: ``popt``, ``info``, ``nfev`` per fit.  Pins ``dphtools_b200.lm`` / ``DPHLM_FLAG_SKIP_PREFIT``.
* ``tests/golden/x_pcov.npz``: ``pcov`` as returned by the reference's ``curve_fit(..., method='mle')`` (``lm.py:672-677``:
  pseudo-inverse from the SVD of the Jacobian at the last accepted point) for the first fits of the c2, c4 and c5 golden
  sets.  Pins the batched ``return_pcov='jtj'``.
"""

import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.environ.get("DPHTOOLS_REF", "/root/reference"))

from dphtools.utils import lm as ref_lm  # noqa: E402  (the live reference)

from dphtools_b200 import models, synth  # noqa: E402
from oracle.golden import GOLDEN_DIR, load_golden  # noqa: E402


def lm_direct(workload, n, seed, method):
    w = synth.WORKLOADS[workload](n, seed)
    model = models.gauss2d if workload.startswith("c2") else models.multi_exp
    xdata = w["xaxis"].astype(np.float64) if w["xaxis"] is not None else models.roi_grid(*w["data"].shape[1:])
    flat = w["data"].reshape(n, -1).astype(np.float64)
    p0 = w["p0"].astype(np.float64)
    popt = np.empty_like(p0)
    info = np.empty(n, np.int32)
    nfev = np.empty(n, np.int32)
    for i in range(n):
        wrap = ref_lm._wrap_func_mle if method == "mle" else ref_lm._wrap_func_ls
        x, _, infodict, _, ier = ref_lm.lm(wrap(model, xdata, flat[i], None), p0[i], Dfun=lambda p: model.jac(xdata, *p),
                                           full_output=1, method=method)
        popt[i], info[i], nfev[i] = x, ier, infodict["nfev"]
    return dict(data=w["data"].astype(np.uint16), p0=w["p0"], popt=popt, info=info, nfev=nfev,
                xaxis=np.zeros(0, np.float32) if w["xaxis"] is None else w["xaxis"])


def pcov_of(name, n):
    g = load_golden(name)
    model = models.resolve(g["model"], g["consts"])
    xdata = g["xaxis"].astype(np.float64) if g["xaxis"] is not None else models.roi_grid(*g["data"].shape[1:])
    out = []
    for i in range(n):
        y = g["data"][i].ravel().astype(np.float64)
        _, pcov = ref_lm.curve_fit(model, xdata, y, g["p0"][i].astype(np.float64), jac=model.jac, method=g["method"])
        out.append(pcov)
    return np.stack(out)


if __name__ == "__main__":
    sets = {}
    for key, (workload, n, seed, method) in {
        "c2_mle": ("c2_gauss2d_11", 400, 21, "mle"), "c2_ls": ("c2_gauss2d_11", 200, 22, "ls"),
        "c5_mle": ("c5_multiexp_256", 200, 23, "mle"),
    }.items():
        r = lm_direct(workload, n, seed, method)
        for k, v in r.items():
            sets[key + "." + k] = v
        vals, cnt = np.unique(r["info"], return_counts=True)
        print(key, "lm() direct: info", dict(zip(vals.tolist(), cnt.tolist())), "nfev mean %.1f max %d" % (r["nfev"].mean(), r["nfev"].max()))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "x_lm.npz"), **sets)
    pc = {name: pcov_of(name, 200) for name in ("c2_mle", "c4_mle", "c5_mle")}
    np.savez_compressed(os.path.join(GOLDEN_DIR, "x_pcov.npz"), **pc)
    print("pcov sets", {k: v.shape for k, v in pc.items()})
