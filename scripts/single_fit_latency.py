"""BASELINE configs[0]: ONE 15x15 symmetric-Gaussian MLE fit through the reference-shaped call curve_fit(...) (host arrays in,
popt/pcov out), against the float64 oracle port on one host core."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402
from oracle import lm_oracle  # noqa: E402  (timed beside it, as the CPU baseline)

w = synth.WORKLOADS["c1_gauss2d_15"](64, seed=0)
xy = models.roi_grid(15, 15)


def gpu(i):
    return dph.curve_fit(models.gauss2d, xy, w["data"][i].ravel().astype(float), w["p0"][i].astype(float), jac=models.gauss2d.jac, method="mle")


def cpu(i):
    return lm_oracle.curve_fit(models.gauss2d, xy, w["data"][i].ravel().astype(float), w["p0"][i].astype(float), models.gauss2d_jac, "mle")


for name, fn in (("B200 curve_fit", gpu), ("float64 oracle port, 1 core", cpu)):
    for i in range(8):
        fn(i)
    ts = []
    for i in range(64):
        t = time.perf_counter()
        fn(i)
        ts.append(time.perf_counter() - t)
    print("%-28s median %.3f ms  min %.3f ms per fit" % (name, 1e3 * np.median(ts), 1e3 * min(ts)))
a, b = gpu(3)[0], cpu(3)[0]
print("max rel diff of popt:", np.abs(a - b).max() / np.abs(b).max())
