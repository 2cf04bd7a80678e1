"""Small mixed workload for compute-sanitizer (memcheck / racecheck): every model, both methods, ragged sizes,
bad inputs, host and device entry points."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

for name, model, n in (("c2_gauss2d_11", models.gauss2d, 777), ("c3_gauss2d_7", models.gauss2d, 1501),
                       ("c4_ellip_15", models.gauss2d_elliptical, 301), ("c5_multiexp_256", models.multi_exp, 257)):
    w = synth.WORKLOADS[name](n, seed=3)
    w["data"][5] = np.nan if w["data"].ndim == 3 else w["data"][5]
    for method in ("mle", "ls"):
        for lanes in (0, 4, 32):
            r = dph.curve_fit_batch(model, torch.from_numpy(w["data"]).cuda(), torch.from_numpy(w["p0"]).cuda(), method, w["xaxis"],
                                    group_lanes=lanes, return_stage1=True, return_counts=True)
            torch.cuda.synchronize()
        rh = dph.curve_fit_batch(model, w["data"], w["p0"], method, w["xaxis"])
        print(name, method, "converged", float(((rh.info >= 1) & (rh.info <= 4)).mean()))
print("done")
