"""Small driver for ncu / timing sweeps: fit one synthetic workload a few times on cuda:0.

    python scripts/profile_run.py c2_gauss2d_11 mle 100000 --lanes 16 --reps 3
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

MODEL = {"c1": models.gauss2d, "c2": models.gauss2d, "c3": models.gauss2d, "c4": models.gauss2d_elliptical, "c5": models.multi_exp,
         "c6": models.gauss2d_integrated, "c7": models.gauss2d_integrated_fixed(1.3), "c9": models.gauss2d_fixed(1.3)}

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("method")
ap.add_argument("n", type=int)
ap.add_argument("--lanes", type=int, nargs="*", default=[0])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
w = synth.WORKLOADS[a.workload](a.n, seed=1)
data, p0 = torch.from_numpy(w["data"]).cuda(), torch.from_numpy(w["p0"]).cuda()
model = models.resolve("multi_exp_convolved", w["consts"]) if a.workload.startswith("c8") else MODEL[a.workload[:2]]
for lanes in a.lanes:
    ts = []
    for r in range(a.reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = dph.curve_fit_batch(model, data, p0, a.method, w["xaxis"], group_lanes=lanes, return_counts=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    c = res.counts.double().mean(0).cpu().numpy()
    info = res.info.cpu().numpy()
    print("%s %s n=%d lanes=%d: best %.3f ms  median %.3f ms -> %.3e fits/s | pre-fit evals %.2f acc %.2f rej %.2f | converged %.5f"
          % (a.workload, a.method, a.n, lanes, min(ts), float(np.median(ts)), a.n / (min(ts) * 1e-3), c[0], c[1], c[2],
             ((info >= 1) & (info <= 4)).mean()))
