"""Trip budgets of the main kernels (DPHLM_BUDGET_A / DPHLM_BUDGET_B) against the rotating-sets step time of bench.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

n = 100000
sets = [synth.WORKLOADS["c2_gauss2d_11"](n, seed=s) for s in range(4)]
dd = [torch.from_numpy(s["data"]).cuda() for s in sets]
dp = [torch.from_numpy(s["p0"]).cuda() for s in sets]


def timed(reps=80):
    for i in range(8):
        dph.curve_fit_batch(models.gauss2d, dd[i % 4], dp[i % 4], "mle")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        dph.curve_fit_batch(models.gauss2d, dd[i % 4], dp[i % 4], "mle")
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


timed(40)
for rep in range(2):
    for a, b in ((16, 24), (12, 24), (10, 24), (12, 20), (14, 22), (20, 28), (16, 20)):
        os.environ["DPHLM_BUDGET_A"] = str(a)
        os.environ["DPHLM_BUDGET_B"] = str(b)
        print("budget A=%2d B=%2d: %.3f ms per step" % (a, b, timed()))
