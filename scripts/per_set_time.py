"""Per-call time of each of bench.py's input sets (one call = 1e5 11x11 MLE fits), and of the sets rotated as the bench does.
Shows how much of a call is the critical path of its slowest straggler (counts: pre-fit evaluations, lm() trips)."""
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


def timed(fn, reps=60):
    for _ in range(10):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for k in range(4):
    r = dph.curve_fit_batch(models.gauss2d, dd[k], dp[k], "mle", return_counts=True)
    c = r.counts.cpu().numpy()
    worst = c[:, 0].argmax()
    ms = timed(lambda i: dph.curve_fit_batch(models.gauss2d, dd[k], dp[k], "mle"))
    print("set %d: %.3f ms per call | max pre-fit evals %d (fit %d), max lm() trips %d, fits > 16 evals: %d, > 24 trips: %d"
          % (k, ms, c[:, 0].max(), worst, (c[:, 1] + c[:, 2]).max(), (c[:, 0] > 16).sum(), ((c[:, 1] + c[:, 2]) > 24).sum()))
print("rotating: %.3f ms per call" % timed(lambda i: dph.curve_fit_batch(models.gauss2d, dd[i % 4], dp[i % 4], "mle")))
