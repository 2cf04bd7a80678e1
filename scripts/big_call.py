"""Development aid: one 7x7 call of n fits (BASELINE configs[2] size by default), repeated; prints the duration of each.
usage: big_call.py [n_fits] [repeats]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
data, p0 = synth.gauss2d_rois_device(n, 7, 77, torch.device("cuda", 0), jitter=0.75)
dph.curve_fit_batch(models.gauss2d, data[:200000], p0[:200000], "mle")
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = dph.curve_fit_batch(models.gauss2d, data, p0, "mle", return_counts=True)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
c = r.counts
print("n=%d POLLER_CTAS=%s: ms %s | pre-fit evals > 16: %d, lm trips >= 24: %d, converged %.6f" % (
    n, os.environ.get("DPHLM_POLLER_CTAS", "8"), " ".join("%.1f" % m for m in ms), int((c[:, 0] > 16).sum()),
    int(((c[:, 1] + c[:, 2]) >= 24).sum()), float(((r.info >= 1) & (r.info <= 4)).double().mean())))
