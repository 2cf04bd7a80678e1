"""Latency of the longest straggler of bench input set 2 (a pre-fit that runs into maxfev = 600) when it is alone on the GPU,
against the same fit inside the full 1e5 batch: how much of its chain is issue-slot contention with the main kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

w = synth.WORKLOADS["c2_gauss2d_11"](100000, seed=2)
d, p = torch.from_numpy(w["data"]).cuda(), torch.from_numpy(w["p0"]).cuda()
r = dph.curve_fit_batch(models.gauss2d, d, p, "mle", return_counts=True)
c = r.counts.cpu().numpy()
k = int(c[:, 0].argmax())
print("fit", k, "pre-fit evals", c[k, 0], "info", int(r.info[k]))


def timed(dd, pp, reps=30, **kw):
    for _ in range(5):
        dph.curve_fit_batch(models.gauss2d, dd, pp, "mle", **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dph.curve_fit_batch(models.gauss2d, dd, pp, "mle", **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("alone, handed to the poller after 16 evals: %.3f ms" % timed(d[k:k + 1].contiguous(), p[k:k + 1].contiguous()))
print("alone, 32 lanes from the start:             %.3f ms" % timed(d[k:k + 1].contiguous(), p[k:k + 1].contiguous(), group_lanes=32))
print("alone, 4 lanes, no hand-off:                %.3f ms" % timed(d[k:k + 1].contiguous(), p[k:k + 1].contiguous(), group_lanes=4) if False else "")
keep = torch.ones(100000, dtype=torch.bool, device="cuda")
keep[k] = False
print("full batch:             %.3f ms" % timed(d, p))
print("full batch without it:  %.3f ms" % timed(d[keep].contiguous(), p[keep].contiguous()))
# the same fit moved to the front of the batch: its chain starts at t = 0
perm = torch.cat([torch.tensor([k], device="cuda"), torch.arange(100000, device="cuda")[keep]])
print("full batch, it first:   %.3f ms" % timed(d[perm].contiguous(), p[perm].contiguous()))
