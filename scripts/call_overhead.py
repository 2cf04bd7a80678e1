"""Fixed cost of one dphlm_fit_batch call on a stream: back-to-back calls on small batches (device-resident)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

w = synth.WORKLOADS["c2_gauss2d_11"](100000, seed=1)
d, p = torch.from_numpy(w["data"]).cuda(), torch.from_numpy(w["p0"]).cuda()
for n in (1, 256, 4096, 14208, 28416, 50000, 100000):
    for _ in range(20):
        dph.curve_fit_batch(models.gauss2d, d[:n], p[:n], "mle")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        dph.curve_fit_batch(models.gauss2d, d[:n], p[:n], "mle")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    print("n=%6d: %.3f ms per call, %.3e fits/s" % (n, ms, n / ms * 1e3))
