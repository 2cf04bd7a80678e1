"""A/B a library environment switch inside one process: alternates the variable between blocks of steps and prints the
device-resident and host-entry (pinned buffers) throughput of each block.  usage: ab_env.py VAR [n_fits] [workload]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

var = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
wl = sys.argv[3] if len(sys.argv) > 3 else "c2_gauss2d_11"
sets = [synth.WORKLOADS[wl](n, seed=s) for s in range(4)]
dd = [torch.from_numpy(s["data"]).cuda() for s in sets]
dp = [torch.from_numpy(s["p0"]).cuda() for s in sets]
hd, hp = dph.pinned_empty(sets[0]["data"].shape), dph.pinned_empty(sets[0]["p0"].shape)
hd[...] = sets[0]["data"]
hp[...] = sets[0]["p0"]
P = sets[0]["p0"].shape[1]
out = dph.BatchResult(dph.pinned_empty((n, P)), dph.pinned_empty(n, np.int32), dph.pinned_empty(n, np.int32), dph.pinned_empty(n), None, None)
t0 = time.perf_counter()
i = 0
while time.perf_counter() - t0 < 1.0:
    dph.curve_fit_batch(models.gauss2d, dd[i % 4], dp[i % 4], "mle")
    i += 1
torch.cuda.synchronize()
res = {0: [], 1: []}
for rep in range(8):
    on = rep % 2
    if on:
        os.environ[var] = "1"
    else:
        os.environ.pop(var, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        dph.curve_fit_batch(models.gauss2d, dd[i % 4], dp[i % 4], "mle")
    e1.record()
    torch.cuda.synchronize()
    dev = n * 100 / (e0.elapsed_time(e1) * 1e-3)
    t = time.perf_counter()
    for i in range(40):
        dph.curve_fit_batch(models.gauss2d, hd, hp, "mle", out=out)
    host = n * 40 / (time.perf_counter() - t)
    res[on].append((dev, host))
    print("%s=%d  resident %.3e  host entry %.3e fits/s" % (var, on, dev, host))
for on in (0, 1):
    a = np.array(res[on])
    print("%s=%d median: resident %.3e  host entry %.3e" % (var, on, np.median(a[:, 0]), np.median(a[:, 1])))
