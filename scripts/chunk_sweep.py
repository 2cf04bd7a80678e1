"""Host-entry throughput against the chunk size of dphlm_fit_batch_host (DPHLM_CHUNK).  usage: chunk_sweep.py [n_fits] sizes..."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

n = int(sys.argv[1])
sizes = [int(v) for v in sys.argv[2:]]
w = synth.WORKLOADS["c2_gauss2d_11"](n, seed=3)
hd, hp = dph.pinned_empty(w["data"].shape), dph.pinned_empty(w["p0"].shape)
hd[...] = w["data"]
hp[...] = w["p0"]
out = dph.BatchResult(dph.pinned_empty((n, 5)), dph.pinned_empty(n, np.int32), dph.pinned_empty(n, np.int32), dph.pinned_empty(n), None, None)
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.0:
    dph.curve_fit_batch(models.gauss2d, hd, hp, "mle", out=out)
for rep in range(3):
    for c in sizes:
        if c:
            os.environ["DPHLM_CHUNK"] = str(c)
        else:
            os.environ.pop("DPHLM_CHUNK", None)
        ts = []
        for i in range(30):
            t = time.perf_counter()
            dph.curve_fit_batch(models.gauss2d, hd, hp, "mle", out=out)
            ts.append(time.perf_counter() - t)
        print("chunk %6d: median %.3f ms -> %.3e fits/s" % (c, 1e3 * np.median(ts), n / np.median(ts)))
