"""Development aid: end-to-end rate of the host entry (page-locked buffers, calls enqueued without waiting, one wait) for
the DPHLM_CHUNK in the environment.  usage: e2e_rate.py [n_fits] [steps] [group_lanes]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sets = [synth.WORKLOADS["c2_gauss2d_11"](n, seed=s) for s in range(2)]
h = []
for s in sets:
    hd, hp = dph.pinned_empty(s["data"].shape, np.uint16 if os.environ.get("U16") else np.float32), dph.pinned_empty(s["p0"].shape)
    hd[...] = s["data"]
    hp[...] = s["p0"]
    h.append((hd, hp))
outs = [dph.BatchResult(dph.pinned_empty((n, 5)), dph.pinned_empty(n, np.int32), dph.pinned_empty(n, np.int32), dph.pinned_empty(n), None, None)
        for _ in range(steps)]
for i in range(3):
    dph.curve_fit_batch(models.gauss2d, h[i % 2][0], h[i % 2][1], "mle", device=0, out=outs[0])
res, enq = [], []
for rep in range(5):
    t0 = time.perf_counter()
    for i in range(steps):
        dph.curve_fit_batch(models.gauss2d, h[i % 2][0], h[i % 2][1], "mle", device=0, out=outs[i], wait=False, group_lanes=lanes)
    t_enq = time.perf_counter() - t0
    dph.host_wait(0)
    res.append(n * steps / (time.perf_counter() - t0))
    enq.append(t_enq / steps * 1e3)
t0 = time.perf_counter()
for i in range(steps):
    dph.curve_fit_batch(models.gauss2d, h[i % 2][0], h[i % 2][1], "mle", device=0, out=outs[i])
blk = n * steps / (time.perf_counter() - t0)
print("lanes=%d DPHLM_CHUNK=%s n=%d: async e2e %.3e fits/s (median of 5; %.1f GB/s H2D; host enqueue %.3f ms per call), blocking %.3e" % (
    lanes, os.environ.get("DPHLM_CHUNK", "auto"), n, np.median(res), np.median(res) * 504 / 1e9, float(np.median(enq)), blk))
