"""Development aid: device-resident throughput of one library build (DPHLM_LIB selects the .so), stream-ordered and pipelined,
plus the per-kernel times of timed calls.  usage: quick_rate.py [n_fits] [workload] [group_lanes] [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import _lib, models, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
wl = sys.argv[2] if len(sys.argv) > 2 else "c2_gauss2d_11"
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 40
method = os.environ.get("METHOD", "mle")
name = {"c1": "gauss2d", "c2": "gauss2d", "c3": "gauss2d", "c4": "gauss2d_elliptical", "c5": "multi_exp", "c6": "gauss2d_integrated",
        "c7": "gauss2d_integrated_fixed", "c8": "multi_exp_convolved", "c9": "gauss2d_fixed"}[wl[:2]]
sets = [synth.WORKLOADS[wl](n, seed=1000 + s) for s in range(4)]
model = models.resolve(name, sets[0].get("consts"))
xa = sets[0]["xaxis"]
dd = [torch.from_numpy(s["data"]).cuda() for s in sets]
dp = [torch.from_numpy(s["p0"]).cuda() for s in sets]


def step(i, **kw):
    return dph.curve_fit_batch(model, dd[i % 4], dp[i % 4], method, xa, group_lanes=lanes, exact_ties=bool(os.environ.get("EXACT_TIES")), **kw)


t0 = time.perf_counter()
i = 0
while time.perf_counter() - t0 < 0.7:
    step(i)
    i += 1
torch.cuda.synchronize()


def rate(pipelined):
    best = []
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i, pipelined=pipelined)
        if pipelined:
            dph.join()
        e1.record()
        torch.cuda.synchronize()
        best.append(n * steps / (e0.elapsed_time(e1) * 1e-3))
    return float(np.median(best))


ser, pip = rate(False), rate(True)
ks = []
for i in range(8):
    torch.cuda.synchronize()
    step(i, timing=True)
    ks.append(_lib.last_kernel_ms())
ks = np.array(ks).mean(0)
r = step(0)
torch.cuda.synchronize()
info = r.info.cpu().numpy()
print("%s lib=%s n=%d G=%d %s: serialized %.3e  pipelined(depth %s) %.3e fits/s | kernels ms: prefit %.3f lm %.3f tail %.3f | converged %.5f" % (
    wl, os.path.basename(_lib.LIB_PATH), n, lanes, method, ser, os.environ.get("DPHLM_PIPE_DEPTH", "2"), pip, ks[0], ks[1], ks[2],
    ((info >= 1) & (info <= 4)).mean()))
