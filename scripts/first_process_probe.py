"""Per-step device times of the bench loop in the FIRST CUDA process of a fresh box (diagnostic)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph
from dphtools_b200 import models, synth
sets = [synth.WORKLOADS["c2_gauss2d_11"](100000, seed=s) for s in range(4)]
d = [torch.from_numpy(s["data"]).cuda() for s in sets]; p = [torch.from_numpy(s["p0"]).cuda() for s in sets]
for i in range(10): dph.curve_fit_batch(models.gauss2d, d[i % 4], p[i % 4], "mle")
torch.cuda.synchronize()
for rnd in range(3):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(61)]
    t0 = time.perf_counter(); host = []
    evs[0].record()
    for i in range(60):
        h0 = time.perf_counter()
        dph.curve_fit_batch(models.gauss2d, d[i % 4], p[i % 4], "mle")
        evs[i + 1].record(); host.append((time.perf_counter() - h0) * 1e3)
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(60)]
    print("round", rnd, "total %.2f ms/step" % (evs[0].elapsed_time(evs[60]) / 60), "| per-step gpu ms:", " ".join("%.2f" % t for t in ts[:24]), "| host enqueue max %.2f mean %.2f" % (max(host), np.mean(host)))
    time.sleep(0.3)
