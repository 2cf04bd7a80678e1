"""Would splitting ONE call into two sub-batches on two streams help?  Each step = two sub-calls (fraction f and 1-f of the
batch) on two streams; the next step starts only after both are done (as one stream-ordered call would behave)."""
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
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
main = torch.cuda.current_stream()


def run(frac, steps=100):
    k = int(n * frac)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        d, p = dd[i % 4], dp[i % 4]
        if frac >= 1.0:
            dph.curve_fit_batch(models.gauss2d, d, p, "mle")
            continue
        s1.wait_stream(main)
        s2.wait_stream(main)
        with torch.cuda.stream(s1):
            dph.curve_fit_batch(models.gauss2d, d[:k], p[:k], "mle")
        with torch.cuda.stream(s2):
            dph.curve_fit_batch(models.gauss2d, d[k:], p[k:], "mle")
        main.wait_stream(s1)
        main.wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return n * steps / (e0.elapsed_time(e1) * 1e-3)


for _ in range(3):
    run(1.0, 50)
for rep in range(2):
    for f in (1.0, 0.5, 0.6, 0.7, 0.8):
        print("fraction %.1f: %.3e fits/s" % (f, run(f)))
