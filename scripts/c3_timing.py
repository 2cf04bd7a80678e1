import sys, numpy as np, torch, os
sys.path.insert(0, '/root/repo')
import dphtools_b200 as dph
from dphtools_b200 import models, synth, _lib
n = int(os.environ.get("C3N", "2000000"))
parts = [synth.WORKLOADS["c3_gauss2d_7"](250000, seed=100 + i) for i in range(n // 250000)]
data = torch.cat([torch.from_numpy(p["data"]) for p in parts]).cuda(); p0 = torch.cat([torch.from_numpy(p["p0"]) for p in parts]).cuda()
dph.curve_fit_batch(models.gauss2d, data[:1000], p0[:1000], "mle"); torch.cuda.synchronize()
for rep in range(2):
    r = dph.curve_fit_batch(models.gauss2d, data, p0, "mle", timing=True, return_counts=True)
    a, b, _tail = _lib.last_kernel_ms()
    c = r.counts.cpu().numpy()
    print("n=%d: prefit(main) %.2f ms, rest (lm main + pollers) %.2f ms -> %.3e fits/s | parked-ish: A>16: %d  B>=24: %d  A>=300: %d" % (n, a, b, n / (a + b) * 1e3, (c[:, 0] > 16).sum(), (r.nfev.cpu().numpy() >= 24).sum(), (c[:, 0] >= 300).sum()))
