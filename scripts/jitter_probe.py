import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import dphtools_b200 as dph
from dphtools_b200 import models, synth
w = synth.WORKLOADS['c2_gauss2d_11'](100000, seed=1)
data, p0 = torch.from_numpy(w['data']).cuda(), torch.from_numpy(w['p0']).cuda()
for trial in range(2):
    ts, hs = [], []
    for r in range(12):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        res = dph.curve_fit_batch(models.gauss2d, data, p0, 'mle')
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append(e0.elapsed_time(e1)); hs.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    print('gpu ms :', ' '.join('%.2f' % t for t in ts))
    print('host enqueue ms:', ' '.join('%.2f' % h[0] for h in hs))
    print('host total ms  :', ' '.join('%.2f' % h[1] for h in hs))
# back-to-back without sync
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(20): res = dph.curve_fit_batch(models.gauss2d, data, p0, 'mle')
e1.record(); torch.cuda.synchronize(); print('20 back-to-back: %.2f ms per call' % (e0.elapsed_time(e1) / 20))
