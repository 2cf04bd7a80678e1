"""BASELINE.json configs 3-5 at full size on cuda:0 (size-independent checks + throughput).

    python scripts/full_size_configs.py            # c3: 1e7 7x7, c4: 1e6 elliptical 15x15 (ls + mle), c5: 1e6 256-bin
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models, synth  # noqa: E402


def run(name, model, n, method, key):
    chunk = 250_000
    parts = [synth.WORKLOADS[name](min(chunk, n - lo), seed=100 + i) for i, lo in enumerate(range(0, n, chunk))]
    data = torch.cat([torch.from_numpy(p["data"]) for p in parts]).cuda()
    p0 = torch.cat([torch.from_numpy(p["p0"]) for p in parts]).cuda()
    truth = np.concatenate([p["truth"] for p in parts])
    xa = parts[0]["xaxis"]
    dph.curve_fit_batch(model, data, p0, method, xa)  # untimed: grows the stream-ordered workspace pool to this batch size
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = dph.curve_fit_batch(model, data, p0, method, xa)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    info, popt = r.info.cpu().numpy(), r.popt.cpu().numpy().astype(float)
    conv = (info >= 1) & (info <= 4)
    err = popt[conv][:, key] - truth[conv][:, key]
    print("%-18s %-3s n=%.0e: %8.2f ms  %.3e fits/s | converged %.5f | info %s | median |err| of params %s: %s" % (
        name, method, n, ms, n / ms * 1e3, conv.mean(), dict(zip(*[a.tolist() for a in np.unique(info, return_counts=True)])), key,
        np.round(np.median(np.abs(err), axis=0), 4).tolist()))


run("c3_gauss2d_7", models.gauss2d, 10_000_000, "mle", [1, 2, 3])
run("c4_ellip_15", models.gauss2d_elliptical, 1_000_000, "mle", [1, 2, 3, 4])
run("c4_ellip_15", models.gauss2d_elliptical, 1_000_000, "ls", [1, 2, 3, 4])
run("c5_multiexp_256", models.multi_exp, 1_000_000, "mle", [1, 3])
run("c2_gauss2d_11", models.gauss2d, 1_000_000, "mle", [1, 2, 3])
