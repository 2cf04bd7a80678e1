"""Driver-style JSON lines for the BASELINE.json configurations other than the headline one (bench.py --workload c3|c4|c5|...).

One step = one pass of the hot path over one batch of synthetic inputs resident in HBM (pipelined calls, joined before the
closing event, like bench.py's `value`); `e2e` = the same through the host entry with page-locked buffers; `roofline` from the
per-fit counters the kernels return and the per-kernel times of timed calls.  Algorithmic flops per pixel: SURVEY.md 8(d).
"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

# key: synth workload, model name, method, fits per step, (F_0, F_acc, F_rej, F_ls) per pixel, metric text, workload text
CONFIGS = {
    "c3": ("c3_gauss2d_7", "gauss2d", "mle", 1_000_000, (75, 93, 36, 73), "Gaussian-PSF MLE fits/sec (7x7 ROI)",
           "1e6 symmetric 2D Gaussian MLE fits per step, 7x7 ROIs (BASELINE configs[2] runs 1e7 of these, sharded)"),
    "c4": ("c4_ellip_15", "gauss2d_elliptical", "mle", 200_000, (125, 150, 52, 110), "Elliptical-Gaussian MLE fits/sec (15x15 ROI)",
           "2e5 elliptical rotated 2D Gaussian (7 parameters) MLE fits per step, 15x15 ROIs (BASELINE configs[3])"),
    "c4ls": ("c4_ellip_15", "gauss2d_elliptical", "ls", 200_000, (125, 150, 52, 110), "Elliptical-Gaussian LS fits/sec (15x15 ROI)",
             "2e5 elliptical rotated 2D Gaussian (7 parameters) least-squares fits per step, 15x15 ROIs (BASELINE configs[3])"),
    "c5": ("c5_multiexp_256", "multi_exp", "mle", 200_000, (71, 89, 35, 69), "Bi-exponential MLE fits/sec (256 bins)",
           "2e5 bi-exponential + offset MLE fits per step, 256-bin histograms (BASELINE configs[4])"),
    "c6": ("c6_gauss2d_int_11", "gauss2d_integrated", "mle", 400_000, (75, 93, 36, 73), "Pixel-integrated Gaussian MLE fits/sec (11x11 ROI)",
           "4e5 pixel-integrated symmetric Gaussian MLE fits per step, 11x11 ROIs (SURVEY 8(f) rank 4)"),
    "c9": ("c9_gauss2d_fixed_11", "gauss2d_fixed", "mle", 400_000, (60, 75, 30, 58), "Fixed-width Gaussian MLE fits/sec (11x11 ROI)",
           "4e5 fixed-width (4 parameters) symmetric Gaussian MLE fits per step, 11x11 ROIs (SURVEY 8(f) rank 4)"),
}


def run(key, steps=10, warmup=3, fits=None, group_lanes=0):
    import torch

    import dphtools_b200 as dph
    from dphtools_b200 import _lib, models, synth

    wl, model_name, method, n_default, (f0, facc, frej, fls), metric, text = CONFIGS[key]
    n = fits or n_default
    sets = [synth.WORKLOADS[wl](n, seed=2000 + s) for s in range(2)]
    model = models.resolve(model_name, sets[0].get("consts"))
    xa = sets[0]["xaxis"]
    dd = [torch.from_numpy(s["data"]).cuda() for s in sets]
    dp = [torch.from_numpy(s["p0"]).cuda() for s in sets]
    M = int(np.prod(sets[0]["data"].shape[1:]))
    P = sets[0]["p0"].shape[1]

    def step(i, **kw):
        return dph.curve_fit_batch(model, dd[i % 2], dp[i % 2], method, xa, group_lanes=group_lanes, **kw)

    t0 = time.perf_counter()
    i = 0
    while i < warmup or time.perf_counter() - t0 < 0.5:
        step(i)
        i += 1
    torch.cuda.synchronize()
    reps = []
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            res = step(i, pipelined=True)
        dph.join()
        e1.record()
        torch.cuda.synchronize()
        reps.append(e0.elapsed_time(e1))
    ms = float(np.median(reps))
    value = n * steps / (ms * 1e-3)
    ks, fl_pre, fl_main = [], [], []
    for i in range(4):
        torch.cuda.synchronize()
        r = step(i, timing=True, return_counts=True)
        ks.append(_lib.last_kernel_ms())
        c = r.counts.sum(0).double().cpu().numpy()
        fl_pre.append(M * fls * c[0])
        fl_main.append(M * (f0 * n + facc * c[1] + frej * c[2]))
    k_pre, k_main, k_tail = (float(v) * 1e-3 for v in np.array(ks).mean(0))
    # end to end through the host entry (page-locked buffers, calls enqueued without waiting, one wait)
    hd, hp = dph.pinned_empty(sets[0]["data"].shape), dph.pinned_empty(sets[0]["p0"].shape)
    hd[...] = sets[0]["data"]
    hp[...] = sets[0]["p0"]
    e2e_steps = max(2, min(steps, 4))
    outs = [dph.BatchResult(dph.pinned_empty((n, P)), dph.pinned_empty(n, np.int32), dph.pinned_empty(n, np.int32), dph.pinned_empty(n), None, None)
            for _ in range(e2e_steps)]
    dph.curve_fit_batch(model, hd, hp, method, xa, device=0, out=outs[0])
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        dph.curve_fit_batch(model, hd, hp, method, xa, device=0, out=outs[i], wait=False)
    dph.host_wait(0)
    e2e = n * e2e_steps / (time.perf_counter() - t0)
    info = res.info.cpu().numpy()
    peak = torch.cuda.get_device_properties(0).multi_processor_count * 128 * 2 * 1.965e9 / 1e12
    measured = _lib.measure_fp32_peak()
    stage_tf = float(np.mean(fl_main)) / (k_main + k_tail) / 1e12
    step_tf = (float(np.mean(fl_pre)) + float(np.mean(fl_main))) / (k_pre + k_main + k_tail) / 1e12
    return {
        "metric": metric, "value": value, "unit": "fits/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": text, "fits_per_step": n, "method": method, "pipeline": "pipelined calls, joined before the closing event",
                   "l2": "two input sets of %.0f MB each" % (n * M * 4 / 1e6), "converged_fraction": float(((info >= 1) & (info <= 4)).mean())},
        "e2e": {"value": e2e, "unit": "fits/s", "h2d_bytes_per_step": n * (M + P) * 4, "d2h_bytes_per_step": n * (P + 3) * 4, "steps": e2e_steps},
        "gpu_launches": 5 * steps,
        "roofline": {"bound": "fp32", "achieved": stage_tf, "peak": peak, "unit": "TFLOP/s", "frac": stage_tf / peak, "traffic": None,
                     "peak_measured": measured, "kernel": "lm() stage of %s (%s)" % (model_name, method), "kernel_ms": k_main * 1e3,
                     "stage_ms": (k_main + k_tail) * 1e3, "prefit_kernel_ms": k_pre * 1e3, "flops_per_fit": float(np.mean(fl_main)) / n,
                     "step": {"achieved": step_tf, "frac": step_tf / peak,
                              "flops_per_fit": (float(np.mean(fl_pre)) + float(np.mean(fl_main))) / n},
                     "hbm_bytes_per_fit": 4 * M + 8 * P + 12},
    }


if __name__ == "__main__":
    for k in sys.argv[1:] or list(CONFIGS):
        print(json.dumps(run(k)), flush=True)
