#!/usr/bin/env python
"""Headline benchmark: Gaussian-PSF MLE fits/s on 11x11 ROIs (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # the reference algorithm on the host cores

One "step" = one pass of the hot path over one batch of 1e5 synthetic Poisson-noised spots per GPU (weak
scaling: each rank fits its own shard, no data-path collective).  `value` is whole-job
fits/s with the inputs resident in HBM; `e2e` is the same through the public API with pinned HOST buffers
(H2D + kernel + D2H inside the timed region).  The reference arm times the reference's own
dphtools.utils.lm.curve_fit (scipy MINPACK pre-fit + lm()) from oracle/_ref (placed there by oracle/build_ref.py at
build time; the float64 port oracle/lm_oracle.py only if that is absent) with one process per host core on a bounded
sample of the same workload.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOAD = "c2_gauss2d_11"
ROI = 11
N_FITS = 100_000          # per GPU and step
N_SETS = 4                # distinct input sets rotated between steps: 4 x 48.4 MB > 126 MB L2
METRIC = "Gaussian-PSF MLE fits/sec (11x11 ROI)"
# algorithmic flops per pixel: setup, accepted / rejected MLE trip, LS pre-fit evaluation
F_0, F_ACC, F_REJ, F_LS = 75, 93, 36, 73
BYTES_PER_FIT = 4 * ROI * ROI + 4 * 5 + 4 * 5 + 12  # 536 B


def _cpu_fit_chunk(args):
    data, p0 = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from dphtools_b200 import models
    from oracle import build_ref, lm_oracle

    xy = models.roi_grid(ROI, ROI)
    ref = build_ref.load()
    if ref is None:  # the float64 port
        out = lm_oracle.fit_batch(models.gauss2d, data, p0, "mle", xy)
        return out["popt"], out["info"], out["nfev"]
    import logging

    logging.getLogger("dphtools.utils.lm").setLevel(logging.ERROR)
    n = data.shape[0]
    popt, info, nfev = np.full((n, 5), np.nan), np.zeros(n, np.int32), np.zeros(n, np.int32)
    for i in range(n):  # per-ROI dphtools.utils.lm.curve_fit, exactly what a user of the reference runs (lm.py:510-685)
        try:
            popt[i], _, infodict, _, info[i] = ref.curve_fit(models.gauss2d, xy, data[i].ravel().astype(float), p0[i].astype(float),
                                                            jac=models.gauss2d.jac, method="mle", full_output=True)
            nfev[i] = infodict["nfev"]
        except RuntimeError:  # "Optimal parameters not found" from either stage
            info[i] = 5
    return popt, info, nfev


def cpu_kind():
    from oracle import build_ref

    return "reference" if build_ref.available() else "port"


def cpu_reference(data, p0, cores):
    """The reference path (oracle/_ref, else the oracle port) on `cores` processes; returns (fits/s, popt, info, nfev)."""
    import multiprocessing as mp

    chunks = [(data[i::cores], p0[i::cores]) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_fit_chunk, [(c[0][:2], c[1][:2]) for c in chunks])  # warm the workers (imports)
        t0 = time.perf_counter()
        res = pool.map(_cpu_fit_chunk, chunks)
        dt = time.perf_counter() - t0
    n = data.shape[0]
    popt = np.empty((n, 5))
    info = np.empty(n, np.int32)
    nfev = np.empty(n, np.int32)
    for i, r in enumerate(res):
        popt[i::cores], info[i::cores], nfev[i::cores] = r
    return n / dt, popt, info, nfev


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled through NVML while the benchmark runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.active = False  # set around the timed regions: only those samples count
        self.ready = threading.Event()  # NVML initialised (nvmlInit takes ~100 ms and must not land in a timed region)

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            for _ in range(3):  # first calls are slow on a freshly booted driver: keep them out of the timed regions
                nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.ready.set()
            while not self.stop_flag:
                # query all the time (a freshly booted driver serves its first NVML queries slowly and stalls kernel
                # launches meanwhile: those must have happened before the timed regions), record only when active
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.active:
                    self.samples.append(mhz)
                    self.reasons |= {v for k, v in names.items() if r & k}
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report that, never fail the run
            self.reasons.add("nvml_unavailable:" + type(e).__name__)
        self.ready.set()

    def summary(self):
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dphtools_b200 import synth

    cores = os.cpu_count()
    kind = cpu_kind()
    per_step = max(256, (100 if kind == "reference" else 400) * cores)  # ~1.5-2 s per step
    w = synth.WORKLOADS[WORKLOAD](per_step * (args.steps + args.warmup), seed=100)
    rates = []
    for s in range(args.steps + args.warmup):
        sl = slice(s * per_step, (s + 1) * per_step)
        rate, _, info, _ = cpu_reference(w["data"][sl], w["p0"][sl], cores)
        if s >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    sample = "%d fits per step (11x11 sym-Gaussian MLE, same generator as the GPU arm), %d steps" % (per_step, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "fits/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per_step / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "1e5 symmetric 2D Gaussian MLE fits, 11x11 ROIs, photons 100-5000 (configs[1]); bounded sample per step", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": cores, "kind": kind, "sample": sample,
                         "what": "dphtools.utils.lm.curve_fit(method='mle') of the reference itself, per ROI (oracle/_ref)" if kind == "reference"
                         else "float64 port oracle/lm_oracle.py (oracle/_ref not built)"},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fits", type=int, default=N_FITS, help="fits per GPU and step")
    ap.add_argument("--group-lanes", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs2", action="store_true", help="skip the informational 1e7 7x7 block (BASELINE configs[2])")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample SM clocks through NVML (diagnostic)")
    ap.add_argument("--workload", default="c2", help="c2 = the headline configuration (default); c3, c4, c4ls, c5, c6, c9: the other "
                    "BASELINE / SURVEY configurations, one GPU, one informational line each (scripts/bench_configs.py)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "c2":
        sys.path.insert(0, os.path.join(REPO, "scripts"))
        import bench_configs

        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps(bench_configs.run(args.workload, steps=min(args.steps, 10), warmup=args.warmup,
                                               fits=None if args.fits == N_FITS else args.fits, group_lanes=args.group_lanes)))
        return

    import torch
    import torch.distributed as dist

    import dphtools_b200 as dph
    from dphtools_b200 import _lib, models, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    try:  # keep the rank (and the pinned buffers it first-touches) on the CPUs next to its GPU
        import pynvml as nv

        nv.nvmlInit()
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = args.fits

    sampler = ClockSampler(local)
    if args.no_clocks:
        sampler.ready.set()
    else:
        sampler.start()

    # ---- synthetic inputs: N_SETS distinct batches per rank (seeded), resident in HBM
    sets = [synth.WORKLOADS[WORKLOAD](n, seed=1000 * rank + s) for s in range(N_SETS)]
    d_data = [torch.from_numpy(s["data"]).to(dev) for s in sets]
    d_p0 = [torch.from_numpy(s["p0"]).to(dev) for s in sets]


    def step(i, counts=False, pipelined=False):
        return dph.curve_fit_batch(models.gauss2d, d_data[i % N_SETS], d_p0[i % N_SETS], "mle", return_counts=counts,
                                   group_lanes=args.group_lanes, pipelined=pipelined)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler.ready.wait(timeout=30)
    for i in range(args.warmup):
        step(i)
    # an idle B200 sits at ~120 MHz and needs tens of milliseconds of load to reach its boost clock: keep the
    # (untimed) warm-up going for at least half a second so that the timed steps do not include the ramp
    t_warm = time.perf_counter()
    i = args.warmup
    while time.perf_counter() - t_warm < 0.5:
        step(i, pipelined=True)
        i += 1
        if i % 8 == 0:
            dph.join()
            torch.cuda.synchronize()
    dph.join()
    barrier()

    def timed_steps(pipelined):
        """K steps, timed with CUDA events on the launch stream and bracketed by barrier + synchronize, max over ranks.
        The block is repeated three times and the median is reported (all three are in the JSON line): the timed region
        is ~0.1 s long and a single host-side hiccup (scheduler, NVML query) would otherwise decide the number."""
        reps, last = [], None
        for _ in range(3):
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for i in range(args.steps):
                last = step(i, pipelined=pipelined)
            if pipelined:
                dph.join()  # the launch stream waits for every pending call: ev1 marks the end of all K steps
            ev1.record()
            barrier()
            ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            reps.append(float(ms.item()))
        return reps, last

    # `value`: the K steps issued as pipelined calls (DPHLM_FLAG_PIPELINED: each call is complete work on its own buffers; the
    # library keeps three in flight on its own streams, so the drain and the stragglers of one step run under the next) and
    # joined before the closing event.  `serialized` below is the same with every call stream-ordered behind the previous one.
    sampler.active = True
    repeats, res = timed_steps(True)
    sampler.active = False
    ms_total = float(np.median(repeats))
    value = world * n * args.steps / (ms_total * 1e-3)
    rep_ser, _ = timed_steps(False)
    value_serialized = world * n * args.steps / (float(np.median(rep_ser)) * 1e-3)

    # ---- the two kernels alone: per-launch durations (CUDA events recorded by the library on the launch
    # stream around each kernel) and algorithmic flops from the per-fit counters the kernels return
    t_pre, t_main, t_tail, fl_pre, fl_main, fl_launch = [], [], [], [], [], []
    for i in range(min(args.steps, 8)):
        torch.cuda.synchronize()
        r = dph.curve_fit_batch(models.gauss2d, d_data[i % N_SETS], d_p0[i % N_SETS], "mle", return_counts=True,
                                group_lanes=args.group_lanes or 2, timing=True)  # two lanes per fit: what a pipelined call of this size picks
        ms_pre, ms_main, ms_tail = _lib.last_kernel_ms()
        t_pre.append(ms_pre * 1e-3)
        t_main.append(ms_main * 1e-3)
        t_tail.append(ms_tail * 1e-3)
        cn = r.counts.double()
        c = cn.sum(0).cpu().numpy()
        fl_pre.append(ROI * ROI * F_LS * c[0])
        fl_main.append(ROI * ROI * (F_0 * n + F_ACC * c[1] + F_REJ * c[2]))
        # the part of those flops the MAIN lm() launch itself executes: a fit whose pre-fit went beyond 16 evaluations runs its
        # whole lm() loop in the poller kernel, one beyond 24 trips only its first 24 there (lm_launch.cuh: kBudgetPrefit, kBudgetMain)
        trips = cn[:, 1] + cn[:, 2]
        part = torch.where(cn[:, 0] > 16, torch.zeros_like(trips), torch.clamp(24.0 / torch.clamp(trips, min=1.0), max=1.0))
        fl_launch.append(ROI * ROI * float((part * (F_0 + F_ACC * cn[:, 1] + F_REJ * cn[:, 2])).sum().item()))
    k_pre, k_main, k_tail = float(np.mean(t_pre)), float(np.mean(t_main)), float(np.mean(t_tail))
    k_s = k_pre + k_main + k_tail  # the whole call: both main kernels and the wait for the pollers' last straggler
    flops = [a + b for a, b in zip(fl_pre, fl_main)]
    # dominant kernel: the lm() main launch -- the algorithmic flops that launch executes over its own duration (the contract's
    # per-launch definition).  `stage` below charges the whole lm() stage instead: all its flops, including the trips the poller
    # finishes, over the main launch plus the wait for the pollers after it (k_tail)
    achieved_tf = float(np.mean(fl_launch)) / k_main / 1e12
    stage_tf = float(np.mean(fl_main)) / (k_main + k_tail) / 1e12
    step_tf = float(np.mean(flops)) / k_s / 1e12               # both kernels of a step
    props = torch.cuda.get_device_properties(dev)
    fp32_measured_tf = _lib.measure_fp32_peak()  # FFMA chains on this GPU, now

    # ---- the platform's ceiling for the end-to-end figure: the bare host-to-device copy of one step's inputs from pinned
    # memory, all ranks at once, nothing else running (GB/s per rank; the slowest rank counts)
    h_probe = torch.from_numpy(dph.pinned_empty(n * (ROI * ROI + 5)))  # the allocator the e2e leg uses (cudaHostAlloc)
    h_probe.zero_()
    d_probe = torch.empty_like(h_probe, device=dev)
    probe_streams = [torch.cuda.Stream(device=dev) for _ in range(8)]  # eight copies in flight, like the host entry's staging slots
    quarter = h_probe.numel() // 8

    def probe_round():
        for k, st_ in enumerate(probe_streams):
            with torch.cuda.stream(st_):
                d_probe[k * quarter:(k + 1) * quarter].copy_(h_probe[k * quarter:(k + 1) * quarter], non_blocking=True)

    def probe_single():  # the same bytes as one copy on one stream
        d_probe.copy_(h_probe, non_blocking=True)

    for _ in range(2):
        probe_round()
        probe_single()
    best_gbs = 0.0
    for rep in range(6):  # best of six blocks of eight rounds, alternating eight copies in flight and one large copy
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if rep % 2 == 0:
            for st_ in probe_streams:
                st_.wait_event(ev0)
            for _ in range(8):
                probe_round()
            for st_ in probe_streams:
                torch.cuda.current_stream().wait_stream(st_)
        else:
            for _ in range(8):
                probe_single()
        ev1.record()
        barrier()
        best_gbs = max(best_gbs, 8 * 8 * quarter * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9)
    h2d_gbs = torch.tensor([best_gbs], device=dev)
    if world > 1:
        dist.all_reduce(h2d_gbs, op=dist.ReduceOp.MIN)
    h2d_gbs = float(h2d_gbs.item())
    del h_probe, d_probe

    # ---- end to end through the public API with pinned host buffers (H2D + fit + D2H per step)
    h_sets = []
    for s in sets[:2]:
        hd, hp = dph.pinned_empty(s["data"].shape), dph.pinned_empty(s["p0"].shape)
        hd[...] = s["data"]
        hp[...] = s["p0"]
        h_sets.append((hd, hp))
    h_out = dph.BatchResult(dph.pinned_empty((n, 5)), dph.pinned_empty(n, np.int32), dph.pinned_empty(n, np.int32),
                            dph.pinned_empty(n), None, None)
    e2e_steps = max(3, min(args.steps, 100))  # one pinned result set (3.2 MB) per step
    # one result set per step in flight: the calls are enqueued without waiting (dphlm_fit_batch_host_async), so the upload of
    # one step overlaps the kernels and the result copies of the one before; every step's H2D and D2H is inside the timed region
    h_outs = [h_out] + [dph.BatchResult(dph.pinned_empty((n, 5)), dph.pinned_empty(n, np.int32), dph.pinned_empty(n, np.int32),
                                        dph.pinned_empty(n), None, None) for _ in range(e2e_steps - 1)]
    for i in range(2):
        dph.curve_fit_batch(models.gauss2d, h_sets[i % 2][0], h_sets[i % 2][1], "mle", device=local, out=h_out)
    for i in range(3):  # and the asynchronous form the timed steps use
        dph.curve_fit_batch(models.gauss2d, h_sets[i % 2][0], h_sets[i % 2][1], "mle", device=local, out=h_outs[i % e2e_steps], wait=False)
    dph.host_wait(local)
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        r_host = dph.curve_fit_batch(models.gauss2d, h_sets[i % 2][0], h_sets[i % 2][1], "mle", device=local, out=h_outs[i], wait=False)
    dph.host_wait(local)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    sampler.active = False
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t_e2e.item())
    # informational: the same steps one blocking call at a time (dphlm_fit_batch_host)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        dph.curve_fit_batch(models.gauss2d, h_sets[i % 2][0], h_sets[i % 2][1], "mle", device=local, out=h_out)
    barrier()
    t_blk = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_blk, op=dist.ReduceOp.MAX)
    e2e_blocking = world * n * e2e_steps / float(t_blk.item())

    # informational: the same call fed the camera's native uint16 counts (half the H2D bytes; widened on the device)
    h_u16 = dph.pinned_empty(sets[0]["data"].shape, np.uint16)
    h_u16[...] = sets[0]["data"]
    for i in range(2):
        dph.curve_fit_batch(models.gauss2d, h_u16, h_sets[0][1], "mle", device=local, out=h_outs[i], wait=False)
    dph.host_wait(local)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        dph.curve_fit_batch(models.gauss2d, h_u16, h_sets[0][1], "mle", device=local, out=h_outs[i], wait=False)
    dph.host_wait(local)
    barrier()
    t_u16 = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_u16, op=dist.ReduceOp.MAX)
    e2e_u16 = world * n * e2e_steps / float(t_u16.item())

    # ---- informational: BASELINE configs[2] -- 1e7 symmetric-Gaussian MLE fits on 7x7 ROIs, sharded over the ranks (strong
    # scaling: every rank fits 1e7 / world of them in ONE call; generated on the device)
    c2_block = None
    if not args.no_configs2:
        n_c2 = 10_000_000 // world
        c2_data, c2_p0 = synth.gauss2d_rois_device(n_c2, 7, 77 + rank, dev, jitter=0.75)
        dph.curve_fit_batch(models.gauss2d, c2_data, c2_p0, "mle")  # untimed: warms the 7x7 kernels and grows the workspace pool
        ms_reps = []
        for _ in range(3):
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            r_c2 = dph.curve_fit_batch(models.gauss2d, c2_data, c2_p0, "mle")
            ev1.record()
            barrier()
            ms_rep = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
            if world > 1:
                dist.all_reduce(ms_rep, op=dist.ReduceOp.MAX)
            ms_reps.append(float(ms_rep.item()))
        ms_c2 = torch.tensor([float(np.median(ms_reps))], device=dev)
        conv_c2 = ((r_c2.info >= 1) & (r_c2.info <= 4)).double().mean().reshape(1)
        if world > 1:
            dist.all_reduce(conv_c2, op=dist.ReduceOp.SUM)
        c2_block = {"workload": "1e7 symmetric 2D Gaussian MLE fits, 7x7 ROIs, sharded across the ranks (BASELINE configs[2]); one call per rank, "
                                "inputs resident, generated on the device", "fits": n_c2 * world, "scaling": "strong", "ms": float(ms_c2.item()), "ms_repeats": ms_reps,
                    "value": n_c2 * world / (float(ms_c2.item()) * 1e-3), "unit": "fits/s", "converged_fraction": float(conv_c2.item()) / world}
        del c2_data, c2_p0, r_c2

    sampler.stop_flag = True
    if sampler.is_alive():
        sampler.join(timeout=2)

    if rank == 0:
        info = res.info.cpu().numpy()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = (sampler.max_mhz or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        fp32_peak_tf = props.multi_processor_count * 128 * 2 * sm_max / 1e12
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        out = {
            "metric": METRIC, "value": value, "unit": "fits/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "ms_per_step_repeats": [r / args.steps for r in repeats], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "serialized": {"value": value_serialized, "ms_per_step": float(np.median(rep_ser)) / args.steps,
                           "note": "informational: the same K steps without DPHLM_FLAG_PIPELINED (each call stream-ordered behind the previous one)"},
            "config": {
                "workload": "1e5 symmetric 2D Gaussian MLE fits per GPU and step, 11x11 ROIs, photons 100-5000 (BASELINE configs[1])",
                "fits_per_gpu_per_step": n, "sharding": "independent fits, contiguous shards, no collective",
                "pipeline": "steps issued as pipelined calls (three in flight on the library's streams), joined before the closing event",
                "l2": "inputs rotate over %d distinct sets (%.0f MB) > 126 MB L2" % (N_SETS, N_SETS * n * ROI * ROI * 4 / 1e6),
                "converged_fraction": float(((info >= 1) & (info <= 4)).mean()),
            },
            "e2e": {"value": e2e_value, "unit": "fits/s", "h2d_bytes_per_step": n * (ROI * ROI + 5) * 4,
                    "d2h_bytes_per_step": n * (5 + 3) * 4, "steps": e2e_steps, "host_memory": "pinned inputs and outputs (dphtools_b200.pinned_empty)",
                    "converged_fraction": float(((r_host.info >= 1) & (r_host.info <= 4)).mean()),
                    "calls": "enqueued without waiting (wait=False), one wait at the end; one result set per step",
                    "blocking": {"value": e2e_blocking, "note": "informational: the same steps, one blocking call at a time"},
                    "h2d_probe": {"gbs_per_rank": h2d_gbs, "ceiling_fits_per_s": world * h2d_gbs * 1e9 / ((ROI * ROI + 5) * 4),
                                  "frac_of_ceiling": e2e_value / (world * h2d_gbs * 1e9 / ((ROI * ROI + 5) * 4)),
                                  "note": "bare pinned host-to-device copy of one step's inputs, all ranks at once (slowest rank): what the platform allows end to end"},
                    "uint16_input": {"value": e2e_u16, "h2d_bytes_per_step": n * (ROI * ROI * 2 + 5 * 4),
                                     "note": "informational: same call with uint16 camera counts as host input"}},
            "gpu_launches": 5 * args.steps,  # per step: the preparation launch, pre-fit + lm() main launches and their two straggler pollers
            "roofline": {
                "bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak_tf,
                # dram__bytes_read+write of the lm() main launch from the ncu --set full capture of this workload
                # (profiles/r2_final_ncu_full_gauss2d_11x11_G2.txt: 50.85 MB read + 0.1 MB written per 1e5 fits; a constant from
                # that capture); algorithmic: 484 B/fit read again + 20 B/fit of pre-fit parameters + 32 B/fit written
                "traffic": 50.95e6 * n / 1e5, "kernel": "dphlm_stage_kernel<StageB<Gauss2D,MLE>> (the lm() loop, lm.py:389-481; the main launch, events around it)",
                "kernel_ms": k_main * 1e3, "stage_ms": (k_main + k_tail) * 1e3, "share_of_step": (k_main + k_tail) / k_s,
                "timing_note": "kernel times come from calls with DPHLM_FLAG_TIMING: events between the launches, so the lm() kernel does not "
                               "start early in the slots the pre-fit kernel frees, and nothing is pipelined; `achieved` = the flops the lm() "
                               "main launch executes / kernel_ms; `stage` = all lm() flops / stage_ms (main launch + straggler tail)",
                "stage": {"achieved": stage_tf, "frac": stage_tf / fp32_peak_tf, "flops_per_fit": float(np.mean(fl_main)) / n},
                "flops_per_fit_in_launch": float(np.mean(fl_launch)) / n,
                "peak_measured": fp32_measured_tf, "frac_of_measured": achieved_tf / fp32_measured_tf,
                "straggler_tail_ms": k_tail * 1e3,  # after the lm() kernel: the pollers finishing the longest straggler (mean over the input sets)
                "peak_source": "SMs x 128 lanes x 2 x max SM clock (no fp32 entry in MEASURED_PEAKS.json)",
                "flops_per_fit": float(np.mean(fl_main)) / n,
                "prefit_kernel": {"kernel": "dphlm_stage_kernel<StageA<Gauss2D>> (MINPACK pre-fit, lm.py:642-648)", "kernel_ms": k_pre * 1e3,
                                  "achieved": float(np.mean(fl_pre)) / k_pre / 1e12, "flops_per_fit": float(np.mean(fl_pre)) / n},
                "step": {"achieved": step_tf, "frac": step_tf / fp32_peak_tf, "flops_per_fit": float(np.mean(flops)) / n},
                # the same flops over the time the headline `value` is measured with (pipelined steps, ms_per_step)
                "step_at_value": {"achieved": float(np.mean(flops)) / (ms_total / args.steps * 1e-3) / 1e12,
                                  "frac": float(np.mean(flops)) / (ms_total / args.steps * 1e-3) / 1e12 / fp32_peak_tf},
                # ncu, sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active of the two main kernels at two lanes per fit
                # (profiles/r2_final_ncu_full_gauss2d_11x11_G2.txt; a constant copied from that capture, not measured in this run):
                # the FP32 pipe is busy 60 % (lm()) / 62 % (pre-fit) of the kernels' active cycles.  A packed FFMA2 holds the pipe
                # for two cycles (profiles/r2_pipes_microbench.txt), and executed flops are ~3x the algorithmic count
                "fma_pipe_cycles_active_ncu": {"lm_kernel": 0.603, "prefit_kernel": 0.618, "source": "profiles/r2_final_ncu_full_gauss2d_11x11_G2.txt"},
                "hbm": {"achieved": BYTES_PER_FIT * n / k_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": BYTES_PER_FIT * n / k_s / 1e9 / hbm_peak, "bytes_per_fit": BYTES_PER_FIT,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback"},
            },
            "clocks": sampler.summary(),
            "configs2": c2_block,
        }
        if not args.no_cpu and world == 1:
            cores = os.cpu_count()
            kind = cpu_kind()
            n_cpu = max(256, (500 if kind == "reference" else 80) * cores)  # ~10-15 core-seconds of the reference
            rate, popt_ref, info_ref, nfev_ref = cpu_reference(sets[0]["data"][:n_cpu], sets[0]["p0"][:n_cpu], cores)
            r0 = step(0)
            torch.cuda.synchronize()
            from oracle.compare import parity_report

            rep = parity_report("gauss2d", dict(popt=r0.popt[:n_cpu].cpu().numpy(), info=r0.info[:n_cpu].cpu().numpy(),
                                                nfev=r0.nfev[:n_cpu].cpu().numpy()), dict(popt=popt_ref, info=info_ref, nfev=nfev_ref))
            what = ("the reference's own dphtools.utils.lm.curve_fit(method='mle') per ROI (oracle/_ref)" if kind == "reference"
                    else "float64 oracle port of curve_fit(method='mle') (oracle/_ref not built)")
            # the same sample with the borderline xtol decisions re-made in float64 (DPHLM_FLAG_EXACT_TIES): the reference's flags
            rx = dph.curve_fit_batch(models.gauss2d, d_data[0][:n_cpu], d_p0[0][:n_cpu], "mle", exact_ties=True)
            torch.cuda.synchronize()
            rep_x = parity_report("gauss2d", dict(popt=rx.popt.cpu().numpy(), info=rx.info.cpu().numpy(), nfev=rx.nfev.cpu().numpy()),
                                  dict(popt=popt_ref, info=info_ref, nfev=nfev_ref))
            out["cpu_baseline"] = {"value": rate, "unit": "fits/s", "cores": cores, "kind": kind,
                                   "sample": "first %d ROIs of input set 0, %s, one process per core" % (n_cpu, what),
                                   "parity_on_sample": rep, "parity_on_sample_exact_ties": rep_x}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
