"""Odd shapes through the CUDA path: rectangular ROIs (height != width), odd pixel counts (the scalar tail of the pixel-pair
loop), pixel counts that do not divide by the group size, batches smaller and slightly larger than one wave, every 2-D model.
Checked against the host build of the same fp32 core on every fit (indexing and tail bugs) and against the float64 oracle
(oracle/lm_oracle.py, the restatement of lm.py pinned to the live reference) on the first fits of every case."""

import numpy as np
import pytest

import dphtools_b200 as dph
from dphtools_b200 import models
from oracle import host_emu, lm_oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _spots(model, h, w, n, seed, sigma_fixed=None):
    rng = np.random.default_rng(seed)
    xy = models.roi_grid(h, w)
    x0 = (w - 1) / 2 + rng.uniform(-0.8, 0.8, n)
    y0 = (h - 1) / 2 + rng.uniform(-0.8, 0.8, n)
    sig = rng.uniform(0.9, 1.4, n) if sigma_fixed is None else np.full(n, sigma_fixed)
    bg = rng.uniform(1, 10, n)
    photons = rng.uniform(300, 4000, n)
    data = np.empty((n, h, w), np.float32)
    for i in range(n):
        if model.model_name in ("gauss2d", "gauss2d_fixed"):
            mu = models.gauss2d(xy, photons[i] / (2 * np.pi * sig[i] ** 2), x0[i], y0[i], sig[i], bg[i])
        else:
            mu = models.gauss2d_integrated(xy, photons[i], x0[i], y0[i], sig[i], bg[i])
        data[i] = rng.poisson(mu).reshape(h, w)
    flat = data.reshape(n, -1)
    mn, mx = flat.min(1), flat.max(1)
    amp = mx - mn if model.model_name in ("gauss2d", "gauss2d_fixed") else np.maximum((flat - mn[:, None]).sum(1), 1.0)
    cols = [amp, np.full(n, (w - 1) / 2), np.full(n, (h - 1) / 2)]
    if model.n_param == 5:
        cols.append(np.full(n, 1.2))
    cols.append(np.maximum(mn, 0.5))
    return data, np.stack(cols, 1).astype(np.float32)


CASES = [
    # model factory, h, w, n, group_lanes
    (lambda: models.gauss2d, 9, 13, 300, 0),
    (lambda: models.gauss2d, 13, 9, 300, 4),
    (lambda: models.gauss2d, 5, 5, 1000, 2),      # 25 pixels: 12 pairs + 1
    (lambda: models.gauss2d, 7, 9, 257, 8),       # 63 pixels: odd, pairs not divisible by the group
    (lambda: models.gauss2d, 17, 17, 129, 16),
    (lambda: models.gauss2d, 6, 4, 33, 32),
    (lambda: models.gauss2d_integrated, 9, 13, 300, 0),
    (lambda: models.gauss2d_integrated, 13, 9, 300, 8),
    (lambda: models.gauss2d_integrated, 7, 7, 1000, 2),
    (lambda: models.gauss2d_integrated_fixed(1.15), 8, 11, 300, 4),
    (lambda: models.gauss2d_integrated_fixed(1.15), 11, 8, 14300, 0),   # a little more than one wave of 4-lane groups
    (lambda: models.gauss2d_fixed(1.15), 10, 7, 300, 0),
    (lambda: models.gauss2d, 25, 25, 200, 0),
    (lambda: models.gauss2d_integrated, 31, 29, 100, 0),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_odd_shapes_match_the_host_build(case):
    make, h, w, n, lanes = CASES[case]
    model = make()
    fixed = model.consts[0] if getattr(model, "consts", None) else None
    data, p0 = _spots(model, h, w, n, seed=100 + case, sigma_fixed=fixed)
    ref = host_emu.fit_batch(model.model_id, "mle", data, p0, None, consts=getattr(model, "consts", None))
    for host in (False, True):
        if host:
            r = dph.curve_fit_batch(model, data, p0, "mle", group_lanes=lanes)
            got = dict(popt=r.popt, info=r.info)
        else:
            r = dph.curve_fit_batch(model, torch.from_numpy(data).cuda(), torch.from_numpy(p0).cuda(), "mle", group_lanes=lanes)
            got = dict(popt=r.popt.cpu().numpy(), info=r.info.cpu().numpy())
        conv_g, conv_r = (got["info"] >= 1) & (got["info"] <= 4), (ref["info"] >= 1) & (ref["info"] <= 4)
        assert (conv_g == conv_r).mean() >= 0.995, (case, host)
        both = conv_g & conv_r
        assert both.mean() > 0.95
        # same fp32 algorithm, different summation order: the rare fit that lands on another side of a stop test aside,
        # the parameters agree to rounding
        rel = np.abs(got["popt"][both] - ref["popt"][both]) / np.maximum(np.abs(ref["popt"][both]), 1e-3)
        assert np.quantile(rel.max(1), 0.99) < 1e-4, (case, host, np.quantile(rel.max(1), 0.99))
        # (the fixed-width model converges quadratically and ends with ftol and xtol both borderline: more 1 <-> 2 swaps)
        assert (got["info"] == ref["info"]).mean() >= (0.93 if fixed else 0.97), (case, host)
        if not host:  # the float64 oracle on the first fits: same convergence, parameters within the north-star tolerance
            k = 24
            orc = lm_oracle.fit_batch(model, data[:k], p0[:k], "mle", models.roi_grid(h, w))
            conv_o = (orc["info"] >= 1) & (orc["info"] <= 4)
            assert (conv_g[:k] == conv_o).mean() >= 0.95, (case, conv_g[:k], orc["info"])
            ok = conv_g[:k] & conv_o
            rel_o = np.abs(got["popt"][:k][ok] - orc["popt"][ok]) / np.maximum(np.abs(orc["popt"][ok]), 1e-3)
            assert (rel_o[:, :-1].max(1) < 1e-4).mean() >= 0.95, (case, rel_o.max(1))  # amplitude, centre, width


def test_long_histograms_and_wide_rois():
    """1024-bin decays (16 and 32 lanes per fit) and a 32x32 ROI are within the shared-memory layout; a 100x100 ROI is refused
    with an error instead of a bad launch."""
    rng = np.random.default_rng(5)
    x = (0.0125 * np.arange(1024)).astype(np.float32)
    truth = np.stack([rng.uniform(200, 2000, 64), rng.uniform(1.5, 3.0, 64), rng.uniform(100, 1000, 64), rng.uniform(0.2, 0.6, 64),
                      rng.uniform(1, 10, 64)], 1)
    xd = x[None, :].astype(float)
    mu = truth[:, 0:1] * np.exp(-truth[:, 1:2] * xd) + truth[:, 2:3] * np.exp(-truth[:, 3:4] * xd) + truth[:, 4:5]
    data = rng.poisson(mu).astype(np.float32)
    p0 = (truth * rng.uniform(0.8, 1.2, truth.shape)).astype(np.float32)
    ref = host_emu.fit_batch(models.multi_exp.model_id, "mle", data, p0, x)
    for lanes in (16, 32):
        r = dph.curve_fit_batch(models.multi_exp, data, p0, "mle", x, group_lanes=lanes)
        ok = (r.info >= 1) & (r.info <= 4) & (ref["info"] >= 1) & (ref["info"] <= 4)
        assert ok.mean() > 0.9
        assert np.quantile(np.abs(r.popt[ok] - ref["popt"][ok]) / np.abs(ref["popt"][ok]), 0.99) < 1e-4
    big, p0b = _spots(models.gauss2d, 32, 32, 40, seed=9)
    r = dph.curve_fit_batch(models.gauss2d, big, p0b, "mle")
    assert ((r.info >= 1) & (r.info <= 4)).mean() > 0.9
    with pytest.raises(RuntimeError, match="too large|ROI"):
        dph.curve_fit_batch(models.gauss2d, np.ones((4, 100, 100), np.float32), np.ones((4, 5), np.float32), "mle", group_lanes=4)
