"""Seeded synthetic workloads for the BASELINE.json configurations.

Every generator returns ``dict(data=[N,H,W] or [N,M] float32 Poisson counts, p0=[N,P] float32,
truth=[N,P] float64, xaxis=None or [M] float32)``.  ``p0`` is computed from the float32 data with
cheap per-ROI statistics (the recipe a caller would use), so that the GPU path and the float64
oracle start from bit-identical initial guesses.
"""

import numpy as np

from . import models


def _poisson_rois(rng, model, xy, truth, shape):
    n = truth.shape[0]
    data = np.empty((n,) + shape, dtype=np.float32)
    step = 65536
    for lo in range(0, n, step):
        t = truth[lo : lo + step]
        mu = model(xy[:, None, :], *[t[:, k : k + 1] for k in range(t.shape[1])])
        data[lo : lo + step] = rng.poisson(mu).reshape((-1,) + shape)
    return data


def gauss2d_p0(data, sigma0=1.3):
    """``(max-min, cx, cy, sigma0, max(min, 0.5))`` per ROI."""
    n, h, w = data.shape
    flat = data.reshape(n, -1)
    mx, mn = flat.max(1), flat.min(1)
    p0 = np.empty((n, 5), dtype=np.float32)
    p0[:, 0] = mx - mn
    p0[:, 1] = (w - 1) / 2
    p0[:, 2] = (h - 1) / 2
    p0[:, 3] = sigma0
    p0[:, 4] = np.maximum(mn, 0.5)
    return p0


def gauss2d_rois(n, size=11, seed=0, photons=(100.0, 5000.0), sigma=(1.0, 1.6), jitter=1.0, bg=(1.0, 10.0)):
    """Symmetric Gaussian spots: C1 (size 15), C2 (size 11), C3 (size 7, jitter 0.75)."""
    rng = np.random.default_rng(seed)
    c = (size - 1) / 2
    nph = rng.uniform(*photons, n)
    sig = rng.uniform(*sigma, n)
    truth = np.stack(
        [
            nph / (2 * np.pi * sig**2),
            c + rng.uniform(-jitter, jitter, n),
            c + rng.uniform(-jitter, jitter, n),
            sig,
            rng.uniform(*bg, n),
        ],
        axis=1,
    )
    xy = models.roi_grid(size, size)
    data = _poisson_rois(rng, models.gauss2d, xy, truth, (size, size))
    return dict(data=data, p0=gauss2d_p0(data), truth=truth, xaxis=None)


def gauss2d_rois_device(n, size, seed, device, photons=(100.0, 5000.0), sigma=(1.0, 1.6), jitter=1.0, bg=(1.0, 10.0), chunk=1 << 20):
    """:func:`gauss2d_rois` generated on a CUDA device with torch (``torch.Generator(device).manual_seed`` +
    ``torch.poisson`` for the large on-device batches, e.g. the 1e7 7x7 ROIs of BASELINE configs[2]).  Returns device tensors
    ``(data [n, size, size] float32, p0 [n, 5] float32)``; same distributions and the same p0 recipe as the host generator."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    c = (size - 1) / 2
    data = torch.empty((n, size, size), dtype=torch.float32, device=device)
    p0 = torch.empty((n, 5), dtype=torch.float32, device=device)
    yy, xx = torch.meshgrid(torch.arange(size, device=device, dtype=torch.float32), torch.arange(size, device=device, dtype=torch.float32), indexing="ij")
    xx, yy = xx.reshape(1, -1), yy.reshape(1, -1)

    def uni(lo, hi, m):
        return lo + (hi - lo) * torch.rand((m, 1), device=device, generator=g, dtype=torch.float32)

    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        nph, sig = uni(*photons, m), uni(*sigma, m)
        x0, y0, off = c + uni(-jitter, jitter, m), c + uni(-jitter, jitter, m), uni(*bg, m)
        mu = nph / (2 * np.pi * sig**2) * torch.exp(-((xx - x0) ** 2 + (yy - y0) ** 2) / (2 * sig**2)) + off
        d = torch.poisson(mu, generator=g)
        data[lo : lo + m] = d.reshape(m, size, size)
        mx, mn = d.max(1).values, d.min(1).values
        p0[lo : lo + m, 0] = mx - mn
        p0[lo : lo + m, 1] = c
        p0[lo : lo + m, 2] = c
        p0[lo : lo + m, 3] = 1.3
        p0[lo : lo + m, 4] = torch.clamp(mn, min=0.5)
    return data, p0


def gauss2d_integrated_rois(n, size=11, seed=0, photons=(100.0, 5000.0), sigma=(1.0, 1.6), jitter=1.0, bg=(1.0, 10.0)):
    """Pixel-integrated Gaussian spots; ``p0 = (sum(data - min), centre, centre, 1.3, max(min, 0.5))``."""
    rng = np.random.default_rng(seed)
    c = (size - 1) / 2
    truth = np.stack(
        [rng.uniform(*photons, n), c + rng.uniform(-jitter, jitter, n), c + rng.uniform(-jitter, jitter, n),
         rng.uniform(*sigma, n), rng.uniform(*bg, n)], axis=1)
    xy = models.roi_grid(size, size)
    data = _poisson_rois(rng, models.gauss2d_integrated, xy, truth, (size, size))
    p0 = gauss2d_p0(data)
    flat = data.reshape(n, -1)
    p0[:, 0] = np.maximum((flat - flat.min(1, keepdims=True)).sum(1), 1.0)
    return dict(data=data, p0=p0, truth=truth, xaxis=None)


def gauss2d_integrated_fixed_rois(n, size=11, seed=0, sigma=1.3):
    """Pixel-integrated spots of one known width, fitted with the width held there: ``p0 = (photons, x0, y0, offset)``."""
    sigma = float(np.float32(sigma))  # the constant reaches the kernels as a float32
    w = gauss2d_integrated_rois(n, size, seed, sigma=(sigma, sigma))
    keep = [0, 1, 2, 4]
    return dict(data=w["data"], p0=np.ascontiguousarray(w["p0"][:, keep]), truth=w["truth"][:, keep], xaxis=None, consts=(sigma,))


def gauss2d_fixed_rois(n, size=11, seed=0, sigma=1.3):
    """Sampled Gaussian spots of one known width, fitted with the width held there: ``p0 = (amp, x0, y0, offset)``."""
    sigma = float(np.float32(sigma))
    w = gauss2d_rois(n, size, seed, sigma=(sigma, sigma))
    keep = [0, 1, 2, 4]
    return dict(data=w["data"], p0=np.ascontiguousarray(w["p0"][:, keep]), truth=w["truth"][:, keep], xaxis=None, consts=(sigma,))


def gauss2d_elliptical_p0(data):
    n, h, w = data.shape
    flat = data.reshape(n, -1)
    mx, mn = flat.max(1), flat.min(1)
    p0 = np.empty((n, 7), dtype=np.float32)
    p0[:, 0] = mx - mn
    p0[:, 1] = (w - 1) / 2
    p0[:, 2] = (h - 1) / 2
    p0[:, 3] = 1.6
    p0[:, 4] = 2.08
    p0[:, 5] = 0.0
    p0[:, 6] = np.maximum(mn, 0.5)
    return p0


def gauss2d_elliptical_rois(n, size=15, seed=0):
    """Rotated elliptical spots (C4)."""
    rng = np.random.default_rng(seed)
    c = (size - 1) / 2
    truth = np.stack(
        [
            rng.uniform(20, 300, n),
            c + rng.uniform(-1, 1, n),
            c + rng.uniform(-1, 1, n),
            rng.uniform(1.2, 2.0, n),
            1.3 * rng.uniform(1.2, 2.0, n),
            rng.uniform(-0.7, 0.7, n),
            rng.uniform(1, 10, n),
        ],
        axis=1,
    )
    xy = models.roi_grid(size, size)
    data = _poisson_rois(rng, models.gauss2d_elliptical, xy, truth, (size, size))
    return dict(data=data, p0=gauss2d_elliptical_p0(data), truth=truth, xaxis=None)


def multi_exp_p0(data):
    """Bi-exponential + offset guess: tail mean as offset, 65/35 amplitude split, fixed rates."""
    n, m = data.shape
    off = data[:, -16:].mean(1)
    amp = np.maximum(data[:, 0] - off, 1.0)
    p0 = np.empty((n, 5), dtype=np.float32)
    p0[:, 0] = 0.65 * amp
    p0[:, 1] = 2.2
    p0[:, 2] = 0.35 * amp
    p0[:, 3] = 0.4
    p0[:, 4] = np.maximum(off, 0.5)
    return p0


def multi_exp_hists(n, bins=256, seed=0):
    """Bi-exponential lifetime histograms (C5), ``x = 0.05 * arange(bins)``."""
    rng = np.random.default_rng(seed)
    x = (0.05 * np.arange(bins)).astype(np.float32)
    truth = np.stack(
        [
            rng.uniform(200, 2000, n),
            rng.uniform(1.5, 3.0, n),
            rng.uniform(100, 1000, n),
            rng.uniform(0.2, 0.6, n),
            rng.uniform(1, 10, n),
        ],
        axis=1,
    )
    xd = x.astype(np.float64)
    data = np.empty((n, bins), dtype=np.float32)
    step = 65536
    for lo in range(0, n, step):
        t = truth[lo : lo + step]
        mu = t[:, 0:1] * np.exp(-t[:, 1:2] * xd) + t[:, 2:3] * np.exp(-t[:, 3:4] * xd) + t[:, 4:5]
        data[lo : lo + step] = rng.poisson(mu)
    return dict(data=data, p0=multi_exp_p0(data), truth=truth, xaxis=x)


def default_irf(taps=16, sigma_bins=1.7):
    """Gaussian instrument response, ``taps`` bins wide, unit sum, float32."""
    q = np.arange(taps)
    g = np.exp(-0.5 * ((q - (taps - 1) / 2) / sigma_bins) ** 2)
    return (g / g.sum()).astype(np.float32)


def multi_exp_irf_hists(n, bins=256, seed=0, dt=0.05, m0=8):
    """Bi-exponential lifetime histograms seen through an instrument response (:func:`default_irf` starting at bin ``m0``);
    the bins before ``m0`` hold only the offset.  Guess: offset from those bins, 65/35 split of the peak height, fixed rates."""
    rng = np.random.default_rng(seed)
    irf = default_irf()
    model = models.multi_exp_convolved(irf, dt, m0)
    x = (dt * np.arange(bins)).astype(np.float32)
    truth = np.stack(
        [rng.uniform(200, 2000, n), rng.uniform(1.5, 3.0, n), rng.uniform(100, 1000, n), rng.uniform(0.2, 0.6, n), rng.uniform(1, 10, n)],
        axis=1,
    )
    lag = (np.arange(bins)[:, None] - m0 - np.arange(irf.size)[None, :]) * float(np.float32(dt))
    ok = lag >= 0
    lag0 = np.where(ok, lag, 0.0)
    w = np.where(ok, irf.astype(float)[None, :], 0.0)
    data = np.empty((n, bins), dtype=np.float32)
    step = 8192
    for lo in range(0, n, step):
        t = truth[lo : lo + step]
        c0 = (w[None] * np.exp(-t[:, 1, None, None] * lag0[None])).sum(2)
        c1 = (w[None] * np.exp(-t[:, 3, None, None] * lag0[None])).sum(2)
        data[lo : lo + step] = rng.poisson(t[:, 0:1] * c0 + t[:, 2:3] * c1 + t[:, 4:5])
    off = data[:, : max(m0 - 2, 1)].mean(1)
    amp = np.maximum(data.max(1) - off, 1.0)
    p0 = np.empty((n, 5), dtype=np.float32)
    p0[:, 0] = 0.65 * amp
    p0[:, 1] = 2.2
    p0[:, 2] = 0.35 * amp
    p0[:, 3] = 0.4
    p0[:, 4] = np.maximum(off, 0.5)
    return dict(data=data, p0=p0, truth=truth, xaxis=x, consts=model.consts)


WORKLOADS = {
    "c1_gauss2d_15": lambda n, seed=0: gauss2d_rois(n, 15, seed, photons=(1500.0, 1500.0), sigma=(1.38, 1.38)),
    "c2_gauss2d_11": lambda n, seed=0: gauss2d_rois(n, 11, seed),
    "c2_gauss2d_12": lambda n, seed=0: gauss2d_rois(n, 12, seed),  # 576-byte ROIs: 16-byte aligned in the batch (128-bit staging)
    "c3_gauss2d_7": lambda n, seed=0: gauss2d_rois(n, 7, seed, jitter=0.75),
    "c4_ellip_15": lambda n, seed=0: gauss2d_elliptical_rois(n, 15, seed),
    "c5_multiexp_256": lambda n, seed=0: multi_exp_hists(n, 256, seed),
    "c6_gauss2d_int_11": lambda n, seed=0: gauss2d_integrated_rois(n, 11, seed),
    "c7_gauss2d_intf_11": lambda n, seed=0: gauss2d_integrated_fixed_rois(n, 11, seed),
    "c8_multiexp_irf_256": lambda n, seed=0: multi_exp_irf_hists(n, 256, seed),
    "c9_gauss2d_fixed_11": lambda n, seed=0: gauss2d_fixed_rois(n, 11, seed),
}
