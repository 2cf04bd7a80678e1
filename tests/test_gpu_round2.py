"""Round-2 surface on the GPU: pipelined calls, lm() alone, the reference's pcov, unfinished-fit sentinel and watchdog
status, automatic group size of the IRF model, the in-process multi-device launcher.  Run with ``-m gpu``."""

import os

import numpy as np
import pytest

import dphtools_b200 as dph
from dphtools_b200 import _lib, models
from oracle.compare import parity_report
from oracle.golden import GOLDEN_DIR

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_pipelined_calls_equal_stream_ordered_calls(golden):
    """DPHLM_FLAG_PIPELINED changes where the kernels run (the library's own streams, three calls in flight), not what they
    compute: bit-identical results once join() has been called."""
    g = golden("c2_mle")
    data, p0 = torch.from_numpy(g["data"]).cuda(), torch.from_numpy(g["p0"]).cuda()
    chunks = [(data[k:k + 2500], p0[k:k + 2500]) for k in range(0, 10000, 2500)]
    plain = [dph.curve_fit_batch(models.gauss2d, d, p, "mle") for d, p in chunks]
    torch.cuda.synchronize()
    for _ in range(3):
        piped = [dph.curve_fit_batch(models.gauss2d, d, p, "mle", pipelined=True) for d, p in chunks]
        dph.join()
        total = sum(r.popt.sum() for r in piped)  # consumer on the current stream, ordered behind the join
        torch.cuda.synchronize()
        assert torch.isfinite(total)
        for a, b in zip(plain, piped):
            assert torch.equal(a.popt, b.popt) and torch.equal(a.info, b.info) and torch.equal(a.nfev, b.nfev)
            assert torch.equal(a.chisq, b.chisq)
    dph.join()  # nothing pending: a no-op
    with pytest.raises(ValueError):
        dph.curve_fit_batch(models.gauss2d, g["data"][:8], g["p0"][:8], "mle", pipelined=True)  # host arrays


def test_lm_alone_matches_the_reference_lm():
    """dphtools.utils.lm.lm() called directly from the raw guesses (tests/golden/x_lm.npz, minted by
    oracle/make_golden_extra.py from the live reference) against DPHLM_FLAG_SKIP_PREFIT."""
    z = np.load(os.path.join(GOLDEN_DIR, "x_lm.npz"))
    for key, model, method in (("c2_mle", models.gauss2d, "mle"), ("c2_ls", models.gauss2d, "ls"), ("c5_mle", models.multi_exp, "mle")):
        xa = z[key + ".xaxis"] if z[key + ".xaxis"].size else None
        data = torch.from_numpy(z[key + ".data"].astype(np.float32)).cuda()
        r = dph.curve_fit_batch(model, data, torch.from_numpy(z[key + ".p0"]).cuda(), method, xa, skip_prefit=True)
        out = dict(popt=r.popt.cpu().numpy(), info=r.info.cpu().numpy(), nfev=r.nfev.cpu().numpy())
        rep = parity_report(model.model_name, out, dict(popt=z[key + ".popt"], info=z[key + ".info"], nfev=z[key + ".nfev"]))
        print(key, rep)
        # from raw guesses the trajectories are long (up to 117 trips on the bi-exponential set): one fit in 200 ends 1.6e-4 away
        assert rep["converged_equal"] == 1.0 and rep["within_rtol"] >= (0.99 if key == "c5_mle" else 1.0) and rep["info_equal"] >= 0.985, rep
        # host entry, same flag
        h = dph.curve_fit_batch(model, z[key + ".data"].astype(np.float32), z[key + ".p0"], method, xa, skip_prefit=True)
        np.testing.assert_array_equal(h.popt, out["popt"])
    # the reference's call surface: lm(func, x0, args, Dfun, full_output, ..., method)
    i = 7
    xy = models.roi_grid(11, 11)
    y = z["c2_mle.data"][i].ravel().astype(float)
    popt, cov_x, infodict, errmsg, ier = dph.lm(models.gauss2d, z["c2_mle.p0"][i], args=(xy, y), Dfun=models.gauss2d.jac,
                                                full_output=True, method="mle")
    assert cov_x is None and ier == z["c2_mle.info"][i] and abs(infodict["nfev"] - z["c2_mle.nfev"][i]) <= 1
    np.testing.assert_allclose(popt[:4], z["c2_mle.popt"][i][:4], rtol=1e-4)
    assert infodict["fvec"].shape == (121,) and infodict["fjac"].shape == (121, 5) and "reductions" in errmsg
    assert len(dph.lm(models.gauss2d, z["c2_mle.p0"][i], args=(xy, y), Dfun=models.gauss2d.jac, method="mle")) == 2
    with pytest.raises(NotImplementedError):
        dph.lm(models.gauss2d, z["c2_mle.p0"][i], args=(xy, y), method="mle")
    with pytest.raises(NotImplementedError, match="Column derivatives"):
        dph.lm(models.gauss2d, z["c2_mle.p0"][i], args=(xy, y), Dfun=models.gauss2d.jac, col_deriv=False)
    with pytest.raises(TypeError, match="not recognized"):
        dph.lm(models.gauss2d, z["c2_mle.p0"][i], args=(xy, y), Dfun=models.gauss2d.jac, method="nope")


@pytest.mark.parametrize("name", ["c2_mle", "c4_mle", "c5_mle"])
def test_batched_pcov_is_the_reference_pcov(golden, name):
    """curve_fit_batch(return_pcov='jtj') against the pcov the live reference returned (tests/golden/x_pcov.npz):
    pinv(J^T J) at the last ACCEPTED point of lm() (lm.py:468-473, 489, 672-677), not at popt."""
    ref = np.load(os.path.join(GOLDEN_DIR, "x_pcov.npz"))[name]
    g = golden(name)
    n = ref.shape[0]
    model = models.resolve(g["model"], g["consts"])
    r = dph.curve_fit_batch(model, torch.from_numpy(g["data"][:n]).cuda(), torch.from_numpy(g["p0"][:n]).cuda(), g["method"],
                            g["xaxis"], return_pcov="jtj")
    got = r.pcov.cpu().numpy().astype(float)
    assert r.x_acc is not None
    scale = np.sqrt(np.einsum("ni,nj->nij", np.diagonal(ref, axis1=1, axis2=2), np.diagonal(ref, axis1=1, axis2=2)))
    err = np.abs((got - ref) / scale).max(axis=(1, 2))
    assert np.quantile(err, 0.99) < 2e-3 and np.median(err) < 2e-4, (np.quantile(err, 0.99), np.median(err))
    h = dph.curve_fit_batch(model, g["data"][:64], g["p0"][:64], g["method"], g["xaxis"], return_pcov="jtj")  # host entry
    np.testing.assert_allclose(h.pcov, got[:64], rtol=1e-5, atol=1e-12)


def test_rank_deficient_jacobian_gives_the_pseudo_inverse():
    """A zero-amplitude point: the centre and width columns of J vanish.  The reference's pinv drops those singular values
    (lm.py:675-677) and returns zeros there; so does dphlm_covariance (it used to return NaN)."""
    p = np.array([[0.0, 5.0, 5.0, 1.3, 4.0]], np.float32)
    got = dph.covariance(models.gauss2d, p, (11, 11), "jtj")[0].astype(float)
    J = models.gauss2d.jac(models.roi_grid(11, 11), *p[0].astype(float))
    _, s, vt = np.linalg.svd(J, full_matrices=False)
    keep = s > np.finfo(float).eps * 121 * s[0]
    ref = (vt[keep].T / s[keep] ** 2) @ vt[keep]
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-9)


def test_every_call_marks_unfinished_fits_and_the_watchdog_is_clean(golden):
    g = golden("c3_mle")
    r = dph.curve_fit_batch(models.gauss2d, torch.from_numpy(g["data"]).cuda(), torch.from_numpy(g["p0"]).cuda(), "mle")
    assert not dph.call_status()                       # synchronises; no straggler kernel gave up
    assert int((r.info == _lib.INFO_UNFINISHED).sum()) == 0
    assert _lib.INFO_UNFINISHED == -2139062144


def test_irf_model_picks_a_wide_group_for_short_histograms():
    """group_lanes = 0 with a 64-bin IRF histogram used to fail (the model is built for 16 and 32 lanes only)."""
    rng = np.random.default_rng(3)
    irf = np.array([0.2, 0.5, 0.3], np.float32)
    model = models.multi_exp_convolved(irf, 0.1, 2)
    x = 0.1 * np.arange(64)
    truth = (800.0, 1.5, 5.0)
    data = rng.poisson(model(x, *truth), size=(500, 64)).astype(np.float32)
    p0 = np.tile(np.array([600.0, 1.0, 3.0], np.float32), (500, 1))
    r = dph.curve_fit_batch(model, torch.from_numpy(data).cuda(), torch.from_numpy(p0).cuda(), "mle", x)
    info, popt = r.info.cpu().numpy(), r.popt.cpu().numpy()
    assert ((info >= 1) & (info <= 4)).mean() > 0.99
    assert abs(np.median(popt[:, 1]) / truth[1] - 1) < 0.02


def test_even_roi_with_refine_is_rejected():
    frames = np.zeros((1, 32, 32), np.float32)
    peaks = np.array([[0, 16, 16]], np.int32)
    with pytest.raises(ValueError, match="odd roi"):
        dph.extract_rois(frames, peaks, roi=10, refine=True)
    rois, p0, valid = dph.extract_rois(frames, peaks, roi=10, refine=False)
    assert rois.shape == (1, 10, 10) and bool(valid[0])
    signed = -5 * np.ones((1, 32, 32), np.int16)  # background-subtracted frames keep their sign
    rois, _, _ = dph.extract_rois(signed, peaks, roi=11)
    assert float(rois.min()) == -5.0


def test_in_process_multi_device_launcher_equals_one_device(golden):
    """curve_fit_batch(device=[...]): contiguous shards, one host thread per GPU, no collective.  With one
    GPU the list holds the same device twice (two shards through the same code path)."""
    g = golden("c3_mle")
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    one = dph.curve_fit_batch(models.gauss2d, g["data"], g["p0"], "mle", device=0)
    many = dph.curve_fit_batch(models.gauss2d, g["data"], g["p0"], "mle", device=devices)
    for k in ("popt", "info", "nfev", "chisq"):
        np.testing.assert_array_equal(getattr(one, k), getattr(many, k))


def test_async_host_calls_equal_blocking_calls(golden):
    """dphlm_fit_batch_host_async + dphlm_host_wait: several batches enqueued back to back from page-locked buffers (the upload of
    one overlaps the kernels of the one before), one wait, same results as one blocking call per batch."""
    g = golden("c2_mle")
    n = 4000
    ins, outs = [], []
    for k in range(3):
        hd, hp = dph.pinned_empty((n, 11, 11)), dph.pinned_empty((n, 5))
        hd[...] = g["data"][k * 2000:k * 2000 + n]
        hp[...] = g["p0"][k * 2000:k * 2000 + n]
        ins.append((hd, hp))
        outs.append(dph.BatchResult(dph.pinned_empty((n, 5)), dph.pinned_empty(n, np.int32), dph.pinned_empty(n, np.int32),
                                    dph.pinned_empty(n), None, None))
    for (hd, hp), o in zip(ins, outs):
        dph.curve_fit_batch(models.gauss2d, hd, hp, "mle", device=0, out=o, wait=False)
    dph.host_wait(0)
    for (hd, hp), o in zip(ins, outs):
        ref = dph.curve_fit_batch(models.gauss2d, hd, hp, "mle", device=0)
        for k in ("popt", "info", "nfev", "chisq"):
            np.testing.assert_array_equal(getattr(o, k), getattr(ref, k))
    dph.host_wait(0)  # nothing pending: returns at once


def test_device_guess_is_the_reference_recipe_and_the_batch_driver_runs_on_the_gpu(golden):
    """dphlm_guess_multi_exp against the host restatement of multi_exp_fit's guess (fitfuncs.py:63-80, 131-152; identical to the
    live reference, tests/test_host_api.py), then multi_exp_fit_batch end to end, including histograms with non-finite bins,
    which the reference masks per histogram (fitfuncs.py:124-126)."""
    from dphtools_b200 import fitfuncs

    g = golden("c5_mle")
    t = g["xaxis"].astype(float)[1:]  # t > 0 for the log-spaced split
    data = g["data"][:300, 1:]
    for comps, off in ((2, True), (1, True), (2, False), (3, True)):
        want = np.stack([fitfuncs.guess_multi_exp(row.astype(float), t, comps, off) for row in data])
        got = fitfuncs.guess_multi_exp_batch(data, t, comps, off).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-5)
    dirty = data.copy()
    dirty[5, 40] = np.nan
    dirty[17, 3] = np.inf
    res, p0 = fitfuncs.multi_exp_fit_batch(dirty, t, 2, method="mle")
    clean, _ = fitfuncs.multi_exp_fit_batch(torch.from_numpy(data).cuda(), t, 2, method="mle")
    assert isinstance(res.popt, np.ndarray) and clean.popt.is_cuda
    keep = np.ones(300, bool)
    keep[[5, 17]] = False
    np.testing.assert_array_equal(res.popt[keep], clean.popt.cpu().numpy()[keep])
    assert ((res.info >= 1) & (res.info <= 4)).mean() > 0.99
    for i in (5, 17):  # fitted alone on their finite bins, like the reference does
        want, _ = fitfuncs.multi_exp_fit(dirty[i].astype(float), t, components=2, method="mle")
        np.testing.assert_allclose(res.popt[i], want, rtol=1e-5)
        np.testing.assert_allclose(res.popt[i], clean.popt[i].cpu().numpy(), rtol=0.2)  # one bin fewer: nearly the same fit


@pytest.mark.parametrize("name", ["c2_mle", "c3_mle", "c5_mle", "c6_mle", "c7_mle", "c9_mle", "c4_mle", "c8_mle"])
def test_exact_ties_give_the_reference_flag_on_gpu(golden, name):
    """DPHLM_FLAG_EXACT_TIES: identical info on >= 99.9 % of every golden set (north_star "identical convergence flags",
    SURVEY 7.3), device and host entry."""
    g = golden(name)
    model = models.resolve(g["model"], g["consts"])
    r = dph.curve_fit_batch(model, torch.from_numpy(g["data"]).cuda(), torch.from_numpy(g["p0"]).cuda(), g["method"], g["xaxis"],
                            exact_ties=True)
    out = dict(popt=r.popt.cpu().numpy(), info=r.info.cpu().numpy(), nfev=r.nfev.cpu().numpy())
    rep = parity_report(g["model"], out, g)
    print(name, rep)
    # (7x7: up to two degenerate fits of the 10 000 -- sigma = 0.27 px and sigma > ROI -- end beyond 1e-4)
    assert rep["converged_equal"] == 1.0 and rep["info_equal"] >= 0.999 and rep["within_rtol"] >= 0.9998, rep
    h = dph.curve_fit_batch(model, g["data"][:512], g["p0"][:512], g["method"], g["xaxis"], exact_ties=True)
    np.testing.assert_array_equal(h.info, out["info"][:512])


def test_128_bit_staging_changes_nothing(monkeypatch):
    """ROIs of a multiple of 16 bytes (12x12; 256-bin histograms) are staged with 16-byte cp.async (lm_launch.cuh: vec16); the
    4-byte path (DPHLM_NO_VEC16) must give the same bits.  A batch pointer that is not 16-byte aligned falls back by itself."""
    from dphtools_b200 import synth

    for wl, model, lanes in (("c2_gauss2d_12", models.gauss2d, 4), ("c5_multiexp_256", models.multi_exp, 16)):
        w = synth.WORKLOADS[wl](3000, seed=5)
        data, p0 = torch.from_numpy(w["data"]).cuda(), torch.from_numpy(w["p0"]).cuda()
        outs = []
        for off in (False, True):
            if off:
                monkeypatch.setenv("DPHLM_NO_VEC16", "1")
            else:
                monkeypatch.delenv("DPHLM_NO_VEC16", raising=False)
            r = dph.curve_fit_batch(model, data, p0, "mle", w["xaxis"], group_lanes=lanes)
            torch.cuda.synchronize()
            outs.append((r.popt.cpu().numpy(), r.info.cpu().numpy(), r.nfev.cpu().numpy()))
        monkeypatch.delenv("DPHLM_NO_VEC16", raising=False)
        for a, b in zip(outs[0], outs[1]):
            np.testing.assert_array_equal(a, b)
        assert ((outs[0][1] >= 1) & (outs[0][1] <= 4)).mean() > 0.99
        # a view that starts one ROI into an allocation of 4-byte alignment only: offset by one float
        flat = torch.empty(data.numel() + 1, device="cuda", dtype=torch.float32)
        shifted = flat[1:].view_as(data)
        shifted.copy_(data)
        r = dph.curve_fit_batch(model, shifted, p0, "mle", w["xaxis"], group_lanes=lanes)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(r.popt.cpu().numpy(), outs[0][0])


def test_device_generator_follows_the_host_recipe():
    """synth.gauss2d_rois_device (the 1e7-ROI generator of bench.py's configs[2] block): seeded, Poisson counts, and the same p0
    recipe as the host generator -- the batch it makes converges like the host-generated one."""
    from dphtools_b200 import synth

    dev = torch.device("cuda", 0)
    a, pa = synth.gauss2d_rois_device(5000, 7, 11, dev, jitter=0.75, chunk=2048)
    b, pb = synth.gauss2d_rois_device(5000, 7, 11, dev, jitter=0.75, chunk=2048)
    assert a.shape == (5000, 7, 7) and pa.shape == (5000, 5) and a.dtype == torch.float32
    assert torch.equal(a, b) and torch.equal(pa, pb)
    assert torch.equal(a, a.round()) and float(a.min()) >= 0.0
    np.testing.assert_array_equal(pa.cpu().numpy(), synth.gauss2d_p0(a.cpu().numpy()))
    r = dph.curve_fit_batch(models.gauss2d, a, pa, "mle")
    torch.cuda.synchronize()
    assert float(((r.info >= 1) & (r.info <= 4)).double().mean()) > 0.999
