"""The float64 oracle restatement against the golden vectors minted from the live reference."""

import numpy as np
import pytest

from dphtools_b200 import models
from oracle import lm_oracle

SETS = ["c1_mle", "c2_mle", "c2_ls", "c3_mle", "c4_mle", "c4_ls", "c5_mle", "c5_ls", "c6_mle", "c7_mle", "c8_mle", "c6_ls", "c7_ls", "c8_ls", "c9_mle"]


def _xdata(g):
    if g["xaxis"] is not None:
        return g["xaxis"].astype(np.float64)
    return models.roi_grid(*g["data"].shape[1:])


@pytest.mark.parametrize("name", SETS)
def test_oracle_matches_reference_goldens(golden, name):
    g = golden(name)
    n = min(250, g["data"].shape[0])
    # always include the hard cases the reference produced (non-info-1 exits, long trajectories)
    hard = np.flatnonzero((g["info"] != 1) | (g["nfev"] > 25))[:50]
    idx = np.union1d(np.arange(n), hard)
    model = models.resolve(g["model"], g["consts"])
    out = lm_oracle.fit_batch(model, g["data"][idx], g["p0"][idx], g["method"], _xdata(g))
    np.testing.assert_array_equal(out["info"], g["info"][idx])
    np.testing.assert_array_equal(out["nfev"], g["nfev"][idx])
    ok = g["info"][idx] > 0
    np.testing.assert_allclose(out["p_ls"][ok], g["p_ls"][idx][ok], rtol=1e-12, atol=0)
    np.testing.assert_allclose(out["popt"][ok], g["popt"][idx][ok], rtol=1e-10, atol=1e-12)


def test_oracle_curve_fit_surface(golden):
    g = golden("c1_mle")
    xy = _xdata(g)
    popt, pcov, infodict, info, p_ls = lm_oracle.curve_fit(
        models.gauss2d, xy, g["data"][0].ravel().astype(float), g["p0"][0].astype(float),
        models.gauss2d_jac, "mle", full_output=True, return_stage1=True,
    )
    assert info == g["info"][0] and infodict["nfev"] == g["nfev"][0]
    np.testing.assert_allclose(popt, g["popt"][0], rtol=1e-10)
    assert pcov.shape == (5, 5) and np.all(np.linalg.eigvalsh(pcov) > 0)
    with pytest.raises(TypeError):
        lm_oracle.curve_fit(models.gauss2d, xy, g["data"][0].ravel(), g["p0"][0], models.gauss2d_jac, "nope")
    with pytest.raises(NotImplementedError):
        lm_oracle.curve_fit(models.gauss2d, xy, g["data"][0].ravel(), g["p0"][0], None, "mle")


@pytest.mark.parametrize("model,p", [
    (models.gauss2d, (50.0, 5.3, 4.6, 1.4, 3.0)),
    (models.gauss2d_elliptical, (50.0, 7.3, 6.6, 1.4, 2.1, 0.3, 3.0)),
    (models.gauss2d_integrated, (1500.0, 7.3, 6.6, 1.4, 3.0)),
    (models.gauss2d_integrated_fixed(1.25), (1500.0, 7.3, 6.6, 3.0)),
    (models.gauss2d_fixed(1.25), (50.0, 7.3, 6.6, 3.0)),
    (models.multi_exp, (900.0, 2.2, 300.0, 0.4, 4.0)),
    (models.multi_exp_convolved([0.1, 0.2, 0.4, 0.2, 0.1], 0.05, 6), (900.0, 2.2, 300.0, 0.4, 4.0)),
    (models.multi_exp_convolved([0.1, 0.2, 0.4, 0.2, 0.1], 0.05, 6), (900.0, 2.2)),
    (models.multi_exp, (900.0, 2.2, 300.0, 0.4)),
])
def test_jacobians_against_central_differences(model, p):
    x = 0.05 * np.arange(256) if model.model_id in models.ONE_D else models.roi_grid(15, 15)
    p = np.array(p)
    jac = model.jac(x, *p)
    for k in range(p.size):
        h = 1e-6 * max(1.0, abs(p[k]))
        pp, pm = p.copy(), p.copy()
        pp[k] += h
        pm[k] -= h
        num = (model(x, *pp) - model(x, *pm)) / (2 * h)
        np.testing.assert_allclose(jac[:, k], num, rtol=1e-6, atol=1e-7)
