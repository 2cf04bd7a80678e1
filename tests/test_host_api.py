"""Host-side logic that needs no GPU: the C ABI exports, argument validation, the curve_fit surface."""

import ctypes
import os
import re

import numpy as np
import pytest

import dphtools_b200 as dph
from dphtools_b200 import _lib, models

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "dphlm.h")).read()
    declared = set(re.findall(r"\b(dphlm_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    lib = _lib.lib()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.dphlm_version() == 200


def test_desc_layout_and_defaults():
    d = _lib.Desc()
    _lib.lib().dphlm_desc_init(ctypes.byref(d))
    assert abs(d.ftol - 1.49012e-8) < 1e-14 and abs(d.xtol - 1.49012e-8) < 1e-14
    assert d.gtol == 0.0 and d.factor == 100.0 and d.maxfev == 0 and d.n_fits == 0 and not d.data


@pytest.mark.parametrize("mutate,msg", [
    (lambda d: setattr(d, "model", 9), "unknown model"),
    (lambda d: setattr(d, "method", 3), "Method not recognized"),
    (lambda d: setattr(d, "n_param", 4), "5 parameters"),
    (lambda d: setattr(d, "dim0", 0), "dim0"),
    (lambda d: setattr(d, "n_fits", -1), "n_fits"),
    (lambda d: setattr(d, "data", None), "null"),
])
def test_validation_errors_come_back_through_last_error(mutate, msg):
    d = _lib.Desc()
    _lib.lib().dphlm_desc_init(ctypes.byref(d))
    buf = np.zeros(16, np.float32)
    d.model, d.method, d.n_param, d.dim0, d.dim1, d.n_fits = 0, 1, 5, 11, 11, 1
    for f in ("data", "p0", "popt", "info", "nfev", "chisq"):
        setattr(d, f, buf.ctypes.data)
    mutate(d)
    assert _lib.lib().dphlm_fit_batch(ctypes.byref(d), None) == -1  # validation happens before any CUDA call
    assert msg in _lib.last_error()


def test_curve_fit_guards_follow_the_reference():
    xy = models.roi_grid(7, 7)
    y = models.gauss2d(xy, 50, 3, 3, 1.2, 2)
    p0 = (40, 3, 3, 1.0, 1)
    with pytest.raises(TypeError, match="Method nope not recognized"):
        dph.curve_fit(models.gauss2d, xy, y, p0, method="nope", jac=models.gauss2d_jac)
    with pytest.raises(NotImplementedError, match="Bounds"):
        dph.curve_fit(models.gauss2d, xy, y, p0, method="mle", jac=models.gauss2d_jac, bounds=(0, 100))
    with pytest.raises(NotImplementedError, match="Weighting"):
        dph.curve_fit(models.gauss2d, xy, y, p0, method="mle", jac=models.gauss2d_jac, sigma=np.ones_like(y))
    with pytest.raises(NotImplementedError, match="Jacobian"):
        dph.curve_fit(models.gauss2d, xy, y, p0, method="mle")
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        dph.curve_fit(lambda x, a: a * x[0], xy, y, (1.0,), method="mle", jac=lambda x, a: x[0][:, None])
    bad = y.copy()
    bad[3] = np.nan
    with pytest.raises(ValueError):
        dph.curve_fit(models.gauss2d, xy, bad, p0, method="mle", jac=models.gauss2d_jac)


def test_curve_fit_delegates_default_methods_to_scipy():
    # the reference's own hot-path tests (tests/test_fitfuncs.py:32-42): noiseless 10 exp(-3x) + 5
    x = np.linspace(0, 10)
    data = models.multi_exp(x, 10, 3, 5)
    popt, pcov = dph.curve_fit(models.multi_exp, x, data, p0=(8.0, 2.0, 4.0), jac=models.multi_exp_jac)
    np.testing.assert_allclose(popt, (10, 3, 5), rtol=1e-3)
    popt, pcov = dph.curve_fit(models.multi_exp, x, -data, p0=(-8.0, 2.0, -4.0), jac=models.multi_exp_jac)
    np.testing.assert_allclose(popt, (-10, 3, -5), rtol=1e-3)
    out = dph.curve_fit(models.multi_exp, x, data, p0=(8.0, 2.0, 4.0), method="trf", full_output=True)
    assert out[3] == "No error" and out[4] == 1


def test_product_models_agree_with_oracle_inputs():
    xy = models.roi_grid(5, 9)
    assert xy.shape == (2, 45) and xy[0, :9].tolist() == list(range(9)) and xy[1, 9] == 1
    assert models.gauss2d.model_id == 0 and models.gauss2d_elliptical.model_id == 1 and models.multi_exp.model_id == 2


def test_multi_exp_fit_driver_reproduces_the_reference_tests_and_guesses():
    """tests/test_fitfuncs.py:32-42 of the reference through our driver, and the guess recipe against the live
    reference's outputs recorded here (generated with dphtools.utils.fitfuncs._estimate_exponent_params)."""
    from dphtools_b200 import fitfuncs

    x = np.linspace(0, 10)
    data = fitfuncs.exponent(x, 10, 3, 5)
    popt, pcov = fitfuncs.exponent_fit(data, x)
    np.testing.assert_allclose(popt, (10, 3, 5), rtol=1e-3)
    popt, pcov = fitfuncs.exponent_fit(-data, x)
    np.testing.assert_allclose(popt, (-10, 3, -5), rtol=1e-3)
    with pytest.raises(RuntimeError, match="Not enough good points"):
        fitfuncs.multi_exp_fit(np.array([1.0, np.nan, 2.0, np.nan]), components=1)
    # two components on a positive axis: segments + dropped inner offset -> (a0, k0, a1, k1, offset)
    t = 0.05 * np.arange(1, 257)
    y = models.multi_exp(t, 900.0, 2.2, 300.0, 0.4, 4.0)
    g = fitfuncs.guess_multi_exp(y, t, 2, offset=True)
    assert g.shape == (5,) and g[1] > g[3] > 0 and g[0] > 0 and g[2] > 0
    assert fitfuncs.guess_multi_exp(y, t, 2, offset=False).shape == (4,)


def test_integration_stub_mirrors_the_descriptor_layout():
    """The ctypes structure shown to a dphtools maintainer in INTEGRATION.md has the fields and offsets of dphlm_desc."""
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    m = re.search(r'class _Desc\(ctypes\.Structure\):[^\n]*\n(    _fields_ = .*?\("consts", ctypes\.c_void_p\)\])', text, re.S)
    assert m, "stub not found"
    ns = {"ctypes": ctypes}
    exec("class _Desc(ctypes.Structure):\n" + m.group(1), ns)
    stub = ns["_Desc"]
    assert [f[0] for f in stub._fields_] == [f[0] for f in _lib.Desc._fields_]
    assert ctypes.sizeof(stub) == ctypes.sizeof(_lib.Desc)
    for name, _ in stub._fields_:
        assert getattr(stub, name).offset == getattr(_lib.Desc, name).offset, name
