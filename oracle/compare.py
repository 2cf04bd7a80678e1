"""Parity report helpers shared by the tests, smoke() and bench.py.  TEST INFRASTRUCTURE."""

import numpy as np

# parameters on which the north-star tolerance (relative 1e-4) is stated: amplitude, centre, width
KEY_PARAMS = {"gauss2d": [0, 1, 2, 3], "gauss2d_integrated": [0, 1, 2, 3], "gauss2d_integrated_fixed": [0, 1, 2], "gauss2d_fixed": [0, 1, 2], "gauss2d_elliptical": [0, 1, 2, 3, 4], "multi_exp": [0, 1, 2, 3],
              "multi_exp_convolved": [0, 1, 2, 3]}


def parity_report(model_name, got, ref, rtol=1e-4):
    """Compare dict(popt, info, nfev) from the CUDA path (or emulation) with the oracle/golden one."""
    gi, ri = np.asarray(got["info"]), np.asarray(ref["info"])
    conv_g, conv_r = (gi >= 1) & (gi <= 4), (ri >= 1) & (ri <= 4)
    both = conv_g & conv_r
    keys = KEY_PARAMS[model_name]
    gp = np.asarray(got["popt"], dtype=np.float64)[:, keys]
    rp = np.asarray(ref["popt"], dtype=np.float64)[:, keys]
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(gp - rp) / np.abs(rp)
    rel = np.where(both[:, None], rel, 0.0)
    worst = rel.max(1)
    return dict(
        n=int(gi.size),
        converged_equal=float((conv_g == conv_r).mean()),
        info_equal=float((gi == ri).mean()),
        nfev_equal=float((np.asarray(got["nfev"]) == np.asarray(ref["nfev"])).mean()),
        within_rtol=float((worst[both] <= rtol).mean()) if both.any() else 1.0,
        max_rel=float(worst.max()),
        p999_rel=float(np.quantile(worst[both], 0.999)) if both.any() else 0.0,
        median_rel=float(np.median(worst[both])) if both.any() else 0.0,
        worst_index=int(worst.argmax()),
    )
