"""Evidence for the central numerical design decision ("fp32 and the 1.49e-8 tolerances"): the reference's loop
evaluated naively in fp32 loses its convergence flags, the differential chi-square evaluation keeps them.  Uses the numpy
prototype of the kernel algorithm (oracle/kernel_emulation.py), which can switch between the two."""

import numpy as np

from oracle import kernel_emulation as ke
from oracle.compare import parity_report


def _run(g, n, differential):
    ap = ke._Approx(np.float32, True)
    model = ke.Gauss2D(11, 11, np.float32, ap)
    out = ke.fit(model, g["data"][:n], g["p0"][:n], "mle", dtype=np.float32, differential=differential)
    return parity_report("gauss2d", out, {k: g[k][:n] for k in ("popt", "info", "nfev")})


def test_naive_fp32_flips_the_flags_differential_fp32_keeps_them(golden):
    g = golden("c2_mle")
    naive = _run(g, 300, differential=False)
    diff = _run(g, 300, differential=True)
    assert naive["info_equal"] < 0.2          # info 1 -> 2 on (almost) every fit
    assert naive["within_rtol"] < 0.995       # and some parameters drift beyond 1e-4
    assert diff["info_equal"] > 0.99 and diff["within_rtol"] == 1.0 and diff["max_rel"] < 1e-5
