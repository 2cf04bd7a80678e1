"""Development aid: the fp32 kernel core (host build, one lane per fit) against every golden set.

    python scripts/host_parity.py [set ...]

Prints info / nfev equality and parameter agreement per set; the CPU test-suite asserts the same numbers
(tests/test_kernel_logic_host.py).  TEST INFRASTRUCTURE.
"""

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dphtools_b200 import models  # noqa: E402
from oracle import host_emu  # noqa: E402
from oracle.compare import parity_report  # noqa: E402
from oracle.golden import GOLDEN_DIR, load_golden  # noqa: E402


def main():
    names = sys.argv[1:] or sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("x_"))
    host_emu.build()
    for name in names:
        g = load_golden(name)
        t0 = time.time()
        out = host_emu.fit_batch(models.resolve(g["model"], g["consts"]).model_id, g["method"], g["data"], g["p0"], g["xaxis"],
                                 consts=g["consts"])
        dt = time.time() - t0
        rep = parity_report(g["model"], out, g)
        swaps = {}
        for a, b in zip(out["info"], g["info"]):
            if a != b:
                swaps[(int(b), int(a))] = swaps.get((int(b), int(a)), 0) + 1
        print("%-8s n=%5d info=%.4f nfev=%.4f conv=%.4f within=%.5f max_rel=%.2e p999=%.2e  swaps(ref->got)=%s  %.1fs" % (
            name, rep["n"], rep["info_equal"], rep["nfev_equal"], rep["converged_equal"], rep["within_rtol"], rep["max_rel"],
            rep["p999_rel"], swaps, dt))


if __name__ == "__main__":
    main()
