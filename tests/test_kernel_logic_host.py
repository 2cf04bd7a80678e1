"""The kernel's fp32 algorithm (dphtools_b200/csrc/lm_core.cuh compiled for the host, one lane per fit)
against the golden vectors minted from the live reference.  Runs without a GPU."""

import numpy as np
import pytest

from dphtools_b200 import models
from oracle import host_emu
from oracle.compare import parity_report

# name, min info-equality, min fraction within rtol 1e-4
CASES = [
    ("c1_mle", 0.99, 1.0), ("c2_mle", 0.995, 1.0), ("c2_ls", 0.99, 1.0), ("c3_mle", 0.995, 0.999),
    ("c4_mle", 0.995, 1.0), ("c4_ls", 0.995, 1.0), ("c5_mle", 0.995, 1.0), ("c5_ls", 0.995, 1.0),
]


@pytest.mark.parametrize("name,min_info,min_within", CASES)
def test_host_build_of_kernel_core_matches_reference(golden, name, min_info, min_within):
    g = golden(name)
    out = host_emu.fit_batch(models.BUILTIN[g["model"]].model_id, g["method"], g["data"], g["p0"], g["xaxis"])
    rep = parity_report(g["model"], out, g)
    assert rep["converged_equal"] == 1.0, rep
    assert rep["info_equal"] >= min_info, rep
    assert rep["within_rtol"] >= min_within, rep
    # stage 1 lands on MINPACK's least-squares optimum
    ok = np.isfinite(g["p_ls"]).all(1)
    rel = np.abs(out["p_ls"][ok] - g["p_ls"][ok]) / np.abs(g["p_ls"][ok])
    assert np.quantile(rel.max(1), 0.99) < 1e-4
