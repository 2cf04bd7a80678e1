"""Loader for the golden vectors in tests/golden/ (written by oracle/make_golden.py).  TEST INFRASTRUCTURE."""

import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["data"] = g["data"].astype(np.float32)
    g["xaxis"] = g["xaxis"] if g["xaxis"].size else None
    g["consts"] = g["consts"] if "consts" in g and g["consts"].size else None  # model constants (fixed-width Gaussian)
    g["workload"], _, _, g["method"], g["model"] = [str(s) for s in g["meta"]]
    return g
