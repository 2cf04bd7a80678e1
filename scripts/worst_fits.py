"""Development aid: the fits of a golden set that are furthest from the reference on the GPU.  usage: worst_fits.py [set] [group_lanes]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dphtools_b200 as dph  # noqa: E402
from dphtools_b200 import models  # noqa: E402
from oracle.golden import load_golden  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3_mle"
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = load_golden(name)
model = models.resolve(g["model"], g["consts"])
for et in (False, True):
    r = dph.curve_fit_batch(model, torch.from_numpy(g["data"]).cuda(), torch.from_numpy(g["p0"]).cuda(), g["method"], g["xaxis"],
                            exact_ties=et, group_lanes=lanes)
    popt, info, nfev = r.popt.cpu().numpy(), r.info.cpu().numpy(), r.nfev.cpu().numpy()
    rel = (np.abs(popt - g["popt"]) / np.maximum(np.abs(g["popt"]), 1e-12))[:, :4].max(1)
    for i in np.argsort(-rel)[:4]:
        print(name, "exact_ties" if et else "plain", "fit", i, "rel %.3g" % rel[i], "info", info[i], g["info"][i], "nfev", nfev[i], g["nfev"][i],
              "popt", np.round(popt[i], 4), "ref", np.round(g["popt"][i], 4))
