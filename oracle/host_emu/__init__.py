"""Host (g++) build of dphtools_b200/csrc/lm_core.cuh, one lane per fit.  TEST INFRASTRUCTURE ONLY."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emulation.cpp")
CORE = os.path.join(HERE, "..", "..", "dphtools_b200", "csrc", "lm_core.cuh")
LIB = os.path.join(HERE, "libdphlm_emu.so")
_lib = None


def build(force=False):
    newest = max(os.path.getmtime(SRC), os.path.getmtime(CORE))
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(
            ["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", LIB]
        )
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emu_fit_batch.restype = ctypes.c_int
    return _lib


def fit_batch(model_id, method, data, p0, xaxis=None, ftol=1.49012e-8, xtol=1.49012e-8, gtol=0.0, factor=100.0, maxfev=0, budget=0, consts=None, skip_prefit=False, exact_ties=False):
    data = np.ascontiguousarray(data, dtype=np.float32)
    p0 = np.ascontiguousarray(p0, dtype=np.float32)
    n, npar = p0.shape
    if data.ndim == 3:
        h, w = data.shape[1:]
    else:
        h, w = data.shape[1], 1
    xa = None if xaxis is None else np.ascontiguousarray(xaxis, dtype=np.float32)
    cs = None if consts is None else np.ascontiguousarray(consts, dtype=np.float32)
    out = dict(
        popt=np.empty((n, npar), np.float32), p_ls=np.empty((n, npar), np.float32), info=np.empty(n, np.int32),
        nfev=np.empty(n, np.int32), chisq=np.empty(n, np.float32), counts=np.empty((n, 4), np.int32),
    )
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib().emu_fit_batch(
        ctypes.c_int(model_id), ctypes.c_int(1 if method == "mle" else 0), ctypes.c_int(npar), ctypes.c_int(h), ctypes.c_int(w),
        ctypes.c_long(n), ptr(data), ptr(xa) if xa is not None else None, ptr(p0), ctypes.c_float(ftol), ctypes.c_float(xtol),
        ctypes.c_float(gtol), ctypes.c_float(factor), ctypes.c_int(maxfev), ctypes.c_int(budget), ptr(out["popt"]), ptr(out["p_ls"]), ptr(out["info"]),
        ptr(out["nfev"]), ptr(out["chisq"]), ptr(out["counts"]), ptr(cs) if cs is not None else None, ctypes.c_int(1 if skip_prefit else 0), ctypes.c_int(1 if exact_ties else 0),
    )
    if rc != 0:
        raise ValueError("unsupported model/n_param")
    return out
