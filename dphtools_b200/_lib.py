"""ctypes binding of the C ABI in ``include/dphlm.h`` (``dphtools_b200/csrc/libdphlm.so``).

There is no CPU fallback: if the CUDA library has not been built, importing this module's ``lib()``
raises, and so does every fit.
"""

import ctypes
import os

LIB_PATH = os.environ.get("DPHLM_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libdphlm.so")

SYMBOLS = [
    "dphlm_desc_init", "dphlm_fit_batch", "dphlm_fit_batch_host", "dphlm_host_alloc", "dphlm_host_free",
    "dphlm_device_count", "dphlm_version", "dphlm_last_error", "dphlm_last_kernel_ms", "dphlm_covariance", "dphlm_extract_rois",
    "dphlm_join", "dphlm_call_status", "dphlm_measure_fp32_peak", "dphlm_fit_batch_host_async", "dphlm_host_wait",
    "dphlm_guess_multi_exp",
]


class Desc(ctypes.Structure):
    """Mirror of ``struct dphlm_desc``."""

    _fields_ = [
        ("model", ctypes.c_int32), ("method", ctypes.c_int32), ("n_param", ctypes.c_int32),
        ("dim0", ctypes.c_int32), ("dim1", ctypes.c_int32), ("group_lanes", ctypes.c_int32),
        ("n_fits", ctypes.c_int64),
        ("data", ctypes.c_void_p), ("xaxis", ctypes.c_void_p), ("p0", ctypes.c_void_p),
        ("popt", ctypes.c_void_p), ("info", ctypes.c_void_p), ("nfev", ctypes.c_void_p), ("chisq", ctypes.c_void_p),
        ("p_ls", ctypes.c_void_p), ("counts", ctypes.c_void_p), ("x_acc", ctypes.c_void_p),
        ("ftol", ctypes.c_float), ("xtol", ctypes.c_float), ("gtol", ctypes.c_float), ("factor", ctypes.c_float),
        ("maxfev", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("consts", ctypes.c_void_p),
    ]


class ExtractDesc(ctypes.Structure):
    """Mirror of ``struct dphlm_extract_desc``."""

    _fields_ = [
        ("frames", ctypes.c_void_p), ("frame_dtype", ctypes.c_int32), ("n_frames", ctypes.c_int32),
        ("height", ctypes.c_int32), ("width", ctypes.c_int32), ("peaks", ctypes.c_void_p), ("n_peaks", ctypes.c_int64),
        ("roi", ctypes.c_int32), ("refine", ctypes.c_int32), ("sigma0", ctypes.c_float),
        ("rois", ctypes.c_void_p), ("p0", ctypes.c_void_p), ("valid", ctypes.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "dphtools_b200: the CUDA library %s has not been built (run `python -m dphtools_b200.build`); "
                "there is no CPU fallback" % LIB_PATH
            )
        L = ctypes.CDLL(LIB_PATH)
        L.dphlm_desc_init.argtypes = [ctypes.POINTER(Desc)]
        L.dphlm_desc_init.restype = None
        L.dphlm_fit_batch.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p]
        L.dphlm_fit_batch.restype = ctypes.c_int
        L.dphlm_fit_batch_host.argtypes = [ctypes.POINTER(Desc), ctypes.c_int]
        L.dphlm_fit_batch_host.restype = ctypes.c_int
        L.dphlm_fit_batch_host_async.argtypes = [ctypes.POINTER(Desc), ctypes.c_int]
        L.dphlm_fit_batch_host_async.restype = ctypes.c_int
        L.dphlm_host_wait.argtypes = [ctypes.c_int]
        L.dphlm_host_wait.restype = ctypes.c_int
        L.dphlm_guess_multi_exp.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.dphlm_guess_multi_exp.restype = ctypes.c_int
        L.dphlm_covariance.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.dphlm_covariance.restype = ctypes.c_int
        L.dphlm_extract_rois.argtypes = [ctypes.POINTER(ExtractDesc), ctypes.c_void_p]
        L.dphlm_extract_rois.restype = ctypes.c_int
        L.dphlm_host_alloc.argtypes = [ctypes.c_size_t]
        L.dphlm_host_alloc.restype = ctypes.c_void_p
        L.dphlm_host_free.argtypes = [ctypes.c_void_p]
        L.dphlm_host_free.restype = None
        L.dphlm_device_count.restype = ctypes.c_int
        L.dphlm_last_kernel_ms.argtypes = [ctypes.POINTER(ctypes.c_float * 3)]
        L.dphlm_last_kernel_ms.restype = ctypes.c_int
        L.dphlm_join.argtypes = [ctypes.c_void_p]
        L.dphlm_join.restype = ctypes.c_int
        L.dphlm_call_status.argtypes = [ctypes.POINTER(ctypes.c_int)]
        L.dphlm_call_status.restype = ctypes.c_int
        L.dphlm_measure_fp32_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
        L.dphlm_measure_fp32_peak.restype = ctypes.c_int
        L.dphlm_version.restype = ctypes.c_int
        L.dphlm_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


FLAG_TIMING = 1
FLAG_NO_HANDOFF = 2
FLAG_DATA_U16 = 4
FLAG_PIPELINED = 8
FLAG_SKIP_PREFIT = 16
FLAG_EXACT_TIES = 32
INFO_UNFINISHED = -2139062144  # DPHLM_INFO_UNFINISHED


def last_kernel_ms():
    """(pre-fit main kernel, lm() main kernel, straggler tail after it) in ms for this thread's last timed device call; blocks
    until it has finished."""
    ms = (ctypes.c_float * 3)()
    if lib().dphlm_last_kernel_ms(ctypes.byref(ms)) != 0:
        raise RuntimeError("dphlm: " + last_error())
    return float(ms[0]), float(ms[1]), float(ms[2])


def last_error():
    return lib().dphlm_last_error().decode()


def measure_fp32_peak():
    """Measured FP32 FFMA throughput of the current device, TFLOP/s."""
    v = ctypes.c_double(0.0)
    if lib().dphlm_measure_fp32_peak(ctypes.byref(v)) != 0:
        raise RuntimeError("dphlm: " + last_error())
    return float(v.value)
