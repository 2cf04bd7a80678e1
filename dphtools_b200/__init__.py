"""B200-native batched Levenberg-Marquardt / Poisson-MLE fitter (drop-in for ``dphtools.utils.lm``)."""

from . import fitfuncs, models, synth  # noqa: F401
from .fitfuncs import exponent_fit, multi_exp_fit, multi_exp_fit_batch  # noqa: F401
from .lm import BatchResult, call_status, covariance, curve_fit, curve_fit_batch, extract_rois, host_wait, join, lm, pinned_empty  # noqa: F401
from .models import gauss2d, gauss2d_elliptical, multi_exp  # noqa: F401

__version__ = "0.2.0"
