"""B200-native batched Levenberg-Marquardt / Poisson-MLE fitter (drop-in for ``dphtools.utils.lm``)."""

from . import models  # noqa: F401
