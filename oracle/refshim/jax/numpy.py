"""jax.numpy -> NumPy (only so that module-level references resolve)."""
from numpy import *  # noqa: F401,F403
import numpy as _np

ndarray = _np.ndarray
linalg = _np.linalg
complex128 = _np.complex128
float64 = _np.float64
newaxis = _np.newaxis
