"""calculateCov(p, v, r) (nucleoatac/multinomial_cov.pyx:20-31) on the device, closed form r*(sum p v^2 - (sum p v)^2)."""
import numpy as np

from .engine import default_engine


def calculateCov(p, v, r):
    p = np.asarray(p)
    v = np.asarray(v)
    if p.dtype != np.float64 or v.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got '%s'" % (p.dtype if p.dtype != np.float64 else v.dtype))
    return default_engine().multinomial_cov(p, v, int(r))  # `int r` in the Cython signature truncates
