"""ctypes loader for oracle/mcov.c and (if built) the reference's compiled .pyx.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import ctypes
import glob
import importlib.util
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.build_mcov()
        lib = ctypes.CDLL(path)
        for name in ("mcov_pairwise", "mcov_closed"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_double
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        _lib = lib
    return _lib


def _prep(p, v):
    p = np.ascontiguousarray(p, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    if p.ndim != 1 or v.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1)")
    if p.shape[0] != v.shape[0]:
        raise ValueError("p and v must be same shape")  # multinomial_cov.pyx:21-22
    return p, v


def calculate_cov(p, v, r, closed=False):
    """nucleoatac/multinomial_cov.pyx:20-31; ``r`` truncated to C int like the Cython wrapper."""
    p, v = _prep(p, v)
    fn = _load().mcov_closed if closed else _load().mcov_pairwise
    return fn(p.ctypes.data, v.ctypes.data, p.shape[0], int(r))


def reference_calculate_cov():
    """The reference's own compiled calculateCov from oracle/_ref, or None if not built."""
    sos = glob.glob(os.path.join(_HERE, "_ref", "multinomial_cov*.so"))
    if not sos:
        return None
    spec = importlib.util.spec_from_file_location("multinomial_cov", sos[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.calculateCov
