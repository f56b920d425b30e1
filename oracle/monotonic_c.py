"""ctypes loader for ``oracle/monotonic.c`` (compiled on demand with gcc).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "monotonic.c")
_LIB = os.path.join(_HERE, "_build", "liboracle_monotonic.so")
_lib = None


def build(force=False):
    """Compile the C restatement (gcc -O2, no fast-math: evaluation order is part of the contract)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _LIB, _SRC])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        for name, ct in (("oracle_monotonic_sweep_f64", ctypes.c_double), ("oracle_monotonic_sweep_f32", ctypes.c_float)):
            fn = getattr(_lib, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                           ctypes.c_int, ctypes.c_int, ct]
    return _lib


def sweep(flat_img, weights, offsets, dist_idx, min_gradient):
    """In-place sweep with the call signature of the reference binding
    (scarlet/operators_pybind11.cc:243-246): ``(flat_img, weights, offsets, dist_idx, min_gradient)``."""
    lib = _load()
    if not (isinstance(flat_img, np.ndarray) and flat_img.flags.c_contiguous and flat_img.ndim == 1):
        raise TypeError("flat_img must be a contiguous 1-D array (mutated in place)")
    if flat_img.dtype == np.float64:
        fn, dt = lib.oracle_monotonic_sweep_f64, np.float64
    elif flat_img.dtype == np.float32:
        fn, dt = lib.oracle_monotonic_sweep_f32, np.float32
    else:
        raise TypeError("flat_img must be float32 or float64")
    w = np.ascontiguousarray(weights, dtype=dt)
    off = np.ascontiguousarray(offsets, dtype=np.int32)
    idx = np.ascontiguousarray(dist_idx, dtype=np.int32)
    fn(flat_img.ctypes.data, w.ctypes.data, off.ctypes.data, int(off.size), idx.ctypes.data, int(idx.size),
       int(flat_img.size), float(min_gradient))
    return None
