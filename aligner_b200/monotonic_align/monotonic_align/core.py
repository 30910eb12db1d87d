"""Stand-in for the reference's compiled module ``monotonic_align.monotonic_align.core``.

``maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9)`` has the calling
convention of the Cython original (monotonic_align/core.pyx:40): numpy arrays in
host memory, ``paths`` pre-zeroed int32 and filled in place, returns None.  The
work happens on the GPU through ``alb200_maximum_path_c``; ``values`` is left
untouched (the reference overwrites it with cumulative scores, which no caller
can observe through ``maximum_path``).
"""
from __future__ import annotations

import numpy as np

from ... import _lib


def _buf(a, dtype, ndim: int, name: str, writable: bool = False) -> np.ndarray:
    # same checks, same exception types as the Cython buffer protocol (core.c:19883-19886, 27929-27995)
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.dtype != dtype:
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'" % (np.dtype(dtype).name, a.dtype.name))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    if not a.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    if writable and not a.flags.writeable:
        raise ValueError("buffer source array is read-only")
    return a


def maximum_path_c(paths, values, t_xs, t_ys, max_neg_val: float = -1e9) -> None:
    paths = _buf(paths, np.int32, 3, "paths", writable=True)
    values = _buf(values, np.float32, 3, "values")
    t_xs = _buf(t_xs, np.int32, 1, "t_xs")
    t_ys = _buf(t_ys, np.int32, 1, "t_ys")
    b, tx, ty = values.shape
    if paths.shape != values.shape or t_xs.shape[0] < b or t_ys.shape[0] < b:
        raise ValueError("paths/values/t_xs/t_ys shapes disagree")
    if b == 0 or tx == 0 or ty == 0:
        return None
    rc = _lib.lib.alb200_maximum_path_c(paths.ctypes.data, values.ctypes.data, t_xs.ctypes.data, t_ys.ctypes.data,
                                        b, tx, ty, float(max_neg_val))
    if rc == _lib.E_LENGTHS:
        raise ValueError(_lib.lib.alb200_last_error().decode())
    _lib.check(rc)
    return None
