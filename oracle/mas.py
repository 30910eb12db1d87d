"""oracle/mas.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy-facing wrappers around the C restatement (``mas_oracle.c``) and loaders
for the compiled, unmodified reference core (``oracle/_ref``).

Reference interfaces restated here (paths relative to the reference root):
  monotonic_align/core.pyx:40      maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9)
  monotonic_align/__init__.py:6-21 maximum_path(value, mask)
"""
from __future__ import annotations

import ctypes
import importlib.machinery
import importlib.util
import sys
from pathlib import Path

import numpy as np

from . import build as _build

HERE = Path(__file__).resolve().parent

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)
_libs: dict[str, ctypes.CDLL] = {}


def _lib(omp: bool) -> ctypes.CDLL:
    key = "omp" if omp else "serial"
    if key not in _libs:
        path = _build.build_restatement()[key]
        lib = ctypes.CDLL(str(path))
        lib.mas_oracle_full.argtypes = [_i32p, _f32p, _i32p, _i32p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_float]
        lib.mas_oracle_full.restype = ctypes.c_int
        lib.mas_oracle_bits.argtypes = [_i32p, _f32p, _i32p, _i32p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_float, _i32p]
        lib.mas_oracle_bits.restype = ctypes.c_int
        lib.mas_oracle_threads.restype = ctypes.c_int
        _libs[key] = lib
    return _libs[key]


def _chk(a: np.ndarray, dtype, ndim: int, name: str) -> np.ndarray:
    # same complaints the Cython buffer protocol raises (core.c:19883-19886)
    if a.dtype != dtype:
        raise ValueError("Buffer dtype mismatch for %s: expected %s got %s" % (name, np.dtype(dtype), a.dtype))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions for %s" % name)
    if not a.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return a


def port_threads(omp: bool = True) -> int:
    return int(_lib(omp).mas_oracle_threads())


def maximum_path_c_port(paths, values, t_xs, t_ys, max_neg_val: float = -1e9, omp: bool = False) -> None:
    """C restatement with the reference's exact calling convention (core.pyx:40):
    fills ``paths`` (pre-zeroed int32) in place and clobbers ``values``."""
    _chk(paths, np.int32, 3, "paths"); _chk(values, np.float32, 3, "values")
    _chk(t_xs, np.int32, 1, "t_xs"); _chk(t_ys, np.int32, 1, "t_ys")
    b, tx, ty = values.shape
    _lib(omp).mas_oracle_full(paths.ctypes.data_as(_i32p), values.ctypes.data_as(_f32p),
                              t_xs.ctypes.data_as(_i32p), t_ys.ctypes.data_as(_i32p),
                              b, tx, ty, ctypes.c_float(max_neg_val))


def mas_bits_port(values, t_xs, t_ys, max_neg_val: float = -1e9, omp: bool = False):
    """Running-column + direction-bit restatement (what the GPU kernel does).
    Returns (paths int32 [b,tx,ty], frame_tok int32 [b,ty]); values untouched."""
    _chk(values, np.float32, 3, "values")
    b, tx, ty = values.shape
    paths = np.zeros((b, tx, ty), np.int32)
    ftok = np.empty((b, ty), np.int32)
    _lib(omp).mas_oracle_bits(paths.ctypes.data_as(_i32p), values.ctypes.data_as(_f32p),
                              np.ascontiguousarray(t_xs, np.int32).ctypes.data_as(_i32p),
                              np.ascontiguousarray(t_ys, np.int32).ctypes.data_as(_i32p),
                              b, tx, ty, ctypes.c_float(max_neg_val), ftok.ctypes.data_as(_i32p))
    return paths, ftok


def maximum_path_port(value, mask, omp: bool = False):
    """Restatement of the reference Python API (``__init__.py:6-21``) on top of the
    C restatement.  torch in, torch out; same dtype/device rules."""
    import torch

    value = value * mask                                   # __init__.py:11
    device, dtype = value.device, value.dtype              # __init__.py:12-13
    v = value.data.cpu().numpy().astype(np.float32)        # __init__.py:14
    path = np.zeros_like(v).astype(np.int32)               # __init__.py:15
    m = mask.data.cpu().numpy()                            # __init__.py:16
    t_x = m.sum(1)[:, 0].astype(np.int32)                  # __init__.py:18
    t_y = m.sum(2)[:, 0].astype(np.int32)                  # __init__.py:19
    maximum_path_c_port(path, np.ascontiguousarray(v), t_x, t_y, omp=omp)
    return torch.from_numpy(path).to(device=device, dtype=dtype)  # __init__.py:21


# ---------------------------------------------------------------------------
# the compiled, unmodified reference (oracle/_ref)
# ---------------------------------------------------------------------------
_ref_cores: dict[str, object] = {}


def load_reference_core(kind: str = "serial"):
    """Return the reference's own compiled ``core`` module (``maximum_path_c``),
    or None if it is neither prebuilt nor buildable here.  kind: 'serial' (as the
    reference ships it: no -fopenmp) or 'omp'."""
    if kind in _ref_cores:
        return _ref_cores[kind]
    outs = _build.build_reference()
    mod = None
    if outs is not None and outs[kind].exists():
        name = "aligner_ref_%s.core" % kind
        loader = importlib.machinery.ExtensionFileLoader(name, str(outs[kind]))
        spec = importlib.util.spec_from_file_location(name, str(outs[kind]), loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    _ref_cores[kind] = mod
    return mod


def load_reference_api(kind: str = "serial"):
    """Import the reference's own ``monotonic_align/__init__.py`` from where it
    lies, bound to the core compiled into oracle/_ref.  Only possible in the
    build container (needs /root/reference); returns None elsewhere."""
    core = load_reference_core(kind)
    init = _build.REF_SRC / "__init__.py"
    if core is None or not init.exists():
        return None
    pkg = "aligner_ref_%s_api" % kind
    if pkg in sys.modules:
        return sys.modules[pkg]
    import types

    # the reference does `from .monotonic_align.core import maximum_path_c` (__init__.py:3)
    spec = importlib.util.spec_from_file_location(pkg, str(init), submodule_search_locations=[])
    mod = importlib.util.module_from_spec(spec)
    sub = types.ModuleType(pkg + ".monotonic_align")
    sub.__path__ = []
    sub.core = core
    sys.modules[pkg] = mod
    sys.modules[pkg + ".monotonic_align"] = sub
    sys.modules[pkg + ".monotonic_align.core"] = core
    spec.loader.exec_module(mod)
    return mod
