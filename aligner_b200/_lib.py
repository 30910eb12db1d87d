"""ctypes binding of libaligner_b200.so (the C ABI declared in include/aligner_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at
import time, and every entry point fails loudly without an sm_100 device.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

_PKG = Path(__file__).resolve().parent
import os as _os
LIB_PATH = Path(_os.environ["ALB200_LIB"]) if _os.environ.get("ALB200_LIB") else _PKG / "libaligner_b200.so"   # ALB200_LIB: A/B builds while tuning

OK = 0
E_INVALID, E_UNSUPPORTED, E_CUDA, E_LENGTHS, E_NO_DEVICE = -1, -2, -3, -4, -5

# mask element types (ALB200_* in the header)
F32, F16, BF16, F64, U8, I8, I16, I32, I64 = range(9)
LAYOUT_VITS = 0x100   # OR into the value dtype of alb200_mas_device_ex: scores and path are [b, t_mel, t_text]

# every symbol include/aligner_b200.h declares (tests check the export list against the header)
SYMBOLS = (
    "alb200_last_error", "alb200_version", "alb200_set_option", "alb200_mas_device", "alb200_mas_device_ex", "alb200_mas_device_masked",
    "alb200_mas_workspace_bytes", "alb200_mas_status", "alb200_mas_describe", "alb200_maximum_path_c",
    "alb200_last_transfer_bytes", "alb200_launch_count",
    "alb200_neg_cent_gaussian", "alb200_neg_cent_ota",
    "alb200_neg_cent_workspace_bytes", "alb200_neg_cent_gaussian_ws", "alb200_neg_cent_ota_ws",
    "alb200_fused_workspace_bytes", "alb200_gaussian_mas_fused", "alb200_neg_cent_ota_bb",
)


class AlignerB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("aligner_b200 error %d: %s" % (code, msg))
        self.code = code


def _load() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            "%s not found: build it with `python build_lib.py` "
            "(aligner_b200 has no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i32, i64, u64, f32, sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64,
                                  ctypes.c_float, ctypes.c_size_t)
    lib.alb200_last_error.restype = ctypes.c_char_p
    lib.alb200_version.restype = ctypes.c_char_p
    lib.alb200_launch_count.restype = u64
    lib.alb200_last_transfer_bytes.argtypes = [ctypes.POINTER(u64), ctypes.POINTER(u64)]
    lib.alb200_last_transfer_bytes.restype = None
    lib.alb200_mas_device.argtypes = [vp, vp, vp, vp, i32, u64, i32, vp, vp, i32, i32, i32, f32, vp, sz, vp]
    lib.alb200_mas_device.restype = i32
    lib.alb200_set_option.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    lib.alb200_set_option.restype = i32
    lib.alb200_mas_device_ex.argtypes = [vp, i32, vp, vp, vp, i32, i64, i64, i64, vp, i32, u64, i32, vp, vp, vp, i32, i32, i32, f32, vp, sz, vp]
    lib.alb200_mas_device_ex.restype = i32
    lib.alb200_mas_device_masked.argtypes = [vp, vp, i32, i64, i64, i64, vp, i32, u64, i32, vp, vp, vp,
                                             i32, i32, i32, f32, vp, sz, vp]
    lib.alb200_mas_device_masked.restype = i32
    lib.alb200_mas_workspace_bytes.argtypes = [i32, i32, i32]
    lib.alb200_mas_workspace_bytes.restype = sz
    lib.alb200_mas_describe.argtypes = [i32, i32, i32, i32, ctypes.c_char_p, sz]
    lib.alb200_mas_describe.restype = i32
    lib.alb200_mas_status.argtypes = [vp, vp]
    lib.alb200_mas_status.restype = i32
    lib.alb200_maximum_path_c.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32]
    lib.alb200_maximum_path_c.restype = i32
    lib.alb200_neg_cent_gaussian.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.alb200_neg_cent_gaussian.restype = i32
    lib.alb200_neg_cent_ota.argtypes = [vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]
    lib.alb200_neg_cent_ota.restype = i32
    lib.alb200_neg_cent_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    lib.alb200_neg_cent_workspace_bytes.restype = sz
    lib.alb200_neg_cent_gaussian_ws.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, sz, vp]
    lib.alb200_neg_cent_gaussian_ws.restype = i32
    lib.alb200_neg_cent_ota_ws.argtypes = [vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp, sz, vp]
    lib.alb200_neg_cent_ota_ws.restype = i32
    lib.alb200_neg_cent_ota_bb.argtypes = [vp, vp, vp, vp, f32, vp, f32, i32, i32, i32, i32, vp, sz, vp]
    lib.alb200_neg_cent_ota_bb.restype = i32
    lib.alb200_fused_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.alb200_fused_workspace_bytes.restype = sz
    lib.alb200_gaussian_mas_fused.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i64, i64, i64, vp, i32, u64, i32, vp, vp,
                                              i32, i32, i32, i32, f32, vp, sz, vp]
    lib.alb200_gaussian_mas_fused.restype = i32
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != OK:
        raise AlignerB200Error(rc, lib.alb200_last_error().decode("utf-8", "replace"))


def set_option(name: str, value=None) -> None:
    """Tuning / test options (include/aligner_b200.h: alb200_set_option); value None restores the default."""
    check(lib.alb200_set_option(name.encode(), None if value is None else str(value).encode()))
    for hook in _option_hooks:
        hook()


_option_hooks: list = []      # callables run after every option change (the Python layer drops its workspace-size memo)


def describe(b: int, tx: int, ty: int, want_durations: bool = False) -> str:
    buf = ctypes.create_string_buffer(256)
    check(lib.alb200_mas_describe(b, tx, ty, int(want_durations), buf, 256))
    return buf.value.decode()


def launch_count() -> int:
    return int(lib.alb200_launch_count())


def last_transfer_bytes() -> tuple[int, int]:
    a, b = ctypes.c_uint64(0), ctypes.c_uint64(0)
    lib.alb200_last_transfer_bytes(ctypes.byref(a), ctypes.byref(b))
    return int(a.value), int(b.value)
