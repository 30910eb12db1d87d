"""Drop-in for the reference package ``monotonic_align`` (monotonic_align/__init__.py:1-21).

``maximum_path(value, mask)`` keeps the reference's signature, return shape,
dtype rule (``torch.result_type(value, mask)``), device rule (``value.device``)
and values (exactly 0 / 1).  CUDA tensors never leave the GPU: lengths are
derived from the mask inside the kernel, the search runs in one sm_100a kernel
launch on torch's current stream, and the dense path is written directly in the
result dtype.  CPU tensors (the reference accepts any device, __init__.py:12-14)
are staged through the host-pointer entry ``alb200_maximum_path_c`` -- the
search still runs on the B200, there is no CPU implementation of it here.
PyTorch is only used for memory and the stream handle.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .monotonic_align.core import maximum_path_c  # noqa: F401  (same import the reference does, __init__.py:3)

__all__ = ["maximum_path", "maximum_path_vits", "maximum_path_lengths", "maximum_path_c", "check_status"]

# bit pattern of 1 in every dtype torch.result_type can produce here
_ONE = {
    torch.float32: (4, 0x3F800000), torch.float16: (2, 0x3C00), torch.bfloat16: (2, 0x3F80),
    torch.float64: (8, 0x3FF0000000000000), torch.int8: (1, 1), torch.uint8: (1, 1), torch.bool: (1, 1),
    torch.int16: (2, 1), torch.int32: (4, 1), torch.int64: (8, 1),
}
_MASK_DTYPE = {
    torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float64: _lib.F64,
    torch.uint8: _lib.U8, torch.bool: _lib.U8, torch.int8: _lib.I8, torch.int16: _lib.I16,
    torch.int32: _lib.I32, torch.int64: _lib.I64,
}

_workspaces: dict = {}
_ws_bytes: dict = {}
_lib._option_hooks.append(_ws_bytes.clear)      # a forced kernel shape changes the scratch size


def _workspace(device: torch.device, stream: int, b: int, tx: int, ty: int) -> torch.Tensor:
    """Zero-initialised scratch, one per (device, stream) so concurrent launches on
    different streams never share the work-stealing counter."""
    skey = (device.index, b, tx, ty)
    need = _ws_bytes.get(skey)
    if need is None:                      # the sizing call probes several kernel configurations: once per shape
        if len(_ws_bytes) > 4096:
            _ws_bytes.clear()
        need = _ws_bytes[skey] = int(_lib.lib.alb200_mas_workspace_bytes(b, tx, ty))
    key = (device.index, stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_VALUE_DTYPE = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _prep_value(value: torch.Tensor) -> torch.Tensor:
    """Contiguous scores in a dtype the kernels read directly: fp32, or fp16 / bf16 as they are (the kernel promotes
    them on load, exactly like .astype(np.float32) in the reference, __init__.py:14); anything else is promoted here."""
    if value.dim() != 3:
        raise ValueError("value must be [b, t_x, t_y], got %s" % (tuple(value.shape),))
    v = value.detach()
    if v.dtype not in _VALUE_DTYPE:
        v = v.to(torch.float32)
    return v.contiguous()


def _launch(v: torch.Tensor, args_after_value: tuple) -> None:
    """alb200_mas_device_ex on v; a half-precision tensor the library has no native kernel shape for
    (ALB200_E_UNSUPPORTED) is promoted to fp32 on the device and retried -- still no CPU anywhere."""
    rc = _lib.lib.alb200_mas_device_ex(v.data_ptr(), _VALUE_DTYPE[v.dtype], *args_after_value)
    if rc == _lib.E_UNSUPPORTED and v.dtype != torch.float32:
        v32 = v.to(torch.float32)
        rc = _lib.lib.alb200_mas_device_ex(v32.data_ptr(), _lib.F32, *args_after_value)
    _lib.check(rc)


def _maximum_path_host(value: torch.Tensor, mask: torch.Tensor, dtype: torch.dtype, apply_mask: bool, return_durations: bool):
    """CPU tensors: what the reference's own staging does (__init__.py:11-21), with the compiled core replaced by the
    host-pointer entry of the library (values go to the B200 in chunks, the token of every frame comes back)."""
    if apply_mask:
        value = value * mask                                           # __init__.py:11
    v = np.ascontiguousarray(value.detach().to(torch.float32).numpy())  # __init__.py:14
    path = np.zeros(v.shape, np.int32)                                 # __init__.py:15
    m = mask.detach()
    if m.dtype == torch.bfloat16:
        m = m.float()
    m = m.numpy()                                                      # __init__.py:16
    t_x = np.ascontiguousarray(m.sum(1)[:, 0].astype(np.int32))        # __init__.py:18
    t_y = np.ascontiguousarray(m.sum(2)[:, 0].astype(np.int32))        # __init__.py:19
    if v.size:
        maximum_path_c(path, v, t_x, t_y)                              # __init__.py:20
    out = torch.from_numpy(path).to(dtype=dtype)                       # __init__.py:21
    return (out, torch.from_numpy(path.sum(-1).astype(np.int32))) if return_durations else out


def maximum_path(value: torch.Tensor, mask: torch.Tensor, *, return_durations: bool = False, apply_mask: bool = False):
    """Monotonic alignment search.  Reference: monotonic_align/__init__.py:6-21.

    value: [b, t_x, t_y] scores (log-likelihoods), any float dtype, any device (the search runs on the B200).
    mask:  [b, t_x, t_y] outer product of the text and mel prefix masks, any dtype.
    Returns the 0/1 path [b, t_x, t_y] in ``torch.result_type(value, mask)`` on
    ``value.device``; inputs are not modified; no autograd history.

    Differences from the reference, all supersets: bf16 and non-contiguous inputs
    are accepted; ``return_durations=True`` also returns ``path.sum(-1)`` as int32.
    For a prefix-shaped mask (ones in [0,t_x) x [0,t_y), zeros elsewhere -- what every
    caller of the reference builds) ``value * mask`` (__init__.py:11) is the identity on
    every cell the search reads, so only mask[:, :, 0] and mask[:, 0, :] are read and the
    product is skipped.  ``apply_mask=True`` performs the multiplication first (one
    extra elementwise pass on value.device), which reproduces the reference for
    arbitrary masks too.
    """
    if mask.shape != value.shape:
        raise ValueError("mask shape %s != value shape %s" % (tuple(mask.shape), tuple(value.shape)))
    if mask.device != value.device:
        raise ValueError("mask and value must be on the same device")
    dtype = torch.result_type(value, mask)
    if dtype not in _ONE or mask.dtype not in _MASK_DTYPE:
        raise TypeError("unsupported dtype combination %s / %s" % (value.dtype, mask.dtype))
    if value.dim() != 3:
        raise ValueError("value must be [b, t_x, t_y], got %s" % (tuple(value.shape),))
    if not value.is_cuda:
        return _maximum_path_host(value, mask, dtype, apply_mask, return_durations)
    if apply_mask:
        value = value.detach() * mask.detach()
    v = _prep_value(value)
    b, tx, ty = v.shape
    esize, one = _ONE[dtype]
    with torch.cuda.device(v.device):
        path = torch.empty((b, tx, ty), dtype=dtype, device=v.device)
        dur = torch.empty((b, tx), dtype=torch.int32, device=v.device) if return_durations else None
        if b == 0 or tx == 0 or ty == 0:
            return (path.zero_(), dur) if return_durations else path.zero_()
        stream = torch.cuda.current_stream(v.device).cuda_stream
        ws = _workspace(v.device, stream, b, tx, ty)
        m = mask.detach()
        sb, sx, sy = m.stride()
        _launch(v, (None, None, m.data_ptr(), _MASK_DTYPE[m.dtype], sb, sx, sy,
                    path.data_ptr(), esize, one, 1, None, dur.data_ptr() if dur is not None else None, None,
                    b, tx, ty, -1e9, ws.data_ptr(), ws.numel(), stream))
    return (path, dur) if return_durations else path


def maximum_path_vits(neg_cent: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """VITS-layout entry: ``neg_cent`` and ``mask`` are ``[b, t_mel, t_text]`` as in VITS'
    ``monotonic_align.maximum_path`` (its core indexes ``value[y, x]``); the result has the same layout, dtype rule and
    device rule as ``maximum_path``.

    The search is layout-agnostic.  fp32 scores in the latency regime (batch <= #SMs, t_text <= 512, t_text % 4 == 0) are read
    in place by kernels whose TMA boxes are taken from the ``[b*t_mel, t_text]`` view, and the path is written directly in
    that layout; every other case hands ``maximum_path`` the transposed view (one device transpose of the scores; mask
    and result are strided views, no copy)."""
    if neg_cent.dim() != 3 or mask.shape != neg_cent.shape:
        raise ValueError("expected neg_cent and mask of shape [b, t_mel, t_text]")
    if mask.device != neg_cent.device:
        raise ValueError("mask and neg_cent must be on the same device")
    dtype = torch.result_type(neg_cent, mask)
    if dtype not in _ONE or mask.dtype not in _MASK_DTYPE:
        raise TypeError("unsupported dtype combination %s / %s" % (neg_cent.dtype, mask.dtype))
    b, ty, tx = neg_cent.shape
    if neg_cent.is_cuda and neg_cent.dtype == torch.float32 and b and tx and ty:
        v = neg_cent.detach().contiguous()
        esize, one = _ONE[dtype]
        with torch.cuda.device(v.device):
            path = torch.empty((b, ty, tx), dtype=dtype, device=v.device)
            stream = torch.cuda.current_stream(v.device).cuda_stream
            ws = _workspace(v.device, stream, b, tx, ty)
            m = mask.detach()
            sb, sy, sx = m.stride()                              # mask is [b, t_mel, t_text]; the kernel wants (b, text, mel) strides
            rc = _lib.lib.alb200_mas_device_ex(v.data_ptr(), _lib.F32 | _lib.LAYOUT_VITS, None, None, m.data_ptr(), _MASK_DTYPE[m.dtype],
                                               sb, sx, sy, path.data_ptr(), esize, one, 1, None, None, None,
                                               b, tx, ty, -1e9, ws.data_ptr(), ws.numel(), stream)
        if rc == 0:
            return path
        if rc != _lib.E_UNSUPPORTED:
            _lib.check(rc)
    return maximum_path(neg_cent.transpose(1, 2), mask.transpose(1, 2)).transpose(1, 2)


def maximum_path_lengths(value: torch.Tensor, x_lengths: torch.Tensor, y_lengths: torch.Tensor, *,
                         out_dtype: torch.dtype | None = None, dense: bool = True,
                         return_durations: bool = False, return_frame_tokens: bool = False,
                         max_neg_val: float = -1e9):
    """Mask-free entry: lengths given directly as int32 [b] CUDA tensors.
    Returns a dict with any of 'path', 'durations' (int32 [b,t_x]), 'frame_tokens' (int32 [b,t_y], -1 past t_y)."""
    if not value.is_cuda:
        raise RuntimeError("maximum_path_lengths takes CUDA tensors (CPU tensors: maximum_path, or maximum_path_c with numpy arrays)")
    v = _prep_value(value)
    b, tx, ty = v.shape
    dtype = out_dtype or value.dtype
    if dtype not in _ONE:
        raise TypeError("unsupported output dtype %s" % dtype)
    esize, one = _ONE[dtype]
    with torch.cuda.device(v.device):
        xl = x_lengths.to(device=v.device, dtype=torch.int32).contiguous()
        yl = y_lengths.to(device=v.device, dtype=torch.int32).contiguous()
        out = {}
        path = torch.empty((b, tx, ty), dtype=dtype, device=v.device) if dense else None
        dur = torch.empty((b, tx), dtype=torch.int32, device=v.device) if return_durations else None
        ftok = torch.empty((b, ty), dtype=torch.int32, device=v.device) if return_frame_tokens else None
        if b > 0 and tx > 0 and ty > 0:
            stream = torch.cuda.current_stream(v.device).cuda_stream
            ws = _workspace(v.device, stream, b, tx, ty)
            _launch(v, (xl.data_ptr(), yl.data_ptr(), None, 0, 0, 0, 0,
                        path.data_ptr() if dense else None, esize, one, 1,
                        ftok.data_ptr() if ftok is not None else None, dur.data_ptr() if dur is not None else None, None,
                        b, tx, ty, max_neg_val, ws.data_ptr(), ws.numel(), stream))
        if dense:
            out["path"] = path
        if dur is not None:
            out["durations"] = dur
        if ftok is not None:
            out["frame_tokens"] = ftok
    return out


def check_status(device=None) -> int:
    """Synchronises and returns (then clears) the kernel status word for the current
    stream of `device`: bit 0 = some item had t_x > t_y or lengths outside the tensor
    (the reference reads out of bounds there; this implementation emits a zero path)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    stream = torch.cuda.current_stream(device).cuda_stream
    ws = _workspaces.get((device.index, stream))
    if ws is None:
        return 0
    with torch.cuda.device(device):
        rc = int(_lib.lib.alb200_mas_status(ws.data_ptr(), stream))
    if rc < 0:
        _lib.check(rc)
    return rc
