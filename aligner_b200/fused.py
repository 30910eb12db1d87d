"""Fused score + search: what a Glow-TTS / VITS training step writes as

    neg_cent = gaussian_neg_cent(z_p, m_p, logs_p)                 # [b, t_text, t_mel]
    attn = monotonic_align.maximum_path(neg_cent, attn_mask)       # reference: monotonic_align/__init__.py:6-21

in one call.  For training-step sized batches (<= one utterance per SM) the score kernel and the search run concurrently:
the search consumes 128-frame tiles out of L2 while later tiles are still being computed (include/aligner_b200.h,
alb200_gaussian_mas_fused).  Results are bit-identical to the two separate calls.  CUDA (sm_100a) only.
"""
from __future__ import annotations

import torch

from . import _lib
from .monotonic_align import _MASK_DTYPE, _ONE

__all__ = ["gaussian_maximum_path"]

_workspaces: dict = {}


def _workspace(device: torch.device, stream: int, need: int) -> torch.Tensor:
    key = (device.index, stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)       # zero on first hand-over (work counters of the search)
        _workspaces[key] = ws
    return ws


def gaussian_maximum_path(z: torch.Tensor, m_p: torch.Tensor, logs_p: torch.Tensor, mask: torch.Tensor | None = None, *,
                          x_lengths: torch.Tensor | None = None, y_lengths: torch.Tensor | None = None,
                          out_dtype: torch.dtype | None = None, return_durations: bool = False):
    """z [b,c,t_mel], m_p / logs_p [b,c,t_text]; lengths either as the reference's mask [b,t_text,t_mel] (outer product of the
    prefix masks) or as int32 x_lengths / y_lengths.  Returns (path, neg_cent) -- or (path, neg_cent, durations).

    path has the dtype rule of the reference API: ``torch.result_type(neg_cent, mask)`` (fp32 when lengths are given)."""
    for name, t in (("z", z), ("m_p", m_p), ("logs_p", logs_p)):
        if not t.is_cuda:
            raise RuntimeError("aligner_b200 runs on sm_100a only: %s must be a CUDA tensor (no CPU fallback)" % name)
    z, m_p, logs_p = (t.detach().float().contiguous() for t in (z, m_p, logs_p))
    if z.dim() != 3 or m_p.shape != logs_p.shape or m_p.dim() != 3 or z.shape[:2] != m_p.shape[:2]:
        raise ValueError("expected z [b,c,t_y], m_p [b,c,t_x], logs_p [b,c,t_x]")
    b, c, ty = z.shape
    tx = m_p.shape[2]
    if (mask is None) == (x_lengths is None or y_lengths is None):
        raise ValueError("give either mask or both x_lengths and y_lengths")
    dtype = out_dtype or (torch.result_type(z, mask) if mask is not None else torch.float32)
    if dtype not in _ONE:
        raise TypeError("unsupported path dtype %s" % dtype)
    esize, one = _ONE[dtype]
    dev = z.device
    with torch.cuda.device(dev):
        neg_cent = torch.empty((b, tx, ty), dtype=torch.float32, device=dev)
        path = torch.empty((b, tx, ty), dtype=dtype, device=dev)
        dur = torch.empty((b, tx), dtype=torch.int32, device=dev) if return_durations else None
        if b and tx and ty:
            stream = torch.cuda.current_stream(dev).cuda_stream
            ws = _workspace(dev, stream, int(_lib.lib.alb200_fused_workspace_bytes(b, c, tx, ty)))
            if mask is not None:
                if tuple(mask.shape) != (b, tx, ty) or mask.dtype not in _MASK_DTYPE:
                    raise ValueError("mask must be [b, t_x, t_y] of a supported dtype")
                m = mask.detach()
                sb, sx, sy = m.stride()
                len_args = (None, None, m.data_ptr(), _MASK_DTYPE[m.dtype], sb, sx, sy)
            else:
                xl = x_lengths.to(device=dev, dtype=torch.int32).contiguous()
                yl = y_lengths.to(device=dev, dtype=torch.int32).contiguous()
                len_args = (xl.data_ptr(), yl.data_ptr(), None, 0, 0, 0, 0)
            _lib.check(_lib.lib.alb200_gaussian_mas_fused(z.data_ptr(), m_p.data_ptr(), logs_p.data_ptr(), neg_cent.data_ptr(), *len_args,
                                                          path.data_ptr(), esize, one, 1, None, dur.data_ptr() if dur is not None else None,
                                                          b, c, tx, ty, -1e9, ws.data_ptr(), ws.numel(), stream))
        else:
            path.zero_()
    return (path, neg_cent, dur) if return_durations else (path, neg_cent)
