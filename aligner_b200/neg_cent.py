"""Score matrices that feed ``monotonic_align.maximum_path`` -- CUDA (sm_100a) only.

The reference snapshot has no code for these: the functions mirror the
upstream expressions users of the reference write in their training step,

    Glow-TTS / VITS:   neg_cent = gaussian_neg_cent(z_p, m_p, logs_p)     # [b, t_text, t_mel]
                       attn = monotonic_align.maximum_path(neg_cent, attn_mask)
    OTA / NeMo:        logp = ota_log_prob(queries, keys, prior=attn_prior)
                       attn_hard = monotonic_align.maximum_path(logp, attn_mask)

PyTorch is used for device memory and the stream handle only.
"""
from __future__ import annotations

import torch

from . import _lib


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("aligner_b200 runs on sm_100a only: %s must be a CUDA tensor (no CPU fallback)" % name)
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _scratch(mode: int, b: int, c: int, tx: int, ty: int, device: torch.device):
    """Device scratch of the TMA / tcgen05 path (text-side operand staged once per utterance), from torch's caching
    allocator -- stream-ordered like every other torch tensor, so it is safe under CUDA-graph capture too."""
    need = int(_lib.lib.alb200_neg_cent_workspace_bytes(mode, b, c, tx, ty))
    if not need:
        return None, 0
    return torch.empty(need, dtype=torch.uint8, device=device), need


def gaussian_neg_cent(z: torch.Tensor, m_p: torch.Tensor, logs_p: torch.Tensor) -> torch.Tensor:
    """z [b,c,t_mel], m_p / logs_p [b,c,t_text]  ->  neg_cent [b,t_text,t_mel] fp32:
    sum_c log N(z[b,c,y]; m_p[b,c,x], exp(logs_p[b,c,x])^2)  (Glow-TTS logp1..4, VITS neg_cent1..4)."""
    z, m_p, logs_p = _f32c(z, "z"), _f32c(m_p, "m_p"), _f32c(logs_p, "logs_p")
    if z.dim() != 3 or m_p.shape != logs_p.shape or m_p.dim() != 3 or z.shape[:2] != m_p.shape[:2]:
        raise ValueError("expected z [b,c,t_y], m_p [b,c,t_x], logs_p [b,c,t_x]; got %s %s %s" % (tuple(z.shape), tuple(m_p.shape), tuple(logs_p.shape)))
    b, c, ty = z.shape
    tx = m_p.shape[2]
    with torch.cuda.device(z.device):
        out = torch.empty((b, tx, ty), dtype=torch.float32, device=z.device)
        if b and tx and ty:
            ws, need = _scratch(0, b, c, tx, ty, z.device)
            _lib.check(_lib.lib.alb200_neg_cent_gaussian_ws(z.data_ptr(), m_p.data_ptr(), logs_p.data_ptr(), out.data_ptr(), b, c, tx, ty,
                                                            ws.data_ptr() if ws is not None else None, need,
                                                            torch.cuda.current_stream(z.device).cuda_stream))
    return out


def beta_binomial_prior(x_lengths: torch.Tensor, y_lengths: torch.Tensor, tx: int, ty: int, scaling: float = 1.0) -> torch.Tensor:
    """Dense [b, t_x, t_y] beta-binomial alignment prior (OTA paper; oracle/neg_cent.py:beta_binomial_prior), zero outside each
    utterance's lengths.  Only needed on the paths that cannot generate it inside the kernel; computed on the device in fp64."""
    dev = x_lengths.device
    xl = x_lengths.to(torch.float64)[:, None, None]
    yl = y_lengths.to(torch.float64)[:, None, None]
    x = torch.arange(tx, device=dev, dtype=torch.float64)[None, :, None]
    y = torch.arange(ty, device=dev, dtype=torch.float64)[None, None, :]
    n, a, bq = xl - 1, scaling * (y + 1), scaling * (yl - y)
    valid = (x < xl) & (y < yl)
    lg = torch.lgamma
    safe = lambda t: torch.where(valid, t, torch.ones_like(t))
    lp = (lg(safe(n + 1)) - lg(safe(x + 1)) - lg(safe(n - x + 1)) + lg(safe(x + a)) + lg(safe(n - x + bq)) - lg(safe(n + a + bq))
          - (lg(safe(a)) + lg(safe(bq)) - lg(safe(a + bq))))
    return torch.where(valid, torch.exp(lp), torch.zeros_like(lp)).to(torch.float32)


def ota_log_prob(queries: torch.Tensor, keys: torch.Tensor, temperature: float = 0.0005, prior: torch.Tensor | None = None,
                 x_lengths: torch.Tensor | None = None, *, y_lengths: torch.Tensor | None = None,
                 prior_scaling: float | None = None) -> torch.Tensor:
    """queries [b,c,t_mel], keys [b,c,t_text]  ->  [b,t_text,t_mel] fp32:
    log_softmax over the text axis of -temperature * ||q - k||^2, plus log(prior + 1e-8) if given
    (OTA aligner, arXiv 2108.10447; NeMo AlignmentEncoder).

    prior_scaling=s (with x_lengths / y_lengths): the beta-binomial prior BetaBinom(x; t_x - 1, s (y + 1), s (t_y - y)) is
    generated inside the kernel instead of being read from a [b, t_x, t_y] tensor."""
    q, k = _f32c(queries, "queries"), _f32c(keys, "keys")
    if q.dim() != 3 or k.dim() != 3 or q.shape[:2] != k.shape[:2]:
        raise ValueError("expected queries [b,c,t_y], keys [b,c,t_x]; got %s %s" % (tuple(q.shape), tuple(k.shape)))
    b, c, ty = q.shape
    tx = k.shape[2]
    pr = None
    if prior is not None:
        pr = _f32c(prior, "prior")
        if tuple(pr.shape) != (b, tx, ty):
            raise ValueError("prior must be [b,t_x,t_y]")
    xl = None
    if x_lengths is not None:
        xl = x_lengths.to(device=q.device, dtype=torch.int32).contiguous()
    yl = None
    if y_lengths is not None:
        yl = y_lengths.to(device=q.device, dtype=torch.int32).contiguous()
    if prior_scaling is not None and prior is not None:
        raise ValueError("give either a prior tensor or prior_scaling, not both")
    with torch.cuda.device(q.device):
        out = torch.empty((b, tx, ty), dtype=torch.float32, device=q.device)
        if b and tx and ty and prior_scaling is not None:
            ws, need = _scratch(1, b, c, tx, ty, q.device)
            rc = _lib.lib.alb200_neg_cent_ota_bb(q.data_ptr(), k.data_ptr(), xl.data_ptr() if xl is not None else None,
                                                 yl.data_ptr() if yl is not None else None, float(prior_scaling), out.data_ptr(),
                                                 float(temperature), b, c, tx, ty, ws.data_ptr() if ws is not None else None, need,
                                                 torch.cuda.current_stream(q.device).cuda_stream)
            if rc == 0:
                return out
            if rc != _lib.E_UNSUPPORTED:
                _lib.check(rc)
            full = lambda t, n: t if t is not None else torch.full((b,), n, dtype=torch.int32, device=q.device)
            pr = beta_binomial_prior(full(xl, tx), full(yl, ty), tx, ty, float(prior_scaling))      # materialised for the other paths
        if b and tx and ty:
            ws, need = _scratch(1, b, c, tx, ty, q.device)
            _lib.check(_lib.lib.alb200_neg_cent_ota_ws(q.data_ptr(), k.data_ptr(), pr.data_ptr() if pr is not None else None,
                                                       xl.data_ptr() if xl is not None else None, out.data_ptr(), float(temperature),
                                                       b, c, tx, ty, ws.data_ptr() if ws is not None else None, need,
                                                       torch.cuda.current_stream(q.device).cuda_stream))
    return out
