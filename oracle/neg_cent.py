"""oracle/neg_cent.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

fp64 restatement of the two score matrices that feed monotonic alignment search.

PARITY UNPINNED.  The reference snapshot (xiaozhah/Aligner) contains no neg_cent code at
all: the MoBo/RoMo/OTA branches its README describes
(README.md:9-25) are not in the tree, and it vendors none of the upstream modules.  What is
restated here are the PUBLISHED formulas of the projects the README links to:

  gaussian_neg_cent   Glow-TTS  models.py, `logp1..logp4` (Kim et al., arXiv 2005.11129, eq. 5-6);
                      VITS models.py, `neg_cent1..4` (same expression; VITS orders the axes
                      [b, t_mel, t_text], this repository uses Glow-TTS order [b, t_text, t_mel],
                      as the reference API documents at monotonic_align/__init__.py:8).
  ota_log_prob        "One TTS Alignment To Rule Them All" (Badlani et al., arXiv 2108.10447,
                      linked at README.md:50), NeMo AlignmentEncoder: L2 distance between
                      projected text keys and mel queries, scaled by a temperature, log-softmax
                      over the text axis, plus log(prior + 1e-8).

Only tests/, bench.py and smoke() may import this module.
"""
from __future__ import annotations

import math

import numpy as np


def gaussian_neg_cent(z, m_p, logs_p):
    """z [b,c,t_y], m_p [b,c,t_x], logs_p [b,c,t_x]  ->  [b,t_x,t_y] float64.

    neg_cent[b,x,y] = sum_c log N(z[b,c,y]; m_p[b,c,x], exp(logs_p[b,c,x])^2), written as the
    four terms the upstream code uses:
        s2 = exp(-2 logs_p)
        n1 = sum_c(-0.5 log(2 pi) - logs_p)            [b,t_x,1]
        n2 = (-0.5 z^2)^T contracted with s2           [b,t_x,t_y]
        n3 = z^T contracted with (m_p s2)              [b,t_x,t_y]
        n4 = sum_c(-0.5 m_p^2 s2)                      [b,t_x,1]
    """
    z = np.asarray(z, np.float64); m_p = np.asarray(m_p, np.float64); logs_p = np.asarray(logs_p, np.float64)
    s2 = np.exp(-2.0 * logs_p)
    n1 = (-0.5 * math.log(2.0 * math.pi) - logs_p).sum(1)[:, :, None]
    n2 = np.einsum("bcx,bcy->bxy", s2, -0.5 * z * z)
    n3 = np.einsum("bcx,bcy->bxy", m_p * s2, z)
    n4 = (-0.5 * m_p * m_p * s2).sum(1)[:, :, None]
    return n1 + n2 + n3 + n4


def ota_log_prob(queries, keys, temperature=0.0005, prior=None, x_lengths=None):
    """queries [b,c,t_y] (mel side), keys [b,c,t_x] (text side)  ->  [b,t_x,t_y] float64.

    d[b,x,y]    = -temperature * sum_c (queries[b,c,y] - keys[b,c,x])^2
    logp        = log_softmax(d, over the text axis x)   (+ log(prior + 1e-8) when a prior is given)
    Text positions >= x_lengths[b] are excluded from the softmax (their output is -inf)."""
    q = np.asarray(queries, np.float64); k = np.asarray(keys, np.float64)
    b, c, ty = q.shape
    tx = k.shape[2]
    diff = q[:, :, None, :] - k[:, :, :, None]                 # [b,c,t_x,t_y]
    d = -float(temperature) * (diff * diff).sum(1)
    if x_lengths is not None:
        valid = np.arange(tx)[None, :, None] < np.asarray(x_lengths)[:, None, None]
        d = np.where(valid, d, -np.inf)
    mx = d.max(1, keepdims=True)
    lse = mx + np.log(np.exp(d - mx).sum(1, keepdims=True))
    out = d - lse
    if prior is not None:
        out = out + np.log(np.asarray(prior, np.float64) + 1e-8)
    return out


def beta_binomial_prior(t_x: int, t_y: int, scaling: float = 1.0):
    """Beta-binomial alignment prior of the OTA paper (NeMo beta_binomial_prior_distribution): [t_x, t_y]."""
    from scipy.stats import betabinom
    out = np.zeros((t_x, t_y))
    xs = np.arange(t_x)
    for y in range(1, t_y + 1):
        a, bb = scaling * y, scaling * (t_y + 1 - y)
        out[:, y - 1] = betabinom(t_x - 1, a, bb).pmf(xs)
    return out
