"""tools/sanitize_case.py -- small MAS + neg_cent calls for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.monotonic_align as ma
import aligner_b200.neg_cent as nc
from aligner_b200 import _lib
rng = np.random.default_rng(0)
CASES = [(None, 5, 150, 260, torch.float32), ("2,32,2,1,1", 5, 150, 260, torch.float32), ("3,16,2,0,0", 5, 150, 260, torch.float32),
         ("2,32,2,0,1,2", 3, 300, 420, torch.float32),        # cluster of 2 CTAs per utterance
         ("1,32,3,0,1,4", 2, 500, 520, torch.float32),        # cluster of 4
         (None, 4, 150, 264, torch.bfloat16), (None, 200, 70, 136, torch.float16),     # native half-precision scores: skewed/TMA and throughput forms
         (None, 200, 70, 133, torch.float32),                 # throughput regime, unaligned rows
         ("4,32,0,-1,1,0,4", 4, 200, 1000, torch.float32),    # 4-frame lag, pre-skewed 3-D boxes: two warps, partly filled last warp
         ("2,32,0,-1,1,0,4", 3, 40, 96, torch.float32),       # ... one warp, tail lanes
         ("3,32,0,-1,1,0,4", 2, 300, 404, torch.float32),     # ... four warps
         (None, 6, 130, 140, torch.float32),                  # two-tile utterances: every tile requested before the lengths are known
         (None, 5, 300, 400, torch.float32), (None, 2, 600, 700, torch.float32)]     # shared zero fill: filler CTAs, filler clusters
for force, b, tx, ty, dt in CASES:
    _lib.set_option("force", force)
    v = torch.randn(b, tx, ty, device="cuda").to(dt)
    t_x = rng.integers(1, min(tx, ty) + 1, b).astype(np.int32); t_y = np.array([rng.integers(t_x[i], ty + 1) for i in range(b)], np.int32)
    out = ma.maximum_path_lengths(v, torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(), return_durations=True, return_frame_tokens=True)
    torch.cuda.synchronize()
    assert int(out["durations"].sum()) == int(t_y.sum())
_lib.set_option("force", None)
# second-generation score kernels (TMA / tcgen05, both accumulator modes), generated prior, fused pipelined entry
import aligner_b200.fused as fused
for (b, c, tx, ty) in [(3, 40, 150, 300), (2, 80, 300, 520)]:
    z2 = torch.randn(b, c, ty, device="cuda"); m2 = torch.randn(b, c, tx, device="cuda"); l2 = torch.rand(b, c, tx, device="cuda") - 0.5
    g2 = nc.gaussian_neg_cent(z2, m2, l2); o2 = nc.ota_log_prob(z2, m2, 0.0005)
    xl = torch.full((b,), tx - 3, dtype=torch.int32, device="cuda"); yl = torch.full((b,), ty - 5, dtype=torch.int32, device="cuda")
    o3 = nc.ota_log_prob(z2, m2, 0.0005, x_lengths=xl, y_lengths=yl, prior_scaling=1.0)
    _lib.set_option("fused_seq", "2")
    pth, sc = fused.gaussian_maximum_path(z2, m2, l2, x_lengths=xl, y_lengths=yl)
    _lib.set_option("fused_seq", None)
    torch.cuda.synchronize()
    assert torch.equal(sc, g2) and int(pth.sum()) == int(yl.sum())
z = torch.randn(2, 40, 300, device="cuda"); m = torch.randn(2, 40, 150, device="cuda"); lg = torch.rand(2, 40, 150, device="cuda") - 0.5
s = nc.gaussian_neg_cent(z, m, lg); q = nc.ota_log_prob(torch.randn(2, 80, 300, device="cuda"), torch.randn(2, 80, 150, device="cuda"))
torch.cuda.synchronize()
print("sanitize case done", float(s.sum()), float(q[torch.isfinite(q)].sum()))
