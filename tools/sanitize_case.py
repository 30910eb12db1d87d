"""tools/sanitize_case.py -- small MAS + neg_cent calls for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.monotonic_align as ma
import aligner_b200.neg_cent as nc
rng = np.random.default_rng(0)
for force in (None, "2,32,2,1,1", "3,16,2,0,0"):
    if force: os.environ["ALB200_FORCE"] = force
    else: os.environ.pop("ALB200_FORCE", None)
    b, tx, ty = 5, 150, 260
    v = torch.randn(b, tx, ty, device="cuda")
    t_x = rng.integers(1, tx + 1, b).astype(np.int32); t_y = np.array([rng.integers(t_x[i], ty + 1) for i in range(b)], np.int32)
    out = ma.maximum_path_lengths(v, torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(), return_durations=True, return_frame_tokens=True)
    torch.cuda.synchronize()
    assert int(out["durations"].sum()) == int(t_y.sum())
os.environ.pop("ALB200_FORCE", None)
z = torch.randn(2, 40, 300, device="cuda"); m = torch.randn(2, 40, 150, device="cuda"); lg = torch.rand(2, 40, 150, device="cuda") - 0.5
s = nc.gaussian_neg_cent(z, m, lg); q = nc.ota_log_prob(torch.randn(2, 80, 300, device="cuda"), torch.randn(2, 80, 150, device="cuda"))
torch.cuda.synchronize()
print("sanitize case done", float(s.sum()), float(q[torch.isfinite(q)].sum()))
