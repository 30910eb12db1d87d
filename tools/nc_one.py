"""tools/nc_one.py gauss|ota [N] -- launch one score kernel N times on its BASELINE shape (for ncu)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.neg_cent as nc
which = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator(device="cuda").manual_seed(0)
if which == "gauss":
    b, c, tx, ty = 64, 192, 200, 1000
    z = torch.randn(b, c, ty, generator=g, device="cuda"); m = torch.randn(b, c, tx, generator=g, device="cuda"); logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
    for _ in range(n): out = nc.gaussian_neg_cent(z, m, logs)
else:
    b, c, tx, ty = 32, 80, 300, 1500
    q = torch.randn(b, c, ty, generator=g, device="cuda"); k = torch.randn(b, c, tx, generator=g, device="cuda")
    for _ in range(n): out = nc.ota_log_prob(q, k)
torch.cuda.synchronize()
print("done", float(out[torch.isfinite(out)].sum()))
