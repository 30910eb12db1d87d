"""tools/flaky_ota.py -- repeat the score kernels on shapes around the t_x = 256 boundary and report sporadic errors (developer aid)."""
import sys, numpy as np, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.neg_cent as nc
from aligner_b200 import _lib
from oracle import neg_cent as nc_oracle
g = torch.Generator(device="cuda").manual_seed(399)
for mode, c, tx, ty in [("ota", 80, 250, 1500), ("ota", 80, 300, 1500), ("ota", 80, 512, 1024), ("ota", 64, 300, 1500), ("gauss", 96, 250, 1024), ("gauss", 96, 300, 1024), ("gauss", 192, 200, 1000)]:
    b = 2
    q = torch.randn(b, c, ty, generator=g, device="cuda"); k = torch.randn(b, c, tx, generator=g, device="cuda"); lg = torch.rand(b, c, tx, generator=g, device="cuda") - 0.5
    if mode == "ota":
        want = nc_oracle.ota_log_prob(q.cpu().numpy(), k.cpu().numpy(), 0.0005); fn = lambda: nc.ota_log_prob(q, k, 0.0005)
    else:
        want = nc_oracle.gaussian_neg_cent(q.cpu().numpy(), k.cpu().numpy(), lg.cpu().numpy()); fn = lambda: nc.gaussian_neg_cent(q, k, lg)
    bad = 0; worst = 0; info = ""
    for it in range(40):
        if it % 3 == 0: junk = torch.randn(32, 1024, 1024, device="cuda")
        got = fn().cpu().numpy()
        d = np.abs(got - want); err = d.max() / np.abs(want).max()
        worst = max(worst, err)
        if err > 1e-5:
            bad += 1
            if not info:
                w = np.argwhere(d > 1e-5 * np.abs(want).max())
                info = "frames %s tokens %s" % (np.unique(w[:, 2])[[0, -1]], np.unique(w[:, 1])[[0, -1]])
    print("%-5s c=%d tx=%d ty=%d: bad %d/40 worst %.2e %s" % (mode, c, tx, ty, bad, worst, info), flush=True)
