"""tools/nc_bench.py -- time the score-matrix kernels on the BASELINE shapes (C2 gaussian, C3 ota)."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.neg_cent as nc

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3

g = torch.Generator(device="cuda").manual_seed(0)
b, c, tx, ty = 64, 192, 200, 1000
z = torch.randn(b, c, ty, generator=g, device="cuda"); m = torch.randn(b, c, tx, generator=g, device="cuda"); logs = torch.rand(b, c, tx, generator=g, device="cuda") - 0.5
us = timeit(lambda: nc.gaussian_neg_cent(z, m, logs))
fl = 2.0 * 2 * c * b * tx * ty
print("gaussian C2 %dx%dx%dx%d: %.1f us  %.2f TFLOP/s (algorithmic 4C flop/cell), out write %.1f GB/s" % (b, c, tx, ty, us, fl / us / 1e6, 4.0 * b * tx * ty / us / 1e3))
b, c, tx, ty = 32, 80, 300, 1500
q = torch.randn(b, c, ty, generator=g, device="cuda"); k = torch.randn(b, c, tx, generator=g, device="cuda")
us = timeit(lambda: nc.ota_log_prob(q, k))
print("ota C3 %dx%dx%dx%d: %.1f us  %.2f TFLOP/s (3C flop/cell), out write %.1f GB/s" % (b, c, tx, ty, us, 3.0 * c * b * tx * ty / us / 1e6, 4.0 * b * tx * ty / us / 1e3))
