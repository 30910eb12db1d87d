"""tools/run_one.py WORKLOAD [N] -- launch the public maximum_path N times on one BASELINE workload (for ncu)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.monotonic_align as ma
from aligner_b200 import _lib
W = {"c1": (16, 100, 800), "c2": (64, 200, 1000), "c3": (32, 300, 1500), "c4": (8, 1000, 6000), "c5b": (4096, 200, 1000), "c5a": (1024, 400, 2000)}
b, tx, ty = W[sys.argv[1]] if sys.argv[1] in W else tuple(int(x) for x in sys.argv[1].split("x"))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
g = torch.Generator(device="cuda").manual_seed(1)
vals = [torch.randn(b, tx, ty, generator=g, device="cuda") for _ in range(3)]
mask = torch.ones(b, tx, ty, device="cuda")
print(_lib.describe(b, tx, ty))
for i in range(n):
    out = ma.maximum_path(vals[i % 3], mask)
torch.cuda.synchronize()
print("done", float(out.sum()))
