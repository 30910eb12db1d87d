import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("ALB200_LIB", str(ROOT / "aligner_b200" / "libaligner_b200_dbg.so"))
sys.path.insert(0, str(ROOT))
import torch
import aligner_b200.fused as fused
from aligner_b200 import _lib
b, c, tx, ty = 64, 192, 200, 1000
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn(b, c, ty, generator=g, device="cuda"); m = torch.randn(b, c, tx, generator=g, device="cuda"); logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
xl = torch.full((b,), tx, dtype=torch.int32, device="cuda"); yl = torch.full((b,), ty, dtype=torch.int32, device="cuda")
for _ in range(3): fused.gaussian_maximum_path(z, m, logs, x_lengths=xl, y_lengths=yl)
torch.cuda.synchronize()
_lib.set_option("dbg", "2")
for _ in range(3): fused.gaussian_maximum_path(z, m, logs, x_lengths=xl, y_lengths=yl)
