"""tools/dbg_timing.py -- run one MAS launch per workload with ALB200_DBG=1 (per-warp clock64 stamps on stderr)."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["ALB200_DBG"] = "1"
import aligner_b200.monotonic_align as ma
from aligner_b200 import _lib
for (b, tx, ty, force) in [(64, 200, 1000, "2,32,4,1,0"), (64, 200, 1000, None), (64, 200, 1000, "4,32,4,1,1"), (32, 300, 1500, None), (8, 1000, 6000, None)]:
    if force: os.environ["ALB200_FORCE"] = force
    else: os.environ.pop("ALB200_FORCE", None)
    v = torch.randn(b, tx, ty, device="cuda")
    xl = torch.full((b,), tx, dtype=torch.int32, device="cuda"); yl = torch.full((b,), ty, dtype=torch.int32, device="cuda")
    for rep in range(2):
        print("== %dx%dx%d force=%s rep%d %s" % (b, tx, ty, force, rep, _lib.describe(b, tx, ty)), file=sys.stderr, flush=True)
        ma.maximum_path_lengths(v, xl, yl)
        torch.cuda.synchronize()
