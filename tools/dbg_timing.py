"""tools/dbg_timing.py -- run one MAS launch per workload with ALB200_DBG=1 (per-warp clock64 stamps on stderr)."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["ALB200_DBG"] = "1"
import aligner_b200.monotonic_align as ma
from aligner_b200 import _lib
cases = [(64, 64, 1000, "2,32,4,1,0"), (64, 64, 1000, "2,32,4,1,1"), (64, 128, 1000, "2,32,4,1,0"), (64, 128, 1000, "2,32,4,1,1"),
         (64, 200, 1000, "2,32,4,1,0"), (64, 200, 1000, "2,32,4,1,1"), (64, 200, 1000, "4,32,4,1,1")]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split("x")[:3]) + (a.split("x")[3] if len(a.split("x")) > 3 else None,) for a in sys.argv[1:]]
for (b, tx, ty, force) in cases:
    if force: os.environ["ALB200_FORCE"] = force
    else: os.environ.pop("ALB200_FORCE", None)
    v = torch.randn(b, tx, ty, device="cuda")
    xl = torch.full((b,), tx, dtype=torch.int32, device="cuda"); yl = torch.full((b,), ty, dtype=torch.int32, device="cuda")
    for rep in range(2):
        if rep == 1:
            print("== %dx%dx%d force=%s %s" % (b, tx, ty, force, _lib.describe(b, tx, ty)), file=sys.stderr, flush=True)
        os.environ["ALB200_DBG"] = "1" if rep == 1 else ""
        if rep == 0: os.environ.pop("ALB200_DBG")
        ma.maximum_path_lengths(v, xl, yl, dense=(os.environ.get('DENSE', '1') == '1'), return_frame_tokens=True)
        torch.cuda.synchronize()
