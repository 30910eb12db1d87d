"""tools/dbg_timing.py -- run one MAS launch per workload with per-warp clock64 stamps on stderr.
Needs a library built with -DALB200_DBG_BUILD=1 (python build_lib.py --dbg writes aligner_b200/libaligner_b200_dbg.so; it is picked
up here through ALB200_LIB).  The shipped library has the stamps compiled out."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("ALB200_LIB", str(Path(__file__).resolve().parent.parent / "aligner_b200" / "libaligner_b200_dbg.so"))
import aligner_b200.monotonic_align as ma
from aligner_b200 import _lib
cases = [(64, 64, 1000, "2,32,4,1,0"), (64, 64, 1000, "2,32,4,1,1"), (64, 128, 1000, "2,32,4,1,0"), (64, 128, 1000, "2,32,4,1,1"),
         (64, 200, 1000, "2,32,4,1,0"), (64, 200, 1000, "2,32,4,1,1"), (64, 200, 1000, "4,32,4,1,1")]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split("x")[:3]) + (a.split("x")[3] if len(a.split("x")) > 3 else None,) for a in sys.argv[1:]]
for (b, tx, ty, force) in cases:
    _lib.set_option("force", force)
    v = torch.randn(b, tx, ty, device="cuda")
    xl = torch.full((b,), tx, dtype=torch.int32, device="cuda"); yl = torch.full((b,), ty, dtype=torch.int32, device="cuda")
    for rep in range(2):
        if rep == 1:
            print("== %dx%dx%d force=%s %s" % (b, tx, ty, force, _lib.describe(b, tx, ty)), file=sys.stderr, flush=True)
        _lib.set_option("dbg", "1" if rep == 1 else None)
        if os.environ.get('MASK'):
            ma.maximum_path(v, torch.ones_like(v) if os.environ['MASK'] == 'f32' else torch.ones(v.shape, dtype=torch.bool, device='cuda'))
        else:
            ma.maximum_path_lengths(v, xl, yl, dense=(os.environ.get('DENSE', '1') == '1'), return_frame_tokens=True)
        torch.cuda.synchronize()
