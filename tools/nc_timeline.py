"""tools/nc_timeline.py gauss|ota -- clock64 timeline of CTA 0's roles in the TMA/tcgen05 score kernel (stderr).
Needs `python build_lib.py --dbg` (libaligner_b200_dbg.so)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("ALB200_LIB", str(ROOT / "aligner_b200" / "libaligner_b200_dbg.so"))
sys.path.insert(0, str(ROOT))
import torch
import aligner_b200.neg_cent as nc
from aligner_b200 import _lib
which = sys.argv[1] if len(sys.argv) > 1 else "gauss"
g = torch.Generator(device="cuda").manual_seed(0)
if which == "gauss":
    b, c, tx, ty = 64, 192, 200, 1000
    args = (torch.randn(b, c, ty, generator=g, device="cuda"), torch.randn(b, c, tx, generator=g, device="cuda"), torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0)
    fn = nc.gaussian_neg_cent
else:
    b, c, tx, ty = 32, 80, 300, 1500
    args = (torch.randn(b, c, ty, generator=g, device="cuda"), torch.randn(b, c, tx, generator=g, device="cuda"))
    fn = nc.ota_log_prob
for _ in range(3): fn(*args)
torch.cuda.synchronize()
_lib.set_option("dbg", "1")
fn(*args)
torch.cuda.synchronize()
