"""tools/mask_vs_lengths.py [BxTXxTY] -- maximum_path(value, mask) against maximum_path_lengths(value, t_x, t_y), CUDA-graph timing."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.monotonic_align as ma
from lag_sweep import graph_time
for cs in sys.argv[1:] or ["64x200x1000", "16x100x800", "32x300x1500"]:
    b, tx, ty = (int(x) for x in cs.split("x"))
    vs = [torch.randn(b, tx, ty, device="cuda") for _ in range(6)]          # > L2 in total for the bench shapes
    masks = {"fp32 mask": torch.ones(b, tx, ty, device="cuda"), "bool mask": torch.ones(b, tx, ty, device="cuda", dtype=torch.bool)}
    xl = torch.full((b,), tx, dtype=torch.int32, device="cuda"); yl = torch.full((b,), ty, dtype=torch.int32, device="cuda")
    i = [0]
    def nxt():
        i[0] = (i[0] + 1) % len(vs); return vs[i[0]]
    out = ["%-14s" % cs, "lengths %6.1f us" % graph_time(lambda: ma.maximum_path_lengths(nxt(), xl, yl))]
    for name, m in masks.items():
        out.append("%s %6.1f us" % (name, graph_time(lambda: ma.maximum_path(nxt(), m))))
    print("   ".join(out), flush=True)
