"""tools/lag_sweep.py -- one-warp and multi-warp shapes under forced 1-frame and 4-frame lag: time per 32-frame unit (developer aid)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.monotonic_align as ma
from aligner_b200 import _lib

def graph_time(fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3

if __name__ == "__main__":
    plans = sys.argv[1:] or ["64x64x1000:2", "64x64x2000:2", "64x128x1000:4", "64x128x2000:4", "64x96x1000:3", "64x96x2000:3", "64x192x1000:6", "64x192x2000:6",
                             "64x200x1000:4", "64x200x2000:4", "64x200x1000:2", "64x200x2000:2"]
    for cs in plans:
        shape, _, rr = cs.partition(":")
        b, tx, ty = (int(x) for x in shape.split("x"))
        R = int(rr)
        v = torch.randn(b, tx, ty, device="cuda")
        xl = torch.full((b,), tx, dtype=torch.int32, device="cuda"); yl = torch.full((b,), ty, dtype=torch.int32, device="cuda")
        dense = os.environ.get('DENSE', '1') == '1'
        run = lambda: ma.maximum_path_lengths(v, xl, yl, dense=dense, return_frame_tokens=not dense)
        res = []
        for lag in (1, 4):
            _lib.set_option("force", "%d,32,0,-1,1,0,%d" % (R, lag))
            try:
                res.append("lag%d %7.1f us" % (lag, graph_time(run)))
            except Exception as e:
                res.append("lag%d     n/a   " % lag)
        _lib.set_option("force", None)
        print("%-16s R=%d  %s" % (shape, R, "   ".join(res)), flush=True)
