"""tools/lag4_check.py [BxTXxTY[:R] ...] -- the 4-frame-lag form against the default configuration: frame-token equality on random
scores with ragged lengths, and CUDA-graph timing of both (developer aid; ALB200_LIB picks the library)."""
import sys
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

cases = sys.argv[1:] or ["64x200x1000:4", "64x200x1000:2", "64x200x1000:5", "16x100x800:2", "32x300x1500:5", "32x300x1500:3", "8x96x300:3", "4x64x64:2", "3x128x128:4", "5x40x333:2", "2x512x2048:4"]
bad = 0
for cs in cases:
    shape, _, rr = cs.partition(":")
    b, tx, ty = (int(x) for x in shape.split("x"))
    R = int(rr) if rr else 4
    gen = torch.Generator(device="cuda").manual_seed(1234 + tx)
    v = torch.randn(b, tx, ty, generator=gen, device="cuda") * 3
    xl = torch.randint(max(1, tx // 2), tx + 1, (b,), generator=gen, device="cuda", dtype=torch.int32)
    yl = torch.randint(max(1, ty // 2), ty + 1, (b,), generator=gen, device="cuda", dtype=torch.int32)
    yl = torch.maximum(yl, xl); xl[0] = tx; yl[0] = ty
    if b > 1: xl[1] = min(tx, ty); yl[1] = min(tx, ty) if ty >= tx else ty
    def run():
        return ma.maximum_path_lengths(v, xl, yl, dense=True, return_frame_tokens=True, return_durations=True)
    _lib.set_option("force", None)
    d0 = _lib.describe(b, tx, ty)
    ref = run(); t0 = graph_time(run)
    _lib.set_option("force", "%d,32,0,-1,1,0,4" % R)
    try:
        d1 = _lib.describe(b, tx, ty)
        out = run(); t1 = graph_time(run)
    except Exception as e:
        print("%-16s R=%d: %s" % (shape, R, e)); _lib.set_option("force", None); continue
    _lib.set_option("force", None)
    ok = all(torch.equal(ref[k], out[k]) for k in ("path", "frame_tokens", "durations"))
    bad += 0 if ok else 1
    print("%-16s default %7.1f us [%s]\n%-16s lag4    %7.1f us [%s]  equal=%s" % (shape, t0, d0.split(" smem")[0], "", t1, d1.split(" smem")[0], ok), flush=True)
print("FAILED" if bad else "ALL EQUAL")
