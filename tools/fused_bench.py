"""tools/fused_bench.py -- C2 training step: separate score + search calls vs the fused (pipelined) entry, CUDA-graph timed."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aligner_b200.neg_cent as nc, aligner_b200.monotonic_align as ma, aligner_b200.fused as fused
def timeit(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side): fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(k): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k * 1e3
for (b, c, tx, ty) in [(64, 192, 200, 1000), (16, 192, 100, 800), (32, 192, 300, 1500)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn(b, c, ty, generator=g, device="cuda"); m = torch.randn(b, c, tx, generator=g, device="cuda"); logs = torch.rand(b, c, tx, generator=g, device="cuda") * 1.5 - 1.0
    xl = torch.full((b,), tx, dtype=torch.int32, device="cuda"); yl = torch.full((b,), ty, dtype=torch.int32, device="cuda")
    keep = [None] * 4; i = [0]
    def sep():
        keep[i[0] % 4] = ma.maximum_path_lengths(nc.gaussian_neg_cent(z, m, logs), xl, yl)["path"]; i[0] += 1
    def fus():
        keep[i[0] % 4] = fused.gaussian_maximum_path(z, m, logs, x_lengths=xl, y_lengths=yl); i[0] += 1
    def only_nc():
        keep[i[0] % 4] = nc.gaussian_neg_cent(z, m, logs); i[0] += 1
    print("%dx%dx%dx%d: score %.1f us, score + search separate %.1f us, fused %.1f us" % (b, c, tx, ty, timeit(only_nc), timeit(sep), timeit(fus)), flush=True)
