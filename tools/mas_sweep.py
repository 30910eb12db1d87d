"""tools/mas_sweep.py -- GPU-side tuning sweep: kernel time of the MAS launch for several
workloads and forced kernel shapes (ALB200_FORCE="R,TF,NS,bits_smem").  Prints a table and
writes gpurun_out/sweep.json.  Not part of the product or the tests."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import aligner_b200.monotonic_align as ma  # noqa: E402
from aligner_b200 import _lib  # noqa: E402

PEAK = 6549.1


def bench(b, tx, ty, force, ragged=False, reps=30, dense=True):
    dense = dense and os.environ.get("DENSE", "1") == "1"
    _lib.set_option("force", force)
    ma._ws_bytes.clear()                     # the workspace size depends on the forced configuration
    dev = torch.device("cuda")
    rng = np.random.default_rng(5)
    if ragged:
        t_x = rng.integers(50, tx + 1, b).astype(np.int32)
        t_y = np.array([rng.integers(max(200, t_x[i]), ty + 1) for i in range(b)], np.int32)
    else:
        t_x, t_y = np.full(b, tx, np.int32), np.full(b, ty, np.int32)
    cells = float((t_x.astype(np.int64) * t_y).sum())
    padded = float(b) * tx * ty
    per_set = padded * 8
    nsets = int(min(max(2, np.ceil(400e6 / per_set) + 1), 8))
    g = torch.Generator(device=dev).manual_seed(1)
    vd = {"f32": (torch.float32, _lib.F32, 4), "f16": (torch.float16, _lib.F16, 2), "bf16": (torch.bfloat16, _lib.BF16, 2)}[os.environ.get("VDTYPE", "f32")]
    vits = os.environ.get("VITS", "0") == "1"            # scores and path stored [b, t_mel, t_text]
    vals = [torch.randn(*((b, ty, tx) if vits else (b, tx, ty)), generator=g, device=dev).to(vd[0]) for _ in range(nsets)]
    outs = [torch.empty(b, tx, ty, device=dev) for _ in range(nsets)]
    xl, yl = torch.from_numpy(t_x).to(dev), torch.from_numpy(t_y).to(dev)
    ws = ma._workspace(dev, torch.cuda.current_stream().cuda_stream, b, tx, ty)
    stream = torch.cuda.current_stream().cuda_stream
    try:
        desc = _lib.describe(b, tx, ty)
    except Exception as e:
        return {"error": str(e)}

    def launch(i):
        _lib.check(_lib.lib.alb200_mas_device_ex(vals[i % nsets].data_ptr(), vd[1] | (_lib.LAYOUT_VITS if vits else 0), xl.data_ptr(), yl.data_ptr(), None, 0, 0, 0, 0,
                                                 outs[i % nsets].data_ptr() if dense else None, 4, 0x3F800000, 1, None, None, None,
                                                 b, tx, ty, -1e9, ws.data_ptr(), ws.numel(), stream))
    for i in range(3):
        launch(i)
    torch.cuda.synchronize()
    ev = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(i); e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b_) for a, b_ in ev]) * 1e3   # us
    med = float(np.median(t))
    algo = vd[2] * cells + 4 * padded
    return {"us_med": med, "us_min": float(t.min()), "cells_per_s": cells / (med * 1e-6), "GBps": algo / (med * 1e-6) / 1e9,
            "frac": algo / (med * 1e-6) / 1e9 / PEAK, "desc": desc}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    rows = []
    plans = {
        "c1": (16, 100, 800, False, [None, "2,32,4,1,0", "2,32,4,1,1", "1,32,4,1,1", "1,32,8,1,1", "1,32,4,1,0", "3,32,4,1,1", "4,32,4,1,1"]),
        "c2": (64, 200, 1000, False, [None, "2,32,4,1,0", "2,32,4,1,1", "2,32,5,1,1", "1,32,4,1,1", "1,32,3,1,1", "1,32,4,0,1,2", "1,32,3,0,1,2",
                                      "2,32,4,0,1,2"]),
        "c3": (32, 300, 1500, False, [None, "3,32,3,0,1", "2,32,4,0,1,2", "2,32,3,0,1,2", "2,32,6,0,1,2", "3,32,3,0,1,2"]),
        "c4": (8, 1000, 6000, False, [None, "8,16,2,0,0,1", "2,32,3,0,1,4", "2,32,2,0,1,4", "3,32,2,0,1,3", "4,32,2,0,1,2", "4,32,3,0,1,2", "1,32,4,0,1,8"]),
        "c4b": (16, 768, 3072, False, [None, "6,16,4,0,0,1", "2,32,3,0,1,3", "3,32,2,0,1,2", "4,32,2,0,1,2"]),
        "c4c": (32, 640, 3200, False, [None, "6,16,4,0,0,1", "2,32,3,0,1,3", "4,32,2,0,1,2"]),
        "m1": (300, 200, 1000, False, [None, "0,0,0,-1,1"]),
        "m2": (500, 200, 1000, False, [None, "0,0,0,-1,1"]),
        "m3": (300, 100, 800, False, [None, "0,0,0,-1,1"]),
        "m4": (600, 400, 2000, True, [None, "0,0,0,-1,1"]),
        "m5": (200, 300, 1500, True, [None, "0,0,0,-1,1"]),
        "m6": (700, 200, 1000, True, [None, "0,0,0,-1,1"]),
        "x1": (900, 200, 1000, True, [None]), "x2": (1200, 200, 1000, True, [None]), "x3": (1600, 200, 1000, True, [None]), "x4": (2400, 200, 1000, True, [None]),
        "x5": (900, 400, 2000, True, [None]), "x6": (1200, 400, 2000, True, [None]),
        "x7": (900, 100, 800, False, [None]), "x8": (1600, 100, 800, False, [None]), "x9": (3000, 100, 800, False, [None]),
        "y1": (4096, 100, 800, False, [None, "2,32,2,1,1", "2,32,3,1,1", "2,32,4,1,1", "2,32,2,0,1", "1,32,3,1,1", "1,32,4,1,1"]),
        "y2": (2048, 400, 2000, True, [None, "4,32,3,0,1", "4,32,2,0,1", "3,32,3,0,1"]),
        "z1": (4096, 200, 300, False, [None]), "z2": (4096, 160, 500, True, [None]), "z3": (2048, 256, 2000, True, [None]), "z4": (3000, 130, 400, True, [None]),
        "q1": (500, 160, 500, True, [None]), "q2": (400, 100, 300, False, [None]), "q3": (600, 300, 500, True, [None]), "q4": (1000, 200, 640, True, [None]),
        "r1": (600, 200, 300, False, [None]), "r2": (1000, 300, 500, True, [None]), "r3": (1500, 160, 500, True, [None]), "r4": (300, 100, 400, False, [None]),
        "r5": (2000, 300, 500, True, [None]), "r6": (450, 400, 560, True, [None]), "r7": (800, 120, 360, False, [None]),
        "s1": (1000, 300, 1500, True, [None]), "s2": (2000, 300, 1500, True, [None]), "s3": (4096, 300, 1000, False, [None]), "s4": (1000, 384, 1200, True, [None]),
        "s5": (3000, 320, 800, True, [None]), "s6": (1200, 512, 1500, True, [None]),
        "c5s": (256, 400, 2000, True, [None, "4,32,2,0,1", "4,32,3,0,1", "4,32,2,0,0"]),
        "c5m": (512, 400, 2000, True, [None, "4,32,2,0,1", "4,32,3,0,1"]),
        "c5l": (1024, 400, 2000, True, [None, "4,32,3,0,1"]),
        "c5a": (2048, 400, 2000, True, [None, "4,16,3,0,0", "8,16,2,0,0", "8,16,3,0,0", "8,8,3,0,0", "8,8,4,0,0", "6,16,2,0,0", "8,32,2,0,0"]),
        "c5b": (4096, 200, 1000, False, [None, "4,16,2,0,0", "4,16,3,0,0", "4,32,2,0,0", "4,32,2,1,0", "6,16,2,0,0", "8,16,2,0,0", "3,16,2,0,0"]),
        "c5c": (4096, 100, 800, False, [None, "2,16,2,0,0", "2,16,3,0,0", "2,32,2,0,0", "4,16,2,0,0", "4,32,2,1,0", "1,32,2,1,0"]),
    }
    for name, (b, tx, ty, ragged, forces) in plans.items():
        if which != "all" and which != name:
            continue
        for f in forces:
            r = bench(b, tx, ty, f, ragged)
            r.update({"workload": name, "force": f})
            rows.append(r)
            if "error" in r:
                print("%-4s %-12s ERROR %s" % (name, f, r["error"]), flush=True)
            else:
                print("%-4s %-12s %9.1f us (min %9.1f)  %6.1f GB/s  frac %.3f  %s" % (name, f, r["us_med"], r["us_min"], r["GBps"], r["frac"], r["desc"]), flush=True)
        torch.cuda.empty_cache()
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "sweep.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
