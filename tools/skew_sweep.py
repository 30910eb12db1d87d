"""tools/skew_sweep.py -- latency regime: lock-step vs skewed forward over a grid of shapes (GPU-side tuning aid)."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from mas_sweep import bench

rows = []
for b in (16, 64, 148):
    for tx, ty in ((32, 200), (64, 400), (100, 800), (128, 640), (200, 1000), (256, 1280), (300, 1500), (400, 2000), (512, 2000), (640, 3200), (768, 3072), (1000, 6000)):
        if b > 16 and tx >= 640:
            continue
        out = {}
        for name, force in (("lockstep", "0,0,0,-1,0"), ("skewed", "0,0,0,-1,1")):
            r = bench(b, tx, ty, force, reps=15)
            out[name] = r
        ls, sk = out["lockstep"], out["skewed"]
        line = "b=%3d tx=%4d ty=%4d  lockstep %8.1f us  skewed %8s us   %s | %s" % (
            b, tx, ty, ls.get("us_med", -1), ("%8.1f" % sk["us_med"]) if "us_med" in sk else "n/a",
            ls.get("desc", ls.get("error", ""))[:60], sk.get("desc", sk.get("error", ""))[:70])
        print(line, flush=True)
        rows.append({"b": b, "tx": tx, "ty": ty, "lockstep": ls, "skewed": sk})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/skew_sweep.json", "w"), indent=1)
