"""tools/c5_sweep.py -- BASELINE config 5: mixed-length batches (t_x ~ U{50..400}, t_y ~ U{max(200,t_x)..2000}), B = 256..8192,
chunked so one chunk's scores + path stay under ~16 GB.  Prints cells/s (true cells = sum t_x*t_y), GB/s of algorithmic bytes
(4*sum t_x*t_y + 4*B*Tx*Ty) and the fraction of the measured HBM peak.  Writes gpurun_out/c5_sweep.json."""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import aligner_b200.monotonic_align as ma
from aligner_b200 import _lib
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
dev = torch.device("cuda")
rows = []
for B in (256, 512, 1024, 2048, 4096, 8192):
    rng = np.random.default_rng(1234 + 5)
    t_x = rng.integers(50, 401, B).astype(np.int32)
    t_y = np.array([rng.integers(max(200, t_x[i]), 2001) for i in range(B)], np.int32)
    tx, ty = 400, 2000
    chunk = min(B, 2048)                                   # 2048 x 400 x 2000 x 8 B = 13 GB
    nchunk = B // chunk
    g = torch.Generator(device=dev).manual_seed(B)
    vals = torch.randn(chunk, tx, ty, generator=g, device=dev)
    out = torch.empty(chunk, tx, ty, device=dev)
    times, cells, algo = [], 0.0, 0.0
    for c in range(nchunk):
        xl = torch.from_numpy(t_x[c * chunk:(c + 1) * chunk]).to(dev); yl = torch.from_numpy(t_y[c * chunk:(c + 1) * chunk]).to(dev)
        ws = ma._workspace(dev, torch.cuda.current_stream().cuda_stream, chunk, tx, ty)
        def launch():
            _lib.check(_lib.lib.alb200_mas_device(vals.data_ptr(), xl.data_ptr(), yl.data_ptr(), out.data_ptr(), 4, 0x3F800000, 1, None, None,
                                                  chunk, tx, ty, -1e9, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
        launch(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); launch(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        times.append(float(np.median(ts)))
        cc = float((t_x[c * chunk:(c + 1) * chunk].astype(np.int64) * t_y[c * chunk:(c + 1) * chunk]).sum())
        cells += cc; algo += 4 * cc + 4.0 * chunk * tx * ty
    ms = sum(times)
    r = {"B": B, "ms": ms, "cells_per_s": cells / (ms * 1e-3), "utt_per_s": B / (ms * 1e-3), "GBps": algo / (ms * 1e-3) / 1e9,
         "frac_of_hbm_peak": algo / (ms * 1e-3) / 1e9 / PEAK, "config": _lib.describe(chunk, tx, ty)}
    rows.append(r)
    print("B=%5d  %8.2f ms  %.3e cells/s  %9.0f utt/s  %7.1f GB/s  %.3f of %.0f GB/s   %s" % (B, ms, r["cells_per_s"], r["utt_per_s"], r["GBps"], r["frac_of_hbm_peak"], PEAK, r["config"]), flush=True)
    del vals, out; torch.cuda.empty_cache()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "c5_sweep.json").write_text(json.dumps(rows, indent=1))
