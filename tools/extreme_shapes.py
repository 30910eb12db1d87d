"""tools/extreme_shapes.py -- ad-hoc parity check of unusual shapes against the C oracle (GPU box)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import aligner_b200.monotonic_align as ma
from aligner_b200 import _lib
from oracle import mas as oracle
rng = np.random.default_rng(3)
for (b, tx, ty) in [(1, 2048, 2304), (1, 3000, 3200), (2, 1, 5000), (3, 513, 516), (2, 1024, 1024), (150, 513, 600), (1, 4, 20000), (5, 33, 33), (2, 2047, 2050)]:
    v = rng.standard_normal((b, tx, ty)).astype(np.float32)
    t_x = rng.integers(1, tx + 1, b).astype(np.int32); t_x[0] = tx
    t_y = np.array([rng.integers(t_x[i], ty + 1) for i in range(b)], np.int32); t_y[0] = ty
    want = np.zeros(v.shape, np.int32)
    oracle.maximum_path_c_port(want, v.copy(), t_x, t_y, omp=True)
    out = ma.maximum_path_lengths(torch.from_numpy(v).cuda(), torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(), out_dtype=torch.int32, return_durations=True)
    torch.cuda.synchronize()
    ok = np.array_equal(out["path"].cpu().numpy(), want) and np.array_equal(out["durations"].cpu().numpy(), want.sum(-1))
    print("%-18s %s   %s" % ((b, tx, ty), "OK " if ok else "MISMATCH", _lib.describe(b, tx, ty)), flush=True)
    assert ok
print("all extreme shapes ok")
