import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    """tests/golden/mas_golden.npz -> {case: {value, mask, path, path_dtype}} (made by oracle/make_golden.py
    from the unmodified reference)."""
    blob = np.load(ROOT / "tests" / "golden" / "mas_golden.npz")
    out = {}
    for key in blob.files:
        case, field = key.split("/")
        out.setdefault(case, {})[field] = blob[key]
    return out


def prefix_mask_np(x_len, y_len, tx, ty, dtype=np.float32):
    m = np.zeros((len(x_len), tx, ty), dtype)
    for i, (a, b) in enumerate(zip(x_len, y_len)):
        m[i, :a, :b] = 1
    return m


def random_lengths(rng, b, tx, ty, full=False):
    if full:
        return np.full(b, tx, np.int32), np.full(b, ty, np.int32)
    t_x = rng.integers(1, tx + 1, b).astype(np.int32)
    t_y = np.array([rng.integers(t_x[i], ty + 1) for i in range(b)], np.int32)
    return t_x, t_y


def make_values(rng, kind, shape):
    if kind == "gauss":
        return rng.standard_normal(shape).astype(np.float32)
    if kind == "ties":
        return rng.integers(-2, 3, shape).astype(np.float32)
    if kind == "sentinel":
        return (rng.standard_normal(shape) * 3e8).astype(np.float32)
    if kind == "negative":
        return (-np.abs(rng.standard_normal(shape)) * 40).astype(np.float32)
    if kind == "zeros":
        return np.zeros(shape, np.float32)
    raise ValueError(kind)
