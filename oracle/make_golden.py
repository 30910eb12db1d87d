"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Regenerates ``tests/golden/mas_golden.npz`` by running the UNMODIFIED reference
(its own ``monotonic_align/__init__.py`` bound to ``core.pyx`` re-cythonized into
``oracle/_ref``, see oracle/build.py) on seeded inputs.  Needs /root/reference,
so it only runs in the build container; the fixture it writes is committed and
is what travels to the GPU box.

The reference has no tests or golden vectors of its own;
these fixtures are therefore "outputs of the reference itself run here".

Run:  python -m oracle.make_golden
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import mas

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "mas_golden.npz"


def prefix_mask(x_len, y_len, tx, ty, dtype=torch.float32):
    m = torch.zeros(len(x_len), tx, ty, dtype=dtype)
    for i, (a, b) in enumerate(zip(x_len, y_len)):
        m[i, :a, :b] = 1
    return m


def cases():
    """name -> (value tensor, mask tensor).  Small on purpose (fixture < 700 KB)."""
    out = {}
    # SURVEY.md 8c known answers (i)-(v)
    out["ka_zeros_4x8"] = (torch.zeros(1, 4, 8), torch.ones(1, 4, 8))
    g = torch.Generator().manual_seed(11)
    out["ka_square_4x4"] = (torch.randn(1, 4, 4, generator=g), torch.ones(1, 4, 4))
    out["ka_single_token_1x6"] = (torch.randn(1, 1, 6, generator=g), torch.ones(1, 1, 6))
    out["ka_sentinel_3x6"] = (torch.full((1, 3, 6), -5e8), torch.ones(1, 3, 6))
    torch.manual_seed(0)
    out["ka_randn_3x5x9"] = (torch.randn(3, 5, 9), prefix_mask([5, 3, 4], [9, 6, 4], 5, 9))
    # differential families: gaussian, dense ties, sentinel crossing, realistic
    g = torch.Generator().manual_seed(1234)
    x_len = [37, 1, 20, 48]
    y_len = [100, 9, 20, 77]
    m = prefix_mask(x_len, y_len, 48, 100)
    out["diff_gauss"] = (torch.randn(4, 48, 100, generator=g), m)
    out["diff_ties"] = (torch.randint(-2, 3, (4, 48, 100), generator=g).float(), m)
    out["diff_sentinel"] = (torch.randn(4, 48, 100, generator=g) * 3e8, m)
    z = torch.randn(4, 8, 100, generator=g)
    mu = torch.randn(4, 8, 48, generator=g)
    logs = torch.rand(4, 8, 48, generator=g) * 1.5 - 1.0
    s2 = torch.exp(-2 * logs)
    nc = ((-0.5 * 1.8378770664093453 - logs).sum(1)[:, :, None]
          + torch.einsum("bcx,bcy->bxy", s2, -0.5 * z * z)
          + torch.einsum("bcx,bcy->bxy", mu * s2, z)
          + (-0.5 * mu * mu * s2).sum(1)[:, :, None])
    out["diff_realistic"] = (nc, m)
    # all-zero mask item -> all-zero path
    m0 = prefix_mask([4, 0, 3], [7, 0, 12], 6, 12)
    out["edge_empty_item"] = (torch.randn(3, 6, 12, generator=g), m0)
    # longer than one 32-frame word and wider than one warp of rows
    out["wide_130x200"] = (torch.randn(2, 130, 200, generator=g), prefix_mask([130, 97], [200, 171], 130, 200))
    # API dtype rules: result dtype = result_type(value, mask)
    out["api_f64_value"] = (torch.randn(2, 7, 19, generator=g, dtype=torch.float64), prefix_mask([7, 4], [19, 11], 7, 19))
    out["api_f16_value"] = (torch.randn(2, 7, 19, generator=g).half(), prefix_mask([7, 4], [19, 11], 7, 19, torch.float16))
    out["api_bool_mask"] = (torch.randn(2, 7, 19, generator=g), prefix_mask([7, 4], [19, 11], 7, 19).bool())
    out["api_int_mask"] = (torch.randn(2, 7, 19, generator=g), prefix_mask([7, 4], [19, 11], 7, 19).to(torch.int64))
    return out


def main() -> None:
    api = mas.load_reference_api("serial")
    if api is None:
        raise SystemExit("reference not available (needs /root/reference)")
    blob = {}
    for name, (value, mask) in cases().items():
        v0 = value.clone()
        path = api.maximum_path(value, mask)
        assert torch.equal(v0, value), "reference mutated its input?"
        blob[name + "/value"] = value.numpy()
        blob[name + "/mask"] = mask.numpy()
        blob[name + "/path"] = path.numpy()
        blob[name + "/path_dtype"] = np.array(str(path.dtype))
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, OUT.stat().st_size, "bytes,", len(blob) // 4, "cases")


if __name__ == "__main__":
    main()
