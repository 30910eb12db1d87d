"""CPU: the C-ABI library loads, exports every symbol include/aligner_b200.h declares, fails loudly
without a device, and the product package never touches the oracle."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib_path():
    import build_lib
    return build_lib.build()


def _declared():
    text = (ROOT / "include" / "aligner_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(alb200_\w+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    names = _declared()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    from aligner_b200 import _lib
    assert sorted(_lib.SYMBOLS) == names, "aligner_b200/_lib.py SYMBOLS out of sync with the header"


def test_nm_shows_no_oracle_symbols(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib_path)], capture_output=True, text=True).stdout
    assert "mas_oracle" not in out
    assert "alb200_mas_device" in out


def test_product_does_not_import_oracle():
    for py in (ROOT / "aligner_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), py
    for cu in (ROOT / "aligner_b200" / "csrc").glob("*.cu*"):
        assert not re.search(r"#\s*include[^\n]*oracle", cu.read_text()), cu   # comments may cite the oracle, code may not include it


def test_sass_uses_bulk_copy_engine(lib_path):
    """UBLKCP = cp.async.bulk (TMA engine) must be in the shipped SASS (B200_PROFILING.md)."""
    out = subprocess.run(["cuobjdump", "-sass", str(lib_path)], capture_output=True, text=True).stdout
    assert "UBLKCP" in out           # bulk shared->global zero fill
    assert "LDGSTS.E.BYPASS.128" in out  # 128-bit async tile loads (lock-step form)
    assert "UTMALDG.2D" in out       # 2-D TMA box loads (skewed form)
    assert "UTMALDG.3D" in out       # pre-skewed 3-D boxes (4-frame-lag form) and the score kernel's operand boxes
    assert "SYNCS" in out          # mbarrier ops
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", str(lib_path)], capture_output=True, text=True).stdout


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aligner_b200 import _lib
    from aligner_b200.monotonic_align.monotonic_align.core import maximum_path_c
    with pytest.raises(_lib.AlignerB200Error) as ei:
        maximum_path_c(np.zeros((1, 2, 3), np.int32), np.zeros((1, 2, 3), np.float32),
                       np.array([2], np.int32), np.array([3], np.int32))
    assert ei.value.code == _lib.E_NO_DEVICE
    import aligner_b200.monotonic_align as ma
    with pytest.raises(_lib.AlignerB200Error) as ei:          # CPU tensors are staged to the GPU: without one this fails loudly too
        ma.maximum_path(torch.zeros(1, 2, 3), torch.ones(1, 2, 3))
    assert ei.value.code == _lib.E_NO_DEVICE


def test_options_are_validated():
    from aligner_b200 import _lib
    _lib.set_option("force", "2,32,3,1,1")
    _lib.set_option("force", None)
    with pytest.raises(_lib.AlignerB200Error):
        _lib.set_option("no_such_option", "1")


def test_host_entry_validates_like_the_cython_buffer_protocol():
    from aligner_b200.monotonic_align.monotonic_align.core import maximum_path_c
    ok = dict(paths=np.zeros((1, 2, 3), np.int32), values=np.zeros((1, 2, 3), np.float32),
              t_xs=np.zeros(1, np.int32), t_ys=np.zeros(1, np.int32))
    for key, bad in [("paths", np.zeros((1, 2, 3), np.int64)), ("values", np.zeros((1, 2, 3), np.float64)),
                     ("t_xs", np.zeros(1, np.int64)), ("values", np.zeros((1, 3, 2), np.float32).transpose(0, 2, 1))]:
        with pytest.raises(ValueError):
            maximum_path_c(**{**ok, key: bad})
    with pytest.raises(TypeError):
        maximum_path_c(**{**ok, "paths": [[0]]})
