"""Build libaligner_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python build_lib.py [--force] [--verbose]

Lives outside the package on purpose: importing aligner_b200 needs the library to exist.

The shared library is the product's only compute path; nothing here builds or
links the CPU oracle.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent / "aligner_b200"
CSRC = PKG / "csrc"
LIB = PKG / "libaligner_b200.so"
OBJ = CSRC / "_obj"
SOURCES = ["mas_api.cu", "neg_cent.cu", "neg_cent_tc.cu", "neg_cent_v2.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _deps(src: Path) -> list[Path]:
    return [src] + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG.parent / "include" / "aligner_b200.h"]


def _stale(out: Path, deps: list[Path]) -> bool:
    return (not out.exists()) or any(d.exists() and d.stat().st_mtime > out.stat().st_mtime for d in deps)


def _compile(src: Path, verbose: bool, extra=(), suffix: str = "") -> Path:
    obj = OBJ / (src.stem + suffix + ".o")
    cmd = [NVCC, *FLAGS, *extra, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src.name, proc.stdout, proc.stderr))
    if verbose:
        sys.stderr.write(proc.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    OBJ.mkdir(exist_ok=True)
    todo = [s for s in srcs if force or _stale(OBJ / (s.stem + ".o"), _deps(s))]
    if todo:
        with ThreadPoolExecutor(max_workers=len(todo)) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    if force or todo or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (proc.stdout, proc.stderr))
    return LIB


def build_variant(name: str, defines=()) -> Path:
    """Developer variants (never shipped, never loaded unless ALB200_LIB points at them): libaligner_b200_<name>.so.
    `dbg` = per-warp clock64 stamps compiled in (tools/dbg_timing.py, tools/nc_timeline.py); others are A/B builds
    (python build_lib.py --variant nospec -DALB_SPEC_ADD=0 -DALB200_FEW_KERNELS=1)."""
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    OBJ.mkdir(exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(lambda s: _compile(s, False, tuple(defines), "_" + name), srcs))
    out = PKG / ("libaligner_b200_%s.so" % name)
    proc = subprocess.run([NVCC, "-shared", "-o", str(out), *map(str, objs), "-lcudart"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (proc.stdout, proc.stderr))
    return out


if __name__ == "__main__":
    if "--dbg" in sys.argv:
        print(build_variant("dbg", ["-DALB200_DBG_BUILD=1"]))
        sys.exit(0)
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")]))
        sys.exit(0)
    out = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(out)
