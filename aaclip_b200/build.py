"""In-tree build of libaaclip_b200.so (nvcc, sm_100a only).

The shared library is the product's only compute path; it is built next to this file so that it travels
with a repository snapshot.  `python -m aaclip_b200.build` (re)builds it; `build()` is incremental
(per-translation-unit object files keyed on source mtimes) and compiles translation units in parallel.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libaaclip_b200.so"
OBJ_DIR = PKG_DIR / "build"

SOURCES = ["gemm_launch.cu", "attn.cu", "vv_attn.cu", "rowops.cu", "head.cu", "head_stream.cu", "preprocess.cu", "engine.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-I", str(INCLUDE),
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libaaclip_b200.so cannot be built")
    return exe


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h"))
    return max(p.stat().st_mtime for p in hdrs)


def _compile(src: Path, obj: Path, verbose: bool) -> str:
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    (obj.with_suffix(".ptxas.log")).write_text(r.stderr)
    return r.stderr if verbose else ""


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA translation unit for sm_100a and link libaaclip_b200.so in-tree."""
    OBJ_DIR.mkdir(exist_ok=True)
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    if not srcs:
        raise RuntimeError("no CUDA sources found")
    hdr_m = _deps_mtime()
    todo = []
    for s in srcs:
        o = OBJ_DIR / (s.stem + ".o")
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, hdr_m):
            todo.append((s, o))
    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for log in ex.map(lambda so: _compile(so[0], so[1], verbose), todo):
                if verbose and log:
                    print(log, file=sys.stderr)
    objs = [OBJ_DIR / (s.stem + ".o") for s in srcs]
    if todo or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(o.stat().st_mtime for o in objs):
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
