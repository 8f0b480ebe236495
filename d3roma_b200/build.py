"""Build the C-ABI shared library ``libd3roma_b200.so`` with nvcc for sm_100a.

The library is built IN-TREE (``d3roma_b200/libd3roma_b200.so``) so that it travels with the
repository snapshot to the GPU box.  It links the CUDA runtime statically and resolves the one
driver entry point it needs (``cuTensorMapEncodeTiled``) at run time, so it loads (``dlopen``)
on a machine with no GPU and no driver — which is what the ``-m "not gpu"`` symbol tests rely on.
"""
from __future__ import annotations

import concurrent.futures as _fut
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "csrc" / "build"
LIB_PATH = PKG_DIR / "libd3roma_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-Xcudafe", "--diag_suppress=177",
]
if os.environ.get("D3_DEV"):   # developer build: kernel experiments (skip loads / epilogue, cycle counters); see conv_igemm.cu
    NVCC_FLAGS.append("-DD3_DEV")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libd3roma_b200.so")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_hash(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((PKG_DIR.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: Path, verbose: bool) -> Path:
    obj = BUILD_DIR / (src.stem + ".o")
    stamp = BUILD_DIR / (src.stem + ".hash")
    digest = _deps_hash(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == digest:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(PKG_DIR.parent / "include"), "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = BUILD_DIR / (src.stem + ".log")
    log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(f"[d3roma_b200.build] compiled {src.name}\n")
    stamp.write_text(digest)
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compile every ``csrc/*.cu`` for sm_100a and link ``libd3roma_b200.so``. Returns its path."""
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    if force:
        for f in BUILD_DIR.glob("*.hash"):
            f.unlink()
    srcs = _sources()
    with _fut.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(f"[d3roma_b200.build] linked {LIB_PATH}\n")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB_PATH)
