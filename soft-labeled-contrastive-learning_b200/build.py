"""Build libslcl.so (sm_100a only) in-tree with nvcc.

    python soft-labeled-contrastive-learning_b200/build.py [--force] [--verbose]

Every .cu under csrc/ is compiled with
``-gencode arch=compute_100a,code=sm_100a -lineinfo -O3`` (cross-compiles on a
GPU-less host) and linked into ``slcl/libslcl.so`` next to the Python package,
so the library travels with the source tree to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "slcl")
OBJ_DIR = os.path.join(HERE, "build", os.environ.get("SLCL_LIB_NAME", "libslcl.so")[:-3])
LIB = os.path.join(OUT_DIR, os.environ.get("SLCL_LIB_NAME", "libslcl.so"))      # tuning variants: other names
EXTRA = os.environ.get("SLCL_EXTRA_NVCC_FLAGS", "").split()

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC, *EXTRA,
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libslcl needs the CUDA 12.9 toolchain")
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for path in _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + \
            [os.path.join(ROOT, "include", "slcl.h"), os.path.abspath(__file__)]:
        h.update(" ".join(EXTRA).encode())
        with open(path, "rb") as fh:
            h.update(path.encode())
            h.update(fh.read())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    cmd = [nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".ptxas.log")
    with open(log, "w") as fh:
        fh.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(lambda s: _compile(s, verbose), _sources()))
    cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
