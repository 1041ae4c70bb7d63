"""Builds libihgnn_b200.so (C ABI, sm_100a only) in-tree with nvcc.

    python -m ihgnn_b200.build [--force]

The shared library and its digest stamp land in ihgnn_b200/lib/ (both git-ignored via `*.so` /
`ihgnn_b200/lib/`, but they travel with the repo snapshot to the GPU box).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import hashlib
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libihgnn_b200.so")
OBJ_DIR = os.path.join(REPO, "build", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libihgnn_b200.so cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(REPO, "include", "ihgnn_b200.h"))
    for f in files:
        h.update(os.path.relpath(f, REPO).encode())      # location-independent: a checkout elsewhere matches
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(nvcc: str, src: str, obj_dir: str = OBJ_DIR, extra=()) -> str:
    obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", os.path.join(REPO, "include"), "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if stale) and return the path of the shared library."""
    stamp = os.path.join(LIB_DIR, "libihgnn_b200.sha256")
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s), sources()))
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs]  # cudart is linked statically (nvcc default)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as f:
        f.write(digest + "\n")
    if verbose:
        print(f"built {LIB_PATH}")
    return LIB_PATH


def build_trace() -> str:
    """Dev-only instrumented copy (clock64 probes, -DIHG_TRACE) -> build/libihgnn_trace.so."""
    nvcc = _nvcc()
    obj_dir = os.path.join(REPO, "build", "obj_trace")
    os.makedirs(obj_dir, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s, obj_dir, ("-DIHG_TRACE",)), sources()))
    out = os.path.join(REPO, "build", "libihgnn_trace.so")
    res = subprocess.run([nvcc, "-shared", "-o", out, *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--trace", action="store_true", help="also build the instrumented dev library")
    args = ap.parse_args()
    print(build(force=args.force, verbose=True))
    if args.trace:
        print(build_trace())
