"""In-tree build of libJoshUpscale.so with nvcc for sm_100a.

    python -m joshupscale_b200.build [--force]

Mirrors the reference's build conventions where they matter (core/CMakeLists.txt:
26-51: one SHARED library, hidden visibility, --exclude-libs,ALL, static cudart)
without its CMake/TensorRT machinery.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""

from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from typing import List

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libJoshUpscale.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-DJoshUpscale_EXPORTS",
          "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
          "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def sources() -> List[str]:
    out = []
    for sub in ("kernels", "host"):
        d = os.path.join(CSRC, sub)
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cc")):
                out.append(os.path.join(d, f))
    return out


def _headers_mtime() -> float:
    newest = 0.0
    for base in (CSRC, os.path.join(ROOT, "include")):
        for dp, _, files in os.walk(base):
            for f in files:
                if f.endswith((".h", ".cuh", ".hpp")):
                    newest = max(newest, os.path.getmtime(os.path.join(dp, f)))
    return newest


def _compile(src: str, obj: str, verbose: bool) -> str:
    cmd = [nvcc()] + ARCH + COMMON
    if src.endswith(".cc"):
        cmd += ["-x", "cu"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    hdr = _headers_mtime()
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        objs.append(obj)
        stale = (force or not os.path.exists(obj)
                 or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr))
        if stale:
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(lambda j: _compile(j[0], j[1], verbose), jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB_PATH] + objs + [
            "-Xlinker", "--exclude-libs,ALL", "-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
