"""Compile art_tts_b200/csrc/*.cu into art_tts_b200/lib/libmas_sm100.so (sm_100a only).

nvcc cross-compiles without a GPU; the built library is git-ignored but travels to the GPU
box with the repo snapshot.  `python -m art_tts_b200.build [--force] [--verbose]`.
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmas_sm100.so")
STAMP = LIB + ".srchash"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--fmad=false",            # parity: the reference never contracts a*b+c (fused kernel opts in per call site)
    "-Xptxas", "-warn-spills",
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libmas_sm100.so cannot be built")
    return cand


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    inc = os.path.join(os.path.dirname(HERE), "include")
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) \
        + sorted(glob.glob(os.path.join(CSRC, "*.h"))) + sorted(glob.glob(os.path.join(inc, "*.h")))


def source_hash() -> str:
    """Content hash of every source the library is built from (+ the flags): robust to the
    mtime changes of a repo snapshot, so a shipped .so is rebuilt only when it is stale."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def _compile_one(args):
    nvcc, src, obj, extra, verbose = args
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", obj, src]
    if verbose:
        print(" ".join(cmd), flush=True)
    return subprocess.run(cmd, capture_output=True, text=True)


def build(force: bool = False, verbose: bool = False, extra=(), out: str | None = None) -> str:
    """`out`: build a variant (e.g. extra=["-DMAS_TIMING"]) next to the production library."""
    global LIB, STAMP
    if out is not None:
        saved = (LIB, STAMP)
        LIB, STAMP = out, out + ".srchash"
        try:
            return build(force=True, verbose=verbose, extra=extra)
        finally:
            LIB, STAMP = saved
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj" + ("_var" if extra else ""))
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    jobs = [(nvcc, s, os.path.join(objdir, os.path.basename(s) + ".o"), tuple(extra), verbose)
            for s in sources()]
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        results = list(ex.map(_compile_one, jobs))
    for res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
    if any(r.returncode != 0 for r in results):
        raise RuntimeError("nvcc failed building libmas_sm100.so")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB,
            *[j[2] for j in jobs]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libmas_sm100.so")
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    extra = ["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else []
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or bool(extra), extra=extra))
