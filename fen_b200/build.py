"""Builds fen_b200/libfen_gpu.so (sm_100a) in-tree with nvcc.  Usage: python -m fen_b200.build [-f]"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# FEN_BUILD_DEFS / FEN_BUILD_TAG: an A/B variant of the library (e.g. FEN_BUILD_DEFS=-DFEN_STRIDED_TWP=1 FEN_BUILD_TAG=twp
# builds fen_b200/libfen_gpu_twp.so, which FEN_GPU_LIB=... makes fen_b200/_lib.py load instead of the default)
TAG = os.environ.get("FEN_BUILD_TAG", "")
EXTRA = os.environ.get("FEN_BUILD_DEFS", "").split()
OBJ = os.path.join(HERE, "csrc", "_obj" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "libfen_gpu%s.so" % ("_" + TAG if TAG else ""))
SOURCES = ["context.cu", "ghost.cu", "stencil.cu", "poisson.cu", "comm.cu", "tma.cu", "io.cu", "multiphase.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + EXTRA


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
        [os.path.join(os.path.dirname(HERE), "include", "fen_gpu.h")]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    spath = os.path.join(CSRC, src)
    if not _stale(obj, [spath] + _deps()):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", spath, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(obj + ".ptxas.log", "w") as fh:
        fh.write(r.stderr)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(_compile, SOURCES))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
