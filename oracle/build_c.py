"""Builds oracle/_build/libfen_oracle_c.so from oracle/fen_oracle_c.c and libfen_oracle_mf_c.so from
oracle/fen_oracle_mf_c.c with gcc (C99 + OpenMP).

Test infrastructure: the C restatement is the second checker and the CPU baseline of bench.py; the product
(fen_b200/) never links it.  Usage: python -m oracle.build_c [-f]"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fen_oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfen_oracle_c.so")
# -ffp-contract=off: the reference (gfortran, no -mfma) rounds every operation; keep that arithmetic
FLAGS = ["-std=c99", "-O3", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math"]


SRC_MF = os.path.join(HERE, "fen_oracle_mf_c.c")
LIB_MF = os.path.join(OUT_DIR, "libfen_oracle_mf_c.so")


def build_mf(force: bool = False) -> str:
    """oracle/_build/libfen_oracle_mf_c.so from oracle/fen_oracle_mf_c.c (the two-phase restatement)."""
    return build(force, SRC_MF, LIB_MF)


def build(force: bool = False, SRC: str = SRC, LIB: str = LIB) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    # the system gcc first ($CC may point at a toolchain without libgomp); without OpenMP as the last resort
    errors = []
    for cc in ("gcc", "/usr/bin/gcc", os.environ.get("CC", "")):
        for flags in (FLAGS, [f for f in FLAGS if f != "-fopenmp"]):
            if not cc:
                continue
            r = subprocess.run([cc] + flags + [SRC, "-o", LIB, "-lm"], capture_output=True, text=True)
            if r.returncode == 0:
                return LIB
            errors.append("%s %s:\n%s" % (cc, " ".join(flags), r.stderr[-400:]))
    raise RuntimeError("could not build the C oracle:\n" + "\n".join(errors))


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
    print(build_mf(force="-f" in sys.argv))
