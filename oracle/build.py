"""Build the oracle's C helpers (oracle/csrc/synth_ref.c) with gcc -> oracle/_build/liboracle_c.so.

    python oracle/build.py

Test infrastructure: only the configuration-golden generator, tests and bench.py's CPU legs
use it (through oracle/fast.py), and only to regenerate inputs faster than NumPy does.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "synth_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_c.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
