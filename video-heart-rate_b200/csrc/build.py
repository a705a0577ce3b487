"""Build libvhr_b200.so in-tree with nvcc for sm_100a (the only target).

    python video-heart-rate_b200/csrc/build.py [--force] [--verbose] [--watchdog]

--watchdog builds a second library, libvhr_b200_wd.so, with -DVHR_WATCHDOG: every hand-rolled mbarrier spin loop
traps with a message after 2 M polls instead of hanging the GPU.  Select it with VHR_LIB=<path> (see _lib.py); the
GPU suite is run against it once per round (profiles/).

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["api.cu", "synth.cu", "pyrdown.cu", "pyrdown_stream.cu", "pyrdown_umma.cu", "bandpass.cu", "collapse_sep.cu", "roi.cu", "bpm.cu", "ica.cu", "degrade.cu", "hostpath.cu"]
LIB = os.path.join(HERE, "libvhr_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [os.path.join(HERE, "common.cuh"), os.path.join(HERE, "pairwise.cuh"),
                                                        os.path.join(HERE, "..", "..", "include", "vhr_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, watchdog: bool = False) -> str:
    lib = LIB.replace(".so", "_wd.so") if watchdog else LIB
    if not force and not watchdog and not needs_build():
        return LIB
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + (["-DVHR_WATCHDOG"] if watchdog else [])
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, s.replace(".cu", "_wd.o" if watchdog else ".o"))
        objs.append(o)
        cmd = [nvcc, *flags, "-c", os.path.join(HERE, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {s} failed ---\n{out}\n")
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, watchdog="--watchdog" in sys.argv))
