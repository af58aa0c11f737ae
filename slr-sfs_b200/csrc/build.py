"""Compile the CUDA sources of this directory into csrc/libslr_splat.so for sm_100a.

Plain nvcc, in-tree, no torch headers: the library exposes the C ABI declared in
include/slr_splat.h and links the static CUDA runtime, so it can be loaded from
Python (ctypes), C or C++ alike.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libslr_splat.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(HERE, "*.h")) + \
        [os.path.join(HERE, "..", "..", "include", "slr_splat.h"), os.path.abspath(__file__)]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    # one process per GPU may arrive here at the same time (torchrun): the first one compiles, the others wait
    # and find the library fresh
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not stale():
            return LIB
        return _compile(verbose)


def _compile(verbose):
    # tuning knobs (see clip_gather.cu / clip_common.cuh); e.g. SLR_DEFINES="-DSLR_LIST_DEPTH=128"
    extra = os.environ.get("SLR_DEFINES", "").split()
    tmp = LIB + ".tmp.%d" % os.getpid()          # never a half-written library at the final path
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + sources()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    os.replace(tmp, LIB)
    if verbose:
        print(proc.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
