#!/usr/bin/env python
"""Prebuild compile-time variants of libslr_splat.so HERE (nvcc cross-compiles without a GPU) so that a
gpurun call spends its minutes on running them, not on compiling:

    python profiles/build_variants.py static:-DSLR_STATIC_TILE_FASTPATH=1 \\
        shift:-DSLR_GATHER_SHIFT_SHARE=1 both:-DSLR_STATIC_TILE_FASTPATH=1,-DSLR_GATHER_SHIFT_SHARE=1

writes gpurun_variants/libslr_splat_<name>.so (git-ignored, but it travels to the GPU box).  There:
    bash profiles/run_variants.sh static shift both        # parity tests + bench per variant"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "slr-sfs_b200", "csrc")
OUT = os.path.join(ROOT, "gpurun_variants")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-Xptxas", "-v"]


def main():
    os.makedirs(OUT, exist_ok=True)
    for spec in sys.argv[1:]:
        name, _, defs = spec.partition(":")
        lib = os.path.join(OUT, "libslr_splat_%s.so" % name)
        cmd = ["nvcc"] + FLAGS + [d for d in defs.split(",") if d] + ["-o", lib] + sorted(glob.glob(os.path.join(CSRC, "*.cu")))
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            sys.exit("nvcc failed for %s:\n%s" % (name, proc.stdout[-4000:]))
        # registers / spills of the two hot kernels
        lines = proc.stdout.splitlines()
        for i, ln in enumerate(lines):
            if "Compiling entry function" in ln and ("rowgather_kernelILi0ELi2ELi2" in ln or "expand_kernel" in ln):
                print(name, ln.split("'")[1][:48], "|", lines[i + 2].strip().replace("ptxas info    : ", ""), "|", lines[i + 1].strip())
        print("built", lib)


if __name__ == "__main__":
    main()
