#!/usr/bin/env python
"""Run bench.py against another build of the library (A/B of compile-time knobs without nvcc on the GPU box):
    SLR_LIB=gpurun_variants/libslr_splat_e13x8.so python profiles/bench_with_lib.py --no-e2e --no-cpu-baseline"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slr_sfs_b200 import _lib

_lib.LIB_PATH = os.path.abspath(os.environ["SLR_LIB"])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
