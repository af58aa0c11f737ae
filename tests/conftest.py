"""pytest configuration: markers, import path, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_addoption(parser):
    parser.addoption("--slr-lib", default=os.environ.get("SLR_TEST_LIB"),
                     help="run the GPU tests against another build of libslr_splat.so (a compile-time variant "
                          "prebuilt by profiles/build_variants.py); default: the product build")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    lib = config.getoption("--slr-lib")
    if lib:
        from slr_sfs_b200 import _lib
        _lib.LIB_PATH = os.path.abspath(lib)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden_softsplat():
    return load_golden("softsplat_ref")


@pytest.fixture(scope="session")
def golden_euler():
    return load_golden("euler_ref")


@pytest.fixture(scope="session")
def golden_joint():
    return load_golden("joint_ref")


def rel_err(a, ref):
    """Deviation as SURVEY.md section 7 defines it: |a-ref| / max(|ref|, rms(ref)), worst element."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    s = float(np.sqrt(np.mean(ref * ref))) if ref.size else 0.0
    den = np.maximum(np.abs(ref), max(s, 1e-30))
    return float(np.max(np.abs(a - ref) / den)) if ref.size else 0.0


# tests/emu (the CUDA sources compiled for the CPU) is importable as `emu`
_TESTS = os.path.dirname(os.path.abspath(__file__))
if _TESTS not in sys.path:
    sys.path.insert(0, _TESTS)


@pytest.fixture(scope="session")
def golden_euler_grad():
    return load_golden("euler_grad_ref")


@pytest.fixture(scope="session")
def golden_warp_flow():
    return load_golden("warp_flow_ref")
