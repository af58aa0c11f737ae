"""CPU: the C-ABI library loads and exports every symbol include/slr_splat.h declares,
and the Python binding covers exactly that set.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "slr_splat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slr_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__
    __graft_entry__.build()
    from slr_sfs_b200 import _lib
    return _lib.LIB_PATH


def test_header_declares_the_survey_boundary():
    syms = declared_symbols()
    for name in ["slr_softsplat_sum_fwd", "slr_softsplat_grad_input", "slr_softsplat_grad_flow",
                 "slr_maxwarpnorm", "slr_euler", "slr_normalize", "slr_version", "slr_last_error_string"]:
        assert name in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), "libslr_splat.so does not export %s" % name


def test_binding_matches_header(lib_path):
    from slr_sfs_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.slr_version() >= 100


def test_argument_errors_are_reported_without_a_gpu(lib_path):
    from slr_sfs_b200 import _lib
    lib = _lib.load()
    rc = lib.slr_softsplat_sum_fwd(None, None, None, 1, 1, 1, 1, 1, None)
    assert rc == -1
    assert b"bad arguments" in lib.slr_last_error_string()
    with pytest.raises(_lib.SlrError):
        _lib.call("slr_euler", None, 1.0, 1, None, None, 4, 4, None)


def test_cpu_tensors_raise_like_the_reference():
    import torch
    import slr_sfs_b200 as pkg
    x, f = torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4)
    with pytest.raises(NotImplementedError):          # softsplat.py:418-419
        pkg.FunctionSoftsplat(x, f, None, "summation")
    with pytest.raises(NotImplementedError):
        pkg.ModuleMaximumWarpNormsplat()(x, f)
    with pytest.raises(NotImplementedError):
        pkg.euler_integration(f, 3)
    with pytest.raises(AssertionError):               # strType assertion, softsplat.py:667
        pkg.FunctionSoftsplat(x, f, None, "bogus")


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "slr-sfs_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), fn
                assert "liboracle" not in text and "libref_softsplat" not in text, fn
