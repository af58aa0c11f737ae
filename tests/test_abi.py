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


@pytest.mark.skipif(not os.path.exists("/root/reference/models/softsplat.py"), reason="reference tree not mounted")
def test_reference_models_import_our_operators_unchanged():
    """Drop-in check at import level: with install_as_reference_modules() the reference's own
    model files (unmodified, /root/reference) bind to this package's operators."""
    import subprocess
    import sys
    code = r'''
import sys, warnings
from unittest import mock
warnings.simplefilter("ignore")
sys.path.insert(0, %r)
sys.path.insert(1, "/root/reference")
for m in ["cv2", "av", "lz4framed", "lpips", "tensorboardX", "matplotlib", "matplotlib.pyplot"]:
    try:
        __import__(m)
    except Exception:
        sys.modules[m] = mock.MagicMock()
import slr_sfs_b200
slr_sfs_b200.install_as_reference_modules()
import importlib
base = importlib.import_module("models.animating_softmax_splating")
two = importlib.import_module("models.animating_softmax_splating_2layers_alpha_seperate")
assert "cupy" not in sys.modules
assert base.softsplat is slr_sfs_b200.softsplat and two.softsplat is slr_sfs_b200.softsplat
assert base.euler_integration is slr_sfs_b200.euler_integration
assert base.EulerIntegration is slr_sfs_b200.EulerIntegration
assert isinstance(base.softsplat.ModuleSoftsplat("summation"), slr_sfs_b200.ModuleSoftsplat)
print("ok")
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
