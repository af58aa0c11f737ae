"""The drop-in claim, EXECUTED: the reference's own model code calling this package's operators.

SURVEY.md section 8 a10: ``for t in range(N): forward_flow(batch)`` (test_animating/
test_v1_4eval_rawsize.py:233-239) must keep working on the replacement operators with zero source
changes.  Three tests:

* ``test_unmodified_forward_flow_on_the_emulated_library`` (CPU, build container): the reference's
  ``AnimatingSoftmaxSplating.forward_flow`` and ``AnimatingSoftmaxSplatingJoint.forward_flow`` are
  imported UNMODIFIED from /root/reference after ``install_as_reference_modules()``, called on a
  stand-in ``self`` exactly like tests/golden/make_golden.py does, and their splats / Euler /
  max-warp-norm calls land in slr_sfs_b200's operator modules, which run the CUDA sources through
  the CPU emulation of the library (tests/emu).  Compared with tests/golden/joint_ref.npz (made by
  the same model code on the reference's own kernels).
* ``test_unmodified_forward_flow_on_cuda`` (GPU): the same on the real library; needs the reference
  tree, which is not on the GPU box -> skipped there, runs wherever both exist.
* ``test_level0_call_pattern_on_cuda`` (GPU): the call pattern of forward_flow (:847-924) restated in
  the test, through the drop-in operators on the device, against the same golden vectors -- what
  the GPU box can check without the reference tree.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
HAVE_REFERENCE = os.path.exists(os.path.join(REFERENCE, "models", "softsplat.py"))
TOL = 1e-4

_DRIVER = r'''
import argparse, contextlib, sys, types, warnings
from unittest import mock
import numpy as np
import torch
warnings.simplefilter("ignore")
ROOT, REFERENCE, MODE = %(root)r, %(reference)r, %(mode)r
sys.path.insert(0, ROOT)
sys.path.insert(1, ROOT + "/tests")
sys.path.insert(2, REFERENCE)
for m in ["cv2", "av", "lz4framed", "lpips", "tensorboardX", "matplotlib", "matplotlib.pyplot"]:
    try:
        __import__(m)
    except Exception:
        sys.modules[m] = mock.MagicMock()
import slr_sfs_b200
from slr_sfs_b200 import _lib

if MODE == "emu":
    # the product has no CPU mode: point the binding at the CPU emulation of the CUDA sources and
    # give torch.cuda's streams inert stand-ins, for this process only (as tests/test_emu_python_layer.py does)
    import emu
    _lib._lib = emu.lib()
    _lib.on_device = lambda t: True
    _lib.current_stream = lambda device: None
    torch.cuda.device = lambda d: contextlib.nullcontext()
    torch.Tensor.cuda = lambda self, *a, **kw: self
    for name in ["linspace", "zeros", "ones"]:
        def strip(fn):
            def wrapped(*a, **kw):
                kw.pop("device", None)
                return fn(*a, **kw)
            return wrapped
        setattr(torch, name, strip(getattr(torch, name)))
    dev = "cpu"
else:
    dev = "cuda"

slr_sfs_b200.install_as_reference_modules()
import importlib
base = importlib.import_module("models.animating_softmax_splating")
two = importlib.import_module("models.animating_softmax_splating_2layers_alpha_seperate")
assert "cupy" not in sys.modules
assert base.softsplat is slr_sfs_b200.softsplat and base.euler_integration is slr_sfs_b200.euler_integration

calls = []
real = _lib._plain_call
def logged(name, *a):
    calls.append(name)
    return real(name, *a)
_lib._plain_call = logged


class Recorder(torch.nn.Module):
    def __init__(self, out_ch):
        super().__init__()
        self.out_ch, self.seen = out_ch, None
    def forward(self, x, *rest):
        self.seen = x.detach().clone()
        return torch.zeros(x.shape[0], self.out_ch, x.shape[2], x.shape[3], device=x.device)


j = np.load(ROOT + "/tests/golden/joint_ref.npz")
N = int(j["N"])
T = lambda k: torch.from_numpy(j[k]).to(dev)
feat, Z, motion = T("feat"), T("Z"), T("motion")
W, C = feat.shape[-1], feat.shape[1]
img = torch.zeros(1, 3, W, W, device=dev)
out = {}
with torch.no_grad():
    for z_mode, flags in [("max", {}), ("v1", {"use_softmax_splatter_v1": True}), ("v2", {"use_softmax_splatter_v2": True})]:
        for t in [0, 3, N - 1]:                      # the per-frame loop of test_v1_4eval_rawsize.py:233-239
            opt = argparse.Namespace(W=W, refine_model_type="resnet_256W8UpDown64", no_clamp_Z=False, **flags)
            me = types.SimpleNamespace(opt=opt, softsplater=base.softsplat.ModuleSoftsplat("summation"), projector=Recorder(3),
                                       maximum_warp_norm_splater=base.softsplat.ModuleMaximumWarpNormsplat())
            batch = {"features": [(feat, Z)], "images": [img], "motions": [motion], "index": torch.tensor([[0, t, N - 1]])}
            base.AnimatingSoftmaxSplating.forward_flow(me, batch)
            out["baseline/%%s/t%%d/gen_fs" %% (z_mode, t)] = me.projector.seen.cpu().numpy()
    a_out = T("alpha_encoder_out")
    for alpha0 in [True, False]:
        for t in [0, 3, N - 1]:
            opt = argparse.Namespace(W=W, ngf=C, use_alpha0_as_blending_weight=alpha0)
            me = types.SimpleNamespace(opt=opt, softsplater=two.softsplat.ModuleSoftsplat("summation"), projector=Recorder(3),
                                       net_alpha_decoder=Recorder(1), net_alpha_encoder=lambda x: a_out)
            batch = {"features": [(feat, Z)], "images": [img], "BGImg": [img], "motions": [motion],
                     "index": torch.tensor([[0, t, N - 1]])}
            two.AnimatingSoftmaxSplatingJoint.forward_flow(me, batch)
            dec_in = me.net_alpha_decoder.seen.cpu().numpy()
            tag = "twolayer/%%s/t%%d" %% ("alpha0" if alpha0 else "plain", t)
            out[tag + "/gen_fs"] = me.projector.seen.cpu().numpy()
            out[tag + "/alpha_fluid"] = dec_in[:, -1:]
assert calls.count("slr_softsplat_sum_fwd") == 2 * 15 and calls.count("slr_euler") == 2 * 15, calls
assert calls.count("slr_maxwarpnorm") == 3
np.savez(%(result)r, **out)
print("ok")
'''


def _run_driver(mode, tmp_path):
    result = str(tmp_path / "drop_in_out.npz")
    code = _DRIVER % dict(root=ROOT, reference=REFERENCE, mode=mode, result=result)
    proc = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert proc.returncode == 0 and proc.stdout.strip().endswith("ok"), proc.stderr[-3000:]
    return np.load(result)


def _compare(got, golden):
    keys = [k for k in golden.files if k.startswith(("baseline/", "twolayer/"))]
    assert len(keys) == 21 and sorted(got.files) == sorted(keys)
    for k in keys:
        assert rel_err(got[k], golden[k]) <= TOL, k
        assert np.all(got[k][golden[k] == 0.0] == 0.0), k         # holes stay exactly zero (decoder mask, architectures.py:369)


@pytest.mark.skipif(not HAVE_REFERENCE, reason="reference tree not mounted")
def test_unmodified_forward_flow_on_the_emulated_library(tmp_path, golden_joint):
    pytest.importorskip("emu", reason="tests/emu")
    _compare(_run_driver("emu", tmp_path), golden_joint)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REFERENCE, reason="reference tree not mounted (it does not travel to the GPU box)")
def test_unmodified_forward_flow_on_cuda(tmp_path, golden_joint):
    _compare(_run_driver("cuda", tmp_path), golden_joint)


@pytest.mark.gpu
def test_level0_call_pattern_on_cuda(golden_joint):
    """forward_flow's call pattern (animating_softmax_splating.py:847-924; 2-layer :921-1045) on the
    drop-in operators: 2 x euler_integration, cat / exp glue, 2 x ModuleSoftsplat('summation'),
    in-place adds on views of the fresh outputs, clamp, divide."""
    import torch
    import __graft_entry__
    __graft_entry__.build()
    import slr_sfs_b200 as pkg
    from slr_sfs_b200 import level0
    j = golden_joint
    N = int(j["N"])
    cu = lambda k: torch.from_numpy(j[k]).cuda()
    feat, Z, motion, a_out = cu("feat"), cu("Z"), cu("motion"), cu("alpha_encoder_out")
    for z_mode in ("max", "v1", "v2"):
        for t in (0, 3, N - 1):
            got = level0.forward_flow_block(feat, Z, motion, (0, t, N - 1), z_mode=z_mode)
            assert rel_err(got.cpu().numpy(), j[f"baseline/{z_mode}/t{t}/gen_fs"]) <= TOL, (z_mode, t)
    a_bg, a_f = torch.sigmoid(a_out[:, 0:1]), a_out[:, 1:2]
    for alpha0 in (True, False):
        for t in (0, 3, N - 1):
            gen, alpha_fluid, mask = level0.forward_flow_block_2layer(feat, Z, a_f, a_bg, motion, (0, t, N - 1), alpha0=alpha0)
            tag = f"twolayer/{'alpha0' if alpha0 else 'plain'}/t{t}"
            assert rel_err(gen.cpu().numpy(), j[f"{tag}/gen_fs"]) <= TOL, tag
            assert rel_err(alpha_fluid.cpu().numpy(), j[f"{tag}/alpha_fluid"]) <= TOL, tag
