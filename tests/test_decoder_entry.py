"""the decoder-exact hole information emitted by the gather.

* CPU (emulated library): ``nnz`` equals the channel sum of ``(gen_fs != 0)`` exactly, including pixels
  where single channels are exactly zero although the pixel is covered; ``decoder_entry.
  partialconv_mask_path(nnz)`` is BIT-IDENTICAL to what the reference's own ``PartialConv2d``
  (imported unmodified from /root/reference/models/layers/partialconv2d.py) computes from the
  per-element mask of ``ResNetDecoderPconv2.forward`` (networks/architectures.py:369).
* GPU: the same through JointSplat.frames(want_nnz=True) at a ragged size."""
import importlib.util
import os

import numpy as np
import pytest
import torch

REF_PCONV = "/root/reference/models/layers/partialconv2d.py"


def _reference_partialconv():
    spec = importlib.util.spec_from_file_location("_ref_partialconv2d", REF_PCONV)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.PartialConv2d


def _scene_with_exact_zeros(H, W, C, seed):
    """Features with whole channels and scattered pixels exactly zero, so that covered pixels have
    fewer than C non-zero outputs, plus a flow that leaves holes."""
    from slr_sfs_b200 import workloads
    feat, Z, motion = workloads.scene(H, W, C, "A", seed=seed)
    feat = feat.clone()
    feat[:, 1] = 0.0
    feat[:, :, ::3, ::5] = 0.0
    motion = motion.clone()
    motion[:, 0, :, : W // 3] = -2.5           # the left third streams out of the frame: holes open behind it
    return feat, Z, motion


def _check(gen, nnz, C):
    from slr_sfs_b200 import decoder_entry
    want_nnz = (gen != 0).float().sum(1, keepdim=True)
    assert torch.equal(nnz, want_nnz)
    assert 0 < float((nnz == 0).float().mean()) < 1 and float(((nnz > 0) & (nnz < C)).float().mean()) > 0.01
    assert torch.equal(decoder_entry.hole_mask(nnz), (gen != 0).any(1, keepdim=True).float())
    return want_nnz


def test_nnz_and_partialconv_mask_path_on_the_emulated_library():
    emu = pytest.importorskip("emu", reason="tests/emu")
    from slr_sfs_b200 import decoder_entry
    H, W, C, N = 29, 70, 6, 7
    feat, Z, motion = _scene_with_exact_zeros(H, W, C, 3)
    sc = emu.Scene(feat.numpy(), Z.numpy(), motion.numpy())
    gen, nnz = sc.frames(0, N - 1, 0, N, want_nnz=True)
    gen, nnz = torch.from_numpy(gen), torch.from_numpy(nnz)
    _check(gen, nnz, C)
    if not os.path.exists(REF_PCONV):
        pytest.skip("reference tree not mounted: the PartialConv2d comparison needs it")
    PartialConv2d = _reference_partialconv()
    for (k, pad, stride, out_c) in [(3, 1, 1, 8), (5, 2, 2, 4)]:
        layer = PartialConv2d(C, out_c, kernel_size=k, stride=stride, padding=pad, bias=True, multi_channel=True, return_mask=True)
        mask = (gen != 0).float()                                   # networks/architectures.py:369
        with torch.no_grad():
            _, ref_update = layer(gen, mask)
        ref_ratio = layer.mask_ratio
        update, ratio = decoder_entry.partialconv_mask_path(nnz, C, kernel_size=k, stride=stride, padding=pad)
        for o in range(out_c):
            assert torch.equal(update[:, 0], ref_update[:, o]) and torch.equal(ratio[:, 0], ref_ratio[:, o])


@pytest.mark.gpu
def test_nnz_on_the_gpu_all_paths():
    import __graft_entry__
    __graft_entry__.build()
    import slr_sfs_b200 as pkg
    H, W, C, N = 83, 150, 20, 9
    feat, Z, motion = _scene_with_exact_zeros(H, W, C, 5)
    for mode in ("ldg", "staged"):
        os.environ["SLR_GATHER_MODE"] = mode
        try:
            js = pkg.JointSplat(feat.cuda(), Z.cuda(), motion.cuda())
            gen, nnz = js.frames(0, N - 1, 0, N, want_nnz=True)
            torch.cuda.synchronize()
            _check(gen.cpu(), nnz.cpu(), C)
        finally:
            os.environ.pop("SLR_GATHER_MODE", None)
