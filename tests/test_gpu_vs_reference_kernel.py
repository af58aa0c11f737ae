"""GPU: our kernels against the REFERENCE's own CUDA kernel running on the same device
(kernel text templated by the reference's cupy_kernel(), compiled by oracle/build.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ref():
    from oracle import refgpu
    if not refgpu.available():
        pytest.skip("oracle/_ref/libref_softsplat_gpu.so not built (needs /root/reference once)")
    return refgpu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__
    __graft_entry__.build()
    import slr_sfs_b200
    return slr_sfs_b200


@pytest.mark.parametrize("shape", [(2, 3, 11, 14), (1, 33, 256, 256), (1, 65, 768, 1024)])
def test_summation_splat_vs_reference_gpu_kernel(pkg, ref, shape):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, H, W, generator=g).cuda()
    flow = (torch.rand(B, 2, H, W, generator=g) * 12 - 6).cuda()
    want = ref.softsplat_sum(x, flow).cpu().numpy()
    got = pkg.FunctionSoftsplat(x, flow, None, "summation").cpu().numpy()
    assert rel_err(got, want) <= TOL
    assert np.array_equal(got == 0, want == 0)


def test_joint_block_vs_reference_style_frame(pkg, ref):
    from slr_sfs_b200 import workloads
    H, W, C, N = 96, 128, 8, 12        # 9 = C + 1 channels is a baked shape
    feat, Z, m = workloads.scene(H, W, C, "A", seed=6)
    feat, Z, m = feat.cuda(), Z.cuda(), m.cuda()
    js = pkg.JointSplat(feat, Z, m)
    for t in (0, 5, N - 1):
        want = ref.reference_frame(feat, Z, m, (0, t, N - 1)).cpu().numpy()
        assert rel_err(js.frame((0, t, N - 1)).cpu().numpy(), want) <= TOL
        assert rel_err(js.frame_scatter((0, t, N - 1)).cpu().numpy(), want) <= TOL


def test_full_size_frame_vs_reference_kernel(pkg, ref):
    from slr_sfs_b200 import workloads
    H, W, C, N = 768, 1024, 64, 60
    feat, Z, m = workloads.scene(H, W, C, "A", seed=0)
    feat, Z, m = feat.cuda(), Z.cuda(), m.cuda()
    js = pkg.JointSplat(feat, Z, m)
    want = ref.reference_frame(feat, Z, m, (0, 20, N - 1)).cpu().numpy()
    got = js.frame((0, 20, N - 1)).cpu().numpy()
    assert rel_err(got, want) <= TOL
