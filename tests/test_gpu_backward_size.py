"""GPU: the backward kernels at the TRAINING shape -- B=2 per GPU, 65 channels
(64 features + the softmax weight), 256x256 (train_alpha_finetuneBG_finetuneFluid_v1.sh) -- and at the
inference shape, against the reference's own kernel_Softsplat_updateGradInput / updateGradFlow
(models/softsplat.py:204-326) compiled for this GPU by oracle/build.py.  1e-4 relative."""
import pytest
import torch

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


def gpu_rel_err(got, want):
    want = want.double()
    s = torch.sqrt(torch.mean(want * want)).clamp(min=1e-30)
    return float(((got.double() - want).abs() / torch.maximum(want.abs(), s)).max())


@pytest.mark.parametrize("shape", [(2, 65, 256, 256), (1, 65, 768, 1024), (2, 3, 11, 14)])
@pytest.mark.parametrize("flow_kind", ["smooth", "random"])
def test_backward_vs_reference_kernels_at_size(pkg, ref, shape, flow_kind):
    B, C, H, W = shape
    if shape not in ref.baked_backward_shapes():
        pytest.skip("reference backward kernels for %s not baked" % (shape,))
    g = torch.Generator().manual_seed(C * H)
    x = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    if flow_kind == "random":
        flow = (torch.rand(B, 2, H, W, generator=g) * 12 - 6).cuda()
    else:
        ys = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).expand(B, 1, H, W)
        xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(B, 1, H, W)
        flow = torch.cat([3.0 * torch.sin(xs / 17.0) * torch.cos(ys / 13.0) + 0.3, 2.0 * torch.cos(xs / 11.0 + 1.0)], 1).cuda()
    flow = flow.contiguous().requires_grad_(True)
    gout = torch.randn(B, C, H, W, generator=g).cuda()
    out = pkg.softsplat._FunctionSoftsplat.apply(x, flow)
    out.backward(gout)
    want_gin, want_gflow = ref.softsplat_backward(x.detach(), flow.detach(), gout)
    assert gpu_rel_err(x.grad, want_gin) <= TOL
    assert gpu_rel_err(flow.grad, want_gflow) <= TOL
