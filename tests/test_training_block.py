"""The fused, differentiable joint block of the TRAINING forward (slr_producer_splat_fwd / _bwd,
slr_sfs_b200.training_block) against the oracle: the pinned restatements of the reference's splat kernels
(softsplat.py:157-326) composed by the chain rule autograd applies in
AnimatingSoftmaxSplating.forward (animating_softmax_splating.py:584-692).

CPU: the kernel sources on the emulator.  GPU: the CUDA library through the autograd Function, also against
torch autograd through the Level-0 drop-in operators on the reference's own expression."""
import numpy as np
import pytest

import oracle
from conftest import rel_err

TOL = 1e-4


def _case(seed, B, C, H, W, amp=3.0):
    r = np.random.default_rng(seed)
    fs = r.standard_normal((B, C, H, W)).astype(np.float32)
    zn = np.clip(r.standard_normal((B, 1, H, W)) * 2 - 2, -20, 20).astype(np.float32)
    flow = r.uniform(-amp, amp, (B, 2, H, W)).astype(np.float32)
    flow[:, :, : H // 3] = 0.0                      # static rows: land exactly on a cell
    alpha = r.uniform(0.05, 0.95, B).astype(np.float32)
    gacc = r.standard_normal((B, C + 1, H, W)).astype(np.float32)
    return fs, zn, flow, alpha, gacc


# ---------------------------------------------------------------------------- CPU: kernel sources on the emulator
emu = pytest.importorskip("emu", reason="tests/emu")


@pytest.mark.parametrize("shape", [(1, 3, 17, 23), (2, 5, 32, 40), (2, 17, 9, 33)])
def test_emu_producer_splat_forward_and_accumulate(shape):
    B, C, H, W = shape
    fs, zn, flow, alpha, _ = _case(3, B, C, H, W)
    acc = np.full((B, C + 1, H, W), np.nan, np.float32)
    emu.call("slr_producer_splat_fwd", emu.p(fs), emu.p(zn), emu.p(flow), emu.p(alpha), emu.p(acc), B, C, H, W, 0, None)
    want, _ = oracle.producer_splat(fs, zn, flow, alpha)
    assert rel_err(acc, want) <= TOL
    # the second direction adds into the same accumulator
    fs2, zn2, flow2, _, _ = _case(4, B, C, H, W, amp=6.0)
    emu.call("slr_producer_splat_fwd", emu.p(fs2), emu.p(zn2), emu.p(flow2), emu.p(1 - alpha), emu.p(acc), B, C, H, W, 1, None)
    want2 = want + oracle.producer_splat(fs2, zn2, flow2, 1 - alpha)[0]
    assert rel_err(acc, want2) <= TOL


@pytest.mark.parametrize("shape", [(1, 3, 17, 23), (2, 5, 32, 40), (2, 17, 9, 33)])
def test_emu_producer_splat_backward(shape):
    B, C, H, W = shape
    fs, zn, flow, alpha, gacc = _case(5, B, C, H, W)
    d_fs, d_zn, d_flow = np.full_like(fs, np.nan), np.full_like(zn, np.nan), np.full_like(flow, np.nan)
    emu.call("slr_producer_splat_bwd", emu.p(fs), emu.p(zn), emu.p(flow), emu.p(alpha), emu.p(gacc),
             emu.p(d_fs), emu.p(d_zn), emu.p(d_flow), B, C, H, W, None)
    w_fs, w_zn, w_flow = oracle.producer_splat_grads(fs, zn, flow, alpha, gacc)
    assert rel_err(d_fs, w_fs) <= TOL and rel_err(d_zn, w_zn) <= TOL and rel_err(d_flow, w_flow) <= TOL
    # outputs that are not wanted are not written
    only = np.full_like(zn, np.nan)
    emu.call("slr_producer_splat_bwd", emu.p(fs), emu.p(zn), emu.p(flow), emu.p(alpha), emu.p(gacc),
             None, emu.p(only), None, B, C, H, W, None)
    assert np.array_equal(only, d_zn)


def test_emu_producer_splat_gradients_are_the_derivative():
    """Finite differences of the oracle's forward (double accumulators) around a point where no footprint
    crosses a cell border within the step."""
    B, C, H, W = 1, 2, 6, 7
    r = np.random.default_rng(9)
    fs = r.standard_normal((B, C, H, W)).astype(np.float32)
    zn = (r.standard_normal((B, 1, H, W)) - 1).astype(np.float32)
    flow = (r.integers(-2, 3, (B, 2, H, W)) + r.uniform(0.25, 0.75, (B, 2, H, W))).astype(np.float32)
    alpha = np.array([0.6], np.float32)
    gacc = r.standard_normal((B, C + 1, H, W)).astype(np.float32)
    d_fs, d_zn, d_flow = np.empty_like(fs), np.empty_like(zn), np.empty_like(flow)
    emu.call("slr_producer_splat_bwd", emu.p(fs), emu.p(zn), emu.p(flow), emu.p(alpha), emu.p(gacc),
             emu.p(d_fs), emu.p(d_zn), emu.p(d_flow), B, C, H, W, None)

    def loss(fs_, zn_, flow_):
        a = alpha.reshape(-1, 1, 1, 1).astype(np.float64)
        ez = np.exp(zn_.astype(np.float64))
        ten = np.concatenate([fs_ * ez * a, ez * a], 1)
        return float((oracle.softsplat_sum_f64(ten, flow_) * gacc).sum())

    eps = 1e-3
    for (name, arr, grad) in (("fs", fs, d_fs), ("zn", zn, d_zn), ("flow", flow, d_flow)):
        for idx in [tuple(r.integers(0, s) for s in arr.shape) for _ in range(6)]:
            hi, lo = arr.astype(np.float64), arr.astype(np.float64)
            hi[idx] += eps
            lo[idx] -= eps
            args_hi = dict(fs=fs, zn=zn, flow=flow)
            args_lo = dict(args_hi)
            args_hi[name], args_lo[name] = hi, lo
            num = (loss(args_hi["fs"], args_hi["zn"], args_hi["flow"]) - loss(args_lo["fs"], args_lo["zn"], args_lo["flow"])) / (2 * eps)
            assert abs(num - grad[idx]) <= 2e-3 * max(1.0, abs(num)), (name, idx, num, grad[idx])


# ---------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__
    __graft_entry__.build()
    import slr_sfs_b200
    return slr_sfs_b200


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 64, 256, 256), (1, 7, 45, 67)])
def test_gpu_joint_block_training_forward_and_backward(pkg, shape):
    import torch
    from slr_sfs_b200 import training_block
    B, C, H, W = shape
    r = np.random.default_rng(11)
    start_fs, end_fs = (r.standard_normal((B, C, H, W)).astype(np.float32) for _ in range(2))
    Z_f, Z_p = (r.standard_normal((B, 1, H, W)).astype(np.float32) * 3 for _ in range(2))
    flow_f, flow_p = (r.uniform(-5, 5, (B, 2, H, W)).astype(np.float32) for _ in range(2))
    alpha = r.uniform(0.1, 0.9, B).astype(np.float32)
    dev = torch.device("cuda")
    t = {k: torch.from_numpy(v).to(dev).requires_grad_(k != "alpha") for k, v in
         dict(start_fs=start_fs, end_fs=end_fs, Z_f=Z_f, Z_p=Z_p, flow_f=flow_f, flow_p=flow_p, alpha=alpha).items()}
    gen = training_block.joint_block_training(t["start_fs"], t["end_fs"], t["Z_f"], t["Z_p"], t["flow_f"], t["flow_p"], t["alpha"])
    want = oracle.joint_block_training(start_fs, end_fs, Z_f, Z_p, flow_f, flow_p, alpha)
    assert rel_err(gen.detach().cpu().numpy(), want) <= TOL
    g = torch.from_numpy(r.standard_normal(want.shape).astype(np.float32)).to(dev)
    gen.backward(g)
    fused = {k: v.grad.clone() for k, v in t.items() if k != "alpha"}

    # the reference's own expression (:586-692) on the Level-0 drop-in operator, walked back by torch autograd
    # through the splat's two backward kernels (pinned to the reference's in tests/test_gpu_ops.py)
    for v in t.values():
        v.grad = None
    a4 = t["alpha"].view(B, 1, 1, 1)
    zf = torch.clamp(t["Z_f"] - t["Z_f"].max(), min=-20.0, max=20.0)
    zp = torch.clamp(t["Z_p"] - t["Z_p"].max(), min=-20.0, max=20.0)
    ten_f = torch.cat([t["start_fs"] * zf.exp() * a4, zf.exp() * a4], 1)
    ten_p = torch.cat([t["end_fs"] * zp.exp() * (1 - a4), zp.exp() * (1 - a4)], 1)
    splat = pkg.softsplat.ModuleSoftsplat("summation")
    ones = t["start_fs"].new_ones(B, 1, H, W)
    acc = splat(tenInput=ten_f, tenFlow=t["flow_f"], tenMetric=ones) + splat(tenInput=ten_p, tenFlow=t["flow_p"], tenMetric=ones)
    ref = acc[:, :-1] / torch.clamp(acc[:, -1:], min=1e-8)
    assert rel_err(gen.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= 1e-5
    ref.backward(g)
    for k, v in fused.items():
        assert rel_err(v.cpu().numpy(), t[k].grad.cpu().numpy()) <= TOL, k


@pytest.mark.gpu
def test_gpu_producer_splat_against_the_oracle_gradients(pkg):
    import torch
    from slr_sfs_b200 import training_block
    B, C, H, W = 2, 9, 40, 56
    fs, zn, flow, alpha, gacc = _case(21, B, C, H, W, amp=7.0)
    dev = torch.device("cuda")
    tf, tz, tl = (torch.from_numpy(a).to(dev).requires_grad_(True) for a in (fs, zn, flow))
    acc = training_block.producer_splat(tf, tz, tl, torch.from_numpy(alpha).to(dev))
    assert rel_err(acc.detach().cpu().numpy(), oracle.producer_splat(fs, zn, flow, alpha)[0]) <= TOL
    acc.backward(torch.from_numpy(gacc).to(dev))
    w_fs, w_zn, w_flow = oracle.producer_splat_grads(fs, zn, flow, alpha, gacc)
    assert rel_err(tf.grad.cpu().numpy(), w_fs) <= TOL
    assert rel_err(tz.grad.cpu().numpy(), w_zn) <= TOL
    assert rel_err(tl.grad.cpu().numpy(), w_flow) <= TOL
    # CPU tensors are refused like everywhere in the package
    with pytest.raises(AssertionError):
        training_block.producer_splat(tf.cpu(), tz.cpu(), tl.cpu(), torch.from_numpy(alpha))
