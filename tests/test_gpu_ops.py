"""GPU: operator-level parity of the sm_100a library (through the reference-shaped
Python API, which calls the C ABI) against the oracle and the reference's golden
vectors.  Tolerance for fp32 sums whose order differs: 1e-4 relative, measured as
|a-b| / max(|b|, rms(b)) (BASELINE.json north_star).  Euler and the max
ops are order-independent and must be bit-exact."""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__
    __graft_entry__.build()
    import slr_sfs_b200
    return slr_sfs_b200


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


SPLAT_CASES = ["zero", "int_shift", "half", "neg_frac", "uniform4", "uniform_big",
               "onto_last_cell", "sink", "sentinel"]


@pytest.mark.parametrize("case", SPLAT_CASES)
def test_summation_splat_vs_reference_golden(pkg, golden_softsplat, case):
    g = golden_softsplat
    inp, flow = g[f"{case}/inp"], g[f"{case}/flow"]
    out = pkg.FunctionSoftsplat(cu(inp), cu(flow), None, "summation").cpu().numpy()
    assert rel_err(out, g[f"{case}/sum"]) <= TOL
    holes = g[f"{case}/sum"] == 0.0
    assert np.all(out[holes] == 0.0)          # cells nobody reaches stay exactly 0


@pytest.mark.parametrize("case", SPLAT_CASES)
def test_summation_splat_through_the_gather_vs_reference_golden(pkg, golden_softsplat, case, monkeypatch):
    """The large-frame path of the operator (slr_softsplat_sum_fwd_gather), forced at the golden sizes."""
    monkeypatch.setattr(pkg.softsplat, "GATHER_MIN_ELEMENTS", 0)
    g = golden_softsplat
    inp, flow = g[f"{case}/inp"], g[f"{case}/flow"]
    out = pkg.FunctionSoftsplat(cu(inp), cu(flow), None, "summation").cpu().numpy()
    assert rel_err(out, g[f"{case}/sum"]) <= TOL
    assert np.all(out[g[f"{case}/sum"] == 0.0] == 0.0)


def test_summation_splat_paths_agree_and_keep_autograd(pkg, monkeypatch):
    """Gather path against atomic scatter at a shape with static rows, a batch of two and ragged tiles; the
    backward (the reference's two gather kernels) is the same whichever forward ran."""
    r = np.random.default_rng(5)
    B, C, H, W = 2, 33, 150, 201
    inp = r.standard_normal((B, C, H, W)).astype(np.float32)
    flow = r.uniform(-6, 6, (B, 2, H, W)).astype(np.float32)
    flow[:, :, :40] = 0.0
    res = {}
    for name, thresh in (("scatter", 1 << 62), ("gather", 0)):
        monkeypatch.setattr(pkg.softsplat, "GATHER_MIN_ELEMENTS", thresh)
        x, f = cu(inp).requires_grad_(True), cu(flow).requires_grad_(True)
        y = pkg.FunctionSoftsplat(x, f, None, "summation")
        y.backward(torch.ones_like(y))
        res[name] = (y.detach().cpu().numpy(), x.grad.cpu().numpy(), f.grad.cpu().numpy())
    assert rel_err(res["gather"][0], oracle.softsplat_sum(inp, flow)) <= TOL
    assert rel_err(res["gather"][0], res["scatter"][0]) <= 1e-5
    assert np.array_equal(res["gather"][1], res["scatter"][1]) and np.array_equal(res["gather"][2], res["scatter"][2])


@pytest.mark.parametrize("case", SPLAT_CASES)
def test_backward_vs_reference_golden(pkg, golden_softsplat, case):
    g = golden_softsplat
    x = cu(g[f"{case}/inp"]).requires_grad_(True)
    f = cu(g[f"{case}/flow"]).requires_grad_(True)
    out = pkg.softsplat._FunctionSoftsplat.apply(x, f)
    out.backward(cu(g[f"{case}/gout"]))
    assert rel_err(x.grad.cpu().numpy(), g[f"{case}/gin"]) <= TOL
    assert rel_err(f.grad.cpu().numpy(), g[f"{case}/gflow"]) <= TOL


def test_backward_respects_needs_input_grad(pkg):
    x = torch.randn(1, 3, 8, 8, device="cuda", requires_grad=True)
    f = torch.rand(1, 2, 8, 8, device="cuda")          # GT motion: no grad wanted
    out = pkg.softsplat._FunctionSoftsplat.apply(x, f)
    out.backward(torch.ones_like(out))
    assert x.grad is not None and f.grad is None
    # like the reference (softsplat.py:438) a non-contiguous gradOutput is an assertion error
    with pytest.raises(AssertionError):
        pkg.softsplat._FunctionSoftsplat.apply(x, f).sum().backward()


@pytest.mark.parametrize("case", SPLAT_CASES)
def test_max_warp_norm_bit_exact(pkg, golden_softsplat, case):
    g = golden_softsplat
    out = pkg.ModuleMaximumWarpNormsplat()(cu(g[f"{case}/z"]), cu(g[f"{case}/flow"])).cpu().numpy()
    assert np.array_equal(out, g[f"{case}/maxwarpnorm"])


def test_maximumsplat_zero_init(pkg):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 2, 9, 12)).astype(np.float32)
    flow = rng.uniform(-3, 3, (2, 2, 9, 12)).astype(np.float32)
    out = pkg.ModuleMaximumsplat()(cu(x), cu(flow)).cpu().numpy()
    assert np.array_equal(out, oracle.maxsplat(x, flow, 0.0))


@pytest.mark.parametrize("mode", ["summation", "average", "linear", "softmax"])
def test_function_softsplat_modes(pkg, mode):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 5, 23, 31)).astype(np.float32)
    flow = rng.uniform(-4, 4, (2, 2, 23, 31)).astype(np.float32)
    metric = None if mode in ("summation", "average") else rng.standard_normal((2, 1, 23, 31)).astype(np.float32)
    if mode == "linear":
        metric = np.abs(metric) + 0.1
    want = oracle.function_softsplat(x, flow, metric, mode)
    got = pkg.FunctionSoftsplat(cu(x), cu(flow), None if metric is None else cu(metric), mode).cpu().numpy()
    assert rel_err(got, want) <= TOL
    mod = pkg.ModuleSoftsplat(mode)(tenInput=cu(x), tenFlow=cu(flow), tenMetric=None if metric is None else cu(metric))
    assert rel_err(mod.cpu().numpy(), want) <= TOL


def test_config1_256x256_softmax(pkg):
    """BASELINE.json configs[0]: single 256x256 frame, 32-ch features, 1 softmax splat."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 32, 256, 256, generator=g)
    flow = (torch.rand(1, 2, 256, 256, generator=g) * 8 - 4)
    metric = torch.randn(1, 1, 256, 256, generator=g)
    want = oracle.function_softsplat(x.numpy(), flow.numpy(), metric.numpy(), "softmax")
    got = pkg.FunctionSoftsplat(x.cuda(), flow.cuda(), metric.cuda(), "softmax").cpu().numpy()
    assert rel_err(got, want) <= TOL


def test_output_is_fresh_and_writable_like_the_reference(pkg):
    # callers take views of the op's output and += into them
    # (animating_softmax_splating.py:918-921)
    x = torch.randn(1, 4, 8, 8, device="cuda")
    f = torch.zeros(1, 2, 8, 8, device="cuda")
    a = pkg.ModuleSoftsplat("summation")(x, f, None)
    b = pkg.ModuleSoftsplat("summation")(x, f, None)
    assert a.data_ptr() != b.data_ptr() and a.data_ptr() != x.data_ptr()
    view = a[:, :-1]
    view += b[:, :-1]
    assert torch.equal(a[:, :-1], 2 * x[:, :-1]) and torch.equal(a[:, -1:], x[:, -1:])


def test_assertions_match_reference(pkg):
    x = torch.randn(1, 4, 8, 8, device="cuda")
    with pytest.raises(AssertionError):
        pkg.FunctionSoftsplat(x, torch.zeros(1, 3, 8, 8, device="cuda"), None, "summation")
    with pytest.raises(AssertionError):
        pkg.FunctionSoftsplat(x, torch.zeros(1, 2, 8, 9, device="cuda"), None, "summation")
    with pytest.raises(AssertionError):
        pkg.FunctionSoftsplat(x.transpose(2, 3), torch.zeros(1, 2, 8, 8, device="cuda"), None, "summation")
    with pytest.raises(AssertionError):
        pkg.FunctionSoftsplat(x, torch.zeros(1, 2, 8, 8, device="cuda"), torch.zeros(1, 2, 8, 8, device="cuda"), "linear")


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 3, 1, 37), (2, 1, 41, 1), (3, 7, 19, 33), (1, 65, 64, 96)])
def test_ragged_shapes(pkg, shape):
    rng = np.random.default_rng(sum(shape))
    B, C, H, W = shape
    x = rng.standard_normal(shape).astype(np.float32)
    flow = rng.uniform(-5, 5, (B, 2, H, W)).astype(np.float32)
    got = pkg.FunctionSoftsplat(cu(x), cu(flow), None, "summation").cpu().numpy()
    assert rel_err(got, oracle.softsplat_sum(x, flow)) <= TOL


def test_euler_bit_exact_vs_reference_golden(pkg, golden_euler):
    e = golden_euler
    n = 0
    for k in e.files:
        if not k.endswith("/disp"):
            continue
        base, T, _ = k.rsplit("/", 2)
        d, v = pkg.euler_integration(cu(e[base + "/motion"]), int(T[1:]))
        assert np.array_equal(d.cpu().numpy(), e[k]), k
        assert np.array_equal(v.cpu().numpy(), e[k[:-4] + "vis"]), k
        n += 1
    assert n > 50
    d, _ = pkg.euler_integration(cu(e["tensorT/motion"]), torch.tensor([4]))
    assert np.array_equal(d.cpu().numpy(), e["tensorT/T4/disp"])


def test_euler_module_and_negated_flow(pkg):
    rng = np.random.default_rng(11)
    m = rng.uniform(-2, 2, (2, 2, 30, 44)).astype(np.float32)
    mod = pkg.EulerIntegration(None)
    d, v = mod(cu(m), torch.tensor([5, 9]), show_visible_pixels=True)
    for b, T in enumerate([5, 9]):
        wd, wv = oracle.euler(m[b:b + 1], T)
        assert np.array_equal(d[b:b + 1].cpu().numpy(), wd) and np.array_equal(v[b:b + 1].cpu().numpy(), wv)
    d2, _ = pkg.euler_integration(-cu(m[:1]), 7)
    assert np.array_equal(d2.cpu().numpy(), oracle.euler(-m[:1], 7)[0])
    assert mod(cu(m), [3, 3]).shape == (2, 2, 30, 44)


def test_euler_non_square_60_steps(pkg):
    H, W = 96, 160
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    m = np.stack([2.5 * np.sin(xs / 17) * np.cos(ys / 13) + 0.5, 2.0 * np.cos(xs / 11) * np.sin(ys / 19)])[None].astype(np.float32)
    for T in (1, 30, 60):
        d, v = pkg.euler_integration(cu(m), T)
        wd, wv = oracle.euler(m, T)
        assert np.array_equal(d.cpu().numpy(), wd) and np.array_equal(v.cpu().numpy(), wv)


def test_euler_return_all_frames(pkg):
    """Broken upstream (euler_integration_manipulator.py:31,:50); ours returns one entry per step count."""
    rng = np.random.default_rng(4)
    m = rng.uniform(-2, 2, (1, 2, 21, 34)).astype(np.float32)
    d, v = pkg.euler_integration(cu(m), 5, return_all_frames=True)
    assert d.shape == (6, 2, 21, 34) and v.shape == (6, 1, 21, 34)
    for T in range(6):
        wd, wv = oracle.euler(m, T)
        assert np.array_equal(d[T:T + 1].cpu().numpy(), wd) and np.array_equal(v[T:T + 1].cpu().numpy(), wv)


def test_euler_backward_vs_reference_autograd_golden(pkg, golden_euler_grad):
    """The gradient the motion regressor receives with --train_motion
    (models/animating_softmax_splating.py:514-580) equals torch autograd through the reference's
    own eager euler_integration (tests/golden/euler_grad_ref.npz)."""
    e = golden_euler_grad
    keys = sorted(k[:-len("/gmotion")] for k in e.files if k.endswith("/gmotion"))
    for key in keys:
        field, T = key.rsplit("/T", 1)
        m = cu(e[field + "/motion"]).requires_grad_(True)
        d, vis = pkg.euler_integration(m, int(T))
        assert not vis.requires_grad
        if d.requires_grad:
            (d * cu(e[key + "/gdisp"])).sum().backward()
        got = np.zeros_like(e[key + "/gmotion"]) if m.grad is None else m.grad.cpu().numpy()
        assert rel_err(got, e[key + "/gmotion"]) <= 1e-6, key


def test_euler_module_passes_gradients_to_the_motion_producer(pkg):
    """EulerIntegration.forward on a batch built from a differentiable producer: gradients reach it
    (the reference's in-place displacements[b:b+1] = ... keeps the graph)."""
    w = torch.randn(2, 2, 12, 16, device="cuda", requires_grad=True)
    motion = 1.5 * torch.tanh(w)
    d = pkg.EulerIntegration()(motion, torch.tensor([3, 5]))
    d.square().sum().backward()
    assert w.grad is not None and torch.isfinite(w.grad).all() and w.grad.abs().sum() > 0
