"""CPU: the oracle against the reference's own outputs (tests/golden) and against
hand-derivable known answers."""
import numpy as np
import pytest

import oracle
from conftest import rel_err

SPLAT_CASES = ["zero", "int_shift", "half", "neg_frac", "uniform4", "uniform_big",
               "onto_last_cell", "sink", "sentinel"]


@pytest.mark.parametrize("case", SPLAT_CASES)
def test_splat_matches_reference_kernels(golden_softsplat, case):
    g = golden_softsplat
    inp, flow, gout, z = (g[f"{case}/{k}"] for k in ("inp", "flow", "gout", "z"))
    # single-threaded reference and oracle visit sources in the same order: bit-exact
    assert np.array_equal(oracle.softsplat_sum(inp, flow), g[f"{case}/sum"])
    assert np.array_equal(oracle.softsplat_grad_input(flow, gout), g[f"{case}/gin"])
    assert np.array_equal(oracle.softsplat_grad_flow(inp, flow, gout), g[f"{case}/gflow"])
    assert np.array_equal(oracle.max_warp_norm(z, flow), g[f"{case}/maxwarpnorm"])


def test_euler_matches_reference_python(golden_euler):
    e = golden_euler
    n = 0
    for k in e.files:
        if not k.endswith("/disp"):
            continue
        base, T, _ = k.rsplit("/", 2)
        d, v = oracle.euler(e[base + "/motion"], int(T[1:]))
        assert np.array_equal(d, e[k]), k
        assert np.array_equal(v, e[k[:-4] + "vis"]), k
        n += 1
    assert n > 50


def test_joint_blocks_match_reference_models(golden_joint):
    j = golden_joint
    N = int(j["N"])
    ao = j["alpha_encoder_out"]
    a_bg = (1.0 / (1.0 + np.exp(-ao[:, 0:1]))).astype(np.float32)
    a_fl = ao[:, 1:2]
    seen = 0
    for k in j.files:
        parts = k.split("/")
        if parts[0] == "baseline":
            t = int(parts[2][1:])
            o = oracle.joint_splat_baseline(j["feat"], j["Z"], j["motion"], (0, t, N - 1), z_mode=parts[1])
            assert rel_err(o, j[k]) < 2e-6, k      # numpy exp vs torch exp: <= 1 ulp
            seen += 1
        elif parts[0] == "twolayer" and parts[-1] == "gen_fs":
            t = int(parts[2][1:])
            gen, af, _ = oracle.joint_splat_2layer(j["feat"], j["Z"], a_fl, a_bg, j["motion"],
                                                    (0, t, N - 1), alpha0=(parts[1] == "alpha0"))
            assert rel_err(gen, j[k]) < 2e-6, k
            assert rel_err(af, j[k.replace("gen_fs", "alpha_fluid")]) < 2e-6, k
            seen += 1
    assert seen == 15


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
def test_threaded_reference_within_tolerance():
    rng = np.random.default_rng(3)
    inp = rng.standard_normal((1, 5, 40, 37)).astype(np.float32)
    flow = rng.uniform(-6, 6, (1, 2, 40, 37)).astype(np.float32)
    seq = oracle.ref_softsplat_sum(inp, flow, threads=1)
    par = oracle.ref_softsplat_sum(inp, flow, threads=4)
    assert np.array_equal(seq, oracle.softsplat_sum(inp, flow))
    assert rel_err(par, seq) < 1e-5          # atomics reorder the fp32 sums


# ---- known answers derivable by hand from softsplat.py:164-200 -----------------
def _rand(shape, seed=0):
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)


def test_zero_flow_is_identity():
    x = _rand((2, 3, 7, 9))
    assert np.array_equal(oracle.softsplat_sum(x, np.zeros((2, 2, 7, 9), np.float32)), x)


def test_integer_shift_drops_out_of_frame():
    x = _rand((1, 2, 6, 8))
    flow = np.zeros((1, 2, 6, 8), np.float32)
    flow[:, 0] = 3.0
    flow[:, 1] = -2.0
    out = oracle.softsplat_sum(x, flow)
    want = np.zeros_like(x)
    want[:, :, 0:4, 3:8] = x[:, :, 2:6, 0:5]
    assert np.array_equal(out, want)


def test_half_pixel_quarters():
    x = np.zeros((1, 1, 5, 5), np.float32)
    x[0, 0, 2, 2] = 8.0
    out = oracle.softsplat_sum(x, np.full((1, 2, 5, 5), 0.5, np.float32))
    assert out[0, 0, 2, 2] == 2.0 and out[0, 0, 2, 3] == 2.0
    assert out[0, 0, 3, 2] == 2.0 and out[0, 0, 3, 3] == 2.0
    assert out.sum() == 8.0


def test_negative_fraction_uses_floor_not_truncation():
    x = np.zeros((1, 1, 4, 4), np.float32)
    x[0, 0, 0, 0] = 1.0
    flow = np.zeros((1, 2, 4, 4), np.float32)
    flow[0, 0, 0, 0] = -0.25            # lands at x = -0.25: NW cell is -1 (dropped), NE cell 0 gets 0.75
    out = oracle.softsplat_sum(x, flow)
    assert out[0, 0, 0, 0] == np.float32(0.75) and out.sum() == np.float32(0.75)


def test_collision_sums_and_holes_stay_zero():
    x = np.ones((1, 1, 3, 3), np.float32)
    ys, xs = np.meshgrid(np.arange(3, dtype=np.float32), np.arange(3, dtype=np.float32), indexing="ij")
    flow = np.stack([1 - xs, 1 - ys])[None]
    out = oracle.softsplat_sum(x, flow)
    assert out[0, 0, 1, 1] == 9.0 and np.count_nonzero(out) == 1


def test_softmax_uniform_metric_equals_average():
    x = _rand((1, 3, 9, 9), 1)
    flow = np.random.default_rng(2).uniform(-2, 2, (1, 2, 9, 9)).astype(np.float32)
    a = oracle.function_softsplat(x, flow, None, "average")
    s = oracle.function_softsplat(x, flow, np.zeros((1, 1, 9, 9), np.float32), "softmax")
    assert np.array_equal(a, s)


def test_hole_normalisation_zero_becomes_one():
    x = np.ones((1, 1, 4, 4), np.float32)
    flow = np.full((1, 2, 4, 4), 100.0, np.float32)     # everything leaves the frame
    out = oracle.function_softsplat(x, flow, None, "average")
    assert np.array_equal(out, np.zeros_like(out))       # 0 / 1, not 0 / 0


def test_grads_match_finite_differences():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 2, 6, 7)).astype(np.float32)
    flow = rng.uniform(-1.5, 1.5, (1, 2, 6, 7)).astype(np.float32)
    flow = np.where(np.abs(flow - np.round(flow)) < 0.05, flow + 0.1, flow).astype(np.float32)
    gout = rng.standard_normal(x.shape).astype(np.float32)

    def loss(xx, ff):
        return float((oracle.softsplat_sum_f64(xx, ff) * gout).sum())

    gin = oracle.softsplat_grad_input(flow, gout)
    gfl = oracle.softsplat_grad_flow(x, flow, gout)
    eps = 1e-3
    for idx in [(0, 0, 2, 3), (0, 1, 5, 6), (0, 1, 0, 0)]:
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        assert abs((loss(xp, flow) - loss(xm, flow)) / (2 * eps) - gin[idx]) < 2e-3
    for idx in [(0, 0, 2, 3), (0, 1, 4, 1)]:
        fp, fm = flow.copy(), flow.copy()
        fp[idx] += eps
        fm[idx] -= eps
        assert abs((loss(x, fp) - loss(x, fm)) / (2 * eps) - gfl[idx]) < 5e-3


# ---- Euler known answers (euler_integration_manipulator.py:18-56) ------------------
def test_euler_zero_steps():
    d, v = oracle.euler(_rand((1, 2, 5, 6)), 0)
    assert not d.any() and v.all()


def test_euler_constant_flow_until_exit():
    H, W, T = 8, 12, 3
    m = np.zeros((1, 2, H, W), np.float32)
    m[0, 0] = 1.25
    m[0, 1] = -0.5
    d, v = oracle.euler(m, T)
    sentinel = max(H, W) + 1
    for y in range(H):
        for x in range(W):
            ok = all(0 <= x + 1.25 * k <= W - 1 and 0 <= y - 0.5 * k <= H - 1 for k in range(1, T + 1))
            if ok:
                assert d[0, 0, y, x] == np.float32(x + 3.75) - np.float32(x) and v[0, 0, y, x] == 1
            else:
                assert d[0, 0, y, x] == sentinel and d[0, 1, y, x] == sentinel and v[0, 0, y, x] == 0


def test_euler_rounds_half_to_even():
    # one row; pixel 0 steps by 0.5: positions 0.5 -> index 0, 1.5 -> 2, 2.5 -> 2
    W = 6
    m = np.zeros((1, 2, 1, W), np.float32)
    m[0, 0, 0, :] = [0.5, 100.0, 1.0, 0.0, 0.0, 0.0]
    d, _ = oracle.euler(m, 2)       # 0 -> 0.5 (reads idx 0 again: rne(0.5)=0) -> 1.0
    assert d[0, 0, 0, 0] == 1.0
    d, _ = oracle.euler(m, 3)       # 1.0 reads idx 1 (=100) -> leaves the frame
    assert d[0, 0, 0, 0] == W + 1
    m[0, 0, 0, :] = [1.5, 7.0, 1.0, 0.0, 0.0, 0.0]
    d, _ = oracle.euler(m, 2)       # 0 -> 1.5; rne(1.5) = 2 -> +1.0 = 2.5
    assert d[0, 0, 0, 0] == 2.5
    d, _ = oracle.euler(m, 3)       # rne(2.5) = 2 -> +1.0 = 3.5
    assert d[0, 0, 0, 0] == 3.5


def test_euler_edge_is_valid_and_invalid_is_sticky():
    W = 5
    m = np.zeros((1, 2, 1, W), np.float32)
    m[0, 0, 0, :] = [4.0, 0, 0, 0, 0]         # pixel 0 lands exactly on W-1: valid (strict >)
    d, v = oracle.euler(m, 1)
    assert d[0, 0, 0, 0] == 4.0 and v[0, 0, 0, 0] == 1
    m[0, 0, 0, :] = [9.0, 0, 0, 0, 0]         # leaves; later steps restart from 0 but it stays invalid
    d, v = oracle.euler(m, 1)
    assert v[0, 0, 0, 0] == 0
    m2 = m.copy()
    m2[0, 0, 0, 0] = 9.0
    d, v = oracle.euler(m2, 4)
    assert v[0, 0, 0, 0] == 0 and d[0, 0, 0, 0] == W + 1 and d[0, 1, 0, 0] == W + 1


def test_euler_accepts_one_element_array(golden_euler):
    d, v = oracle.euler(golden_euler["tensorT/motion"], np.array([4]))
    assert np.array_equal(d, golden_euler["tensorT/T4/disp"])


def test_euler_grad_matches_reference_autograd(golden_euler_grad):
    """orc_euler_grad_motion against torch autograd through the reference's own euler_integration
    (tests/golden/make_golden.py::make_euler_grad).  Sums of few fp32 terms in a different order:
    1e-6 relative."""
    e = golden_euler_grad
    keys = sorted(k[:-len("/gmotion")] for k in e.files if k.endswith("/gmotion"))
    assert len(keys) >= 40
    nonzero = 0
    for key in keys:
        field, T = key.rsplit("/T", 1)
        got = oracle.euler_grad_motion(e[field + "/motion"], int(T), e[key + "/gdisp"])
        want = e[key + "/gmotion"]
        assert rel_err(got, want) <= 1e-6, key
        assert np.array_equal(got == 0.0, want == 0.0), key
        nonzero += int(np.any(want != 0.0))
    assert nonzero >= 30


def test_warp_flow_block_matches_reference_model(golden_warp_flow):
    """oracle.warp_flow_block against AnimatingSoftmaxSplating.warp_flow imported unmodified
    (tests/golden/make_golden.py::make_warp_flow)."""
    g = golden_warp_flow
    N = int(g["N"])
    for t in (0, 3, N - 2):
        got = oracle.warp_flow_block(g["img"], g["flow_f"][t:t + 1], g["flow_p"][N - 1 - t:N - t], (0, t, N - 1),
                                     splat=oracle.ref_softsplat_sum if oracle.ref_available() else oracle.softsplat_sum)
        assert rel_err(got, g[f"t{t}/PredImg"]) <= 2e-6, t
