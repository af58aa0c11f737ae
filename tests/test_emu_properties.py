"""CPU: property-based and edge-case tests of the kernel sources (through tests/emu) against the
oracle -- SURVEY.md section 8c's "hypothesis-driven random shapes / flows", plus degenerate
geometry the GPU suite does not spend device time on."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
from conftest import rel_err

emu = pytest.importorskip("emu", reason="tests/emu")
TOL = 1e-4
COMMON = dict(deadline=None, max_examples=60, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large,
                                                                  HealthCheck.function_scoped_fixture])


def _motion(kind, H, W, rng, amp):
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    if kind == "random":
        m = rng.uniform(-amp, amp, (1, 2, H, W))
    elif kind == "smooth":
        m = np.stack([amp * np.sin(xs / 5.0 + 0.3) * np.cos(ys / 4.0), amp * np.cos(xs / 6.0) * np.sin(ys / 3.0 + 0.1)])[None]
    elif kind == "constant":
        m = np.broadcast_to(np.array([amp, -amp / 2]).reshape(1, 2, 1, 1), (1, 2, H, W))
    elif kind == "integer":
        m = np.round(rng.uniform(-amp, amp, (1, 2, H, W)))
    elif kind == "half":
        m = np.round(rng.uniform(-amp, amp, (1, 2, H, W))) + 0.5
    else:                               # "patchy": half of the pixels exactly static
        m = rng.uniform(-amp, amp, (1, 2, H, W)) * (rng.uniform(0, 1, (1, 1, H, W)) > 0.5)
    return np.ascontiguousarray(m, dtype=np.float32)


@settings(**COMMON)
@given(H=st.integers(1, 40), W=st.integers(1, 70), C=st.integers(1, 9), B=st.integers(1, 2),
       kind=st.sampled_from(["random", "smooth", "constant", "integer", "half", "patchy"]),
       amp=st.sampled_from([0.0, 0.75, 3.0, 9.0, 100.0]), seed=st.integers(0, 2 ** 16))
def test_operator_kernels_match_the_oracle(H, W, C, B, kind, amp, seed):
    rng = np.random.default_rng(seed)
    inp = rng.standard_normal((B, C, H, W)).astype(np.float32)
    flow = np.concatenate([_motion(kind, H, W, rng, amp) for _ in range(B)], 0)
    gout = rng.standard_normal((B, C, H, W)).astype(np.float32)
    out, gin, gflow = np.empty_like(inp), np.empty_like(inp), np.empty_like(flow)
    emu.call("slr_softsplat_sum_fwd", emu.p(inp), emu.p(flow), emu.p(out), B, C, H, W, 1, None)
    emu.call("slr_softsplat_grad_input", emu.p(flow), emu.p(gout), emu.p(gin), B, C, H, W, None)
    emu.call("slr_softsplat_grad_flow", emu.p(inp), emu.p(flow), emu.p(gout), emu.p(gflow), B, C, H, W, None)
    assert rel_err(out, oracle.softsplat_sum(inp, flow)) <= TOL
    assert rel_err(gin, oracle.softsplat_grad_input(flow, gout)) <= TOL
    assert rel_err(gflow, oracle.softsplat_grad_flow(inp, flow, gout)) <= TOL
    scratch, mw = np.empty_like(inp), np.empty_like(inp)
    emu.call("slr_maxwarpnorm", emu.p(inp), emu.p(flow), emu.p(scratch), emu.p(mw), B, C, H, W, None)
    assert np.array_equal(mw, oracle.max_warp_norm(inp, flow))


@settings(**COMMON)
@given(H=st.integers(1, 33), W=st.integers(1, 50), T=st.integers(0, 40),
       kind=st.sampled_from(["random", "smooth", "constant", "integer", "half", "patchy"]),
       amp=st.sampled_from([0.5, 2.5, 7.0]), seed=st.integers(0, 2 ** 16), sign=st.sampled_from([1.0, -1.0]))
def test_euler_kernel_is_bit_exact(H, W, T, kind, amp, seed, sign):
    motion = _motion(kind, H, W, np.random.default_rng(seed), amp)
    disp = np.empty((1, 2, H, W), dtype=np.float32)
    vis = np.empty((1, 1, H, W), dtype=np.float32)
    emu.call("slr_euler", emu.p(motion), sign, T, emu.p(disp), emu.p(vis), H, W, None)
    want_d, want_v = oracle.euler(np.float32(sign) * motion, T)
    assert np.array_equal(disp, want_d) and np.array_equal(vis, want_v)


@settings(**dict(COMMON, max_examples=40))
@given(H=st.integers(1, 26), W=st.integers(1, 70), C=st.integers(1, 9), N=st.integers(1, 12),
       kind=st.sampled_from(["random", "smooth", "constant", "integer", "half", "patchy"]),
       amp=st.sampled_from([0.0, 0.75, 3.0, 9.0]), seed=st.integers(0, 2 ** 16), data=st.data(),
       shape=st.sampled_from(["1x4", "2x2"]))
def test_clip_pipeline_matches_the_oracle(H, W, C, N, kind, amp, seed, data, shape, monkeypatch):
    monkeypatch.setenv("SLR_GATHER_SHAPE", shape)
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((1, C, H, W)).astype(np.float32)
    Z = rng.standard_normal((1, 1, H, W)).astype(np.float32)
    motion = _motion(kind, H, W, rng, amp)
    start = data.draw(st.integers(0, 3))
    end = start + N - 1
    t0 = data.draw(st.integers(start, end))
    n = data.draw(st.integers(1, min(4, end - t0 + 1)))
    sc = emu.Scene(feat, Z, motion, z_mode=data.draw(st.sampled_from(["max", "v1"])))
    use_table = data.draw(st.booleans())
    tab = sc.table(start, end, start, N) if use_table else None
    got = sc.frames(start, end, t0, n, table=tab, split=not use_table)
    for i in range(n):
        want = oracle.joint_splat_baseline(feat, Z, motion, (start, t0 + i, end), z_mode=sc.zsub is not None and "max" or "v1")
        assert rel_err(got[i:i + 1], want) <= TOL
        assert np.all(got[i:i + 1][want == 0.0] == 0.0)


def test_one_pixel_and_one_row_images():
    for (H, W) in [(1, 1), (1, 9), (9, 1), (2, 2)]:
        feat = np.arange(3 * H * W, dtype=np.float32).reshape(1, 3, H, W) + 1.0
        Z = np.zeros((1, 1, H, W), dtype=np.float32)
        for m in (np.zeros((1, 2, H, W), np.float32), np.full((1, 2, H, W), 0.25, np.float32)):
            sc = emu.Scene(feat, Z, m)
            got = sc.frames(0, 3, 0, 4)
            for t in range(4):
                want = oracle.joint_splat_baseline(feat, Z, m, (0, t, 3))
                assert rel_err(got[t:t + 1], want) <= TOL, (H, W, t)


def test_frame_past_the_end_of_the_clip():
    """t = end + 1 is legal at the C ABI (backward chain of zero steps)."""
    H, W, C = 12, 20, 3
    rng = np.random.default_rng(3)
    feat = rng.standard_normal((1, C, H, W)).astype(np.float32)
    Z = rng.standard_normal((1, 1, H, W)).astype(np.float32)
    m = _motion("smooth", H, W, rng, 2.0)
    got = emu.Scene(feat, Z, m).frames(0, 5, 5, 2)
    for i, t in enumerate((5, 6)):
        assert rel_err(got[i:i + 1], oracle.joint_splat_baseline(feat, Z, m, (0, t, 5))) <= TOL


def test_non_finite_flow_is_dropped_not_crashing():
    """The reference's behaviour on NaN / inf flow is undefined (float -> int casts); ours drops
    such pixels.  Finite pixels must be unaffected."""
    H, W, C = 10, 16, 2
    rng = np.random.default_rng(5)
    inp = rng.standard_normal((1, C, H, W)).astype(np.float32)
    flow = rng.uniform(-2, 2, (1, 2, H, W)).astype(np.float32)
    clean = flow.copy()
    bad = [(0, 3), (4, 7), (9, 15)]
    vals = [np.nan, np.inf, -np.inf]
    for (y, x), v in zip(bad, vals):
        flow[0, 0, y, x] = v
    out = np.empty_like(inp)
    emu.call("slr_softsplat_sum_fwd", emu.p(inp), emu.p(flow), emu.p(out), 1, C, H, W, 1, None)
    masked = inp.copy()
    for (y, x) in bad:
        masked[0, :, y, x] = 0.0
    assert rel_err(out, oracle.softsplat_sum(masked, clean)) <= TOL


def test_non_finite_motion_invalidates_the_chain_not_the_context():
    """NaN / inf motion (a diverging motion network): the reference raises an index error at the
    next step; ours marks every chain that meets such a value invalid (sentinel displacement, not
    visible) and must never index the field out of range -- slr_euler and slr_clip_table alike.
    Chains that never meet a bad pixel are bit-identical to the clean field's."""
    H, W, T = 12, 18, 6
    rng = np.random.default_rng(7)
    motion = rng.uniform(-1.5, 1.5, (1, 2, H, W)).astype(np.float32)
    bad = [(2, 3, np.nan), (5, 9, np.inf), (8, 14, -np.inf)]
    dirty = motion.copy()
    for (y, x, v) in bad:
        dirty[0, 0, y, x] = v
    dirty[0, 1, 10, 1] = np.nan
    disp = np.empty((1, 2, H, W), dtype=np.float32)
    vis = np.empty((1, 1, H, W), dtype=np.float32)
    for sign in (1.0, -1.0):
        emu.call("slr_euler", emu.p(dirty), sign, T, emu.p(disp), emu.p(vis), H, W, None)
        assert np.isfinite(disp).all()
        want_d, want_v = oracle.euler(np.float32(sign) * motion, T)
        # chains that pass through a bad pixel are invalid; find them by perturbing the clean field there
        probe = motion.copy()
        for (y, x, _) in bad:
            probe[0, 0, y, x] += 1000.0
        probe[0, 1, 10, 1] += 1000.0
        other_d, _ = oracle.euler(np.float32(sign) * probe, T)
        untouched = np.all(other_d == want_d, axis=1, keepdims=True)
        assert np.array_equal(disp[np.broadcast_to(untouched, disp.shape)], want_d[np.broadcast_to(untouched, disp.shape)])
        assert np.array_equal(vis[untouched], want_v[untouched])
        for (y, x, _) in bad:
            assert vis[0, 0, y, x] == 0.0 and disp[0, 0, y, x] == max(H, W) + 1
    # the clip table walks the same chains: it must survive the same field (ASan build checks the bounds)
    feat = rng.standard_normal((1, 3, H, W)).astype(np.float32)
    Z = rng.standard_normal((1, 1, H, W)).astype(np.float32)
    out = emu.Scene(feat, Z, dirty).frames(0, T, 0, T + 1)
    assert np.isfinite(out).all()


@settings(**dict(COMMON, max_examples=40))
@given(H=st.integers(1, 40), W=st.integers(1, 70), C=st.integers(1, 9), B=st.integers(1, 2),
       kind=st.sampled_from(["random", "smooth", "constant", "integer", "half", "patchy"]),
       amp=st.sampled_from([0.0, 0.75, 3.0, 9.0, 100.0]), seed=st.integers(0, 2 ** 16))
def test_summation_splat_through_the_gather_matches_the_oracle(H, W, C, B, kind, amp, seed):
    """slr_softsplat_sum_fwd_gather (scene prep + flow table + insert + un-normalised gather) on random shapes /
    flows: 1-pixel images, all-static and all-leaving flows, integer and half-pixel landings."""
    rng = np.random.default_rng(seed)
    inp = rng.standard_normal((B, C, H, W)).astype(np.float32)
    flow = np.concatenate([_motion(kind, H, W, rng, amp) for _ in range(B)], 0)
    nb = emu.lib().slr_softsplat_gather_scratch_bytes(C, H, W)
    scratch = emu.aligned(nb)
    scratch[:] = 0xA5
    out = np.full_like(inp, np.nan)
    emu.call("slr_softsplat_sum_fwd_gather", emu.p(inp), emu.p(flow), emu.p(out), B, C, H, W, emu.p(scratch), nb, None)
    want = oracle.softsplat_sum(inp, flow)
    assert rel_err(out, want) <= TOL
    assert np.all(out[want == 0.0] == 0.0)


@settings(**dict(COMMON, max_examples=40))
@given(H=st.integers(1, 30), W=st.integers(1, 50), C=st.integers(1, 7), B=st.integers(1, 2),
       kind=st.sampled_from(["random", "smooth", "constant", "integer", "half", "patchy"]),
       amp=st.sampled_from([0.0, 0.75, 3.0, 9.0, 100.0]), seed=st.integers(0, 2 ** 16))
def test_producer_splat_matches_the_oracle(H, W, C, B, kind, amp, seed):
    """slr_producer_splat_fwd / _bwd (the fused training producer) on random shapes / flows."""
    rng = np.random.default_rng(seed)
    fs = rng.standard_normal((B, C, H, W)).astype(np.float32)
    zn = np.clip(rng.standard_normal((B, 1, H, W)) * 3 - 2, -20, 20).astype(np.float32)
    flow = np.concatenate([_motion(kind, H, W, rng, amp) for _ in range(B)], 0)
    alpha = rng.uniform(0.0, 1.0, B).astype(np.float32)
    gacc = rng.standard_normal((B, C + 1, H, W)).astype(np.float32)
    acc = np.full((B, C + 1, H, W), np.nan, np.float32)
    emu.call("slr_producer_splat_fwd", emu.p(fs), emu.p(zn), emu.p(flow), emu.p(alpha), emu.p(acc), B, C, H, W, 0, None)
    assert rel_err(acc, oracle.producer_splat(fs, zn, flow, alpha)[0]) <= TOL
    d_fs, d_zn, d_flow = np.full_like(fs, np.nan), np.full_like(zn, np.nan), np.full_like(flow, np.nan)
    emu.call("slr_producer_splat_bwd", emu.p(fs), emu.p(zn), emu.p(flow), emu.p(alpha), emu.p(gacc),
             emu.p(d_fs), emu.p(d_zn), emu.p(d_flow), B, C, H, W, None)
    w_fs, w_zn, w_flow = oracle.producer_splat_grads(fs, zn, flow, alpha, gacc)
    assert rel_err(d_fs, w_fs) <= TOL and rel_err(d_zn, w_zn) <= TOL and rel_err(d_flow, w_flow) <= TOL


def test_non_finite_flow_through_the_new_paths():
    """NaN / inf flow: those sources miss the frame (as in the reference's range tests), nothing else is disturbed."""
    rng = np.random.default_rng(3)
    B, C, H, W = 1, 3, 12, 20
    inp = rng.standard_normal((B, C, H, W)).astype(np.float32)
    flow = rng.uniform(-3, 3, (B, 2, H, W)).astype(np.float32)
    flow[0, 0, 2, 3] = np.nan
    flow[0, 1, 5, 7] = np.inf
    flow[0, 0, 8, 1] = -np.inf
    nb = emu.lib().slr_softsplat_gather_scratch_bytes(C, H, W)
    out = np.full_like(inp, np.nan)
    emu.call("slr_softsplat_sum_fwd_gather", emu.p(inp), emu.p(flow), emu.p(out), B, C, H, W, emu.p(emu.aligned(nb)), nb, None)
    ref = np.full_like(inp, np.nan)
    emu.call("slr_softsplat_sum_fwd", emu.p(inp), emu.p(flow), emu.p(ref), B, C, H, W, 1, None)
    assert np.isfinite(out).all() and rel_err(out, ref) <= 1e-5
    zn = rng.standard_normal((B, 1, H, W)).astype(np.float32)
    alpha = np.array([0.4], np.float32)
    acc = np.full((B, C + 1, H, W), np.nan, np.float32)
    emu.call("slr_producer_splat_fwd", emu.p(inp), emu.p(zn), emu.p(flow), emu.p(alpha), emu.p(acc), B, C, H, W, 0, None)
    assert np.isfinite(acc).all()
    d_fs, d_zn = np.full_like(inp, np.nan), np.full_like(zn, np.nan)
    emu.call("slr_producer_splat_bwd", emu.p(inp), emu.p(zn), emu.p(flow), emu.p(alpha), emu.p(np.ones_like(acc)),
             emu.p(d_fs), emu.p(d_zn), None, B, C, H, W, None)
    assert np.isfinite(d_fs).all() and np.isfinite(d_zn).all()
    assert d_fs[0, :, 2, 3].tolist() == [0.0] * C and d_fs[0, :, 5, 7].tolist() == [0.0] * C
