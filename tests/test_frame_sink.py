"""the frame sink against the reference's per-frame host code
(test_animating/test_v1_4eval_rawsize.py:240-242, 284-286), restated with the same torch / numpy calls
(cv2.imwrite's float -> uint8 conversion is saturate_cast<uchar>(cvRound(x)): round half to even, clamp)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def reference_sink(pred, out_size, alpha=False):
    x = F.interpolate(pred, out_size, mode='bilinear')                        # :241
    if alpha:
        img = x.permute(0, 2, 3, 1).cpu().numpy() * 255                          # :245
    else:
        img = (x.permute(0, 2, 3, 1).cpu().numpy() * 0.5 + 0.5) * 255            # :242
    img = img[..., ::-1]                                                         # cv2.cvtColor(RGB2BGR), :286
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)                        # cv2.imwrite


def test_emulated_kernel_matches_the_reference_sink():
    emu = pytest.importorskip("emu", reason="tests/emu")
    g = torch.Generator().manual_seed(3)
    pred = torch.tanh(torch.randn(2, 3, 19, 31, generator=g) * 1.5)
    pred[0, :, 0, :4] = torch.tensor([1.0, -1.0, 0.0, 0.999])[None]
    pred[1, 0, 1, 1] = 3.0            # out of range saturates
    pred[1, 1, 1, 1] = -3.0
    for out_size in [(19, 31), (38, 62), (25, 40), (10, 17)]:
        out = np.zeros((2, out_size[0], out_size[1], 3), np.uint8)
        x = emu.f32(pred.numpy())
        emu.call("slr_frame_sink_u8", emu.p(x), emu.p(out), 2, 19, 31, out_size[0], out_size[1], 0.5, 0.5, 1, None)
        want = reference_sink(pred, out_size)
        diff = np.abs(out.astype(np.int16) - want.astype(np.int16))
        if out_size == (19, 31):
            assert np.array_equal(out, want)
        else:                          # a product that lands within an ulp of .5 may round the other way
            assert diff.max() <= 1 and (diff != 0).mean() < 1e-3


@pytest.mark.gpu
def test_frame_sink_ring_on_the_gpu():
    import __graft_entry__
    __graft_entry__.build()
    import slr_sfs_b200 as pkg
    H, W, out_size = 96, 128, (150, 200)
    g = torch.Generator().manual_seed(5)
    sink = pkg.FrameSink(H, W, "cuda", out_size=out_size, group=4, slots=3)
    frames = [torch.tanh(torch.randn(4 if i < 4 else 2, 3, H, W, generator=g)) for i in range(5)]
    got = {}
    for i, fr in enumerate(frames):
        if len(sink.pending) == 3:
            first, imgs = sink.pop()
            got[first] = imgs.copy()
        sink.push(fr.cuda())
    while True:
        item = sink.pop()
        if item is None:
            break
        got[item[0]] = item[1].copy()
    first = 0
    for fr in frames:
        want = reference_sink(fr, out_size)
        diff = np.abs(got[first].astype(np.int16) - want.astype(np.int16))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
        first += fr.shape[0]
    # no resize, alpha scaling, RGB order: exact
    a = torch.rand(1, 3, H, W, generator=g)
    out = pkg.frame_sink.to_u8(a.cuda(), mul=1.0, add=0.0, bgr=False).cpu().numpy()
    assert np.array_equal(out, np.clip(np.rint(a.permute(0, 2, 3, 1).numpy() * 255), 0, 255).astype(np.uint8))
