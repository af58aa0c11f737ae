"""The decoder's entry, as far as it depends only on WHERE the synthesised features are zero
(SURVEY.md section 8 f1).

``ResNetDecoderPconv2.forward`` starts with ``mask = (x != 0).float()`` over all C channels
(models/networks/architectures.py:369) and hands it to its first block, whose first
``PartialConv2d`` (multi_channel) turns it into

    update_mask = conv2d(mask, ones[out_c, in_c, k, k])          layers/partialconv2d.py:61
    mask_ratio  = slide_winsize / (update_mask + 1e-8)            :64
    update_mask = clamp(update_mask, 0, 1);  mask_ratio *= update_mask      :66-67

-- two full passes over a [C,H,W] tensor and an all-ones convolution over C channels, all of which
only ever see the mask SUMMED over the channels and the window.  The gather knows that sum when it
writes a pixel: ``JointSplat.frames(..., want_nnz=True)`` returns ``nnz[p]`` = number of non-zero
channels of gen_fs at pixel p (exact: counted on the values that are stored).  From it:

    update_mask[:, o] = boxsum_kxk(nnz)          (identical for every output channel o)

so the mask path of the decoder's first partial convolution costs one k x k box filter over ONE plane.
Bit-exact with the reference's layer: sums of at most C*k*k ones are exact in fp32.
"""
import torch
import torch.nn.functional as F


def hole_mask(nnz):
    """Per-pixel hole mask [n,1,H,W]: 1 where at least one channel is non-zero."""
    return (nnz > 0).to(nnz.dtype)


def partialconv_mask_path(nnz, in_channels, kernel_size=3, stride=1, padding=1, dilation=1):
    """(update_mask, mask_ratio), each [n,1,H',W'] and valid for every output channel, of a
    multi-channel ``PartialConv2d(in_channels, *, kernel_size, stride, padding, dilation)`` applied to
    ``mask = (x != 0)`` where ``nnz = mask.sum(1, keepdim=True)`` (layers/partialconv2d.py:41-67)."""
    assert nnz.dim() == 4 and nnz.shape[1] == 1
    k = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
    ones = torch.ones(1, 1, k[0], k[1], dtype=nnz.dtype, device=nnz.device)
    update_mask = F.conv2d(nnz, ones, bias=None, stride=stride, padding=padding, dilation=dilation)      # :61
    slide_winsize = in_channels * k[0] * k[1]                                                             # :37
    mask_ratio = slide_winsize / (update_mask + 1e-8)                                                     # :64
    update_mask = torch.clamp(update_mask, 0, 1)                                                          # :66
    mask_ratio = torch.mul(mask_ratio, update_mask)                                                       # :67
    return update_mask, mask_ratio
