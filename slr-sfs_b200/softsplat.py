"""Drop-in replacement for the reference's ``models/softsplat.py``.

Same public names, argument names and error behaviour; the cupy/NVRTC kernels
are replaced by the precompiled sm_100a library (csrc/, C ABI in
include/slr_splat.h).  Shapes and strides are runtime arguments, so there is no
per-shape JIT (the reference re-templates and recompiles per tensor shape,
softsplat.py:328-386).

    FunctionSoftsplat(tenInput, tenFlow, tenMetric, strType)   softsplat.py:665-690
    ModuleSoftsplat(strType)(tenInput, tenFlow, tenMetric)      softsplat.py:692-702
    ModuleMaximumsplat()(tenInput, tenFlow)                     softsplat.py:705-712
    ModuleMaximumWarpNormsplat()(tenInput, tenFlow)             softsplat.py:716-724
    _FunctionSoftsplat.apply(input, flow)                       softsplat.py:388-479

As in the reference, CPU tensors raise NotImplementedError (softsplat.py:418-419):
there is no CPU path in this package.
"""
import os

import torch

from . import _lib


def _check_pair(input, flow):
    # same assertions as the reference forward (softsplat.py:397-402)
    assert (flow.shape[1] == 2)
    assert (input.shape[2] == flow.shape[2])
    assert (input.shape[3] == flow.shape[3])
    assert (input.is_contiguous() == True)
    assert (flow.is_contiguous() == True)
    if _lib.on_device(input) == False:
        raise NotImplementedError()
    assert input.dtype == torch.float32 and flow.dtype == torch.float32
    assert _lib.on_device(flow) and flow.device == input.device and flow.shape[0] == input.shape[0]


#: elements per batch item (C * H * W) from which the summation splat goes through the gather pipeline instead of
#: fp32 atomics.  Measured with the benchmark scene's mid-clip displacement, 65 channels (profiles/r02/forward_op.json):
#: 768x1024 0.29 against 0.40 ms, 1536x2048 0.87 against 1.15 ms, but 512x512 0.25 against 0.15 ms and 2 x 256x256 0.33
#: against 0.09 ms (the gather's dozen launches and the input interleave cost more than they save on small frames).
#: SLR_SPLAT_GATHER_MIN overrides.
GATHER_MIN_ELEMENTS = int(os.environ.get("SLR_SPLAT_GATHER_MIN", str(1 << 25)))


def _gather_scratch_bytes(C, H, W):
    """Scratch bytes of slr_softsplat_sum_fwd_gather, or 0 when the scatter kernel is the better choice (small
    frames; a non-default SLR_GATHER_MODE; more than 2^27 pixels)."""
    if C * H * W < GATHER_MIN_ELEMENTS or H * W >= (1 << 27):
        return 0
    return _lib.load().slr_softsplat_gather_scratch_bytes(C, H, W)


class _FunctionSoftsplat(torch.autograd.Function):
    """Summation splat with the reference's autograd contract: saves (input, flow),
    returns (gradInput | None, gradFlow | None) by needs_input_grad
    (softsplat.py:388-479)."""

    @staticmethod
    def forward(self, input, flow):
        self.save_for_backward(input, flow)
        _check_pair(input, flow)
        B, C, H, W = input.shape
        # freshly allocated and caller-owned: callers take views of it and += into them
        output = input.new_empty([B, C, H, W])
        with torch.cuda.device(input.device):
            scratch_bytes = _gather_scratch_bytes(C, H, W)
            if scratch_bytes:
                # large frames: bin by destination + gather instead of 4 * C * H * W fp32 atomics (csrc/clip_gather.cu)
                scratch = torch.empty(scratch_bytes // 4 + 64, dtype=torch.float32, device=input.device)
                off = (-scratch.data_ptr() % 256) // 4
                _lib.call("slr_softsplat_sum_fwd_gather", _lib.ptr(input), _lib.ptr(flow), _lib.ptr(output), B, C, H, W,
                          _lib.ptr(scratch[off:]), scratch_bytes, _lib.current_stream(input.device))
            else:
                _lib.call("slr_softsplat_sum_fwd", _lib.ptr(input), _lib.ptr(flow), _lib.ptr(output),
                          B, C, H, W, 1, _lib.current_stream(input.device))
        return output

    @staticmethod
    def backward(self, gradOutput):
        input, flow = self.saved_tensors
        assert (gradOutput.is_contiguous() == True)       # softsplat.py:438
        if _lib.on_device(input) == False:
            raise NotImplementedError()
        B, C, H, W = input.shape
        gradInput = input.new_empty([B, C, H, W]) if self.needs_input_grad[0] == True else None
        gradFlow = input.new_empty([B, 2, H, W]) if self.needs_input_grad[1] == True else None
        with torch.cuda.device(input.device):
            stream = _lib.current_stream(input.device)
            if gradInput is not None:
                _lib.call("slr_softsplat_grad_input", _lib.ptr(flow), _lib.ptr(gradOutput),
                          _lib.ptr(gradInput), B, C, H, W, stream)
            if gradFlow is not None:
                _lib.call("slr_softsplat_grad_flow", _lib.ptr(input), _lib.ptr(flow), _lib.ptr(gradOutput),
                          _lib.ptr(gradFlow), B, C, H, W, stream)
        return gradInput, gradFlow


class _FunctionMaximumsplat(torch.autograd.Function):
    """Max splat into a zero-initialised buffer; forward only, like the reference
    (softsplat.py:482-518 defines no backward)."""

    @staticmethod
    def forward(self, input, flow):
        self.save_for_backward(input, flow)
        _check_pair(input, flow)
        B, C, H, W = input.shape
        output = input.new_empty([B, C, H, W])
        with torch.cuda.device(input.device):
            _lib.call("slr_maxsplat_fwd", _lib.ptr(input), _lib.ptr(flow), _lib.ptr(output), 0.0,
                      B, C, H, W, _lib.current_stream(input.device))
        return output


def _FunctionMaximumWarpNormsplat(input, flow):
    """softsplat.py:576-624: max-splat into a -1000 buffer, then for every source
    pixel the max over its own value and its landing cells.  Plain function (not
    differentiable) in the reference too."""
    _check_pair(input, flow)
    B, C, H, W = input.shape
    scratch = input.new_empty([B, C, H, W])
    output = input.new_empty([B, C, H, W])
    with torch.cuda.device(input.device):
        _lib.call("slr_maxwarpnorm", _lib.ptr(input), _lib.ptr(flow), _lib.ptr(scratch), _lib.ptr(output),
                  B, C, H, W, _lib.current_stream(input.device))
    return output


def FunctionSoftsplat(tenInput, tenFlow, tenMetric, strType):
    assert (tenMetric is None or tenMetric.shape[1] == 1)
    assert (strType in ['summation', 'average', 'linear', 'softmax'])

    if strType == 'average':
        tenInput = torch.cat([tenInput, tenInput.new_ones(tenInput.shape[0], 1, tenInput.shape[2], tenInput.shape[3])], 1)
    elif strType == 'linear':
        tenInput = torch.cat([tenInput * tenMetric, tenMetric], 1)
    elif strType == 'softmax':
        tenInput = torch.cat([tenInput * tenMetric.exp(), tenMetric.exp()], 1)

    tenOutput = _FunctionSoftsplat.apply(tenInput, tenFlow)

    if strType != 'summation':
        tenNormalize = tenOutput[:, -1:, :, :]
        tenNormalize[tenNormalize == 0.0] = 1.0      # exact-zero holes divide by 1 (softsplat.py:684)
        tenOutput = tenOutput[:, :-1, :, :] / tenNormalize
    return tenOutput


class ModuleSoftsplat(torch.nn.Module):
    def __init__(self, strType):
        super(ModuleSoftsplat, self).__init__()
        self.strType = strType

    def forward(self, tenInput, tenFlow, tenMetric):
        return FunctionSoftsplat(tenInput, tenFlow, tenMetric, self.strType)


class ModuleMaximumsplat(torch.nn.Module):
    def __init__(self):
        super(ModuleMaximumsplat, self).__init__()

    def forward(self, tenInput, tenFlow):
        return _FunctionMaximumsplat.apply(tenInput, tenFlow)


class ModuleMaximumWarpNormsplat(torch.nn.Module):
    def __init__(self):
        super(ModuleMaximumWarpNormsplat, self).__init__()

    def forward(self, tenInput, tenFlow):
        return _FunctionMaximumWarpNormsplat(tenInput, tenFlow)
