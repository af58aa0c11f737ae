"""Level 0 of the drop-in: the joint block of ``forward_flow`` as the reference's models call it,
on this package's operator modules -- what a user of the reference gets with ZERO source changes
(``install_as_reference_modules()`` and nothing else).

The reference's own code is the authority for this call pattern
(models/animating_softmax_splating.py:847-924, 2-layer variant
models/animating_softmax_splating_2layers_alpha_seperate.py:921-1045); tests/test_reference_drop_in.py
executes that code unmodified.  This module restates the pattern so that it can be timed
(``bench.py --algo level0``) and parity-checked on a box that does not have the reference tree:
two Euler integrations from zero, eager cat / exp glue, two ``ModuleSoftsplat('summation')`` calls,
in-place adds on views of the fresh outputs, clamp, divide.  ``synthesis.JointSplat`` is the fused
path (Level 1) that replaces all of it.
"""
import torch

from .euler_integration_manipulator import euler_integration
from .softsplat import ModuleMaximumWarpNormsplat, ModuleSoftsplat

_softsplater = ModuleSoftsplat('summation')
_maximum_warp_norm_splater = ModuleMaximumWarpNormsplat()


def _alpha(start, mid, end, device):
    # fp32 arithmetic on 0-d tensors like the reference (:860)
    a = 1.0 - torch.tensor(float(mid - start)) / torch.tensor(float(end - start + 1))
    return a.view(1, 1, 1, 1).to(device)


def forward_flow_block(start_fs, Z_f, flow, index, z_mode="max"):
    """gen_fs [1,C,H,W] for index = (start, mid, end); animating_softmax_splating.py:847-924."""
    start_index, middle_index, end_index = [int(v) for v in index]
    forward_flow, _ = euler_integration(flow, middle_index - start_index)              # :847
    backward_flow, _ = euler_integration(-flow, end_index - middle_index + 1)          # :848
    if z_mode == "v2":                                                                  # :849-851
        Z_f_norm = Z_f - _maximum_warp_norm_splater(tenInput=Z_f.contiguous().detach().clone(), tenFlow=forward_flow)
    elif z_mode == "v1":                                                                # :852-853
        Z_f_norm = Z_f
    else:                                                                               # :855
        Z_f_norm = Z_f - Z_f.max()
    alpha = _alpha(start_index, middle_index, end_index, start_fs.device)               # :860
    tenInput_f = torch.cat([start_fs * Z_f_norm.exp() * alpha, Z_f_norm.exp() * alpha], 1)          # :862
    gen_fs_f = _softsplater(tenInput=tenInput_f, tenFlow=forward_flow, tenMetric=None)               # :884
    gen_fs = gen_fs_f[:, :-1, :, :]
    tenNormalize = gen_fs_f[:, -1:, :, :]
    tenInput_p = torch.cat([start_fs * Z_f_norm.exp() * (1 - alpha), Z_f_norm.exp() * (1 - alpha)], 1)  # :895
    gen_fs_p = _softsplater(tenInput=tenInput_p, tenFlow=backward_flow, tenMetric=None)              # :916
    gen_fs += gen_fs_p[:, :-1, :, :]                                                    # :920
    tenNormalize += gen_fs_p[:, -1:, :, :]                                              # :921
    tenNormalize = torch.clamp(tenNormalize, min=1e-8)                                  # :923
    return gen_fs / tenNormalize                                                        # :924


def forward_flow_block_2layer(start_fs, Z_f, alpha_fluid_f, alpha_bg_f, flow, index, alpha0=True):
    """(gen_fs, alpha_fluid, alpha_fluid_mask); 2layers...py:921-1045.  ``alpha_fluid_f``: raw fluid
    channel of the alpha encoder (:946), ``alpha_bg_f``: background alpha after the sigmoid (:948)."""
    start_index, middle_index, end_index = [int(v) for v in index]
    forward_flow, _ = euler_integration(flow, middle_index - start_index)              # :921
    backward_flow, _ = euler_integration(-flow, end_index - middle_index + 1)          # :922
    alpha = torch.clamp(_alpha(start_index, middle_index, end_index, start_fs.device),
                        min=1.0 / 600.0, max=599.0 / 600.0)                             # :950-952
    Z_f_norm = Z_f - Z_f.max()                                                          # :961
    if alpha0:                                                                          # :963-972
        alpha_0_norm = torch.clamp(torch.sigmoid(alpha_fluid_f) + alpha_bg_f, min=1e-8)
        A = torch.sigmoid(alpha_fluid_f) / alpha_0_norm
        chans = [alpha_fluid_f * A.exp(), A.exp(), Z_f_norm.exp()]
    else:                                                                               # :974-976
        chans = [alpha_fluid_f * Z_f_norm.exp(), Z_f_norm.exp()]
    n = len(chans)
    tenInput_f = torch.cat([start_fs * Z_f_norm.exp() * alpha] + [c * alpha for c in chans], 1)
    acc = _softsplater(tenInput=tenInput_f, tenFlow=forward_flow, tenMetric=None)      # :987
    tenInput_p = torch.cat([start_fs * Z_f_norm.exp() * (1 - alpha)] + [c * (1 - alpha) for c in chans], 1)
    acc_p = _softsplater(tenInput=tenInput_p, tenFlow=backward_flow, tenMetric=None)   # :1024
    acc += acc_p                                                                        # :1028-1036
    gen_fs = acc[:, :-n, :, :]
    alpha_fluid = acc[:, -n:-n + 1, :, :]
    tenNormalize = torch.clamp(acc[:, -1:, :, :], min=1e-8)                             # :1038
    alpha_fluid_mask = (tenNormalize > 1e-8).float()                                    # :1039
    gen_fs = gen_fs / tenNormalize                                                      # :1040
    if alpha0:
        alpha_fluid = alpha_fluid / torch.clamp(acc[:, -2:-1, :, :], min=1e-8)          # :1042-1043
    else:
        alpha_fluid = alpha_fluid / tenNormalize                                        # :1045
    return gen_fs, alpha_fluid, alpha_fluid_mask
