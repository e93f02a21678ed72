"""Observer module and the reference's hand-written simulation helpers
(mirror of myQL/quan_modules.py:14-201).

myMinMaxObserver keeps running min/max buffers like the reference's (quan_modules.py:14-40); on CUDA
tensors the reduction is libsesrq's sesrq_minmax_update (vectorised loads, warp shuffles, one atomic
per CTA).  conv_forward_naive / sesr_forward_sim are the reference's unused debugging loops: they are
provided for API completeness as plain tensor code on the input's device (not on the hot path).
"""
import ctypes as C

import torch
from torch import nn
import torch.nn.functional as F

from define import PE, BIAS_BIT, PE_ACC_BIT, PE_ADD_BIT, QUAN_BIT  # noqa: F401
from myQL.quan_func import (quantize_asymmetrical_by_tensor, reshape_input_for_hardware_pe, PEs_and_bias_adder,  # noqa: F401
                            requan_conv2d_output, quantize_model_weight)
from myQL.state import state
from sesr_pytorch_quantize_b200 import _lib


class myMinMaxObserver(nn.Module):
    """Records the running minimum and maximum of everything passed through it."""
    min_val: torch.Tensor
    max_val: torch.Tensor

    def __init__(self) -> None:
        super().__init__()
        self.register_buffer("min_val", torch.tensor(float("inf")))
        self.register_buffer("max_val", torch.tensor(float("-inf")))

    def forward(self, x_orig):
        if x_orig.numel() == 0:
            return x_orig
        x = x_orig.detach()
        if not x.is_cuda:
            raise _lib.SesrqError("myMinMaxObserver needs a CUDA tensor: the sesrq ops have no CPU fallback")
        x = x.contiguous().float()
        mm = torch.stack([self.min_val.to(x.device, torch.float32), self.max_val.to(x.device, torch.float32)])
        lib = _lib.load_library()
        with torch.cuda.device(x.device):
            _lib.check(lib.sesrq_minmax_update(C.c_void_p(x.data_ptr()), x.numel(), C.c_void_p(mm.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                       "sesrq_minmax_update")
        self.min_val.copy_(mm[0])
        self.max_val.copy_(mm[1])
        return x_orig

    def reset_min_max_vals(self):
        self.min_val.copy_(torch.tensor(float("inf")))
        self.max_val.copy_(torch.tensor(float("-inf")))


def conv_forward_naive(x, w, stride, pad_num, pad_value):
    """Convolution whose accumulator is clamped to PE_ACC_BIT after EVERY multiply-accumulate, in
    (channel, row, column) order -- the reference's step-wise overflow model (quan_modules.py:43-99).
    x: (N, C, H, W), w: (F, C, HH, WW); returns (N, F, H', W')."""
    N, Cc, H, W = x.shape
    Fo, _, HH, WW = w.shape
    Ho = int(1 + (H + 2 * pad_num - HH) / stride)
    Wo = int(1 + (W + 2 * pad_num - WW) / stride)
    xp = torch.full((N, Cc, H + 2 * pad_num, W + 2 * pad_num), fill_value=float(pad_value), device=x.device)
    xp[:, :, pad_num:pad_num + H, pad_num:pad_num + W] = x
    hi = 2 ** (PE_ACC_BIT - 1) - 1
    lo = -(2 ** (PE_ACC_BIT - 1))
    acc = torch.zeros((N, Fo, Ho, Wo), device=x.device)
    wv = w.to(x.device)
    for c in range(Cc):
        for kh in range(HH):
            for kw in range(WW):
                patch = xp[:, c, kh:kh + (Ho - 1) * stride + 1:stride, kw:kw + (Wo - 1) * stride + 1:stride]
                acc = torch.clamp(acc + patch[:, None] * wv[None, :, c, kh, kw, None, None], min=lo, max=hi)
    return acc


def sesr_forward_sim(input):
    """Hand-unrolled integer simulation of the collapsed SESR x4 network built on conv_forward_naive
    (reference quan_modules.py:101-201; like the reference it applies ReLU after every layer, including the
    last one, and PixelShuffle(4)).  Needs thresholds and weight codes in myQL.state."""
    exe_mode = 1
    t = input
    pads = (2, 1, 1, 1, 2)
    for func_id in range(5):
        t = quantize_asymmetrical_by_tensor(tensor_input=t, width=QUAN_BIT, exe_mode=exe_mode, func_id=func_id)
        weight = state.weight_codes[func_id].to(t.device, torch.float32)
        bias = state.bias[func_id].tolist()
        t = reshape_input_for_hardware_pe(t)
        t = conv_forward_naive(x=t, w=weight, stride=1, pad_num=pads[func_id], pad_value=0)
        t = PEs_and_bias_adder(input_tensor=t, bias=bias, pe_acc_width=PE_ACC_BIT, pe_add_width=PE_ADD_BIT,
                               bias_width=BIAS_BIT, func_id=func_id, pe_num=4, exe_mode=exe_mode)
        t = requan_conv2d_output(input_tensor=t, func_id=func_id, exe_mode=exe_mode)
        t = F.relu(t)
    return nn.PixelShuffle(4)(t)
