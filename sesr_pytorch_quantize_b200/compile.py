"""Host-side "compile" step: collapsed float convs + calibration thresholds -> integer layer
descriptors for libsesrq.

Everything here is one-time, tiny (<= 6400 weights per layer) host arithmetic.  It is done with the
same torch-CPU float32 / Python-double operations the reference uses, so the constants are
bit-identical by construction (checked against the reference's own dumps in tests/golden):

  weights   quantize_symmetrical_by_tensor      myQL/quan_func.py:58-71
  bias      quantize_bias_with_scale            myQL/quan_func.py:391-416, :448-449
  add const clamp16(bias_q - sum(W) * z)        myQL/quan_func.py:481-486
  (M16, n)  quan_layer_between_const            myQL/quan_func.py:495-515, call sites :527/:557/:582/:603/:259
  scale/zero from running min/max               test.py:186-217
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence

import torch

PE = 4
QUAN_BIT = 8
BIAS_BIT = 16
PE_ACC_BIT = 18
PE_ADD_BIT = 20
REQUAN_BIT = 16
REQUAN_N_MAX = 32

# domain each layer's output is requantised into (quan_func.py:525, :601, :556 -> 1, :579 -> 5)
NEXT_DOMAIN = (1, 2, 3, 1, 5)


def quan_layer_between_const(value: float, data_bit: int = REQUAN_BIT, shift_max: int = REQUAN_N_MAX):
    """Fixed-point split value ~= M * 2**-n with M < 2**data_bit (quan_func.py:495-515)."""
    assert data_bit < shift_max, "requan data bit must be less than shift_max"
    value = float(value)
    whole = int(value)
    if whole != 0:
        shift = data_bit - math.ceil(math.log2(whole + 1))
    else:
        shift = data_bit
        probe = value * 2
        while int(probe) == 0:
            shift += 1
            probe *= 2
        shift = min(shift, shift_max)
    return int(value * (2 ** shift)), shift


def symmetric_weight_codes(w: torch.Tensor, width: int = QUAN_BIT):
    """Per-tensor symmetric int codes + Python-double scale (quan_func.py:58-71)."""
    w = w.detach().to("cpu", torch.float32)
    hi = torch.max(w).item()
    lo = torch.min(w).item()
    amax = max(abs(hi), abs(lo))
    assert amax > 0, "Conv2d weight tensor is all zero"
    qmax = 2 ** (width - 1) - 1
    qmin = -(2 ** (width - 1))
    scale = (amax - (0 - amax)) / (qmax - qmin)
    codes = torch.clamp(torch.round(w / scale), min=qmin, max=qmax)
    return codes.to(torch.int8), scale


def bias_codes(bias: torch.Tensor, bias_scale: float, width: int = BIAS_BIT) -> torch.Tensor:
    """clamp(round(b / (s_x * s_w))) to `width` bits (quan_func.py:391-402)."""
    b = torch.as_tensor(bias, dtype=torch.float32).detach().cpu()
    assert torch.max(b).item() != torch.min(b).item(), "bias elements are all equal"
    lim = 2 ** (width - 1)
    return torch.clamp(torch.round(b / bias_scale), min=-lim, max=lim - 1).to(torch.int64)


def thresholds_from_minmax(mins: Sequence[float], maxs: Sequence[float], width: int = QUAN_BIT,
                           force_zero_min_last: bool = True):
    """Running (min, max) per domain -> (scales, zeros) as test.py:186-217 does, including the
    `inp_min = 0` override for the PixelShuffle input domain (test.py:207)."""
    qmax = 2 ** (width - 1) - 1
    qmin = -(2 ** (width - 1))
    scales, zeros = [], []
    last = len(mins) - 1
    for i, (lo, hi) in enumerate(zip(mins, maxs)):
        lo = float(lo)
        hi = float(hi)
        if force_zero_min_last and i == last:
            lo = 0
        s = (hi - lo) / (qmax - qmin)
        scales.append(s)
        zeros.append(int(qmin - round(lo / s)))
    return scales, zeros


@dataclass
class CompiledLayer:
    ksize: int
    cin: int
    cout: int
    relu: bool
    weight: torch.Tensor        # int8 [cout, cin, k, k]
    w_scale: float
    bias: torch.Tensor          # float32 [cout] (as held by the collapsed conv)
    bias_q: torch.Tensor        # int64 [cout]
    add_const: List[int]        # clamp16(bias_q - sum(W) * z_in)
    z_in: int
    m16: int
    shift_n: int


@dataclass
class CompiledNet:
    layers: List[CompiledLayer]
    scales: List[float]
    zeros: List[int]
    res_m16: int
    res_shift_n: int
    r_shuffle: int
    pe_num: int = PE
    pe_acc_bits: int = PE_ACC_BIT
    pe_add_bits: int = PE_ADD_BIT
    quan_bits: int = QUAN_BIT
    bias_bits: int = BIAS_BIT

    @property
    def cin(self) -> int:
        return self.layers[0].cin

    @property
    def cout(self) -> int:
        return self.layers[-1].cout


def compile_network(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                    scales: Sequence[float], zeros: Sequence[int], r_shuffle: int,
                    pe_acc_bits: int = PE_ACC_BIT, pe_add_bits: int = PE_ADD_BIT,
                    bias_bits: int = BIAS_BIT, quan_bits: int = QUAN_BIT) -> CompiledNet:
    """weights[i]: collapsed float conv weight [cout, cin, k, k]; biases[i]: float [cout];
    scales/zeros: the six calibrated activation domains input.{0..5}."""
    if len(weights) != 5 or len(biases) != 5:
        raise ValueError("the SESR/NRDM hot path has exactly 5 collapsed conv layers")
    if len(scales) != 6 or len(zeros) != 6:
        raise ValueError("need 6 activation domains (input.0 .. input.5)")
    pairs = [symmetric_weight_codes(w, quan_bits) for w in weights]
    return compile_from_codes([p[0] for p in pairs], [p[1] for p in pairs], biases, scales, zeros, r_shuffle,
                              pe_acc_bits, pe_add_bits, bias_bits, quan_bits)


def compile_from_codes(codes: Sequence[torch.Tensor], w_scales: Sequence[float], biases: Sequence[torch.Tensor],
                       scales: Sequence[float], zeros: Sequence[int], r_shuffle: int,
                       pe_acc_bits: int = PE_ACC_BIT, pe_add_bits: int = PE_ADD_BIT,
                       bias_bits: int = BIAS_BIT, quan_bits: int = QUAN_BIT) -> CompiledNet:
    """Same as compile_network for weights that are already integer codes (what a graph holds after
    myQL.quan_func.quantize_model_weight)."""
    if len(codes) != 5 or len(biases) != 5 or len(w_scales) != 5:
        raise ValueError("the SESR/NRDM hot path has exactly 5 collapsed conv layers")
    if len(scales) != 6 or len(zeros) != 6:
        raise ValueError("need 6 activation domains (input.0 .. input.5)")
    scales = [float(s) for s in scales]
    zeros = [int(z) for z in zeros]
    lim = 2 ** (bias_bits - 1)
    layers = []
    for i in range(5):
        c = codes[i].detach().to("cpu", torch.int8).contiguous()
        cout, cin, k, k2 = c.shape
        if k != k2 or k not in (3, 5):
            raise ValueError("unsupported kernel size %dx%d" % (k, k2))
        w_scale = float(w_scales[i])
        b = torch.as_tensor(biases[i], dtype=torch.float32).detach().cpu()
        bq = bias_codes(b, scales[i] * w_scale, bias_bits)
        wsum = c.to(torch.int64).sum(dim=(1, 2, 3))
        add_const = torch.clamp(bq - wsum * zeros[i], min=-lim, max=lim - 1)
        m16, n = quan_layer_between_const(scales[i] / scales[NEXT_DOMAIN[i]] * w_scale)
        layers.append(CompiledLayer(ksize=k, cin=cin, cout=cout, relu=(i < 4), weight=c, w_scale=w_scale, bias=b,
                                    bias_q=bq, add_const=[int(v) for v in add_const.tolist()], z_in=zeros[i],
                                    m16=m16, shift_n=n))
    res_m16, res_n = quan_layer_between_const(scales[1] / scales[4])
    return CompiledNet(layers=layers, scales=scales, zeros=zeros, res_m16=res_m16, res_shift_n=res_n,
                       r_shuffle=int(r_shuffle), pe_acc_bits=pe_acc_bits, pe_add_bits=pe_add_bits,
                       quan_bits=quan_bits, bias_bits=bias_bits)


def collapsed_convs(model: torch.nn.Module):
    """(weights, biases) of the nn.Conv2d modules of a collapsed model in state-dict order -- the
    order quantize_model_weight numbers them in (quan_func.py:136-148)."""
    ws, bs = [], []
    for _, mod in model.named_modules():
        if type(mod) is torch.nn.Conv2d:
            ws.append(mod.weight.detach().cpu().float())
            if mod.bias is None:
                raise ValueError("collapsed conv without bias")
            bs.append(mod.bias.detach().cpu().float())
    return ws, bs
