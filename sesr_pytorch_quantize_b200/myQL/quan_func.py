"""Quantisation ops inserted around every nn.Conv2d (mirror of the reference's myQL/quan_func.py).

Same public names, signatures and conventions as the reference (exe_mode 0 = fake-quant "staged" float,
exe_mode 1 = integer sim; tensors are float32 NCHW holding integer codes in mode 1; `func_id` is the conv
index in graph order, 5 = the output domain).  Differences:
  * every tensor op runs in libsesrq's CUDA kernels (one per reference callable and exe_mode, include/sesrq.h
    "stand-alone myQL ops"); tensors must live on a CUDA device -- there is no CPU fallback;
  * inter-node state lives in myQL.state instead of ./output_pt files (files stay as a compat mirror);
  * graphs built by myQL.graph_modify normally do not execute these node functions at all: the complete
    5-layer chain is recognised and replaced by one fused network call (graph_modify.FusedGraphModule).
The one-time host-side pieces (weight / bias quantisation, (M16, n) split) are shared with
sesr_pytorch_quantize_b200.compile.
"""
import copy
import ctypes as C
import math
import os

import torch
from torch import Tensor, nn

from define import (WEIGHT_W_FLG, QUAN_BIT, WEIGHT_W_HIST_PNG, INPUT_W_HIST_PNG, OUTPUT_PE_W_FLG,  # noqa: F401
                    OUTPUT_PE_ADD_W_FLG, INPUT_W_FLG, BIAS_W_FLG, REQUAN_BIT, REQUAN_N_MAX, BIAS_QUAN_W_FLG,
                    REQUAN_FACTOR_W_FLG, MFLAG)
from myQL.state import state
from sesr_pytorch_quantize_b200 import _lib
from sesr_pytorch_quantize_b200 import compile as _cp

quan_layer_between_const = _cp.quan_layer_between_const


def remove_suffix(string):
    return string.rsplit(".", 1)[0] if "." in string else ""


def float_to_hex(item, bit_width):
    """signed integer value -> two's-complement hex string of ceil(bit_width / 4) digits (min 2)."""
    value = int(item)
    if value < 0:
        value += 2 ** bit_width
    digits = math.ceil(bit_width / 4)
    if digits not in (3, 4, 5):
        digits = 2
    return '{:0{w}x}'.format(value, w=digits)


# ---------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------
def _cuda_f32(t: Tensor, what: str) -> Tensor:
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise _lib.SesrqError("%s needs a CUDA tensor: the sesrq ops have no CPU fallback" % what)
    return t.contiguous().float()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _tensor_minmax(x: Tensor):
    """(min, max) of a CUDA tensor as Python floats (sesrq_minmax_update; one D2H of two floats, like the
    reference's two .item() calls, quan_func.py:177-178)."""
    mm = torch.tensor([float("inf"), float("-inf")], dtype=torch.float32, device=x.device)
    lib = _lib.load_library()
    with torch.cuda.device(x.device):
        _lib.check(lib.sesrq_minmax_update(_p(x), x.numel(), _p(mm), _stream(x)), "sesrq_minmax_update")
    lo, hi = mm.tolist()
    return lo, hi


def _save(obj, folder, name):
    os.makedirs(folder, exist_ok=True)
    torch.save(obj, os.path.join(folder, name))


# ---------------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------------
def quantize_symmetrical_by_tensor(tensor_input: Tensor, width: int, exe_mode: int, func_id: int = None,
                                   filename: str = None) -> Tensor:
    """Per-tensor symmetric weight quantisation (reference :44-126).  exe_mode 0 returns codes * scale,
    exe_mode 1 the integer codes (as float32, like the reference)."""
    codes, scale = _cp.symmetric_weight_codes(tensor_input, width)
    state.weight_codes[func_id] = codes
    state.weight_scale[func_id] = scale
    codes_f = codes.to(tensor_input.device, torch.float32)
    if WEIGHT_W_FLG:
        _save(scale, "output_pt/weight", "conv.weight.{}.scale.pt".format(func_id))
        _save(codes_f.cpu(), "output_pt/weight", "conv.weight.{}.pt".format(func_id))
        _write_weight_hex(codes, func_id)
    return codes_f * scale if exe_mode == 0 else codes_f


def _write_weight_hex(codes: Tensor, func_id):
    """Hardware weight image: 4 oc x 4 ic blocks, one line per (block, tap, ic) (reference :82-111)."""
    oc_real, ic_real, kh, kw = codes.shape
    oc = math.ceil(oc_real / 4) * 4
    ic = math.ceil(ic_real / 4) * 4
    padded = torch.zeros(oc, ic, kh, kw, dtype=torch.int32)
    padded[:oc_real, :ic_real] = codes.to(torch.int32)
    os.makedirs("output_txt/weight", exist_ok=True)
    with open("output_txt/weight/conv.weight.{}.txt".format(func_id), "w") as f:
        f.write('{}\n'.format("{:02x}".format(int(oc * ic * kh * kw / 16))))
        for bo in range(0, oc, 4):
            for bi in range(0, ic, 4):
                for y in range(kh):
                    for x in range(kw):
                        for i in range(4):
                            f.write("".join(float_to_hex(padded[bo + o, bi + i, y, x].item(), QUAN_BIT) for o in range(4)))
                        f.write('\n')


def _exact_int_conv_forward(conv: nn.Conv2d):
    """Forward of a Conv2d that holds integer codes: exact int32 accumulation in CUDA instead of the
    library float conv (cuDNN may pick non-exact algorithms)."""
    def forward(x):
        x = _cuda_f32(x, "integer Conv2d")
        w = conv.weight.detach().to(x.device, torch.float32).contiguous()
        B, cin, H, W = x.shape
        out = torch.empty((B, w.shape[0], H, W), dtype=torch.float32, device=x.device)
        lib = _lib.load_library()
        with torch.cuda.device(x.device):
            _lib.check(lib.sesrq_op_conv_int(_p(x), B, cin, H, W, _p(w), w.shape[0], w.shape[2], _p(out), _stream(x)),
                       "sesrq_op_conv_int")
        if conv.bias is not None:
            out = out + conv.bias.detach().to(x.device, torch.float32)[None, :, None, None]
        return out
    return forward


def _fq_conv_forward(conv: nn.Conv2d, conv_id: int):
    """Forward of a Conv2d of the exe_mode 0 graph (weights = codes * s_w, input = PE-split (q - z) * s of THIS call's
    quantiser): exact integer accumulation scaled once by f32(s) * f32(s_w) -- the arithmetic of the fused fake-quant
    network, so the per-node graph and the fused call agree bit for bit (the reference's library float conv differs in
    summation order; parity with it is tolerance-based)."""
    def forward(x):
        x = _cuda_f32(x, "fake-quant Conv2d")
        codes = state.weight_codes[conv_id].to(x.device, torch.float32).contiguous()
        B, cin, H, W = x.shape
        out = torch.empty((B, codes.shape[0], H, W), dtype=torch.float32, device=x.device)
        lib = _lib.load_library()
        with torch.cuda.device(x.device):
            _lib.check(lib.sesrq_op_conv_fq(_p(x), B, cin, H, W, _p(codes), codes.shape[0], codes.shape[2],
                                            state.scale(conv_id), state.weight_scale[conv_id], _p(out), _stream(x)),
                       "sesrq_op_conv_fq")
        if conv.bias is not None:
            out = out + conv.bias.detach().to(x.device, torch.float32)[None, :, None, None]
        return out
    return forward


def quantize_model_weight(model_input: nn.Module, weight_width: int, exe_mode: int):
    """Quantise every nn.Conv2d weight of a (collapsed) model; conv index = order in the state dict
    (reference :128-159)."""
    model = copy.deepcopy(model_input)
    params = model.state_dict()
    modules = dict(model.named_modules())
    conv_id = 0
    for name in params:
        owner = modules[remove_suffix(name)]
        if type(owner) is not nn.Conv2d:
            continue
        if 'weight' in name:
            params[name] = quantize_symmetrical_by_tensor(params[name], weight_width, exe_mode, func_id=conv_id)
            conv_id += 1
        elif 'bias' in name:
            pass                                            # biases are quantised later, with the input scale
        else:
            raise KeyError('Unsupported state dict type found. (%s)' % name)
    model.load_state_dict(params)
    conv_id = 0
    for m in model.modules():
        if type(m) is nn.Conv2d:                              # module order == state-dict order of the weights
            m.forward = _exact_int_conv_forward(m) if exe_mode == 1 else _fq_conv_forward(m, conv_id)
            conv_id += 1
    return model


# ---------------------------------------------------------------------------------------------------
# activations
# ---------------------------------------------------------------------------------------------------
def quantize_asymmetrical_by_tensor(tensor_input: Tensor, width: int, exe_mode: int, func_id: int = None) -> Tensor:
    """Input quantiser of conv `func_id` (reference :161-293).  exe_mode 1 returns q - max(z, -128)."""
    x = _cuda_f32(tensor_input, "quantize_asymmetrical_by_tensor")
    lib = _lib.load_library()
    n = x.numel()
    if exe_mode == 0:
        # dynamic per-call range (reference :175-215): observe min/max, running min/max, scale/zero, fake-quant
        lo, hi = _tensor_minmax(x)
        state.observe(func_id, lo, hi)
        assert hi != lo, "Input tensor is all equal,{}".format(func_id)
        quan_max = 2 ** (width - 1) - 1
        quan_min = 0 - 2 ** (width - 1)
        scale = (hi - lo) / (quan_max - quan_min)
        zero = quan_min - round(lo / scale)
        state.set_threshold(func_id, scale, zero)                       # the reference rewrites the files every call (:211-212)
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.sesrq_op_fakequant(_p(x), n, scale, int(zero), _p(y), _stream(x)), "sesrq_op_fakequant")
        if INPUT_W_FLG:
            _save(y, "output_pt/input", "input.{}.pt".format(func_id))
        return y
    q = torch.empty_like(x)
    u = torch.empty_like(x)
    zero = state.zero(func_id)
    with torch.cuda.device(x.device):
        if func_id == 0:
            rc = lib.sesrq_op_quant_first(_p(x), n, state.scale(0), zero, _p(q), _p(u), _stream(x))
        elif func_id == 4:
            if state.shortcut is None:
                raise KeyError("residual shortcut missing: requan_conv2d_output(func_id=0) must run first")
            sc = _cuda_f32(state.shortcut, "shortcut")
            m16, n16 = quan_layer_between_const(state.scale(1) / state.scale(4), REQUAN_BIT, REQUAN_N_MAX)
            state.requan["res"] = (m16, n16)
            ic = torch.empty_like(x) if INPUT_W_FLG else None
            rc = lib.sesrq_op_quant_residual(_p(x), _p(sc), n, m16, n16, zero, _p(ic), _p(q), _p(u), _stream(x))
            if INPUT_W_FLG:
                _save(ic, "output_pt/input", "input.{}.spcial.pt".format(func_id))
            if REQUAN_FACTOR_W_FLG:
                _save(m16, "output_pt/requan_factor", "requan_res.pt")
                _save(n16, "output_pt/requan_factor", "n_res.pt")
        else:
            rc = lib.sesrq_op_quant_next(_p(x), n, zero, _p(q), _p(u), _stream(x))
    _lib.check(rc, "sesrq_op_quant_*")
    if INPUT_W_FLG:
        _save(q, "output_pt/input", "input.{}.pt".format(func_id))
    return u


def reshape_input_for_hardware_pe(input_tensor, pe_num: int = 4):
    """[B,C,H,W] -> [4B,C,H,W]: entry p keeps channels c == p (mod 4) (reference :298-318)."""
    assert len(input_tensor.shape) == 4, 'Expect input tensor dimension: 4, but get %d' % len(input_tensor.shape)
    x = _cuda_f32(input_tensor, "reshape_input_for_hardware_pe")
    B, Cc, H, W = x.shape
    out = torch.empty((B * pe_num, Cc, H, W), dtype=torch.float32, device=x.device)
    lib = _lib.load_library()
    with torch.cuda.device(x.device):
        _lib.check(lib.sesrq_op_pe_split(_p(x), B, Cc, H * W, pe_num, _p(out), _stream(x)), "sesrq_op_pe_split")
    return out


def _pe_merge(input_tensor, conv_weight, acc_bits, add_bits, input_zero, add_const, func_id, pe_num, dumps=True):
    assert len(input_tensor.shape) == 4, 'Expect input tensor dimension: 4, but get %d' % len(input_tensor.shape)
    x = _cuda_f32(input_tensor, "PE merge")
    batch, Cout, H, W = x.shape
    assert batch % pe_num == 0, 'Input batch size %d is not multiples of PE number %d' % (batch, pe_num)
    B = batch // pe_num
    w = conv_weight.to(torch.int64)
    zp = max(int(input_zero), -128)                                      # reference :351-352
    corr = torch.zeros(Cout, 4, dtype=torch.float32)
    for p in range(4):
        corr[:, p] = (w[:, p::4].sum(dim=(1, 2, 3)) * zp).to(torch.float32)
    corr = corr.contiguous().to(x.device)
    addc = add_const.to(x.device, torch.float32).contiguous()
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    pe_out = torch.empty_like(x) if (OUTPUT_PE_W_FLG and dumps) else None
    pe_add = torch.empty_like(out) if (OUTPUT_PE_ADD_W_FLG and dumps) else None
    ovf = torch.zeros(1, dtype=torch.int32, device=x.device)
    lib = _lib.load_library()
    with torch.cuda.device(x.device):
        _lib.check(lib.sesrq_op_pe_merge(_p(x), B, Cout, H * W, _p(corr), _p(addc), acc_bits, add_bits, _p(pe_out),
                                         _p(pe_add), _p(out), _p(ovf), _stream(x)), "sesrq_op_pe_merge")
    flags = int(ovf.item())
    if flags & 1:
        print('max_overflow')                                            # reference :358-361
    if flags & 2:
        print('min_overflow')
    if pe_out is not None:
        assert batch == 4, "only support input batch = 1"
        for p in range(pe_num):
            _save(pe_out[p], "output_pt/pe_out", "pe_output{}_{}.pt".format(func_id, p))
    if pe_add is not None:
        _save(pe_add, "output_pt/pe_add", "pe_add_output{}.pt".format(func_id))
    return out


def _float_clamp_bounds(bits, input_zero, input_scale, weight_scale):
    """(min, max) of the float-domain emulation of a `bits`-wide accumulator (reference :331-332, :432-433)."""
    hi = 2 ** (bits - 1) - 1
    lo = 0 - 2 ** (bits - 1)
    return (lo - input_zero) * input_scale * weight_scale, (hi - input_zero) * input_scale * weight_scale


def _pe_merge_fq(input_tensor, acc_bounds, add_bounds, bias, pe_num):
    assert len(input_tensor.shape) == 4, 'Expect input tensor dimension: 4, but get %d' % len(input_tensor.shape)
    x = _cuda_f32(input_tensor, "PE merge")
    batch, Cout, H, W = x.shape
    assert batch % pe_num == 0, 'Input batch size %d is not multiples of PE number %d' % (batch, pe_num)
    B = batch // pe_num
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    b = None if bias is None else bias.to(x.device, torch.float32).contiguous()
    lib = _lib.load_library()
    inf = float("inf")
    lo_a, hi_a = acc_bounds
    lo_s, hi_s = add_bounds if add_bounds is not None else (-inf, inf)
    with torch.cuda.device(x.device):
        _lib.check(lib.sesrq_op_pe_merge_fq(_p(x), B, Cout, H * W, lo_a, hi_a, lo_s, hi_s, _p(b), _p(out), _stream(x)),
                   "sesrq_op_pe_merge_fq")
    return out


def reshape_ouput_for_hardware_pe(input_tensor: Tensor, conv_weight, width_accum, input_scale, input_zero,
                                  weight_scale, exe_mode, func_id, pe_num: int = 4):
    """Clamp each PE's partial sum to `width_accum` bits and add the four PEs (reference :320-389); exe_mode 0 clamps
    in the float domain (:330-333)."""
    if exe_mode == 0:
        return _pe_merge_fq(input_tensor, _float_clamp_bounds(width_accum, input_zero, input_scale, weight_scale),
                            None, None, pe_num)
    zeros = torch.zeros(input_tensor.shape[1])
    return _pe_merge(input_tensor, conv_weight, width_accum, 31, input_zero, zeros, func_id, pe_num)


def quantize_bias_with_scale(tensor_input: Tensor, bias_scale, bias_width: int, exe_mode: int, func_id) -> Tensor:
    """clamp(round(b / (s_x * s_w))) to bias_width bits (reference :391-416)."""
    assert isinstance(tensor_input, Tensor)
    codes = _cp.bias_codes(tensor_input, bias_scale, bias_width).to(torch.float32)
    if BIAS_W_FLG:
        _save(codes, "output_pt/bias", "conv.bias.{}.pt".format(func_id))
        _save(bias_scale, "output_pt/bias", "conv.bias.{}.scale.pt".format(func_id))
    return codes * bias_scale if exe_mode == 0 else codes


def PEs_and_bias_adder(input_tensor, bias, pe_add_width, pe_acc_width, bias_width, func_id, pe_num, exe_mode):
    """PE adder tree + bias (reference :418-493): clamp18 per PE, sum, clamp20, then exe_mode 1:
    + clamp16(bias_q - sum(W) * z); exe_mode 0: float-domain clamps and + bias_q * s_x * s_w."""
    s_x, z_x = state.scale(func_id), state.zero(func_id)
    if exe_mode == 0:
        s_w = state.weight_scale[func_id]
        bias_t = torch.tensor(bias, dtype=torch.float32)
        bias_fq = quantize_bias_with_scale(bias_t, s_x * s_w, bias_width, exe_mode, func_id)
        state.bias[func_id] = bias_t
        out = _pe_merge_fq(input_tensor, _float_clamp_bounds(pe_acc_width, z_x, s_x, s_w),
                           _float_clamp_bounds(pe_add_width, z_x, s_x, s_w), bias_fq, pe_num)
        import define
        if getattr(define, "MFLAG", MFLAG) == 3 and func_id == 4:          # NRDM: observe the output as domain 5 (:460-479)
            lo, hi = _tensor_minmax(out)
            state.observe(func_id + 1, lo, hi)
        return out
    codes = state.weight_codes[func_id]
    bias_t = torch.tensor(bias, dtype=torch.float32)
    bq = quantize_bias_with_scale(bias_t, s_x * state.weight_scale[func_id], bias_width, exe_mode, func_id)
    lim = 2 ** (bias_width - 1)
    wsum = codes.to(torch.int64).sum(dim=(1, 2, 3))
    add_const = torch.clamp(bq.to(torch.int64) - wsum * z_x, min=-lim, max=lim - 1)
    state.bias[func_id], state.bias_quan[func_id], state.add_const[func_id] = bias_t, bq, add_const
    if BIAS_QUAN_W_FLG:
        _save(add_const.to(torch.float32)[None, :, None, None], "output_pt/bias", "conv.bias.quan{}.pt".format(func_id))
    return _pe_merge(input_tensor, codes, pe_acc_width, pe_add_width, z_x, add_const, func_id, pe_num)


def requan_conv2d_output(input_tensor, func_id, exe_mode):
    """Requantise the conv accumulator into the next layer's domain with a 16-bit multiplier and a shift
    (reference :517-611)."""
    if exe_mode == 0:
        return input_tensor
    x = _cuda_f32(input_tensor, "requan_conv2d_output")
    nxt = _cp.NEXT_DOMAIN[func_id]
    ratio = state.scale(func_id) / state.scale(nxt) * state.weight_scale[func_id]
    m16, n16 = quan_layer_between_const(ratio, REQUAN_BIT, REQUAN_N_MAX)
    tag = "{}_{}".format(func_id, 4 if func_id == 3 else nxt)
    state.requan[tag] = (m16, n16)
    if REQUAN_FACTOR_W_FLG:
        _save(m16, "output_pt/requan_factor", "requan_{}.pt".format(tag))
        _save(n16, "output_pt/requan_factor", "n_{}.pt".format(tag))
    lib = _lib.load_library()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        if func_id == 4:
            q5 = torch.empty_like(x) if INPUT_W_FLG else None
            rc = lib.sesrq_op_requant_final(_p(x), x.numel(), m16, n16, state.zero(5), state.scale(5), _p(q5), _p(y), _stream(x))
            if INPUT_W_FLG:
                _save(q5, "output_pt/input", "input.{}.pt".format(func_id + 1))
        else:
            sc = torch.empty_like(x) if func_id == 0 else None
            rc = lib.sesrq_op_requant(_p(x), x.numel(), m16, n16, _p(y), _p(sc), _stream(x))
            if func_id == 0:
                state.shortcut = sc                                    # reference :530,:549
                import define
                if getattr(define, "COMPAT_FILES", True) or getattr(define, "RESIDUAL_W_FLG", False):   # reference :546-549
                    _save(sc, "output_pt/residual", "shortcut_tensor.pt")
    _lib.check(rc, "sesrq_op_requant")
    return y
