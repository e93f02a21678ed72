"""Quantisation-aware-training library (mirror of the reference's models/quantize_utils_cuda.py:1-835: same class
names, constructor arguments, buffer names -- so `*_qat_G.pth` checkpoints load -- and the same `prepare()` /
`add_quant_op()` module swap that sim.py:66 / test.py:63 / train.py call as `quantize.prepare(...)`).

What differs from the reference:
  * the arithmetic runs in libsesrq's CUDA kernels (include/sesrq.h "QAT library"): one fused fake-quant forward
    (divide, offset, round-half-away, clamp, de-quantise) and one straight-through backward kernel replace the
    ~10 elementwise torch ops + 3 boolean-mask passes of Round / Quantizer.forward (:150-246); observers reduce with
    sesrq_qat_observe (per tensor or per output channel) and (scale, zero_point) come from sesrq_qat_qparams;
    HistogramObserver's k-th value is an exact radix select (sesrq_kth_abs) instead of torch.kthvalue's sort;
  * buffers are created on the CPU and follow the module (`.cuda()`, `.to()`); `add_quant_op` places every new
    module on the device of the module it replaces (the reference hard-codes `.cuda()` in the constructors);
  * the observer / quantizer choice is table driven (`_make_quantizer`) instead of the reference's nested if-trees.
There is no CPU fallback: a quantizer called with a CPU tensor raises SesrqError.
"""
import copy
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from models.model_utils_pt import AddOp
from sesr_pytorch_quantize_b200 import _lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.SesrqError("%s needs a CUDA tensor: the sesrq QAT kernels have no CPU fallback" % what)


def _rows_of(level, out_channels):
    return 1 if level == "L" else int(out_channels)


def _range_shape(level, out_channels):
    return {"L": (1,), "C": (out_channels, 1, 1, 1), "FC": (out_channels, 1)}[level]


# ********************* observers *********************
class ObserverBase(nn.Module):
    """Tracks the (min, max) range of what passes through: per tensor ("L") or per output channel ("C" conv
    weights, "FC" linear weights).  Reference :14-35."""
    _MODE = 1                                   # sesrq_qat_observe mode after the first call

    def __init__(self, q_level):
        super().__init__()
        self.q_level = q_level
        self.num_flag = 0
        self.momentum = 0.0

    def _register_range(self, out_channels):
        shape = _range_shape(self.q_level, out_channels)
        self.register_buffer("min_val", torch.zeros(shape, dtype=torch.float32))
        self.register_buffer("max_val", torch.zeros(shape, dtype=torch.float32))

    def update_range(self, min_val, max_val):
        raise NotImplementedError

    @torch.no_grad()
    def forward(self, input):
        _need_cuda(input, type(self).__name__)
        x = input.detach().contiguous().float()
        rows = self.min_val.numel()
        cols = x.numel() // rows
        if self.min_val.device != x.device:
            self.to(x.device)
        tmp = torch.empty(2 * rows, dtype=torch.float32, device=x.device)
        mode = 0 if self.num_flag == 0 else self._MODE
        with torch.cuda.device(x.device):
            _lib.check(_lib.load_library().sesrq_qat_observe(_p(x), rows, cols, mode, float(self.momentum), _p(self.min_val),
                                                             _p(self.max_val), _p(tmp), _stream(x)), "sesrq_qat_observe")
        self.num_flag += 1 if self.num_flag == 0 else 0


class MinMaxObserver(ObserverBase):
    """Running min / max (reference :45-80)."""
    _MODE = 1

    def __init__(self, q_level, out_channels):
        super().__init__(q_level)
        self.out_channels = out_channels
        self._register_range(out_channels)


class MovingAverageMinMaxObserver(ObserverBase):
    """Exponential moving average of the per-call min / max (reference :83-119)."""
    _MODE = 2

    def __init__(self, q_level, out_channels, momentum=0.1):
        super().__init__(q_level)
        self.momentum = momentum
        self.out_channels = out_channels
        self._register_range(out_channels)


class HistogramObserver(nn.Module):
    """Percentile calibrator: max_val = moving average of the `percentile` quantile of |x| (reference :122-145)."""

    def __init__(self, q_level, momentum=0.1, percentile=0.9999):
        super().__init__()
        self.q_level = q_level
        self.momentum = momentum
        self.percentile = percentile
        self.num_flag = 0
        self.register_buffer("min_val", torch.zeros((1), dtype=torch.float32))
        self.register_buffer("max_val", torch.zeros((1), dtype=torch.float32))

    @torch.no_grad()
    def forward(self, input):
        _need_cuda(input, "HistogramObserver")
        x = input.detach().contiguous().float().view(-1)
        if self.max_val.device != x.device:
            self.to(x.device)
        k = int(self.percentile * x.size(0))
        cur = torch.empty(1, dtype=torch.float32, device=x.device)
        ws = torch.empty((65536 + 32768 + 2) * 4, dtype=torch.uint8, device=x.device)
        lib = _lib.load_library()
        with torch.cuda.device(x.device):
            _lib.check(lib.sesrq_kth_abs(_p(x), x.numel(), k, _p(cur), _p(ws), ws.numel(), _stream(x)), "sesrq_kth_abs")
            tmp = torch.empty(2, dtype=torch.float32, device=x.device)
            mode = 0 if self.num_flag == 0 else 2
            # moving average of the quantile: reuse the observer combine on a 1-element "tensor" (min_val stays 0)
            keep_min = self.min_val.clone()
            _lib.check(lib.sesrq_qat_observe(_p(cur), 1, 1, mode, float(self.momentum), _p(self.min_val), _p(self.max_val),
                                             _p(tmp), _stream(x)), "sesrq_qat_observe")
            self.min_val.copy_(keep_min)
        self.num_flag += 1 if self.num_flag == 0 else 0


# ********************* quantizers *********************
class _FakeQuantSTE(Function):
    """(clamp(round_half_away(x / s - zp), qmin, qmax) + zp) * s with the reference's straight-through gradient:
    Round.backward's range mask (:169-174) times torch.clamp's mask, scaled by s and divided by s again as autograd
    does for `(...) * scale` and `input / scale`."""

    @staticmethod
    def forward(ctx, x, scale, zero_point, obs_min, obs_max, qmin, qmax, q_type):
        _need_cuda(x, "Quantizer")
        xc = x.contiguous()
        rows = scale.numel()
        inner = max(xc.numel() // max(xc.shape[0], 1), 1) if rows > 1 else 1
        if rows > 1 and xc.shape[0] != rows:
            raise _lib.SesrqError("per-channel quantizer: tensor has %d rows, quantizer %d" % (xc.shape[0], rows))
        y = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(_lib.load_library().sesrq_qat_fakequant_fwd(_p(xc), xc.numel(), inner, rows, _p(scale), _p(zero_point),
                                                                   qmin, qmax, _p(y), _stream(xc)), "sesrq_qat_fakequant_fwd")
        ctx.save_for_backward(xc, scale.clone(), zero_point.clone(), obs_min.clone(), obs_max.clone())
        ctx.cfg = (inner, rows, qmin, qmax, q_type)
        return y

    @staticmethod
    def backward(ctx, grad_out):
        xc, scale, zero_point, obs_min, obs_max = ctx.saved_tensors
        inner, rows, qmin, qmax, q_type = ctx.cfg
        g = grad_out.contiguous().float()
        gx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(_lib.load_library().sesrq_qat_fakequant_bwd(_p(xc), _p(g), xc.numel(), inner, rows, _p(scale),
                                                                   _p(zero_point), _p(obs_min), _p(obs_max), q_type, qmin, qmax,
                                                                   _p(gx), _stream(xc)), "sesrq_qat_fakequant_bwd")
        return gx, None, None, None, None, None, None, None


class Round(Function):
    """Round half away from zero with the saturating straight-through gradient (reference :150-174).  Kept for API
    parity; Quantizer.forward uses the fused kernel instead of composing this with clamp / scale ops."""

    @staticmethod
    def forward(self, input, observer_min_val, observer_max_val, q_type):
        if q_type == 0:
            max_val = torch.max(torch.abs(observer_min_val), torch.abs(observer_max_val))
            min_val = -max_val
        else:
            max_val, min_val = observer_max_val, observer_min_val
        self.save_for_backward(input, min_val, max_val)
        return torch.sign(input) * torch.floor(torch.abs(input) + 0.5)

    @staticmethod
    def backward(self, grad_output):
        input, min_val, max_val = self.saved_tensors
        keep = ~(input.gt(max_val) | input.lt(min_val))
        return grad_output * keep, None, None, None


class Quantizer(nn.Module):
    """Fake-quantises its input with (scale, zero_point) derived from an observer (reference :177-246)."""
    _Q_TYPE = 0

    def __init__(self, bits, observer, activation_weight_flag, qaft=False, union=False):
        super().__init__()
        self.bits = bits
        self.observer = observer
        self.activation_weight_flag = activation_weight_flag
        self.qaft = qaft
        self.union = union
        self.q_type = 0
        shape = _range_shape(observer.q_level, getattr(observer, "out_channels", None))
        self.register_buffer("scale", torch.ones(shape, dtype=torch.float32))
        self.register_buffer("zero_point", torch.zeros(shape, dtype=torch.float32))
        self.register_buffer("eps", torch.tensor(torch.finfo(torch.float32).eps, dtype=torch.float32))
        lo, hi = self._quant_range()
        self.register_buffer("quant_min_val", torch.tensor(lo, dtype=torch.float32))
        self.register_buffer("quant_max_val", torch.tensor(hi, dtype=torch.float32))
        self._qlim = (float(lo), float(hi))

    def _quant_range(self):
        raise NotImplementedError

    def update_qparams(self):
        """(scale, zero_point) from the observer range (reference :299-327), on device."""
        obs = self.observer
        _need_cuda(obs.min_val, "Quantizer.update_qparams")
        self.q_type = self._Q_TYPE
        with torch.cuda.device(obs.min_val.device):
            _lib.check(_lib.load_library().sesrq_qat_qparams(_p(obs.min_val), _p(obs.max_val), self.scale.numel(), self._Q_TYPE,
                                                             self._qlim[0], self._qlim[1], float(torch.finfo(torch.float32).eps),
                                                             _p(self.scale), _p(self.zero_point), _stream(obs.min_val)),
                       "sesrq_qat_qparams")

    def round(self, input, observer_min_val, observer_max_val, q_type):
        return Round.apply(input, observer_min_val, observer_max_val, q_type)

    def forward(self, input):
        if self.bits == 32:
            return input
        assert self.bits != 1, "Binary quantization is not supported"
        if self.scale.device != input.device:
            self.to(input.device)
        if not self.qaft and self.training:
            if not self.union:
                self.observer(input)
            self.update_qparams()
        return _FakeQuantSTE.apply(input, self.scale, self.zero_point, self.observer.min_val, self.observer.max_val,
                                   self._qlim[0], self._qlim[1], self.q_type)


class SignedQuantizer(Quantizer):
    def _quant_range(self):
        hi = (1 << (self.bits - 1)) - 1
        if self.activation_weight_flag == 0:      # weight: symmetric range without -2^(b-1)
            return -hi, hi
        if self.activation_weight_flag == 1:      # activation
            return -(1 << (self.bits - 1)), hi
        raise ValueError("activation_weight_flag error")


class UnsignedQuantizer(Quantizer):
    def _quant_range(self):
        if self.activation_weight_flag == 0:
            return 0, (1 << self.bits) - 2
        if self.activation_weight_flag == 1:
            return 0, (1 << self.bits) - 1
        raise ValueError("activation_weight_flag error")


class SymmetricQuantizer(SignedQuantizer):
    _Q_TYPE = 0


class AsymmetricQuantizer(UnsignedQuantizer):
    _Q_TYPE = 1


def _make_quantizer(bits, q_type, level, out_channels, flag, qaft, ptq, percentile, moving, union=False):
    """One place for the reference's observer / quantizer selection (:365-496, :538-563, :657-695):
    ptq -> symmetric + percentile histogram for activations (weights keep min/max observers);
    q_type 0 / 1 -> symmetric / asymmetric; `moving` -> moving-average instead of running min/max."""
    if ptq and flag == 1:
        return SymmetricQuantizer(bits=bits, observer=HistogramObserver(q_level="L", percentile=percentile),
                                  activation_weight_flag=flag, qaft=qaft, union=union)
    obs_cls = MovingAverageMinMaxObserver if moving else MinMaxObserver
    q_cls = SymmetricQuantizer if (q_type == 0 or ptq) else AsymmetricQuantizer
    return q_cls(bits=bits, observer=obs_cls(q_level=level, out_channels=out_channels), activation_weight_flag=flag,
                 qaft=qaft, union=union)


# ********************* quantised modules *********************
class QuantConv2d(nn.Conv2d):
    """Conv2d whose input and weight are fake-quantised (reference :331-513)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode="zeros", a_bits=8, w_bits=8, q_type=0, q_level=0, weight_observer=0, quant_inference=False,
                 qaft=False, ptq=False, percentile=0.9999):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        self.quant_inference = quant_inference
        self.activation_quantizer = _make_quantizer(a_bits, q_type, "L", None, 1, qaft, ptq, percentile, moving=True)
        per_channel = (q_level == 0)              # the reference tests `q_level == 0` (:376); prepare() passes "C"
        self.weight_quantizer = _make_quantizer(w_bits, q_type, "C" if per_channel else "L",
                                                out_channels if per_channel else None, 0, qaft, ptq, percentile,
                                                moving=(weight_observer != 0))

    def forward(self, input):
        quant_input = self.activation_quantizer(input)
        quant_weight = self.weight if self.quant_inference else self.weight_quantizer(self.weight)
        return F.conv2d(quant_input, quant_weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


def reshape_to_activation(input):
    return input.reshape(1, -1, 1, 1)


def reshape_to_weight(input):
    return input.reshape(-1, 1, 1, 1)


def reshape_to_bias(input):
    return input.reshape(-1)


class QuantReLU(nn.ReLU):
    def __init__(self, inplace=False, a_bits=8, q_type=0, qaft=False, ptq=False, percentile=0.9999):
        super().__init__(inplace)
        self.activation_quantizer = _make_quantizer(a_bits, q_type, "L", None, 1, qaft, ptq, percentile, moving=True)

    def forward(self, input):
        return F.relu(self.activation_quantizer(input), self.inplace)


class QuantLeakyReLU(nn.LeakyReLU):
    def __init__(self, negative_slope=0.01, inplace=False, a_bits=8, q_type=0, qaft=False, ptq=False, percentile=0.9999):
        super().__init__(negative_slope, inplace)
        self.activation_quantizer = _make_quantizer(a_bits, q_type, "L", None, 1, qaft, ptq, percentile, moving=True)

    def forward(self, input):
        return F.leaky_relu(self.activation_quantizer(input), self.negative_slope, self.inplace)


class QuantAdaptiveAvgPool2d(nn.AdaptiveAvgPool2d):
    def __init__(self, output_size, a_bits=8, q_type=0, qaft=False, ptq=False, percentile=0.9999):
        super().__init__(output_size)
        self.activation_quantizer = _make_quantizer(a_bits, q_type, "L", None, 1, qaft, ptq, percentile, moving=True)

    def forward(self, input):
        return F.adaptive_avg_pool2d(self.activation_quantizer(input), self.output_size)


class QuantAdd(nn.Module):
    """res + shortcut with both operands quantised in ONE shared range: the union of their observed ranges
    (reference :654-711)."""

    def __init__(self, a_bits=8, q_type=0, qaft=False, ptq=False, percentile=0.9999):
        super().__init__()
        if ptq:
            self.observer_res = HistogramObserver(q_level="L", percentile=percentile)
            self.observer_shortcut = HistogramObserver(q_level="L", percentile=percentile)
        else:
            self.observer_res = MovingAverageMinMaxObserver(q_level="L", out_channels=None)
            self.observer_shortcut = MovingAverageMinMaxObserver(q_level="L", out_channels=None)
        self.activation_quantizer = _make_quantizer(a_bits, q_type, "L", None, 1, qaft, ptq, percentile, moving=True, union=True)

    def forward(self, res, shortcut):
        self.observer_res(res)
        self.observer_shortcut(shortcut)
        obs = self.activation_quantizer.observer
        if obs.min_val.device != res.device:
            self.activation_quantizer.to(res.device)
        # union of the two ranges (two 1-element tensors: host-level plumbing, not a kernel-worthy reduction)
        obs.min_val = torch.min(self.observer_res.min_val, self.observer_shortcut.min_val)
        obs.max_val = torch.max(self.observer_res.max_val, self.observer_shortcut.max_val)
        return self.activation_quantizer(res) + self.activation_quantizer(shortcut)


# ********************* module swap *********************
def _device_of(module, default):
    for t in list(module.parameters(recurse=False)) + list(module.buffers(recurse=False)):
        return t.device
    return default


def add_quant_op(module, a_bits=8, w_bits=8, q_type=0, q_level=0, weight_observer=0, bn_fuse=False, bn_fuse_calib=False,
                 quant_inference=False, pretrained_model=False, qaft=False, ptq=False, percentile=0.9999, _device=None):
    """Replace every nn.Conv2d by QuantConv2d (sharing weight / bias) and every AddOp by QuantAdd, recursively
    (reference :714-798)."""
    if _device is None:
        _device = next((p.device for p in module.parameters()), torch.device("cpu"))
    for name, child in module.named_children():
        if isinstance(child, nn.Conv2d):
            qc = QuantConv2d(child.in_channels, child.out_channels, child.kernel_size, stride=child.stride,
                             padding=child.padding, dilation=child.dilation, groups=child.groups,
                             bias=child.bias is not None, padding_mode=child.padding_mode, a_bits=a_bits, w_bits=w_bits,
                             q_type=q_type, q_level=q_level, weight_observer=weight_observer,
                             quant_inference=quant_inference, qaft=qaft, ptq=ptq, percentile=percentile)
            qc.to(_device_of(child, _device))
            if child.bias is not None:
                qc.bias.data = child.bias
            qc.weight.data = child.weight
            module._modules[name] = qc
        elif isinstance(child, AddOp):
            module._modules[name] = QuantAdd(a_bits=a_bits, q_type=q_type, qaft=qaft, ptq=ptq,
                                             percentile=percentile).to(_device)
        else:
            add_quant_op(child, a_bits=a_bits, w_bits=w_bits, q_type=q_type, q_level=q_level,
                         weight_observer=weight_observer, bn_fuse=bn_fuse, bn_fuse_calib=bn_fuse_calib,
                         quant_inference=quant_inference, pretrained_model=pretrained_model, qaft=qaft, ptq=ptq,
                         percentile=percentile, _device=_device)


def prepare(model, inplace=False, a_bits=8, w_bits=8, q_type=0, q_level=0, weight_observer=0, bn_fuse=False,
            bn_fuse_calib=False, quant_inference=False, pretrained_model=False, qaft=False, ptq=False, percentile=0.9999):
    """Entry point used by the scripts: `quantize.prepare(model, inplace=True, a_bits=8, w_bits=8, q_type=0,
    q_level="C")` (sim.py:66, test.py:63; reference :801-834)."""
    if not inplace:
        model = copy.deepcopy(model)
    add_quant_op(model, a_bits=a_bits, w_bits=w_bits, q_type=q_type, q_level=q_level, weight_observer=weight_observer,
                 bn_fuse=bn_fuse, bn_fuse_calib=bn_fuse_calib, quant_inference=quant_inference,
                 pretrained_model=pretrained_model, qaft=qaft, ptq=ptq, percentile=percentile)
    return model
