"""Golden vectors for the QAT library: runs the UNMODIFIED reference models/quantize_utils_pt.py (the CPU
spelling of quantize_utils_cuda.py) on seeded inputs in this container and stores inputs, outputs, gradients and the
observer / quantizer buffers.  Test infrastructure: only tests/ read the fixture; nothing on the GPU box reads
/root/reference.

    python oracle/gen_qat_golden.py            -> tests/golden/qat_golden.npz
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
from models import quantize_utils_pt as Q   # noqa: E402  (the reference's module)
from models import nrdm_3                   # noqa: E402

out = {}


def put(name, t):
    out[name] = t.detach().cpu().numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t)


def buffers(prefix, mod):
    for k, v in mod.state_dict().items():
        if any(s in k for s in ("scale", "zero_point", "min_val", "max_val")):
            put(prefix + ".sd." + k, v)


def conv_case(tag, q_type, q_level, weight_observer, steps=2):
    g = torch.Generator().manual_seed(100 + len(out))
    m = Q.QuantConv2d(3, 8, 3, padding=1, a_bits=8, w_bits=8, q_type=q_type, q_level=q_level, weight_observer=weight_observer)
    with torch.no_grad():
        m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.2)
        m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    put(tag + ".w", m.weight)
    put(tag + ".b", m.bias)
    m.train()
    for s in range(steps):
        x = (torch.randn(2, 3, 12, 20, generator=g) * (1.0 + s)).requires_grad_(True)
        gy = torch.randn(2, 8, 12, 20, generator=g)
        m.zero_grad()
        # the two fake-quantised operands (what the kernels must reproduce exactly) and the conv result
        qx = m.activation_quantizer(x)
        qw = m.weight_quantizer(m.weight)
        y = torch.nn.functional.conv2d(qx, qw, m.bias, m.stride, m.padding, m.dilation, m.groups)
        y.backward(gy)
        for n, t in (("x", x), ("gy", gy), ("qx", qx), ("qw", qw), ("y", y), ("gx", x.grad), ("gw", m.weight.grad), ("gb", m.bias.grad)):
            put("%s.s%d.%s" % (tag, s, n), t)
        buffers("%s.s%d" % (tag, s), m)
    m.eval()
    x = torch.randn(1, 3, 12, 20, generator=g)
    put(tag + ".eval.x", x)
    put(tag + ".eval.y", m(x))


def act_case(tag, ctor, q_type, ptq=False):
    g = torch.Generator().manual_seed(200 + len(out))
    m = ctor(a_bits=8, q_type=q_type, ptq=ptq)
    m.train()
    for s in range(2):
        x = (torch.randn(2, 4, 16, 16, generator=g) * (0.5 + s) + 0.3).requires_grad_(True)
        gy = torch.randn(2, 4, 16, 16, generator=g)
        y = m(x)
        y.backward(gy)
        for n, t in (("x", x), ("gy", gy), ("y", y), ("gx", x.grad)):
            put("%s.s%d.%s" % (tag, s, n), t)
        buffers("%s.s%d" % (tag, s), m)


def add_case(tag, q_type):
    g = torch.Generator().manual_seed(300 + len(out))
    m = Q.QuantAdd(a_bits=8, q_type=q_type)
    m.train()
    for s in range(2):
        a = (torch.randn(2, 4, 16, 16, generator=g) * (0.5 + s)).requires_grad_(True)
        b = (torch.randn(2, 4, 16, 16, generator=g) * 2.0 + 1.0).requires_grad_(True)
        gy = torch.randn(2, 4, 16, 16, generator=g)
        y = m(a, b)
        y.backward(gy)
        for n, t in (("a", a), ("b", b), ("gy", gy), ("y", y), ("ga", a.grad), ("gb", b.grad)):
            put("%s.s%d.%s" % (tag, s, n), t)
        buffers("%s.s%d" % (tag, s), m)


def model_case(tag):
    """prepare() on the un-collapsed NRDM_3 training model exactly as train.py / test.py:63 call it."""
    torch.manual_seed(7)
    model = nrdm_3.nr()
    put(tag + ".init_sd_keys", np.array(sorted(model.state_dict().keys())))
    for k, v in model.state_dict().items():
        put(tag + ".init." + k, v)
    Q.prepare(model, inplace=True, a_bits=8, w_bits=8, q_type=0, q_level="C")
    model.train()
    g = torch.Generator().manual_seed(11)
    x = torch.rand(1, 3, 24, 32, generator=g).requires_grad_(True)
    gy = torch.randn(1, 3, 24, 32, generator=g)
    y = model(x)
    y.backward(gy)
    put(tag + ".x", x)
    put(tag + ".gy", gy)
    put(tag + ".y", y)
    put(tag + ".gx", x.grad)
    put(tag + ".g_first_expand", model.conv_first.conv_expand.weight.grad)
    put(tag + ".g_last_squeeze", model.conv_last.conv_squeeze.weight.grad)
    put(tag + ".sd_keys", np.array(sorted(model.state_dict().keys())))
    buffers(tag, model)


conv_case("conv_sym_pc", q_type=0, q_level=0, weight_observer=0)          # per-channel weights (q_level == 0)
conv_case("conv_sym_pt", q_type=0, q_level="C", weight_observer=0)        # what the scripts pass: per-tensor
conv_case("conv_asym_pc_ma", q_type=1, q_level=0, weight_observer=1)      # asymmetric, moving-average weight observer
act_case("relu_sym", Q.QuantReLU, 0)
act_case("relu_asym", Q.QuantReLU, 1)
act_case("lrelu_sym", Q.QuantLeakyReLU, 0)
act_case("relu_ptq", Q.QuantReLU, 0, ptq=True)
add_case("add_sym", 0)
add_case("add_asym", 1)
model_case("nrdm3_prepared")

path = os.path.join(ROOT, "tests", "golden", "qat_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, "%d arrays, %.1f KB" % (len(out), os.path.getsize(path) / 1024))
