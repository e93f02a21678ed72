"""QAT library mirror (models/quantize_utils_cuda.py) against vectors produced by the UNMODIFIED reference module
(oracle/gen_qat_golden.py -> tests/golden/qat_golden.npz).  The fake-quant forward / straight-through backward kernels
restate fp32 IEEE operations step by step, so the quantised operands and their gradients must match the reference
(CPU torch) exactly; results that pass through a library convolution (cuDNN vs oneDNN) are compared to tolerance."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD

Z = np.load(os.path.join(GOLD, "qat_golden.npz"), allow_pickle=False)


def T(name, dev="cuda"):
    return torch.from_numpy(Z[name]).to(dev)


def buffers_match(prefix, mod, atol=0.0):
    sd = mod.state_dict()
    n = 0
    for k in Z.files:
        if k.startswith(prefix + ".sd."):
            key = k[len(prefix) + 4:]
            got = sd[key].detach().cpu().numpy().reshape(Z[k].shape)
            np.testing.assert_allclose(got, Z[k], rtol=0, atol=atol, err_msg=k)
            n += 1
    assert n > 0


# ---- structure (CPU): the module swap produces the reference's module tree / state-dict keys ------------------------
def test_prepare_builds_reference_state_dict_keys():
    from models import nrdm_3, quantize_utils_cuda as Q
    model = nrdm_3.nr()
    assert sorted(model.state_dict().keys()) == list(Z["nrdm3_prepared.init_sd_keys"])
    Q.prepare(model, inplace=True, a_bits=8, w_bits=8, q_type=0, q_level="C")
    assert sorted(model.state_dict().keys()) == list(Z["nrdm3_prepared.sd_keys"])
    assert type(model.add_residual).__name__ == "QuantAdd"
    assert type(model.conv_first.conv_expand).__name__ == "QuantConv2d"
    # q_level "C" is what the scripts pass; the reference tests `q_level == 0`, so weights end up per tensor
    assert model.conv_first.conv_expand.weight_quantizer.scale.shape == (1,)


def test_quantizer_ranges_and_per_channel_shapes():
    from models import quantize_utils_pt as Q
    c = Q.QuantConv2d(3, 8, 3, q_type=0, q_level=0)
    assert c.weight_quantizer.scale.shape == (8, 1, 1, 1)
    assert (float(c.weight_quantizer.quant_min_val), float(c.weight_quantizer.quant_max_val)) == (-127.0, 127.0)
    assert (float(c.activation_quantizer.quant_min_val), float(c.activation_quantizer.quant_max_val)) == (-128.0, 127.0)
    a = Q.QuantConv2d(3, 8, 3, q_type=1, q_level=0)
    assert (float(a.weight_quantizer.quant_min_val), float(a.weight_quantizer.quant_max_val)) == (0.0, 254.0)
    assert (float(a.activation_quantizer.quant_min_val), float(a.activation_quantizer.quant_max_val)) == (0.0, 255.0)
    assert type(Q.QuantReLU(ptq=True).activation_quantizer.observer).__name__ == "HistogramObserver"


def test_cpu_tensor_raises():
    from models import quantize_utils_cuda as Q
    from sesr_pytorch_quantize_b200 import SesrqError
    with pytest.raises(SesrqError):
        Q.QuantReLU().train()(torch.randn(1, 2, 4, 4))


# ---- numerics (GPU) -------------------------------------------------------------------------------------------------
CONV_CASES = [("conv_sym_pc", 0, 0, 0), ("conv_sym_pt", 0, "C", 0), ("conv_asym_pc_ma", 1, 0, 1)]


@pytest.mark.gpu
@pytest.mark.parametrize("tag,q_type,q_level,wobs", CONV_CASES)
def test_quant_conv2d_matches_reference(tag, q_type, q_level, wobs):
    from models import quantize_utils_cuda as Q
    m = Q.QuantConv2d(3, 8, 3, padding=1, a_bits=8, w_bits=8, q_type=q_type, q_level=q_level, weight_observer=wobs).cuda()
    with torch.no_grad():
        m.weight.copy_(T(tag + ".w"))
        m.bias.copy_(T(tag + ".b"))
    m.train()
    for s in range(2):
        p = "%s.s%d." % (tag, s)
        x = T(p + "x").requires_grad_(True)
        m.zero_grad()
        qx = m.activation_quantizer(x)
        qw = m.weight_quantizer(m.weight)
        # exact: the fake-quantised operands
        assert torch.equal(qx, T(p + "qx")), "quantised input differs"
        assert torch.equal(qw, T(p + "qw")), "quantised weight differs"
        y = torch.nn.functional.conv2d(qx, qw, m.bias, m.stride, m.padding, m.dilation, m.groups)
        y.backward(T(p + "gy"))
        torch.testing.assert_close(y, T(p + "y"), rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(x.grad, T(p + "gx"), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(m.weight.grad, T(p + "gw"), rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(m.bias.grad, T(p + "gb"), rtol=1e-4, atol=1e-3)
        buffers_match("%s.s%d" % (tag, s), m)
    m.eval()
    torch.testing.assert_close(m(T(tag + ".eval.x")), T(tag + ".eval.y"), rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,cls,q_type,ptq", [("relu_sym", "QuantReLU", 0, False), ("relu_asym", "QuantReLU", 1, False),
                                                ("lrelu_sym", "QuantLeakyReLU", 0, False), ("relu_ptq", "QuantReLU", 0, True)])
def test_quant_activations_bit_exact(tag, cls, q_type, ptq):
    from models import quantize_utils_cuda as Q
    m = getattr(Q, cls)(a_bits=8, q_type=q_type, ptq=ptq).cuda().train()
    for s in range(2):
        p = "%s.s%d." % (tag, s)
        x = T(p + "x").requires_grad_(True)
        y = m(x)
        y.backward(T(p + "gy"))
        assert torch.equal(y, T(p + "y")), "forward differs (step %d)" % s
        assert torch.equal(x.grad, T(p + "gx")), "straight-through gradient differs (step %d)" % s
        buffers_match("%s.s%d" % (tag, s), m)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,q_type", [("add_sym", 0), ("add_asym", 1)])
def test_quant_add_bit_exact(tag, q_type):
    from models import quantize_utils_cuda as Q
    m = Q.QuantAdd(a_bits=8, q_type=q_type).cuda().train()
    for s in range(2):
        p = "%s.s%d." % (tag, s)
        a, b = T(p + "a").requires_grad_(True), T(p + "b").requires_grad_(True)
        y = m(a, b)
        y.backward(T(p + "gy"))
        assert torch.equal(y, T(p + "y"))
        assert torch.equal(a.grad, T(p + "ga"))
        assert torch.equal(b.grad, T(p + "gb"))
        buffers_match("%s.s%d" % (tag, s), m)


@pytest.mark.gpu
def test_prepared_model_forward_backward():
    from models import nrdm_3, quantize_utils_cuda as Q
    tag = "nrdm3_prepared"
    model = nrdm_3.nr()
    model.load_state_dict({k: torch.from_numpy(Z[tag + ".init." + k]) for k in Z[tag + ".init_sd_keys"]})
    model = model.cuda()
    Q.prepare(model, inplace=True, a_bits=8, w_bits=8, q_type=0, q_level="C")
    model.train()
    x = T(tag + ".x").requires_grad_(True)
    y = model(x)
    y.backward(T(tag + ".gy"))
    # 14 quantised convs in sequence: a different conv summation order can flip single codes -> tolerance in quantisation steps
    ref = T(tag + ".y")
    err = (y - ref).abs().max().item()
    assert err <= 0.05 * ref.abs().max().item(), err      # a handful of 8-bit steps after 14 quantisers
    g = T(tag + ".gx")
    assert (x.grad - g).abs().mean().item() <= 0.05 * g.abs().mean().item()
    buffers_match(tag, model, atol=1e-2)


@pytest.mark.gpu
def test_kth_abs_equals_torch_kthvalue():
    import ctypes as C
    from sesr_pytorch_quantize_b200 import _lib
    lib = _lib.load_library()
    g = torch.Generator().manual_seed(5)
    for n in (1, 7, 1000, 123457):
        x = (torch.randn(n, generator=g) * 3).cuda()
        x[::5] = 0.0
        for k in sorted({1, max(1, n // 2), max(1, int(0.9999 * n)), n}):
            out = torch.empty(1, device="cuda")
            ws = torch.empty((65536 + 32768 + 2) * 4, dtype=torch.uint8, device="cuda")
            _lib.check(lib.sesrq_kth_abs(C.c_void_p(x.data_ptr()), n, k, C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()),
                                         ws.numel(), None))
            assert out.item() == torch.kthvalue(x.abs().cpu(), k)[0].item(), (n, k)
