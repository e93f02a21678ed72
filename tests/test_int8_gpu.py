"""Parity of the CUDA integer-sim path (through the C ABI) against the committed golden vectors of
the unmodified reference and against the oracle on seeded inputs.  Bit-exact (integer work)."""
import hashlib

import numpy as np
import pytest
import torch

import sesr_oracle as orc
from conftest import load_golden, ALL_CASES
from sesr_pytorch_quantize_b200 import Int8Net, compile_network, IMPL_GENERIC, IMPL_AUTO, IMPL_UMMA, SesrqError

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build(g, impl):
    """impl: a kernel family, or (family, fuse) -- fuse=False keeps layers 1 and 2 as separate launches (sesrq_net_set_fusion)."""
    impl, fuse = impl if isinstance(impl, tuple) else (impl, None)
    cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf],
                           g.scales, g.zeros, g.r, pe_add_bits=g.pe_add_bit)
    return Int8Net(cnet, impl=impl, fuse=fuse)


def impl_ids():
    # "auto" is the shipped default: tcgen05 kernels, layers 1 + 2 fused into one launch where the layouts allow it
    return [pytest.param(IMPL_GENERIC, id="generic"), pytest.param(IMPL_AUTO, id="auto"),
            pytest.param((IMPL_AUTO, False), id="auto_unfused")]


@pytest.mark.parametrize("impl", impl_ids())
@pytest.mark.parametrize("case", ALL_CASES)
def test_golden_bit_exact(case, impl):
    g = load_golden(case)
    net = build(g, impl)
    x = torch.from_numpy(g.x).cuda()
    q5, y = net(x)
    torch.cuda.synchronize()
    assert np.array_equal(q5.cpu().numpy(), g["q5"]), "q5 differs from the reference's input.5.pt"
    assert sha(y.cpu().numpy()) == g.hashes["yout"], "float output differs from the reference"
    for dom in range(0, 5):
        try:
            q = net.read_activation(dom).cpu().numpy()
        except SesrqError:
            # never written to HBM: the tcgen05 layer-0 kernel quantises the frame on the fly (domain 0), and layers
            # 1 + 2 run as one launch when fusion is on (domain 2)
            assert dom == 0 or (dom == 2 and net.fused), "domain %d should be readable" % dom
            continue
        assert sha(q) == g.hashes["q%d" % dom], "q%d differs" % dom


@pytest.mark.parametrize("impl", impl_ids())
@pytest.mark.parametrize("case,shape", [
    ("x4_fixture", (3, 67, 301)), ("x2_fixture", (2, 130, 257)), ("nrdm_fixture", (2, 33, 129)),
    ("x4_fixture", (1, 1, 1)), ("x2_fixture", (1, 2, 3)), ("x4_fixture", (1, 5, 1)), ("nrdm_fixture", (1, 1, 700)),
    ("synth_clamp", (2, 40, 150)), ("synth_zero", (2, 37, 131)), ("synth_neg", (1, 50, 260)),
])
def test_oracle_bit_exact_seeded(case, shape, impl):
    g = load_golden(case)
    B, H, W = shape
    cin = g.wf[0].shape[1]
    gen = torch.Generator().manual_seed(1000 + H * W)
    x = torch.rand(B, cin, H, W, generator=gen)
    if int(g["signed_input"]):
        x = x * 2 - 1
    onet = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r, pe_add_bit=g.pe_add_bit)
    q5_ref, y_ref = orc.forward_int(onet, x.numpy())
    net = build(g, impl)
    q5, y = net(x.cuda())
    assert np.array_equal(q5.cpu().numpy(), q5_ref)
    assert np.array_equal(y.cpu().numpy(), y_ref)


@pytest.mark.parametrize("impl", impl_ids())
def test_host_entry_point_and_empty_batch(impl):
    g = load_golden("x2_small")
    net = build(g, impl)
    x = torch.from_numpy(g.x).pin_memory()
    (qs, ys) = net.out_shapes(*[x.shape[0], x.shape[2], x.shape[3]])
    q5 = torch.empty(qs, dtype=torch.int8).pin_memory()
    y = torch.empty(ys, dtype=torch.float32).pin_memory()
    net.forward_host(x, q5, y)
    assert np.array_equal(q5.numpy(), g["q5"])
    assert np.array_equal(y.numpy(), g["yout"])
    # empty batch is a no-op, wrong channel count is an error
    e = torch.empty((0, x.shape[1], 8, 8), device="cuda")
    q, yy = net(e)
    assert q.shape[0] == 0 and yy.shape[0] == 0
    with pytest.raises(SesrqError):
        net(torch.zeros((1, x.shape[1] + 1, 8, 8), device="cuda"))


@pytest.mark.parametrize("impl", impl_ids())
def test_chunked_batch_equals_single_frames(impl, monkeypatch):
    # frames are independent units: a chunked multi-frame launch must equal frame-by-frame results
    monkeypatch.setenv("SESRQ_CHUNK_MB", "1")
    g = load_golden("x4_fixture")
    net = build(g, impl)
    gen = torch.Generator().manual_seed(7)
    x = torch.rand(5, 1, 48, 160, generator=gen).cuda()
    q_all, y_all = net(x)
    for b in range(5):
        q1, y1 = net(x[b:b + 1])
        assert torch.equal(q1, q_all[b:b + 1]) and torch.equal(y1, y_all[b:b + 1])


@pytest.mark.parametrize("impl", impl_ids())
@pytest.mark.parametrize("case", ["x2_fixture", "x4_fixture", "synth_neg"])
def test_input_quantiser_rounding_boundaries(case, impl):
    """Layer-0 quantiser is clamp(rint(x / s0 + z0)) with TRUE fp32 division (quan_func.py:225):
    feed inputs sitting on / next to the half-way points, plus zeros, denormals, huge values and negatives."""
    g = load_golden(case)
    cin = g.wf[0].shape[1]
    s0 = np.float32(g.scales[0])
    rng = np.random.default_rng(11)
    H, W = 48, 256
    k = rng.integers(-40, 300, size=(1, cin, H, W)).astype(np.float32)
    x = (k + np.float32(0.5)) * s0                                  # half-way points of the code grid
    ulps = rng.integers(-3, 4, size=x.shape)
    x = (x.view(np.int32) + ulps.astype(np.int32)).view(np.float32)  # +-3 ulp around them
    x[0, :, 0, 0:8] = [0.0, -0.0, 1e-45, -1e-40, 1e-38, 3e38, -3e38, 1.0]
    x[0, :, 1, :] = rng.random((cin, W), dtype=np.float32) * np.float32(1e-30)
    x = np.ascontiguousarray(x, dtype=np.float32)
    onet = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r, pe_add_bit=g.pe_add_bit)
    q5_ref, y_ref = orc.forward_int(onet, x)
    net = build(g, impl)
    q5, y = net(torch.from_numpy(x).cuda())
    assert np.array_equal(q5.cpu().numpy(), q5_ref)
    assert np.array_equal(y.cpu().numpy(), y_ref)


@pytest.mark.parametrize("impl", impl_ids())
@pytest.mark.parametrize("case,shape", [
    ("x2_fixture", (2, 130, 257)), ("x4_fixture", (2, 67, 301)), ("nrdm_fixture", (2, 33, 129)),
    ("x2_fixture", (1, 40, 512)), ("synth_clamp", (1, 40, 150)),
])
def test_single_output_calls_match_both(case, shape, impl):
    """The last layer has separate kernels for "q5 only" and "y only"; each must give what the two-output call gives
    and leave the other buffer alone (canary)."""
    g = load_golden(case)
    B, H, W = shape
    cin = g.wf[0].shape[1]
    gen = torch.Generator().manual_seed(77 + H * W)
    x = torch.rand(B, cin, H, W, generator=gen)
    if int(g["signed_input"]):
        x = x * 2 - 1
    x = x.cuda()
    net = build(g, impl)
    q5_both, y_both = net(x)
    q5_both, y_both = q5_both.clone(), y_both.clone()
    q5_only, none_y = net(x, want_q5=True, want_y=False)
    assert none_y is None
    assert torch.equal(q5_only, q5_both)
    none_q, y_only = net(x, want_q5=False, want_y=True)
    assert none_q is None
    assert torch.equal(y_only, y_both)
    # caller-provided buffers: the unwanted one is not written
    q5_buf = torch.full_like(q5_both, 77)
    y_buf = torch.full_like(y_both, -5.0)
    net(x, want_q5=False, want_y=True, y=y_buf)
    assert torch.equal(y_buf, y_both)
    net(x, want_q5=True, want_y=False, q5=q5_buf)
    assert torch.equal(q5_buf, q5_both)


def test_call_without_any_output_is_rejected():
    g = load_golden("x2_small")
    net = build(g, IMPL_AUTO)
    x = torch.from_numpy(g.x).cuda()
    with pytest.raises(SesrqError):
        net(x, want_q5=False, want_y=False)
