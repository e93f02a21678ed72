"""The fused pair (layers 1 + 2 in one launch, csrc/sesrq_chain.cu) against the per-layer path and the C oracle.

The fused kernel streams strips of rows through rings of 16 rows with 4- and 8-row barrier groups, splits tall frames
into segments and the width into 124-pixel columns, so the shapes below sit on and around every one of those boundaries.
Bit-exact (integer work): the intermediate codes only change where they travel, not what they are."""
import numpy as np
import pytest
import torch

import c_oracle
import sesr_oracle as orc
from conftest import load_golden
from sesr_pytorch_quantize_b200 import Int8Net, SesrqError, compile_network

pytestmark = pytest.mark.gpu


def pair(case):
    g = load_golden(case)
    cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r)
    return g, Int8Net(cnet, fuse=True), Int8Net(cnet, fuse=False)


def test_reference_networks_take_the_fused_path():
    for case in ("x2_fixture", "x4_fixture", "nrdm_fixture"):
        g, fused, plain = pair(case)
        assert fused.fused and not plain.fused, case
        x = torch.rand(1, g.wf[0].shape[1], 20, 40).cuda()
        fused(x)
        with pytest.raises(SesrqError):
            fused.read_activation(2)            # never written to HBM
        plain(x)
        plain.read_activation(2)
        assert fused.set_fusion(False) is True and not fused.fused      # returns the previous setting
        fused(x)
        assert torch.equal(fused.read_activation(2), plain.read_activation(2))


HEIGHTS = [1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 14, 15, 16, 17, 18, 19, 29, 30, 31, 32, 33, 34, 46, 47, 48, 49, 63, 64, 65, 66, 100]


@pytest.mark.parametrize("case", ["x2_fixture", "nrdm_fixture"])
def test_every_row_count_around_the_ring_boundaries(case):
    g, fused, plain = pair(case)
    cin = g.wf[0].shape[1]
    gen = torch.Generator().manual_seed(7)
    for H in HEIGHTS:
        x = torch.rand(2, cin, H, 150, generator=gen).cuda()
        q1, y1 = fused(x)
        q0, y0 = plain(x)
        assert torch.equal(q1, q0) and torch.equal(y1, y0), "H = %d" % H


@pytest.mark.parametrize("W", [1, 2, 3, 60, 122, 123, 124, 125, 126, 127, 128, 129, 247, 248, 249, 250, 372, 373, 500])
def test_every_width_around_the_column_boundaries(W):
    g, fused, plain = pair("x4_fixture")
    gen = torch.Generator().manual_seed(W)
    x = torch.rand(3, 1, 37, W, generator=gen).cuda()
    q1, y1 = fused(x)
    q0, y0 = plain(x)
    assert torch.equal(q1, q0) and torch.equal(y1, y0)


@pytest.mark.parametrize("B,H,W", [
    (1, 300, 100),      # one column, few strips: the frame is cut into many row segments
    (1, 1000, 124),     # segment boundaries at arbitrary rows
    (2, 515, 260),
    (5, 97, 700),
    (40, 33, 250),      # more strips than CTAs, several per CTA: ring phases carried across strips
    (200, 18, 124),
])
def test_segments_and_many_strips_vs_c_oracle(B, H, W):
    g, fused, plain = pair("x2_fixture")
    onet = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r)
    cn = c_oracle.CNet(onet)
    gen = torch.Generator().manual_seed(B * 131 + H)
    x = torch.rand(B, 3, H, W, generator=gen)
    q5_ref, y_ref = cn.forward(x.numpy())
    q1, y1 = fused(x.cuda())
    assert np.array_equal(q1.cpu().numpy(), q5_ref)
    assert np.array_equal(y1.cpu().numpy(), y_ref)


@pytest.mark.parametrize("case,B,H,W", [
    ("x2_fixture", 16, 1080, 1920),        # the bench step (BASELINE config 5)
    ("x4_fixture", 64, 540, 960),          # config 3 at its full batch
    ("nrdm_fixture", 512, 80, 960),        # config 4 at its largest batch
])
def test_fused_equals_per_layer_at_baseline_sizes(case, B, H, W):
    """Size-independent property at BASELINE.json's full sizes: the fused pair and the per-layer path give the same bytes."""
    g, fused, plain = pair(case)
    gen = torch.Generator().manual_seed(B + H)
    x = torch.rand(B, g.wf[0].shape[1], H, W, generator=gen).cuda()
    q1, y1 = fused(x)
    h1 = (int(q1.view(torch.uint8).to(torch.int64).sum()), float(y1.double().sum()))
    assert torch.equal(q1, plain(x, want_y=False)[0])
    del q1
    y0 = plain(x, want_q5=False)[1]
    assert torch.equal(y1, y0)
    assert h1[0] > 0


def test_extreme_inputs_and_determinism():
    g, fused, plain = pair("x2_fixture")
    for fill in (0.0, 1.0, -3.0, 1e6):
        x = torch.full((2, 3, 50, 130), fill).cuda()
        q1, y1 = fused(x)
        q0, y0 = plain(x)
        assert torch.equal(q1, q0) and torch.equal(y1, y0), fill
    x = torch.rand(4, 3, 270, 480).cuda()
    a = fused(x)
    b = fused(x)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_networks_the_fused_kernel_cannot_take_fall_back():
    """A zero point above -128 after layer 1 needs the explicit ReLU floor and a non-zero pad byte: per-layer kernels."""
    g = load_golden("synth_zero")
    cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r,
                           pe_add_bits=g.pe_add_bit)
    net = Int8Net(cnet, fuse=True)
    onet = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r, pe_add_bit=g.pe_add_bit)
    x = torch.rand(2, g.wf[0].shape[1], 30, 140)
    q5_ref, y_ref = orc.forward_int(onet, x.numpy())
    q5, y = net(x.cuda())
    assert np.array_equal(q5.cpu().numpy(), q5_ref) and np.array_equal(y.cpu().numpy(), y_ref)
