"""Host-side compile step of the product against the reference's dumped constants.  CPU only."""
import numpy as np
import torch

from sesr_pytorch_quantize_b200 import compile as cp


def test_compile_matches_reference_dumps(golden):
    g = golden
    net = cp.compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf],
                             g.scales, g.zeros, g.r, pe_add_bits=g.pe_add_bit)
    for i, L in enumerate(net.layers):
        assert np.array_equal(L.weight.numpy(), g["W%d" % i])
        assert L.w_scale == float(g["sw%d" % i])
        assert L.add_const == [int(v) for v in g["K%d" % i]]
    assert [L.m16 for L in net.layers] + [net.res_m16] == [int(v) for v in g["M"]]
    assert [L.shift_n for L in net.layers] + [net.res_shift_n] == [int(v) for v in g["n"]]


def test_thresholds_tail(golden):
    g = golden
    sc, zr = cp.thresholds_from_minmax(g["mins"].tolist(), g["maxs"].tolist())
    assert sc == g.scales
    if g.name != "synth_zero":
        assert zr == g.zeros


def test_quan_layer_between_const_cases():
    # >= 1 branch, < 1 branch, shift cap at 32 (quan_func.py:495-515)
    assert cp.quan_layer_between_const(1.0) == (32768, 15)
    assert cp.quan_layer_between_const(3.5) == (int(3.5 * 2 ** 14), 14)
    assert cp.quan_layer_between_const(0.5) == (32768, 16)
    assert cp.quan_layer_between_const(0.3) == (int(0.3 * 2 ** 17), 17)
    m, n = cp.quan_layer_between_const(1e-9)
    assert n == 32 and m == int(1e-9 * 2 ** 32)
