"""Randomised parity: networks with random weights, scales, zero points, channel counts and clamp widths (seeded, so the
run is reproducible) through every accumulator layout of the tcgen05 path (single set, spare columns, two sets, four
sets) and the dp4a fallback, bit-exact against the numpy oracle.  The generator biases some weight tensors so that the
18-bit PE clamp and the adder-tree clamp really fire (checked via the oracle's trace)."""
import numpy as np
import pytest
import torch

import sesr_oracle as orc
from sesr_pytorch_quantize_b200 import Int8Net, compile_network, IMPL_AUTO, IMPL_GENERIC

pytestmark = pytest.mark.gpu


def random_net(seed):
    rng = np.random.default_rng(seed)
    cin = int(rng.integers(1, 5))
    r = int(rng.choice([1, 1, 2, 4]))
    cout = {1: int(rng.choice([2, 3, 5, 12, 16])), 2: int(rng.choice([4, 8, 12, 16])), 4: 16}[r]
    chans = [cin, 16, 16, 16, 16, cout]
    ks = [5, 3, 3, 3, 5] if rng.random() < 0.8 else [3, 3, 3, 3, 3]
    wf, bf = [], []
    for i in range(5):
        shift = rng.choice([0.0, 0.0, 0.35, -0.35, 0.5])             # same-sign sums -> PE clamps become reachable
        w = (rng.random((chans[i + 1], chans[i], ks[i], ks[i])) - 0.5 + shift) * rng.uniform(0.01, 0.08)
        if rng.random() < 0.5:
            # "peaky" tensor like the trained checkpoints: one large tap sets the scale, the rest are a few codes wide,
            # so the static bound proves the PE clamps inert (single-accumulator layout)
            w = w * 0.03
            w.flat[int(rng.integers(w.size))] = 0.08 * rng.choice([-1.0, 1.0])
        wf.append(w.astype(np.float32))
        bf.append(((rng.random(chans[i + 1]) - 0.5) * rng.uniform(0.05, 3.0)).astype(np.float32))
    scales = [float(rng.uniform(0.003, 0.01))] + [float(rng.uniform(0.002, 0.05)) for _ in range(5)]
    if rng.random() < 0.7:
        zeros = [-128] * 6
    else:
        zeros = [int(rng.choice([-128, -128, -131, -120, -100])) for _ in range(6)]
    pe_add_bit = int(rng.choice([20, 20, 19, 18]))
    return wf, bf, scales, zeros, r, pe_add_bit


@pytest.mark.parametrize("seed", list(range(24)))
def test_random_network_bit_exact(seed):
    wf, bf, scales, zeros, r, pe_add_bit = random_net(seed)
    rng = np.random.default_rng(1000 + seed)
    B, H, W = int(rng.integers(1, 4)), int(rng.integers(1, 60)), int(rng.choice([16, 48, 130, 257, 96, 33]))
    x = rng.random((B, wf[0].shape[1], H, W), dtype=np.float32)
    onet = orc.derive_net(wf, bf, scales, zeros, r, pe_add_bit=pe_add_bit)
    q5_ref, y_ref = orc.forward_int(onet, x)
    cnet = compile_network([torch.from_numpy(w) for w in wf], [torch.from_numpy(b) for b in bf], scales, zeros, r,
                           pe_add_bits=pe_add_bit)
    for impl in (IMPL_AUTO, IMPL_GENERIC):
        net = Int8Net(cnet, impl=impl)
        q5, y = net(torch.from_numpy(x).cuda())
        assert np.array_equal(q5.cpu().numpy(), q5_ref), "q5 differs (seed %d, impl %d, layers %s)" % (
            seed, impl, [net.layer_info(i) for i in range(5)])
        assert np.array_equal(y.cpu().numpy(), y_ref), "y differs (seed %d, impl %d)" % (seed, impl)


def test_generator_reaches_every_layout():
    """The seeds above must cover: tcgen05 layers with and without provable clamp inertness, and the dp4a fallback."""
    seen = set()
    for seed in range(24):
        wf, bf, scales, zeros, r, pe_add_bit = random_net(seed)
        cnet = compile_network([torch.from_numpy(w) for w in wf], [torch.from_numpy(b) for b in bf], scales, zeros, r,
                               pe_add_bits=pe_add_bit)
        net = Int8Net(cnet)
        for i in range(5):
            seen.add(net.layer_info(i))
    impls = {s[0] for s in seen}
    assert 1 in impls and 2 in impls                      # generic and tcgen05 both exercised
    assert (2, 0) in seen and (2, 1) in seen              # tcgen05 with inert clamps and with clamps that can fire


def test_every_accumulator_layout_is_reached_and_bit_exact():
    """NS = 1, the spare-column layout, NS = 2 and NS = 4 (csrc/sesrq_umma.cu header) must EACH be selected by some network
    of the test set -- the random networks above plus the reference fixtures -- so that the bit-exactness tests cover all
    four epilogue / weight-image variants, not only the common one."""
    from conftest import load_golden, ALL_CASES
    layouts = {}
    for seed in range(24):
        wf, bf, scales, zeros, r, pe_add_bit = random_net(seed)
        cnet = compile_network([torch.from_numpy(w) for w in wf], [torch.from_numpy(b) for b in bf], scales, zeros, r,
                               pe_add_bits=pe_add_bit)
        net = Int8Net(cnet)
        for i in range(5):
            layouts.setdefault(net.layer_layout(i), []).append("seed%d/L%d" % (seed, i))
    for case in ALL_CASES:
        g = load_golden(case)
        cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r,
                               pe_add_bits=g.pe_add_bit)
        net = Int8Net(cnet)
        for i in range(5):
            layouts.setdefault(net.layer_layout(i), []).append("%s/L%d" % (case, i))
    print({k: len(v) for k, v in layouts.items()}, {k: v[:2] for k, v in layouts.items()})
    for want in ("ns1", "spare", "ns2", "ns4", "generic"):
        assert want in layouts, "no network of the test set selects the %s layout: %s" % (want, {k: len(v) for k, v in layouts.items()})
