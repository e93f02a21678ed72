"""The oracle against the committed golden vectors (generated from the unmodified reference by
oracle/gen_golden.py).  CPU only."""
import hashlib

import numpy as np

import sesr_oracle as orc
from conftest import load_golden


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_constants_match_reference(golden):
    g = golden
    net = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r, pe_add_bit=g.pe_add_bit)
    for i, L in enumerate(net.layers):
        assert np.array_equal(L.W, g["W%d" % i])
        assert L.sw == float(g["sw%d" % i])
        assert np.array_equal(L.K, g["K%d" % i].astype(np.int64))
    assert [L.M for L in net.layers] + [net.Mr] == [int(v) for v in g["M"]]
    assert [L.n for L in net.layers] + [net.nr] == [int(v) for v in g["n"]]


def test_oracle_forward_matches_reference(golden):
    g = golden
    net = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r, pe_add_bit=g.pe_add_bit)
    trace = {}
    q5, y = orc.forward_int(net, g.x, trace=trace)
    assert np.array_equal(q5, g["q5"])
    assert sha(y.astype(np.float32)) == g.hashes["yout"]
    for i in range(5):
        assert sha(trace["q%d" % i]) == g.hashes["q%d" % i]
        assert sha(trace["pe_add%d" % i]) == g.hashes["pe_add%d" % i]
        if "pe%d_0" % i in g:
            for p in range(4):
                assert np.array_equal(trace["pe%d_%d" % (i, p)].reshape(-1), g["pe%d_%d" % (i, p)].reshape(-1))
    if "shortcut" in g:
        assert np.array_equal(trace["shortcut"].reshape(-1), g["shortcut"].reshape(-1))
        assert np.array_equal(y, g["yout"])


def test_minmax_tail_matches_reference(golden):
    g = golden
    if g.name == "synth_zero":      # zeros were hand-set for that case
        sc, _ = orc.minmax_to_thresholds(g["mins"], g["maxs"])
        assert sc == g.scales
        return
    sc, zr = orc.minmax_to_thresholds(g["mins"], g["maxs"])
    assert sc == g.scales and zr == g.zeros


def test_survey_anchor_requant_factors():
    # SURVEY.md section 8c table (torch 2.11 CPU, seed 1234)
    g = load_golden("x4_fixture")
    assert list(zip(g["M"].tolist(), g["n"].tolist())) == [(33775, 22), (64755, 22), (37038, 22), (35941, 19), (35380, 21), (45681, 18)]
    g = load_golden("x2_fixture")
    assert list(zip(g["M"].tolist(), g["n"].tolist())) == [(41136, 21), (40925, 21), (63706, 22), (56879, 20), (63645, 21), (33086, 18)]


def test_static_clamp_bounds():
    # 18/20-bit clamps provably inert for x4 and NRDM; x2 conv_last exceeds the 18-bit bound
    for name, expect in (("x4_fixture", [True] * 5), ("nrdm_fixture", [True] * 5), ("x2_fixture", [True, True, True, True, False])):
        g = load_golden(name)
        got = [orc.pe_clamp_bounds(g["W%d" % i])[0] for i in range(5)]
        assert got == expect, (name, got)
