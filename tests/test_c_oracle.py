"""The C restatement of the oracle against the numpy oracle and the reference goldens.  CPU only."""
import hashlib

import numpy as np

import c_oracle
import sesr_oracle as orc


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_c_oracle_matches_reference_goldens(golden):
    g = golden
    net = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r, pe_add_bit=g.pe_add_bit)
    q5, y, tr = c_oracle.CNet(net).forward(g.x, want_trace=True)
    assert np.array_equal(q5, g["q5"])
    assert sha(y) == g.hashes["yout"]
    for i in range(5):
        cin = net.layers[i].cin
        assert sha(tr[i, :cin][None]) == g.hashes["q%d" % i]


def test_c_oracle_matches_numpy_oracle_batch():
    from conftest import load_golden
    g = load_golden("x2_fixture")
    net = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r)
    x = np.random.default_rng(0).random((3, 3, 37, 91), dtype=np.float32)
    q_np, y_np = orc.forward_int(net, x)
    q_c, y_c = c_oracle.CNet(net).forward(x)
    assert np.array_equal(q_np, q_c) and np.array_equal(y_np, y_c)
