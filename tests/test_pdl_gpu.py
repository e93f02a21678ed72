"""Programmatic dependent launch (csrc/sesrq_umma_dev.cuh: launch_pdl / pdl_wait): every tcgen05 kernel may start its
prologue while the previous kernel of the stream is still running and must not touch an activation before
griddepcontrol.wait.  A hazard would show as a wrong or run-to-run varying result when launches follow each other without
host synchronisation, so: many back-to-back forwards of alternating shapes (buffers reused across calls, the last layer
of one frame still running when the first layer of the next starts) must reproduce, bit for bit, what a process with
SESRQ_PDL=0 (plain stream-ordered launches) computes for the same inputs."""
import hashlib
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_golden
from sesr_pytorch_quantize_b200 import Int8Net, compile_network

pytestmark = pytest.mark.gpu
SHAPES = [(1, 80, 960), (3, 37, 131), (2, 270, 500), (1, 16, 124), (5, 64, 250)]


def digests(rounds):
    g = load_golden("x2_fixture")
    cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r)
    net = Int8Net(cnet, device="cuda:0")
    gen = torch.Generator().manual_seed(2024)
    xs = [torch.rand(B, 3, H, W, generator=gen).cuda() for (B, H, W) in SHAPES]
    outs = []
    for _ in range(rounds):                               # no host synchronisation inside the loop
        for x in xs:
            outs.append(net(x))
    torch.cuda.synchronize()
    return [hashlib.sha256(q.cpu().numpy().tobytes() + y.cpu().numpy().tobytes()).hexdigest() for q, y in outs]


def test_back_to_back_launches_equal_plain_stream_order():
    env = dict(os.environ, SESRQ_PDL="0", PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]))
    code = "import test_pdl_gpu as t; print('\\n'.join(t.digests(1)))"
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600, cwd=os.path.join(ROOT, "tests"))
    assert p.returncode == 0, p.stderr[-2000:]
    want = p.stdout.split()
    assert len(want) == len(SHAPES)
    got = digests(8)
    for i, d in enumerate(got):
        assert d == want[i % len(SHAPES)], "forward %d (shape %s) differs from the plain-launch result" % (i, SHAPES[i % len(SHAPES)])
