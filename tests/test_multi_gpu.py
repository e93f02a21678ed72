"""Multi-GPU parity on the GPU box: frames sharded over N ranks give bit-identical results to
one GPU and to the oracle, and the all-reduced calibration statistics give exactly the single-process thresholds.
One process per rank through torch.distributed.run -- NCCL with one GPU per rank when the box has that many GPUs,
otherwise the ranks share cuda:0 and the collectives run over gloo (tests/multi_gpu_worker.py).  The pure host logic of
the merge is covered on CPU by tests/test_dist_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return sorted({2, max(n, 2)})


@pytest.mark.parametrize("world", _worlds())
def test_sharded_run_and_merged_calibration_equal_single_gpu(world):
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-4000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    r = json.loads(lines[0])
    print(r)
    assert r["world"] == world
    assert r["backend"] == ("nccl" if torch.cuda.device_count() >= world else "gloo")
    assert r["covered_every_frame_once"] and r["sharded_equals_single"]
    assert all(r["oracle_ok_per_rank"]) and len(r["oracle_ok_per_rank"]) == world
    assert r["minmax_thresholds_equal"] and r["kl_indices_equal"]
