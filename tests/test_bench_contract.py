"""bench.py contract checks that need no GPU: the reference arm (the C restatement of the reference path timed on the
host cores) prints exactly one JSON line with the agreed keys, and the default arm refuses to run without a GPU instead
of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=300, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "sesr_x2_int8_sim_80x960")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line"
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["unit"] == "Mpix/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "sesr_x2_int8_sim_80x960"
    # same config as the product arm's default for this workload (frames per step included), all host cores
    assert line["config"]["frames_per_gpu_per_step"] == 256
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    # the real Python reference, timed from the staged oracle/_ref (or reported unavailable where it is not staged)
    pr = line["python_reference"]
    assert "unavailable" in pr or (pr["kind"] == "reference" and 0 < pr["value"] < line["value"])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_default_arm_needs_a_gpu():
    p = run_bench("--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e", "--calib-frames", "0")
    assert p.returncode != 0, "the product arm must fail loudly without a GPU, not fall back"
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")]


def test_reference_arm_under_torchrun_prints_once():
    """Launched the way the driver launches N > 1: rank 0 alone runs and prints, the other rank exits 0."""
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--no-python-ref",
                        "--workload", "sesr_x2_int8_sim_80x960"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2
    # torchrun exports OMP_NUM_THREADS=1: the arm must still use every core it may run on
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
