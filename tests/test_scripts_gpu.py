"""The drop-in boundary, proven: the reference's own test.py and sim.py run UNMODIFIED (runpy, __main__) on this
package's import root (define / myQL / models), in the order a user runs them -- test.py calibrates (exe_mode 0 fake-quant
+ running min/max + its calibration tail writing output_pt/input/input.{i}.{scale,zero}.pt), sim.py then simulates in
integers (exe_mode 1).  The scripts come from the staged oracle/_ref copy (oracle/stage_ref.py; never imported from:
only the two script files are executed, every `import` resolves to the mirror or to tests/stubs).

Checked: both scripts exit 0; the graphs ran as fused networks; sim.py's output equals the numpy oracle on the input
the script used, with the constants the run itself produced; thresholds agree with the reference-calibrated goldens.
"""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref")
LAUNCH = os.path.join(ROOT, "tests", "run_ref_script.py")
CASES = {3: ("nrdm_fixture", 3), 5: ("x4_fixture", 4), 6: ("x2_fixture", 3)}


def run_script(name, mflag, work, frames):
    env = dict(os.environ, SESRQ_STUB_FRAMES=str(frames))
    env.pop("CUDA_VISIBLE_DEVICES", None)
    res = os.path.join(work, name + ".result.npz")
    p = subprocess.run([sys.executable, LAUNCH, os.path.join(REF, name + ".py"), str(mflag), work, res],
                       capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, "%s (MFLAG %d) failed:\n%s\n%s" % (name, mflag, p.stdout[-3000:], p.stderr[-3000:])
    return np.load(res), p.stdout


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "sim.py")), reason="oracle/_ref not staged (run __graft_entry__.build())")
@pytest.mark.parametrize("mflag", [3, 5, 6])
def test_reference_scripts_run_unmodified(mflag, tmp_path):
    import sesr_oracle as orc
    case, frames = CASES[mflag]
    g = load_golden(case)
    work = str(tmp_path)
    shutil.copytree(os.path.join(REF, "model_params"), os.path.join(work, "model_params"))
    for f in ("rand_DM_Input_80x960.pt", "rand_SR_Input_80x960.pt"):
        shutil.copy(os.path.join(REF, f), os.path.join(work, f))

    # ---- test.py: fake-quant evaluation + calibration -------------------------------------------------------------
    cal, out = run_script("test", mflag, work, frames)
    assert str(cal["plan"]) == "fq", "test.py's graph did not run as the fused fake-quant network"
    assert "calibrate end" in out and "mean psnr is" in out
    assert int(cal["totalnum"]) == frames and np.isfinite(cal["totalpsnr"]) and np.isfinite(cal["totalssim"])
    scales = [float(cal["scale%d" % i]) for i in range(6)]
    zeros = [int(cal["zero%d" % i]) for i in range(6)]
    # same frames as the golden's calibration set.  The script collapses the expand/squeeze convs on the GPU
    # (model.cuda() precedes collapse(), sim.py:57,77) with torch's library conv -- TF32 by default on this class of
    # GPU -- while the golden run collapsed on the CPU: the folded weights differ by ~1e-3 relative, weight codes near a
    # rounding boundary flip, and the ranges of the later domains move by up to 2e-2 (observed); domain 0 is the input
    np.testing.assert_allclose(scales, g.scales, rtol=5e-2)
    assert abs(scales[0] - g.scales[0]) <= 1e-6 * g.scales[0]
    assert zeros == g.zeros

    # ---- sim.py: integer hardware simulation with the thresholds test.py just wrote -----------------------------------
    sim, out = run_script("sim", mflag, work, frames)
    assert str(sim["plan"]) == "int8", "sim.py's graph did not run as the fused int8 network"
    assert "SIM_mflag: %d" % mflag in out
    x = sim["inps"]
    codes = [sim["W%d" % i] for i in range(5)]
    net = orc.derive_net(None, [sim["b%d" % i] for i in range(5)], scales, zeros, g.r,
                         codes=codes, w_scales=[float(sim["sw%d" % i]) for i in range(5)])
    q5_ref, y_ref = orc.forward_int(net, x)
    assert np.array_equal(sim["gfake"], y_ref), "sim.py output differs from the oracle"
    # the weights the script collapsed on the GPU quantise to the reference's codes (collapsed on the CPU) up to
    # isolated codes on a rounding boundary
    for i in range(5):
        assert np.mean(codes[i] != g["W%d" % i]) <= 0.01
    if mflag == 5:            # loads the shipped SR fixture; MFLAG 3 / 6 run on the unseeded torch.rand the script itself
        assert np.array_equal(x, g.x)     # draws and saves over rand_DM_Input_80x960.pt (sim.py:197-203)
