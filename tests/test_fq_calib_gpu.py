"""Fake-quant network and on-device calibration through the C ABI, against the oracles:
  * FakeQuantNet (sesrq_fq_forward) vs oracle/fq_oracle.py -- same formulation, so compared tightly -- and vs the
    reference's own mode-0 output (tolerance);
  * sesrq_minmax_update / sesrq_hist_update: exact; sesrq_kl_threshold: index-exact vs oracle/kl_oracle.py
    (parity unpinned w.r.t. the reference, which has no KL code);
  * MinMaxCalibrator over the reference's calibration frames (torch.rand, seed 1234) reproduces the reference's
    running min/max to rel 1e-5 and its thresholds -> identical (M16, n) requant constants."""
import ctypes as C

import numpy as np
import pytest
import torch

import fq_oracle as fo
import kl_oracle as ko
from conftest import load_golden, SMALL_CASES
from sesr_pytorch_quantize_b200 import _lib, compile_network, thresholds_from_minmax
from sesr_pytorch_quantize_b200.fq import FakeQuantNet
from sesr_pytorch_quantize_b200.calib import MinMaxCalibrator, KLCalibrator, thresholds_from_kl

pytestmark = pytest.mark.gpu


def make_fq(g, **kw):
    mflag = int(g["mflag"])
    return FakeQuantNet([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.r, residual=True,
                        observe_out=(mflag == 3), pe_add_bits=g.pe_add_bit, **kw)


@pytest.mark.parametrize("case", SMALL_CASES)
def test_fq_forward_vs_oracle_and_reference(case):
    g = load_golden(case)
    net = make_fq(g)
    y = net(torch.from_numpy(g.x).cuda()).cpu().numpy()
    y_o, mins_o, maxs_o = fo.forward_fq(g.wf, g.bf, g.x, g.r, observe_out=(int(g["mflag"]) == 3), pe_add_bit=g.pe_add_bit)
    mm = net.frame_minmax.cpu().numpy()[0]
    # same integer-exact formulation as the oracle: identical per-domain ranges and output
    assert np.array_equal(mm[:, 0], np.array(mins_o, dtype=np.float32))
    assert np.array_equal(mm[:, 1], np.array(maxs_o, dtype=np.float32))
    assert np.array_equal(y, y_o)
    ref = g["fq_out"]
    mse = float(((y.astype(np.float64) - ref) ** 2).mean())
    s5 = (g["fq_maxs"][5] - g["fq_mins"][5]) / 255
    assert mse == 0 or 10 * np.log10(1.0 / mse) >= 60.0
    assert float(np.abs(y - ref).max()) <= 8 * s5


def test_fq_batch_frames_with_unusual_ranges():
    """Frames whose minimum is far from code 0 (operand offset dmin != 0: zero padding is then not byte 0), a signed
    frame and a tiny-range frame, batched with ordinary ones: every frame must equal the oracle's single-frame result,
    and the per-frame ranges are reported per frame."""
    g = load_golden("x2_small")
    net = make_fq(g)
    gen = torch.Generator().manual_seed(21)
    base = torch.rand(5, 3, 19, 23, generator=gen)
    x = base.clone()
    x[1] = base[1] * 0.05 + 0.9            # [0.9, 0.95]: min / s ~ 4600 -> large positive dmin
    x[2] = base[2] * 2 - 1                 # signed input: negative dmin (z0 > -128)
    x[3] = base[3] * 1e-3 + 0.25
    y = net(x.cuda()).cpu().numpy()
    fm = net.frame_minmax.cpu().numpy()
    assert fm.shape == (5, 6, 2)
    for b in range(5):
        y_o, mins_o, maxs_o = fo.forward_fq(g.wf, g.bf, x[b:b + 1].numpy(), g.r)
        assert np.array_equal(fm[b, :, 0], np.array(mins_o, dtype=np.float32)), "frame %d mins" % b
        assert np.array_equal(fm[b, :, 1], np.array(maxs_o, dtype=np.float32)), "frame %d maxs" % b
        assert np.array_equal(y[b:b + 1], y_o), "frame %d" % b


def test_fq_batch_is_frame_by_frame_and_running_minmax():
    g = load_golden("x2_small")
    net = make_fq(g)
    gen = torch.Generator().manual_seed(9)
    x = torch.rand(3, 3, 31, 45, generator=gen)
    y = net(x.cuda()).cpu().numpy()
    per = [fo.forward_fq(g.wf, g.bf, x[b:b + 1].numpy(), g.r) for b in range(3)]
    for b in range(3):
        assert np.array_equal(y[b:b + 1], per[b][0])
    mins, maxs = fo.running_minmax([(p[1], p[2]) for p in per])
    rm, rx = net.running_minmax()
    assert rm == [float(np.float32(v)) for v in mins] and rx == [float(np.float32(v)) for v in maxs]


def test_minmax_and_histogram_kernels_exact():
    lib = _lib.load_library()
    gen = torch.Generator().manual_seed(3)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for n in (1, 31, 1000, 4097, 1 << 20, (1 << 22) + 3):
        x = (torch.randn(n + 1, generator=gen) * 2).cuda()[1:]               # deliberately 4-byte (not 16) aligned
        mm = torch.tensor([float("inf"), float("-inf")]).cuda()
        _lib.check(lib.sesrq_minmax_update(C.c_void_p(x.data_ptr()), n, C.c_void_p(mm.data_ptr()), st))
        assert mm.tolist() == [float(x.min()), float(x.max())]
        lo, hi = 0.0, float(x.max())
        h = torch.zeros(2048, dtype=torch.int64).cuda()
        for _ in range(2):                                                    # accumulates across calls
            _lib.check(lib.sesrq_hist_update(C.c_void_p(x.data_ptr()), n, lo, hi, C.c_void_p(h.data_ptr()), st))
        want = ko.histogram(x.cpu().numpy(), lo, hi).astype(np.int64) * 2
        assert np.array_equal(h.cpu().numpy(), want)
    # empty input is a no-op; inverted range is an error
    assert lib.sesrq_minmax_update(None, 0, C.c_void_p(mm.data_ptr()), st) == 0
    assert lib.sesrq_hist_update(C.c_void_p(x.data_ptr()), 10, 1.0, 1.0, C.c_void_p(h.data_ptr()), st) == -1


def _kl_device(hist_np):
    lib = _lib.load_library()
    h = torch.from_numpy(hist_np.astype(np.int64)).cuda()
    idx = torch.zeros(1, dtype=torch.int32).cuda()
    kl = torch.full((2049,), float("inf"), dtype=torch.float64).cuda()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.sesrq_kl_threshold(C.c_void_p(h.data_ptr()), 2048, 128, C.c_void_p(idx.data_ptr()),
                                      C.c_void_p(kl.data_ptr()), st))
    return int(idx.item()), kl.cpu().numpy()


def test_kl_search_index_exact_vs_oracle():
    rng = np.random.default_rng(7)
    hists = []
    v = np.abs(rng.standard_normal(500000)).astype(np.float32)
    hists.append(ko.histogram(v, 0.0, float(v.max())))
    v = np.maximum(rng.standard_normal(300000), 0).astype(np.float32) ** 2          # ReLU-like spike at 0 + long tail
    hists.append(ko.histogram(v, 0.0, float(v.max())))
    hists.append(ko.histogram(rng.random(200000).astype(np.float32), 0.0, 1.0))      # uniform
    h = np.zeros(2048, dtype=np.uint64); h[:128] = 10; hists.append(h)               # exact zero-KL plateau
    h = h.copy(); h[2047] = 1; hists.append(h)                                       # far outlier
    hists.append(np.zeros(2048, dtype=np.uint64))                                    # empty
    h = np.zeros(2048, dtype=np.uint64); h[5] = 1 << 40; hists.append(h)             # one huge bin (64-bit counts)
    for hist in hists:
        idx, curve = _kl_device(hist)
        want = ko.kl_curve(hist)
        assert idx == ko.kl_threshold_index(hist)
        fin = np.isfinite(want)
        assert np.array_equal(fin[128:], np.isfinite(curve[128:]))
        assert np.allclose(curve[fin], want[fin], rtol=1e-9, atol=1e-13)


M16_EXCEPTIONS = {}          # case -> edges (0..4 = layers, 5 = residual) whose M16 is off by one; empty = 6 of 6 identical


@pytest.mark.parametrize("case", ["x2_fixture", "x4_fixture", "nrdm_fixture"])
def test_minmax_calibrator_reproduces_reference_thresholds(case):
    g = load_golden(case)
    net = make_fq(g)
    cal = MinMaxCalibrator(net)
    gen = torch.Generator().manual_seed(1234)                        # the goldens' calibration frames (gen_golden.py)
    h, w = [int(v) for v in g["calib_hw"]]
    cin = g.wf[0].shape[1]
    frames = torch.stack([torch.rand(1, cin, h, w, generator=gen)[0] for _ in range(int(g["calib_n"]))])
    cal.observe(frames.cuda())
    scales, zeros, mins, maxs = cal.finish()
    assert np.allclose(mins, g["mins"], rtol=1e-5, atol=1e-7) and np.allclose(maxs, g["maxs"], rtol=1e-5, atol=1e-7)
    assert zeros == g.zeros
    assert np.allclose(scales, g.scales, rtol=1e-5)
    cn = compile_network([torch.from_numpy(w_) for w_ in g.wf], [torch.from_numpy(b) for b in g.bf], scales, zeros, g.r)
    got = [(L.m16, L.shift_n) for L in cn.layers] + [(cn.res_m16, cn.res_shift_n)]
    want = list(zip([int(v) for v in g["M"]], [int(v) for v in g["n"]]))
    # (M16, n) come from truncation of s_i / s_j * s_w * 2^n: a last-ulp difference in a range can move M16 by one
    assert all(a[1] == b[1] and abs(a[0] - b[0]) <= 1 for a, b in zip(got, want))
    differing = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    print("identical (M16,n) %s:" % case, 6 - len(differing), "of 6; differing edges", differing,
          [(got[i], want[i]) for i in differing])
    # the edges whose M16 differs by one from the reference's (observed ranges agree to rel 1e-5, not bit for bit: the
    # reference's min/max come from an fp32 library conv, ours from the integer-exact one) -- pinned per case
    assert differing == M16_EXCEPTIONS.get(case, []), "edges with a differing M16 changed: %s" % differing


def test_kl_calibrator_two_pass_vs_oracle():
    g = load_golden("x2_small")
    net = make_fq(g)
    cal = KLCalibrator(net)
    gen = torch.Generator().manual_seed(77)
    frames = torch.rand(3, 3, 40, 64, generator=gen)
    cal.observe_pass1(frames.cuda())
    cal.begin_pass2()
    cal.observe_pass2(frames.cuda())
    hist_dev = net.hist.cpu().numpy().copy()
    scales, zeros, idx = cal.finish()
    # oracle: replay the per-frame tensors of the fake-quant path and histogram every domain over the pass-1 ranges
    doms = [[] for _ in range(6)]
    for b in range(3):
        trace = []
        fo.forward_fq(g.wf, g.bf, frames[b:b + 1].numpy(), g.r, trace=trace)
        for d in range(6):
            doms[d].append(trace[d])
    for d in range(6):
        lo, hi = cal.lows[d], cal.maxs[d]
        want = sum(ko.histogram(t, lo, hi).astype(np.int64) for t in doms[d])
        assert np.array_equal(hist_dev[d], want), "histogram of domain %d" % d
        assert idx[d] == ko.kl_threshold_index(want.astype(np.uint64))
    s2, z2 = thresholds_from_kl(cal.mins, cal.maxs, idx)
    assert s2 == scales and z2 == zeros
    assert all(s > 0 for s in scales)


def _float_model(g, x):
    """The collapsed FLOAT network (models/sesr.py:40-50 after collapse()): the quality yardstick, computed with plain
    torch convolutions (checker only)."""
    import torch.nn.functional as F
    w = [torch.from_numpy(v).cuda() for v in g.wf]
    b = [torch.from_numpy(v).cuda() for v in g.bf]
    x1 = F.relu(F.conv2d(x, w[0], b[0], padding=2))
    t = x1
    for i in (1, 2, 3):
        t = F.relu(F.conv2d(t, w[i], b[i], padding=1))        # the residual identity is folded into the centre tap
    t = F.conv2d(t + x1, w[4], b[4], padding=2)
    return F.pixel_shuffle(t, g.r) if g.r > 1 else t


@pytest.mark.parametrize("case", ["x2_fixture", "x4_fixture", "nrdm_fixture"])
def test_kl_thresholds_versus_minmax_quality_against_float_model(case, capsys):
    """Is the KL calibrator useful?  Calibrate both ways on the same frames, run the INT8 SIM with either threshold set
    on the shipped fixture and compare with the float model.  Uniform-noise activations have no outlier tail, so the
    entropy search is expected to keep (almost) the full range; the test records the PSNRs and fails if KL costs more
    than 1 dB anywhere.  Numbers are quoted in DESIGN.md."""
    from sesr_pytorch_quantize_b200 import Int8Net
    g = load_golden(case)
    cin = g.wf[0].shape[1]
    frames = torch.rand(8, cin, 80, 960, generator=torch.Generator().manual_seed(1234)).cuda()
    net = make_fq(g)
    mmc = MinMaxCalibrator(net)
    mmc.observe(frames)
    s_mm, z_mm, _, _ = mmc.finish()
    net.reset_observers()
    klc = KLCalibrator(net)
    klc.observe_pass1(frames)
    klc.begin_pass2()
    klc.observe_pass2(frames)
    s_kl, z_kl, idx = klc.finish()
    x = torch.from_numpy(g.x).cuda()
    ref = _float_model(g, x)
    out = {}
    for name, (s, z) in (("minmax", (s_mm, z_mm)), ("kl", (s_kl, z_kl))):
        cn = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], s, z, g.r)
        _, y = Int8Net(cn)(x, want_q5=False)
        mse = float(((y.double() - ref.double()) ** 2).mean())
        out[name] = 10 * np.log10(float(ref.abs().max()) ** 2 / mse)
    with capsys.disabled():
        print("\\n[kl-vs-minmax] %s: chosen bins %s; int8-sim PSNR vs float model: min/max %.2f dB, KL %.2f dB" %
              (case, idx, out["minmax"], out["kl"]))
    assert all(128 <= i <= 2048 for i in idx)
    assert len(set((i + 1) % 128 for i in idx)) > 1 or all(i == 2048 for i in idx), "the 128k - 1 artefact is back"
    assert out["kl"] >= out["minmax"] - 1.0
