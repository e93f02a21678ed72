"""Pins of the two extra oracles (CPU only):
  * oracle/fq_oracle.py (fake-quant / exe_mode 0 path) against the UNMODIFIED reference's mode-0 output and
    per-domain min/max stored in tests/golden/*_small.npz (tolerance-based, see the oracle's header);
  * oracle/kl_oracle.py (parity unpinned: the reference has no KL code) against hand-computable cases."""
import numpy as np
import pytest

import fq_oracle as fo
import kl_oracle as ko
from conftest import load_golden, SMALL_CASES

PSNR_MIN_DB = 60.0          # at range 1.0
MAXABS_LSB = 8.0            # in units of the frame's output step s_5


@pytest.mark.parametrize("case", SMALL_CASES)
def test_fq_oracle_against_reference_mode0(case):
    g = load_golden(case)
    y, mins, maxs = fo.forward_fq(g.wf, g.bf, g.x, g.r, residual=True, observe_out=(int(g["mflag"]) == 3),
                                  pe_add_bit=g.pe_add_bit)
    ref = g["fq_out"]
    assert y.shape == ref.shape
    mse = float(((y.astype(np.float64) - ref) ** 2).mean())
    psnr = float("inf") if mse == 0 else 10 * np.log10(1.0 / mse)
    s5 = (g["fq_maxs"][5] - g["fq_mins"][5]) / 255
    assert psnr >= PSNR_MIN_DB
    assert float(np.abs(y - ref).max()) <= MAXABS_LSB * s5
    assert np.allclose(mins, g["fq_mins"], rtol=1e-5, atol=1e-7)
    assert np.allclose(maxs, g["fq_maxs"], rtol=1e-5, atol=1e-7)


def test_histogram_known_answers():
    x = np.array([0.0, 0.49, 0.5, 0.999, 1.0, 2.0, -1.0], dtype=np.float32)
    h = ko.histogram(x, 0.0, 1.0, bins=4)
    # clamp to [0,1]; x == hi lands in the last bin; bins of width 0.25
    assert h.tolist() == [2, 1, 1, 3]
    rng = np.random.default_rng(0)
    v = rng.random(100000, dtype=np.float32) * 3
    h = ko.histogram(v, 0.0, 3.0)
    assert int(h.sum()) == v.size and h.shape == (2048,)


def test_kl_search_known_answers():
    # all mass inside the first 128 bins, uniform: every candidate reproduces P exactly -> KL = 0 from i = 128 on;
    # the published tie-break keeps the LAST minimum
    h = np.zeros(2048, dtype=np.uint64)
    h[:128] = 10
    c = ko.kl_curve(h)
    assert c[128] == 0.0 and c[2048] == 0.0 and ko.kl_threshold_index(h) == 2048
    assert np.isinf(c[:128]).all()
    # a far outlier bin: candidates that end on an empty bin while outliers exist are infinite (P > 0, Q = 0)
    h2 = h.copy()
    h2[2047] = 1
    c2 = ko.kl_curve(h2)
    assert np.isinf(c2[129]) and np.isfinite(c2[128]) and np.isfinite(c2[2048])
    # empty histogram -> keep the whole range
    assert ko.kl_threshold_index(np.zeros(2048, dtype=np.uint64)) == 2048
    # long-tailed distribution: threshold strictly inside the range, curve finite there
    rng = np.random.default_rng(1)
    v = np.abs(rng.standard_normal(400000)).astype(np.float32)
    hv = ko.histogram(v, 0.0, float(v.max()))
    i = ko.kl_threshold_index(hv)
    assert 128 < i < 2048
    assert ko.threshold_value(i, 0.0, 1.0) == i / 2048               # calib_bin_edges[i]
    # scale invariance: multiplying every count by a constant leaves the curve (almost) unchanged
    ca, cb = ko.kl_curve(hv), ko.kl_curve(hv * np.uint64(3))
    fin = np.isfinite(ca)
    assert np.array_equal(fin, np.isfinite(cb)) and np.allclose(ca[fin], cb[fin], rtol=1e-9, atol=1e-12)


def test_closed_form_equals_published_algorithm_restated():
    """kl_curve / kl_threshold_index (what the CUDA kernels follow) against the statement-by-statement restatement of
    NVIDIA pytorch-quantization's `_compute_amax_entropy` (python loops, np.digitize, scipy.stats.entropy) -- on
    histograms with a ReLU spike in bin 0, empty bins, a heavy tail, and remainders in the group mapping."""
    rng = np.random.default_rng(3)
    cases = []
    for t in range(3):
        x = np.abs(rng.standard_normal(60000)).astype(np.float32) * (1 + t)
        x[rng.random(x.size) < 0.35] = 0                                     # post-ReLU zeros
        cases.append((ko.histogram(x, 0.0, float(x.max()), bins=512), 512, 32))
    h = ko.histogram(rng.random(20000).astype(np.float32) ** 3, 0.0, 1.0, bins=300)   # 300 = not a multiple of the target
    h[40:60] = 0
    cases.append((h, 300, 16))        # target counts are powers of two (128 in production): np.linspace edges are then exact
    for hist, bins, tgt in cases:
        i_ref, div, args = ko.nvidia_entropy_reference(hist, tgt=tgt, start_bin=tgt)
        curve = ko.kl_curve(hist, bins=bins, tgt=tgt)
        assert ko.kl_threshold_index(hist, bins=bins, tgt=tgt) == i_ref
        fin = np.isfinite(div)
        assert np.array_equal(np.isfinite(curve[tgt:]), fin)
        assert np.allclose(curve[tgt:][fin], div[fin], rtol=1e-9, atol=1e-12)


def test_remainder_is_spread_not_dumped():
    """The r1 artefact: with m = i // 128 and the remainder in the last group every chosen i was 128k - 1.  The published
    mapping floor(k * 128 / i) gives group sizes that differ by at most one bin."""
    for i in (255, 383, 1000, 2047):
        g = (np.arange(i) * 128) // i
        sizes = np.bincount(g, minlength=128)
        assert sizes.max() - sizes.min() <= 1 and sizes.sum() == i
