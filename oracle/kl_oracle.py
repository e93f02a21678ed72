"""CPU ORACLE (numpy) for the activation histogram and the KL-entropy threshold search.
TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import it; the product path never does.

The reference abandoned KL calibration -- define.py:16-20 keeps only the commented constants USE_KL_CALI,
BINS_NUM = 2048, TGT_BINS_NUM = 128 ("KL discarded"); no histogram or entropy code exists in the tree, so
there is nothing in the REFERENCE to pin this file against ("parity unpinned" w.r.t. the reference).
Instead the search is pinned to a PUBLISHED entropy calibrator: NVIDIA pytorch-quantization 2.1
(`pytorch_quantization/calib/histogram.py::_compute_amax_entropy`, the TensorRT "entropy calibration 2"
algorithm), sized by the two surviving constants (2048 source bins, 128 target bins = its `1 << (8 - 1)` for
signed 8-bit).  The package is not installable here (no network, not in the wheelhouse), so
`nvidia_entropy_reference()` below restates that function line by line from its published source -- python
loops, np.digitize, scipy.stats.entropy -- and `kl_curve()` / `kl_threshold_index()` are the closed-form
equivalent the CUDA kernels follow; tests/test_fq_kl_oracle.py checks that the two agree on the chosen bin
and on the whole divergence curve.  The CUDA kernels (csrc/sesrq_calib.cu) must reproduce the histogram
counts exactly and the chosen bin count index-exactly.

  histogram   bin = min(B-1, max(0, floor((clamp(x, lo, hi) - lo) * (B / (hi - lo)))))   all in fp32
  search      (published algorithm)
              h[0] := h[1]                          the zero bin of a post-ReLU tensor must not dominate
              for i in [tgt, B]:   P = h[0:i], outlier mass sum(h[i:]) added to P[i-1]
                                   Q = h[0:i] merged into tgt groups, source bin k -> group floor(k * tgt / i)
                                       (np.digitize on np.linspace(0, i, tgt + 1): the remainder is spread over the
                                       groups, not dumped into the last one), each group's mass spread uniformly over
                                       its non-empty bins
                                   KL_i = sum_k P_k/total * log((P_k/total) / (Q_k/inside))    in fp64
              a bin with P_k > 0 but h[k] == 0 (only possible for k = i-1) makes KL_i infinite;
              answer = LARGEST i attaining the minimum (`len - 1 - argmin(reversed)`); all-infinite -> B
  threshold   T = lo + i * (hi - lo) / B            (`calib_bin_edges[i]`: the upper edge of the last kept bin)
"""
from __future__ import annotations

import numpy as np

BINS = 2048
TARGET_BINS = 128


def histogram(x: np.ndarray, lo: float, hi: float, bins: int = BINS) -> np.ndarray:
    f = np.float32
    lo32, hi32 = f(lo), f(hi)
    assert hi32 > lo32
    scale = f(f(bins) / f(hi32 - lo32))
    c = np.minimum(np.maximum(x.astype(f).reshape(-1), lo32), hi32)
    b = np.floor(((c - lo32).astype(f) * scale).astype(f)).astype(np.int64)
    b = np.clip(b, 0, bins - 1)
    return np.bincount(b, minlength=bins).astype(np.uint64)


def _published_hist(hist: np.ndarray) -> np.ndarray:
    h = hist.astype(np.int64).copy()
    if h.size > 1:
        h[0] = h[1]                               # `bins[0] = bins[1]`
    return h


def kl_curve(hist: np.ndarray, bins: int = BINS, tgt: int = TARGET_BINS) -> np.ndarray:
    """float64 [bins + 1]; entries below tgt are +inf (never candidates).  Closed form of the published loop."""
    h = _published_hist(hist)
    out = np.full(bins + 1, np.inf)
    total = int(h.sum())
    if total == 0:
        return out
    pref = np.concatenate([[0], np.cumsum(h)])
    nz = np.concatenate([[0], np.cumsum(h != 0)])
    inv_total = 1.0 / float(total)
    for i in range(tgt, bins + 1):
        inside = int(pref[i])
        if inside == 0:
            continue
        outliers = total - inside
        k = np.arange(i)
        hk = h[:i]
        pk = hk.copy()
        pk[i - 1] += outliers
        live = pk > 0
        if np.any(live & (hk == 0)):
            continue
        g = (k * tgt) // i                         # np.digitize(range(i), np.linspace(0, i, tgt + 1)) - 1
        start = (g * i + tgt - 1) // tgt           # first source bin of group g: ceil(g * i / tgt)
        stop = ((g + 1) * i + tgt - 1) // tgt
        gs = (pref[stop] - pref[start]).astype(np.float64)
        gn = (nz[stop] - nz[start]).astype(np.float64)
        inv_inside = 1.0 / float(inside)
        with np.errstate(divide="ignore", invalid="ignore"):
            p = pk.astype(np.float64) * inv_total
            q = (gs / gn) * inv_inside
            term = p * np.log(p / q)
        out[i] = float(np.sum(term[live]))
    return out


def kl_threshold_index(hist: np.ndarray, bins: int = BINS, tgt: int = TARGET_BINS) -> int:
    curve = kl_curve(hist, bins, tgt)
    if not np.isfinite(curve).any():
        return bins
    rev = curve[::-1]
    return int(len(curve) - 1 - np.argmin(rev))      # the published tie-break: the last (largest i) minimum


def threshold_value(idx: int, lo: float, hi: float, bins: int = BINS) -> float:
    return lo + idx * (hi - lo) / bins


def nvidia_entropy_reference(hist: np.ndarray, tgt: int = TARGET_BINS, start_bin: int = TARGET_BINS, stride: int = 1):
    """`_compute_amax_entropy` of NVIDIA pytorch-quantization 2.1 (calib/histogram.py), restated statement by
    statement (num_bits = 8, unsigned = False -> nbins = 128 = tgt).  Returns (chosen bin index i, divergences,
    arguments); the published function returns calib_bin_edges[i].  Slow (python loops): small inputs only."""
    from collections import Counter
    from scipy.stats import entropy
    bins = np.asarray(hist, dtype=np.float64).copy()
    bins[0] = bins[1]
    total_data = np.sum(bins)
    divergences, arguments = [], []
    nbins = tgt
    stop = len(bins)
    new_density_counts = np.zeros(nbins, dtype=np.float64)
    for i in range(start_bin, stop + 1, stride):
        new_density_counts.fill(0)
        space = np.linspace(0, i, num=nbins + 1)
        digitized_space = np.digitize(range(i), space) - 1
        digitized_space[bins[:i] == 0] = -1
        for idx, digitized in enumerate(digitized_space):
            if digitized != -1:
                new_density_counts[digitized] += bins[idx]
        counter = Counter(digitized_space)
        for key, val in counter.items():
            if key != -1:
                new_density_counts[key] = new_density_counts[key] / val
        new_density = np.zeros(i, dtype=np.float64)
        for idx, digitized in enumerate(digitized_space):
            if digitized != -1:
                new_density[idx] = new_density_counts[digitized]
        total_counts_new = np.sum(new_density) + np.sum(bins[i:])
        new_density = new_density / max(np.sum(new_density), 1e-300)
        reference_density = np.array(bins[:len(digitized_space)])
        reference_density[-1] += np.sum(bins[i:])
        total_counts_old = np.sum(reference_density)
        if round(total_counts_new) != total_data or round(total_counts_old) != total_data:
            raise RuntimeError("Count mismatch! total_counts_new={}, total_counts_old={}, total_data={}".format(
                total_counts_new, total_counts_old, total_data))
        reference_density = reference_density / max(np.sum(reference_density), 1e-300)
        divergences.append(entropy(reference_density, new_density))
        arguments.append(i)
    divergences = np.array(divergences)
    last_argmin = len(divergences) - 1 - np.argmin(divergences[::-1])
    return int(last_argmin * stride + start_bin), divergences, arguments
