"""CPU ORACLE (numpy) for the activation histogram and the KL-entropy threshold search.
TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import it; the product path never does.

PARITY UNPINNED: the reference abandoned KL calibration -- define.py:16-20 keeps only the commented
constants USE_KL_CALI, BINS_NUM = 2048, TGT_BINS_NUM = 128 ("KL discarded"), and no histogram or entropy
code exists anywhere in the tree.  There is therefore nothing in the reference to pin this file against;
the specification is SURVEY.md App. C (a TensorRT-style entropy calibrator sized by the two surviving
constants) and this file is its executable form.  The CUDA kernels (csrc/sesrq_calib.cu) must reproduce
the histogram counts exactly and the chosen bin count index-exactly.

  histogram   bin = min(B-1, max(0, floor((clamp(x, lo, hi) - lo) * (B / (hi - lo)))))   all in fp32
  search      for i in [tgt, B]:   P = hist[0:i], outlier mass sum(hist[i:]) added to P[i-1]
                                   Q = hist[0:i] merged into tgt groups of m = i // tgt bins (the last group
                                       takes the remainder), each group's mass spread uniformly over its
                                       non-empty bins
                                   KL_i = sum_k P_k/total * log((P_k/total) / (Q_k/inside))    in fp64
              a bin with P_k > 0 but hist[k] == 0 (only possible for k = i-1) makes KL_i infinite;
              answer = smallest i attaining the minimum; all-infinite -> B
  threshold   T = lo + (i + 0.5) * (hi - lo) / B
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


def kl_curve(hist: np.ndarray, bins: int = BINS, tgt: int = TARGET_BINS) -> np.ndarray:
    """float64 [bins + 1]; entries below tgt are +inf (never candidates)."""
    h = hist.astype(np.int64)
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
        m = i // tgt
        k = np.arange(i)
        hk = h[:i]
        pk = hk.copy()
        pk[i - 1] += outliers
        live = pk > 0
        if np.any(live & (hk == 0)):
            continue
        g = np.minimum(k // m, tgt - 1)
        start = g * m
        stop = np.where(g == tgt - 1, i, start + m)
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
    return int(np.argmin(curve))          # numpy argmin returns the first (smallest i) minimum


def threshold_value(idx: int, lo: float, hi: float, bins: int = BINS) -> float:
    return lo + (idx + 0.5) * (hi - lo) / bins
