"""Activation calibration on device: reference-exact running min/max, and the KL-entropy calibrator the
reference only sketched (define.py:16-20: BINS_NUM = 2048, TGT_BINS_NUM = 128).

Data-parallel: every rank observes its shard of the calibration frames; the statistics are merged with
one collective each -- MAX over [max_i, -min_i] (12 floats) for the min/max calibrator, SUM over the
int64 [6, 2048] histograms for KL -- over NCCL on GPUs (gloo in the CPU tests of the merge logic).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .compile import QUAN_BIT, thresholds_from_minmax
from .fq import FakeQuantNet


# ---- merge logic (device agnostic: exercised with gloo on CPU) -------------------------------------------
def merge_minmax(minmax: torch.Tensor, group=None) -> torch.Tensor:
    """minmax: float32 [6, 2] (min, max) per domain -> all-rank running min/max.  One all_reduce(MAX)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return minmax
    packed = torch.stack([-minmax[:, 0], minmax[:, 1]], dim=1).contiguous()
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
    return torch.stack([-packed[:, 0], packed[:, 1]], dim=1)


def merge_histograms(hist: torch.Tensor, group=None) -> torch.Tensor:
    """hist: int64 [6, bins] -> summed over ranks.  One all_reduce(SUM)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return hist
    h = hist.contiguous()
    dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
    return h


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_items units (frames) over `world` ranks; the first n % world ranks get one more."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def kl_threshold_from_index(idx: int, lo: float, hi: float, bins: int = _lib.KL_BINS) -> float:
    """Upper clipping threshold for the chosen bin count: the upper edge of the last kept bin (`calib_bin_edges[i]` of the
    published calibrator, pytorch-quantization calib/histogram.py)."""
    return lo + idx * (hi - lo) / bins


def thresholds_from_kl(mins: Sequence[float], maxs: Sequence[float], idx: Sequence[int], width: int = QUAN_BIT):
    """scale/zero per domain from the KL-chosen upper thresholds; lower bounds follow the reference rule
    (observed min, 0 for the output domain, test.py:207)."""
    lo = [float(m) for m in mins]
    lo[-1] = 0.0
    t = [kl_threshold_from_index(int(i), l, float(h)) for i, l, h in zip(idx, lo, maxs)]
    return thresholds_from_minmax(lo, t, width, force_zero_min_last=True)


# ---- calibrators -------------------------------------------------------------------------------------
class MinMaxCalibrator:
    """Reference-exact: running min/max per domain over the fake-quant forward of every calibration frame
    (quan_func.py:175-196), then scale = (max - min) / 255, zero = -128 - round(min / scale)
    (test.py:186-217)."""

    def __init__(self, net: FakeQuantNet, group=None):
        self.net = net
        self.group = group

    def observe(self, frames: torch.Tensor):
        self.net.forward(frames)

    def finish(self):
        mm = merge_minmax(self.net.minmax, self.group).cpu()
        mins = [float(v) for v in mm[:, 0]]
        maxs = [float(v) for v in mm[:, 1]]
        for i, (a, b) in enumerate(zip(mins, maxs)):
            assert a != b, "Input tensor is all equal,{}".format(i)      # quan_func.py:196
        scales, zeros = thresholds_from_minmax(mins, maxs)
        return scales, zeros, mins, maxs


class KLCalibrator:
    """Two passes: (1) min/max as above fixes the histogram range [lo_i, max_i] of every domain
    (lo_5 = 0); (2) 2048-bin histograms accumulated on device by warp-aggregated shared-memory atomics,
    merged over ranks, and the TensorRT-style threshold search run on device (sesrq_kl_threshold)."""

    def __init__(self, net: FakeQuantNet, group=None, bins: int = _lib.KL_BINS, target_bins: int = _lib.KL_TARGET_BINS):
        if bins != _lib.KL_BINS:
            raise _lib.SesrqError("libsesrq histograms have %d bins" % _lib.KL_BINS)
        self.net, self.group, self.bins, self.target = net, group, bins, target_bins
        self.mins: Optional[List[float]] = None
        self.maxs: Optional[List[float]] = None

    def observe_pass1(self, frames: torch.Tensor):
        self.net.forward(frames)

    def begin_pass2(self):
        mm = merge_minmax(self.net.minmax, self.group).cpu()
        self.mins = [float(v) for v in mm[:, 0]]
        self.maxs = [float(v) for v in mm[:, 1]]
        lows = list(self.mins)
        lows[-1] = 0.0
        self.lows = lows
        self.net.reset_observers()
        self.net.enable_histograms([[l, h] for l, h in zip(lows, self.maxs)])

    def observe_pass2(self, frames: torch.Tensor):
        self.net.forward(frames)

    def search(self, hist: torch.Tensor) -> Tuple[List[int], torch.Tensor]:
        """hist: int64 [6, bins] on the device -> chosen bin counts and the KL curves (float64 [6, bins+1])."""
        lib = self.net.lib
        dev = hist.device
        idx = torch.zeros(hist.shape[0], dtype=torch.int32, device=dev)
        kl = torch.full((hist.shape[0], self.bins + 1), float("inf"), dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            for d in range(hist.shape[0]):
                rc = lib.sesrq_kl_threshold(C.c_void_p(hist[d].data_ptr()), self.bins, self.target,
                                            C.c_void_p(idx[d:d + 1].data_ptr()), C.c_void_p(kl[d].data_ptr()), st)
                _lib.check(rc, "sesrq_kl_threshold")
        return [int(v) for v in idx.cpu()], kl

    def finish(self):
        hist = merge_histograms(self.net.hist, self.group)
        idx, _ = self.search(hist)
        self.net.disable_histograms()
        scales, zeros = thresholds_from_kl(self.mins, self.maxs, idx)
        return scales, zeros, idx
