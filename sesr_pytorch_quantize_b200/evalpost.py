"""Evaluation post-step of the reference's test.py:128-170 on device: for the x2 model the
nearest-upsampled input is added OUTSIDE the network (:149-155), the result is clipped to [0, 1] (:157) and compared with
the ground truth by PSNR -- plain (data_range 1), channel 0 * 255 for the x4 model (:165-166) or BT.601 luma for the x2
model (rgb_to_yuv :133-136, :167-168).  `ssim` is the structural similarity the script takes from skimage (test.py:172-175),
restated as a CUDA kernel with skimage's defaults (7 x 7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _lib

METRIC_PLAIN, METRIC_CH0_255, METRIC_LUMA = 0, 1, 2
METRIC_OF_MFLAG = {1: METRIC_PLAIN, 2: METRIC_PLAIN, 3: METRIC_PLAIN, 4: METRIC_PLAIN, 5: METRIC_CH0_255, 6: METRIC_LUMA}


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def sr_post(y: torch.Tensor, x: Optional[torch.Tensor] = None, gt: Optional[torch.Tensor] = None,
            metric: int = METRIC_PLAIN, eps: float = 1e-8) -> Tuple[torch.Tensor, Optional[float]]:
    """y: network output float32 CUDA [B,C,Hr,Wr]; x: network input [B,C,H,W] to be nearest-upsampled and added (x2 model) or
    None; gt: ground truth [B,C,Hr,Wr] or None.  Returns (clip(y + up(x), 0, 1), PSNR in dB over the whole batch or None)."""
    if not y.is_cuda or y.dtype != torch.float32 or y.dim() != 4:
        raise _lib.SesrqError("sr_post needs a float32 CUDA NCHW tensor (no CPU fallback)")
    lib = _lib.load_library()
    y = y.contiguous()
    B, Cc, Hr, Wr = y.shape
    r = 1
    if x is not None:
        x = x.contiguous().float()
        r = Hr // x.shape[2]
        if x.shape[0] != B or x.shape[1] != Cc or x.shape[2] * r != Hr or x.shape[3] * r != Wr:
            raise _lib.SesrqError("input %s is not a 1/r-size version of the output %s" % (tuple(x.shape), tuple(y.shape)))
    out = torch.empty_like(y)
    sq = torch.zeros(1, dtype=torch.float64, device=y.device) if gt is not None else None
    if gt is not None:
        gt = gt.contiguous().float()
        if gt.shape != y.shape:
            raise _lib.SesrqError("ground truth shape %s != output shape %s" % (tuple(gt.shape), tuple(y.shape)))
    with torch.cuda.device(y.device):
        rc = lib.sesrq_op_sr_post(_p(y), _p(x), _p(gt), B, Cc, Hr, Wr, r, metric, _p(out), _p(sq),
                                  C.c_void_p(torch.cuda.current_stream(y.device).cuda_stream))
    _lib.check(rc, "sesrq_op_sr_post")
    psnr = None
    if gt is not None:
        mse = float(sq.item()) / (B * Hr * Wr * (Cc if metric == METRIC_PLAIN else 1))
        if metric == METRIC_PLAIN:                       # skimage compare_psnr(data_range=1.0), test.py:170
            psnr = float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)
        else:                                            # compute_psnr(data_range=255, eps), test.py:128-131
            psnr = 10.0 * math.log10((255.0 ** 2) / (mse + eps))
    return out, psnr


def ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 1.0) -> float:
    """Mean structural similarity of two float32 CUDA images [..., H, W] (any leading plane dimensions: channels and / or
    batch): skimage.metrics.structural_similarity(a, b, data_range=..., channel_axis=<planes>) with its default window
    (test.py:172-175 passes HWC numpy arrays with channel_axis=2, or a single plane with multichannel=False)."""
    if not a.is_cuda or not b.is_cuda:
        raise _lib.SesrqError("ssim needs CUDA tensors (no CPU fallback)")
    if a.shape != b.shape or a.dim() < 2:
        raise _lib.SesrqError("ssim: shapes differ or are not images: %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    a = a.contiguous().float()
    b = b.contiguous().float()
    H, W = a.shape[-2], a.shape[-1]
    planes = a.numel() // (H * W)
    if H < 7 or W < 7:
        raise _lib.SesrqError("win_size exceeds image extent")          # skimage's own message
    total = torch.zeros(1, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load_library().sesrq_op_ssim(_p(a), _p(b), planes, H, W, float(data_range), _p(total),
                                                C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream))
    _lib.check(rc, "sesrq_op_ssim")
    return float(total.item()) / (planes * (H - 6) * (W - 6))
