"""FakeQuantNet: the exe_mode 0 ("staged" float) network of test.py:79-106,146 as one libsesrq handle.

Per frame, every activation domain is quantised with THAT frame's min/max (quan_func.py:177-207); the
running min/max over all frames is what the calibration tail (test.py:186-217) turns into thresholds.
A batch is ONE C call (sesrq_fq_forward_batch: frames are a batch dimension of every kernel, each frame keeps
its own ranges).  The observer state (running min/max, per-frame min/max of the last call, optional 2048-bin
histograms for the KL calibrator) stays on the device; nothing is synchronised unless the caller asks.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from .compile import PE_ACC_BIT, PE_ADD_BIT, QUAN_BIT, symmetric_weight_codes


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class FakeQuantNet:
    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], r_shuffle: int,
                 residual: bool = True, observe_out: bool = False, device=None,
                 pe_acc_bits: int = PE_ACC_BIT, pe_add_bits: int = PE_ADD_BIT,
                 codes: Optional[Sequence[torch.Tensor]] = None, w_scales: Optional[Sequence[float]] = None):
        """weights: collapsed float conv weights (quantised here per tensor), or pass `codes` + `w_scales`.
        residual: AddOp(block output, conv_first output) before conv_last (float model variants);
        observe_out: no PixelShuffle node -> only observe conv_last's output as domain 5 (MFLAG 3)."""
        self.lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise _lib.SesrqError("sesrq needs a CUDA device (B200 / sm_100); no CPU fallback exists")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if codes is None:
            pairs = [symmetric_weight_codes(w, QUAN_BIT) for w in weights]
            codes = [p[0] for p in pairs]
            w_scales = [p[1] for p in pairs]
        self.codes = [c.to("cpu", torch.int8).contiguous() for c in codes]
        self.w_scales = [float(s) for s in w_scales]
        self.biases = [torch.as_tensor(b, dtype=torch.float32).detach().cpu().contiguous() for b in biases]
        self.r = int(r_shuffle)
        self.cin = self.codes[0].shape[1]
        self.cout = self.codes[-1].shape[0]
        d = _lib.FqDesc()
        d.abi_version, d.n_layers, d.r_shuffle = _lib.ABI_VERSION, 5, self.r
        d.residual, d.observe_out = int(residual), int(observe_out)
        d.pe_acc_bits, d.pe_add_bits = pe_acc_bits, pe_add_bits
        for i in range(5):
            co, ci, k, _ = self.codes[i].shape
            d.ksize[i], d.cin[i], d.cout[i], d.relu[i] = k, ci, co, int(i < 4)
            d.h_weight[i] = self.codes[i].data_ptr()
            d.w_scale[i] = self.w_scales[i]
            d.h_bias[i] = self.biases[i].data_ptr()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sesrq_fq_create(C.byref(d), C.byref(h)), "sesrq_fq_create")
        self._h = h
        self._ws = None
        self.reset_observers()
        self.hist: Optional[torch.Tensor] = None           # uint64 [6, 2048] as int64 storage
        self.hist_range: Optional[torch.Tensor] = None     # float32 [6, 2]

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sesrq_fq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- observers ----------------------------------------------------------------------------------
    def reset_observers(self):
        mm = torch.empty((_lib.NUM_DOMAINS, 2), dtype=torch.float32)
        mm[:, 0] = float("inf")
        mm[:, 1] = float("-inf")
        self.minmax = mm.to(self.device)                    # running (min, max) per domain, on device
        self.frame_minmax = torch.zeros((1, _lib.NUM_DOMAINS, 2), dtype=torch.float32, device=self.device)

    def enable_histograms(self, ranges: Sequence[Sequence[float]]):
        """Second calibration pass: accumulate 2048-bin histograms of every domain over [lo, hi]."""
        self.hist_range = torch.tensor(ranges, dtype=torch.float32).reshape(_lib.NUM_DOMAINS, 2).to(self.device)
        self.hist = torch.zeros((_lib.NUM_DOMAINS, _lib.KL_BINS), dtype=torch.int64, device=self.device)

    def disable_histograms(self):
        self.hist = None
        self.hist_range = None

    def running_minmax(self):
        mm = self.minmax.cpu()
        return [float(v) for v in mm[:, 0]], [float(v) for v in mm[:, 1]]

    # ---- forward ------------------------------------------------------------------------------------
    def out_shape(self, H, W):
        return (1, self.cout // (self.r * self.r), H * self.r, W * self.r)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: float32 CUDA [B, Cin, H, W]; all frames in one call, every frame with its own dynamic ranges.
        self.frame_minmax: float32 [B, 6, 2] of this call."""
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != self.cin:
            raise _lib.SesrqError("FakeQuantNet.forward needs a float32 CUDA [B,%d,H,W] tensor" % self.cin)
        x = x.contiguous()
        B, _, H, W = x.shape
        y = torch.empty((B,) + self.out_shape(H, W)[1:], dtype=torch.float32, device=x.device)
        self.frame_minmax = torch.empty((B, _lib.NUM_DOMAINS, 2), dtype=torch.float32, device=x.device)
        if B == 0:
            return y
        need = int(self.lib.sesrq_fq_ws_bytes_batch(self._h, B, H, W))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            rc = self.lib.sesrq_fq_forward_batch(self._h, _ptr(x), B, H, W, _ptr(y), _ptr(self.minmax),
                                                 _ptr(self.frame_minmax), _ptr(self.hist), _ptr(self.hist_range),
                                                 _ptr(self._ws), self._ws.numel(), st)
        _lib.check(rc, "sesrq_fq_forward_batch")
        return y

    __call__ = forward
