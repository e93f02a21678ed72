"""CPU ORACLE (numpy) for the fake-quant ("staged", exe_mode 0) path of test.py and for the running
min/max calibrator.  TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it; the product path never does.

What the reference does per frame (exe_mode 0):
  quantize_asymmetrical_by_tensor   myQL/quan_func.py:175-215   per-FRAME min/max -> scale, zero;
                                    x_fq = (clamp(round(x / s + z)) - z) * s; running min/max files
  nn.Conv2d on x_fq with W_fq = codes * s_w (quan_func.py:72-75) on the PE-split tensor (fp32, library
                                    summation order)
  reshape_ouput_for_hardware_pe     :330-333, :386   float-domain 18-bit clamp per PE, sum of 4 PEs
  PEs_and_bias_adder                :431-434, :448-459   float-domain 20-bit clamp, + fake-quant bias
                                    (:402,:412); MFLAG 3 records the output range as domain 5 (:460-479)
  activation / AddOp                models/sesr.py:45 (residual add before conv_last), ReLU on layers 0-3
  the PixelShuffle input is quantised as domain 5 (test.py:90-91)
  calibration tail                  test.py:186-217 (restated in sesr_oracle.minmax_to_thresholds)

Parity pin: this restatement carries the convolution out EXACTLY on the integer operands (q - z) and W
and scales once by f32(f32(s) * f32(s_w)); the reference's conv is fp32 in oneDNN's summation order, so
the two agree only to rounding noise.  tests/test_fq_kl_oracle.py pins it against the reference's own
mode-0 output (tests/golden/*_small.npz: fq_out, fq_mins, fq_maxs, produced by oracle/gen_golden.py
from the UNMODIFIED reference) at PSNR >= 60 dB (range 1.0) and max-abs <= 8 * s_5, and the per-domain
min/max at rel 1e-5.  The CUDA path follows this restatement op for op and is compared to it tightly.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from sesr_oracle import F32, quantize_weight, pixel_shuffle

QMIN, QMAX = -128, 127


def frame_quant(mn: float, mx: float):
    """(scale, zero) of one frame's tensor, Python doubles (quan_func.py:198-205)."""
    assert mx != mn, "Input tensor is all equal"
    s = (float(mx) - float(mn)) / (QMAX - QMIN)
    z = QMIN - round(float(mn) / s)
    return s, z


def fq_tensor_codes(x: np.ndarray, s: float, z: int) -> np.ndarray:
    """q - z as int64 (quan_func.py:207, :215 before the multiply by s)."""
    t = (x.astype(F32) / F32(s)).astype(F32) + F32(z)
    q = np.clip(np.rint(t.astype(F32)), QMIN, QMAX)
    return q.astype(np.int64) - int(z)


def conv_pe_int(d: np.ndarray, W: np.ndarray) -> np.ndarray:
    """Exact integer conv per PE: d [C,H,W] int64 (zero padded), W [Co,C,k,k] -> [4, Co, H, W] int64 where
    PE p takes channels c == p (mod 4) (reshape_input_for_hardware_pe, quan_func.py:312-316)."""
    Co, C, k, _ = W.shape
    pad = (k - 1) // 2
    _, H, Wd = d.shape
    dp = np.zeros((C, H + 2 * pad, Wd + 2 * pad), dtype=np.int64)
    dp[:, pad:pad + H, pad:pad + Wd] = d
    out = np.zeros((4, Co, H, Wd), dtype=np.int64)
    Wl = W.astype(np.int64)
    for c in range(C):
        for dy in range(k):
            for dx in range(k):
                out[c % 4] += Wl[:, c, dy, dx][:, None, None] * dp[c, dy:dy + H, dx:dx + Wd][None]
    return out


def forward_fq(weights_f: Sequence[np.ndarray], biases_f: Sequence[np.ndarray], x: np.ndarray, r: int,
               residual: bool = True, observe_out: bool = False, pe_acc_bit: int = 18, pe_add_bit: int = 20,
               codes: Optional[Sequence[np.ndarray]] = None, w_scales: Optional[Sequence[float]] = None,
               trace: Optional[list] = None):
    """One frame x [1,Cin,H,W] float32 through the mode-0 graph.  Returns (y, mins[6], maxs[6]) where the
    min/max are this frame's per-domain ranges (what the running observer is updated with).  `trace` (a list)
    receives the six domain tensors X_0..X_5 (what each quantiser sees)."""
    assert x.shape[0] == 1
    if codes is None:
        pairs = [quantize_weight(w) for w in weights_f]
        codes = [p[0] for p in pairs]
        w_scales = [p[1] for p in pairs]
    t = x[0].astype(F32)
    mins: List[float] = []
    maxs: List[float] = []
    first = None
    for i in range(5):
        if trace is not None:
            trace.append(t.copy())
        mn, mx = float(t.min()), float(t.max())
        mins.append(mn)
        maxs.append(mx)
        s, z = frame_quant(mn, mx)
        d = fq_tensor_codes(t, s, z)
        pe = conv_pe_int(d, np.asarray(codes[i]))
        sw = float(w_scales[i])
        cscale = F32(float(F32(s)) * float(F32(sw)))
        hi_a, lo_a = float(2 ** (pe_acc_bit - 1) - 1), -float(2 ** (pe_acc_bit - 1))
        hi_s, lo_s = float(2 ** (pe_add_bit - 1) - 1), -float(2 ** (pe_add_bit - 1))
        fmax_a, fmin_a = F32((hi_a - z) * s * sw), F32((lo_a - z) * s * sw)          # :331-332
        fmax_s, fmin_s = F32((hi_s - z) * s * sw), F32((lo_s - z) * s * sw)          # :432-433
        acc = np.zeros(pe.shape[1:], dtype=F32)
        for p in range(4):
            yp = (pe[p].astype(F32) * cscale).astype(F32)
            acc = (acc + np.clip(yp, fmin_a, fmax_a)).astype(F32)                      # :333, :386
        acc = np.clip(acc, fmin_s, fmax_s)                                             # :434
        bs = F32(s * sw)                                                               # :448
        bq = np.clip(np.rint((np.asarray(biases_f[i], dtype=F32) / bs).astype(F32)), -32768, 32767).astype(F32)  # :402
        bias_fq = (bq * bs).astype(F32)                                                # :412
        acc = (acc + bias_fq[:, None, None]).astype(F32)                               # :459
        if i < 4:
            acc = np.maximum(acc, F32(0))
        if i == 0:
            first = acc
        if i == 3 and residual:
            acc = (acc + first).astype(F32)                                            # AddOp
        t = acc
    if trace is not None:
        trace.append(t.copy())
    mn, mx = float(t.min()), float(t.max())
    mins.append(mn)
    maxs.append(mx)
    if not observe_out:
        s, z = frame_quant(mn, mx)
        q = np.clip(np.rint(((t / F32(s)).astype(F32) + F32(z)).astype(F32)), QMIN, QMAX).astype(F32)
        t = ((q - F32(z)).astype(F32) * F32(s)).astype(F32)                            # :215
    y = pixel_shuffle(t[None], r) if r > 1 else t[None]
    return y.astype(F32), mins, maxs


def running_minmax(frames_minmax):
    """Fold per-frame (mins, maxs) the way the observer files do (quan_func.py:184-195)."""
    mins = [min(f[0][i] for f in frames_minmax) for i in range(6)]
    maxs = [max(f[1][i] for f in frames_minmax) for i in range(6)]
    return mins, maxs
