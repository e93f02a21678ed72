"""CPU ORACLE (numpy) for the SESR / NRDM int8 hardware-sim hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain-numpy restatement of the reference algorithm (gui-yupeng/sesr-pytorch-quantize).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
it, and only as the checker.  The product path (sesr_pytorch_quantize_b200) never imports it.

Parity pin: oracle/gen_golden.py runs the UNMODIFIED reference on CPU in the build container and
checks this restatement bit-for-bit against every tensor the reference dumps (input.{0..5}.pt,
pe_output{i}_{p}.pt, pe_add_output{i}.pt, shortcut_tensor.pt, conv.weight.{i}.pt,
conv.bias.quan{i}.pt, requan factors, final float output) for SESR x4 / x2 / NRDM_3 on the
shipped fixtures plus synthetic clamp- and zero-point-stress cases; the resulting vectors are
committed under tests/golden/ (see tests/golden/ORACLE_PIN.json).  The KL calibrator has no
reference implementation (define.py:16-20 "KL discarded") -> that part is "parity unpinned".

All `file:line` citations are into the reference tree (myQL/quan_func.py unless stated).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

F32 = np.float32
QMIN, QMAX = -128, 127

# define.py:6-14
PE = 4
QUAN_BIT = 8
BIAS_BIT = 16
PE_ACC_BIT = 18
PE_ADD_BIT = 20
REQUAN_BIT = 16
REQUAN_N_MAX = 32


# --------------------------------------------------------------------------------------
# host-side constants
# --------------------------------------------------------------------------------------
def quan_layer_between_const(x: float, data_bit: int = REQUAN_BIT, shift_max: int = REQUAN_N_MAX):
    """(M16, n) with x ~= M16 * 2**-n.  quan_func.py:495-515 (Python double arithmetic)."""
    assert data_bit < shift_max
    x = float(x)
    if int(x) != 0:
        before_point_bit = math.ceil(math.log2(int(x) + 1))
        n = data_bit - before_point_bit
        return int(x * (2 ** n)), n
    data = x * 2
    times = 0
    while int(data) == 0:
        times += 1
        data = data * 2
    n = min(times + data_bit, shift_max)
    return int(x * (2 ** n)), n


def quantize_weight(wf: np.ndarray, width: int = QUAN_BIT):
    """Per-tensor symmetric weight quantisation, quan_func.py:58-71.
    Returns (int8 weights, scale as Python double)."""
    wf = np.asarray(wf, dtype=F32)
    max_val = float(wf.max())
    min_val = float(wf.min())
    absmax = max(abs(max_val), abs(min_val))
    assert absmax > 0, "Conv2d weight tensor is all zero"          # :63
    qmax = 2 ** (width - 1) - 1
    qmin = -(2 ** (width - 1))
    scale = (absmax - (0 - absmax)) / (qmax - qmin)                   # :70 (double)
    q = np.clip(np.rint(wf / F32(scale)), qmin, qmax)                 # :71 fp32 true division
    return q.astype(np.int8), scale


def quantize_bias(bias_f: np.ndarray, bias_scale: float, width: int = BIAS_BIT):
    """quan_func.py:391-416 (exe_mode 1): clamp(round(b / (s_x*s_w)), 16 bit)."""
    b = np.asarray(bias_f, dtype=F32)
    assert float(b.max()) != float(b.min()), "bias elements are all equal"   # :397
    qmax = 2 ** (width - 1) - 1
    qmin = -(2 ** (width - 1))
    return np.clip(np.rint(b / F32(bias_scale)), qmin, qmax).astype(np.int64)


@dataclass
class OracleLayer:
    k: int
    cin: int
    cout: int
    W: np.ndarray            # int8 [cout, cin, k, k]
    sw: float                # weight scale (double)
    bq: np.ndarray           # int64 [cout] quantised bias
    K: np.ndarray            # int64 [cout] clamp16(bq - sum(W)*z)          (:481-486)
    z: int                   # input zero point (unclamped)
    M: int                   # requant multiplier  (:527/557/582/603)
    n: int
    relu: bool


@dataclass
class OracleNet:
    layers: List[OracleLayer]
    scales: List[float]      # s_0..s_5 (Python doubles)
    zeros: List[int]         # z_0..z_5
    Mr: int                  # residual-add requant (:259-260)
    nr: int
    r: int                   # PixelShuffle factor (1 = none)
    pe_acc_bit: int = PE_ACC_BIT
    pe_add_bit: int = PE_ADD_BIT


def derive_net(weights_f: Sequence[np.ndarray], biases_f: Sequence[np.ndarray],
               scales: Sequence[float], zeros: Sequence[int], r: int,
               pe_acc_bit: int = PE_ACC_BIT, pe_add_bit: int = PE_ADD_BIT,
               codes: Optional[Sequence[np.ndarray]] = None, w_scales: Optional[Sequence[float]] = None) -> OracleNet:
    """All host-side constants of SURVEY App. A from collapsed float weights + calibration.  With `codes` / `w_scales`
    the weights are taken as already quantised (what a graph holds after quantize_model_weight); weights_f is unused."""
    assert len(weights_f if codes is None else codes) == 5 and len(scales) == 6 and len(zeros) == 6
    nxt = [1, 2, 3, 1, 5]                      # requant target domain per layer (:525,:601,:556,:579)
    layers = []
    for i in range(5):
        if codes is not None:
            W, sw = np.asarray(codes[i]).astype(np.int8), float(w_scales[i])
        else:
            W, sw = quantize_weight(weights_f[i])
        cout, cin, k, _ = W.shape
        s_i = float(scales[i])
        z_i = int(zeros[i])
        bq = quantize_bias(biases_f[i], s_i * sw)                     # :448-449
        wsum = W.astype(np.int64).sum(axis=(1, 2, 3))
        K = np.clip(bq - wsum * z_i, -(2 ** (BIAS_BIT - 1)), 2 ** (BIAS_BIT - 1) - 1)
        x = s_i / float(scales[nxt[i]]) * sw                          # :527 evaluation order
        M, n = quan_layer_between_const(x)
        layers.append(OracleLayer(k=k, cin=cin, cout=cout, W=W, sw=sw, bq=bq, K=K, z=z_i,
                                  M=M, n=n, relu=(i < 4)))
    Mr, nr = quan_layer_between_const(float(scales[1]) / float(scales[4]))   # :259-260
    return OracleNet(layers=layers, scales=[float(s) for s in scales], zeros=[int(z) for z in zeros],
                     Mr=Mr, nr=nr, r=r, pe_acc_bit=pe_acc_bit, pe_add_bit=pe_add_bit)


# --------------------------------------------------------------------------------------
# integer forward
# --------------------------------------------------------------------------------------
def _clampq(a):
    return np.clip(a, QMIN, QMAX)


def _requant(acc_i64: np.ndarray, M: int, n: int) -> np.ndarray:
    """y = RN_f32(f32(A) * f32(M)) * 2**-n   (:529; the product is NOT exact in fp32)."""
    a = acc_i64.astype(F32)                     # |A| < 2**24 -> exact
    return (a * F32(M)) * F32(2.0 ** (-n))


def conv_pe(q: np.ndarray, layer: OracleLayer, pe_acc_bit: int, pe_add_bit: int, want_pe=False):
    """q: int64 [B, cin, H, W] raw codes.  Returns A = clamp20(sum_p clamp18(P_p)) + K  (:338-370,
    :437, :491) where P_p is the convolution restricted to input channels c == p (mod 4), with the
    frame padded by z' = max(z, -128)  (equivalently: u = q - z' zero padded, plus sum(W_p)*z')."""
    B, C, H, Wd = q.shape
    k = layer.k
    pad = (k - 1) // 2
    zp = max(layer.z, -128)                                           # :351-352
    qp = np.full((B, C, H + 2 * pad, Wd + 2 * pad), zp, dtype=np.float64)
    qp[:, :, pad:pad + H, pad:pad + Wd] = q
    Wt = layer.W.astype(np.float64)
    lo18, hi18 = -(2 ** (pe_acc_bit - 1)), 2 ** (pe_acc_bit - 1) - 1
    lo20, hi20 = -(2 ** (pe_add_bit - 1)), 2 ** (pe_add_bit - 1) - 1
    total = np.zeros((B, layer.cout, H, Wd), dtype=np.int64)
    pes = []
    for p in range(PE):
        ch = list(range(p, C, PE))                                    # :315
        acc = np.zeros((B, layer.cout, H, Wd), dtype=np.float64)      # exact: |sum| << 2**53
        if ch:
            for dy in range(k):
                for dx in range(k):
                    patch = qp[:, ch, dy:dy + H, dx:dx + Wd]          # [B, len(ch), H, W]
                    acc += np.einsum('oc,bchw->bohw', Wt[:, ch, dy, dx], patch, optimize=True)
        P = np.clip(acc.astype(np.int64), lo18, hi18)                 # :370
        if want_pe:
            pes.append(P)
        total += P
    pe_add = np.clip(total, lo20, hi20)                               # :437
    A = pe_add + layer.K[None, :, None, None]                         # :491
    if want_pe:
        return A, pe_add, pes
    return A, pe_add


def pixel_shuffle(x: np.ndarray, r: int) -> np.ndarray:
    if r == 1:
        return x
    B, C, H, W = x.shape
    c = C // (r * r)
    x = x.reshape(B, c, r, r, H, W).transpose(0, 1, 4, 2, 5, 3)
    return np.ascontiguousarray(x.reshape(B, c, H * r, W * r))


def forward_int(net: OracleNet, x: np.ndarray, trace: Optional[dict] = None):
    """Integer-sim forward (sim.py path).  x: float32 [B, Cin, H, W].
    Returns (q5 int8 [B, Cout, H, W], y float32 pixel-shuffled [B, Cout/r^2, rH, rW]).
    If `trace` is a dict it receives q0..q5, pe_add0..4, pe{i}_{p}, shortcut."""
    x = np.asarray(x, dtype=F32)
    L = net.layers
    s, z = net.scales, net.zeros
    # layer-0 input quantisation, :222-225 (true fp32 division, then fp32 add)
    q = _clampq(np.rint(x / F32(s[0]) + F32(z[0]))).astype(np.int64)
    shortcut = None
    for i in range(5):
        if trace is not None:
            trace['q%d' % i] = q.astype(np.int8)
            A, pe_add, pes = conv_pe(q, L[i], net.pe_acc_bit, net.pe_add_bit, want_pe=True)
            trace['pe_add%d' % i] = pe_add.astype(np.int32)
            for p, P in enumerate(pes):
                trace['pe%d_%d' % (i, p)] = P.astype(np.int32)
        else:
            A, _ = conv_pe(q, L[i], net.pe_acc_bit, net.pe_add_bit)
        y = _requant(A, L[i].M, L[i].n)                               # :529/:605/:560/:584
        if i == 0:
            shortcut = np.maximum(y, F32(0))                          # :530,:549 float shortcut
            if trace is not None:
                trace['shortcut'] = shortcut
        if i < 3:
            a = np.maximum(y, F32(0))                                 # ReLU
            q = _clampq(np.rint(a + F32(z[i + 1]))).astype(np.int64)  # :280
        elif i == 3:
            a = np.maximum(y, F32(0))
            rc = _clampq(np.rint(shortcut - F32(128)))                # :250
            ic = _clampq(np.rint(a - F32(128)))                       # :251
            t = (rc + ic + F32(256))                                  # :252 (exact small ints)
            t = (t * F32(net.Mr)) * F32(2.0 ** (-net.nr))             # :269
            q = _clampq(np.rint(t + F32(z[4]))).astype(np.int64)      # :270
        else:
            q5 = _clampq(np.rint(y + F32(z[5])))                      # :585
            out = (q5 - F32(z[5])) * F32(s[5])                        # :594
            if trace is not None:
                trace['q5'] = q5.astype(np.int8)
            return q5.astype(np.int8), pixel_shuffle(out.astype(F32), net.r)
    raise AssertionError


# --------------------------------------------------------------------------------------
# static clamp analysis: is the 18-bit per-PE / 20-bit PE-sum clamp provably inert?
# --------------------------------------------------------------------------------------
def pe_clamp_bounds(W: np.ndarray, pe_acc_bit: int = PE_ACC_BIT, pe_add_bit: int = PE_ADD_BIT):
    """Exact range of P_p and sum_p P_p over all codes q in [-128,127] (pad value z' is in range).
    Returns (inert18: bool, inert20: bool, worst |P_p|, worst |sum|)."""
    Wl = W.astype(np.int64)
    cout, cin = Wl.shape[:2]
    hi_t = np.maximum(Wl * QMIN, Wl * QMAX)
    lo_t = np.minimum(Wl * QMIN, Wl * QMAX)
    worst18 = 0
    ok18 = True
    for p in range(PE):
        ch = list(range(p, cin, PE))
        if not ch:
            continue
        hi = hi_t[:, ch].sum(axis=(1, 2, 3))
        lo = lo_t[:, ch].sum(axis=(1, 2, 3))
        worst18 = max(worst18, int(hi.max()), int(-lo.min()))
        ok18 &= bool(hi.max() <= 2 ** (pe_acc_bit - 1) - 1 and lo.min() >= -(2 ** (pe_acc_bit - 1)))
    hi = hi_t.sum(axis=(1, 2, 3))
    lo = lo_t.sum(axis=(1, 2, 3))
    ok20 = bool(hi.max() <= 2 ** (pe_add_bit - 1) - 1 and lo.min() >= -(2 ** (pe_add_bit - 1)))
    return ok18, ok20, worst18, max(int(hi.max()), int(-lo.min()))


# --------------------------------------------------------------------------------------
# min/max calibration tail (test.py:186-217)
# --------------------------------------------------------------------------------------
def minmax_to_thresholds(mins: Sequence[float], maxs: Sequence[float], width: int = QUAN_BIT):
    """Running min/max (Python floats) -> (scales, zeros); id 5 forces min := 0 (test.py:207)."""
    qmax = 2 ** (width - 1) - 1
    qmin = -(2 ** (width - 1))
    scales, zeros = [], []
    for i, (mn, mx) in enumerate(zip(mins, maxs)):
        mn = float(mn)
        mx = float(mx)
        if i == 5:
            mn = 0
        sc = (mx - mn) / (qmax - qmin)
        zr = qmin - round(mn / sc)              # Python round(): half-to-even on the double
        scales.append(sc)
        zeros.append(int(zr))
    return scales, zeros
