// Stand-alone myQL ops on float32 NCHW device tensors (exe_mode 1: tensors hold integer codes; exe_mode 0: float
// fake-quant values), exactly as the reference's un-fused fx graph passes them from node to node; plus the fake-quant
// forward / straight-through backward of the QAT library (models/quantize_utils_cuda.py).  Used when a graph is not the full
// 5-layer chain the fused network takes, and for per-node dumps.  One kernel per reference callable.
#include <algorithm>
#include "sesrq_internal.h"

namespace sesrq {

constexpr int EW_THREADS = 256;
static inline unsigned ew_grid(size_t n) { return (unsigned)std::min<size_t>((n + EW_THREADS - 1) / EW_THREADS, 148u * 16u); }
#define EW_LOOP(i, n) \
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

// quantize_asymmetrical_by_tensor, func_id 0 (quan_func.py:222-225, :290-293)
__global__ void op_quant_first_kernel(const float* __restrict__ x, size_t n, float s, float z, float zret,
                                      float* __restrict__ q, float* __restrict__ u) {
  EW_LOOP(i, n) {
    float qv = (float)quant_input(x[i], s, z);
    if (q) q[i] = qv;
    if (u) u[i] = __fsub_rn(qv, zret);
  }
}

// quantize_asymmetrical_by_tensor, func_id 1..3 (:275-280)
__global__ void op_quant_next_kernel(const float* __restrict__ x, size_t n, float z, float zret,
                                     float* __restrict__ q, float* __restrict__ u) {
  EW_LOOP(i, n) {
    float qv = (float)quant_add_zero(x[i], z);
    if (q) q[i] = qv;
    if (u) u[i] = __fsub_rn(qv, zret);
  }
}

// quantize_asymmetrical_by_tensor, func_id 4 (:249-270)
__global__ void op_quant_residual_kernel(const float* __restrict__ x, const float* __restrict__ sc, size_t n,
                                         float mf, float z, float zret, float* __restrict__ ic_out,
                                         float* __restrict__ q, float* __restrict__ u) {
  EW_LOOP(i, n) {
    int rc = quant_add_zero(sc[i], -128.f);
    int ic = quant_add_zero(x[i], -128.f);
    float t = __fmul_rn((float)(rc + ic + 256), mf);
    float qv = (float)quant_add_zero(t, z);
    if (ic_out) ic_out[i] = (float)ic;
    if (q) q[i] = qv;
    if (u) u[i] = __fsub_rn(qv, zret);
  }
}

// reshape_input_for_hardware_pe (:298-318): out[b*pe + p, c] = (c % 4 == p) ? x[b, c] : 0
__global__ void op_pe_split_kernel(const float* __restrict__ x, int C, int HW, int pe, size_t n_out,
                                   float* __restrict__ out) {
  EW_LOOP(i, n_out) {
    size_t plane = i / HW;
    int hw = (int)(i - plane * HW);
    int c = (int)(plane % C);
    size_t bp = plane / C;
    int p = (int)(bp % pe);
    size_t b = bp / pe;
    out[i] = ((c % 4) == p) ? x[(b * C + c) * (size_t)HW + hw] : 0.f;    // the reference strides by 4 (:315)
  }
}

// PEs_and_bias_adder exe_mode 1 (:418-493) incl. reshape_ouput_for_hardware_pe (:335-389)
__global__ void op_pe_merge_kernel(const float* __restrict__ conv, int Cout, int HW, size_t n_out,
                                   const float* __restrict__ pe_corr, const float* __restrict__ add_const,
                                   float lo_acc, float hi_acc, float lo_add, float hi_add,
                                   float* __restrict__ pe_out, float* __restrict__ pe_add,
                                   float* __restrict__ out, int32_t* __restrict__ overflow) {
  int ovf = 0;
  EW_LOOP(i, n_out) {
    size_t plane = i / HW;
    int hw = (int)(i - plane * HW);
    int co = (int)(plane % Cout);
    size_t b = plane / Cout;
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      size_t src = ((b * 4 + p) * Cout + co) * (size_t)HW + hw;
      float v = __fadd_rn(conv[src], pe_corr[co * 4 + p]);       // :353-356
      if (v > hi_acc) ovf |= 1;                                    // :358-361 'max_overflow'
      if (v < lo_acc) ovf |= 2;
      v = fminf(fmaxf(v, lo_acc), hi_acc);                         // :370
      if (pe_out) pe_out[src] = v;
      s = __fadd_rn(s, v);                                         // :386 (exact: small integers)
    }
    s = fminf(fmaxf(s, lo_add), hi_add);                           // :437
    if (pe_add) pe_add[i] = s;
    out[i] = __fadd_rn(s, add_const[co]);                          // :491
  }
  if (ovf && overflow) atomicOr(overflow, ovf);
}

// requan_conv2d_output exe_mode 1 (:529/:560/:605)
__global__ void op_requant_kernel(const float* __restrict__ x, size_t n, float m, float sh,
                                  float* __restrict__ y, float* __restrict__ relu_copy) {
  EW_LOOP(i, n) {
    float v = __fmul_rn(__fmul_rn(x[i], m), sh);
    y[i] = v;
    if (relu_copy) relu_copy[i] = fmaxf(v, 0.f);                   // :530 shortcut
  }
}

// requan_conv2d_output func_id 4 (:584-594)
__global__ void op_requant_final_kernel(const float* __restrict__ x, size_t n, float m, float sh, float z,
                                        float s_out, float* __restrict__ q5, float* __restrict__ y) {
  EW_LOOP(i, n) {
    float v = __fmul_rn(__fmul_rn(x[i], m), sh);
    float q = (float)quant_add_zero(v, z);
    if (q5) q5[i] = q;
    if (y) y[i] = __fmul_rn(__fsub_rn(q, z), s_out);
  }
}

// the nn.Conv2d node: stride-1 zero-padded conv of float-held integer codes, accumulated in int32
// (exact, like the reference's fp32 conv whose partial sums stay below 2^24)
__global__ void op_conv_int_kernel(const float* __restrict__ x, int Cin, int H, int W,
                                   const float* __restrict__ w, int Cout, int k, size_t n_out,
                                   float* __restrict__ out) {
  const int pad = (k - 1) / 2;
  EW_LOOP(i, n_out) {
    int xx = (int)(i % W);
    size_t t = i / W;
    int yy = (int)(t % H);
    t /= H;
    int co = (int)(t % Cout);
    size_t b = t / Cout;
    int acc = 0;
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xp = x + (b * Cin + ci) * (size_t)H * W;
      const float* wp = w + ((size_t)co * Cin + ci) * k * k;
      for (int dy = 0; dy < k; ++dy) {
        int sy = yy + dy - pad;
        if (sy < 0 || sy >= H) continue;
        for (int dx = 0; dx < k; ++dx) {
          int sx = xx + dx - pad;
          if (sx < 0 || sx >= W) continue;
          acc += __float2int_rn(xp[(size_t)sy * W + sx]) * __float2int_rn(wp[dy * k + dx]);
        }
      }
    }
    out[i] = (float)acc;
  }
}


// ---- exe_mode 0 (fake-quant, "staged" float) stand-alone ops -------------------------------------------------------
// quantize_asymmetrical_by_tensor exe_mode 0 (:207, :215): y = (clamp(rint(x / s + z)) - z) * s, all steps fp32
__global__ void op_fakequant_kernel(const float* __restrict__ x, size_t n, float s, float z, float* __restrict__ y) {
  EW_LOOP(i, n) {
    const float q = (float)quant_add_zero(__fdiv_rn(x[i], s), z);
    y[i] = __fmul_rn(__fsub_rn(q, z), s);
  }
}

// The nn.Conv2d node of the mode-0 graph: its input holds (q - z) * s (PE-split, zeros elsewhere), its weights are
// codes * s_w.  The integer operand d = q - z is recovered exactly (|d| <= 2^16, relative error of the product 2^-24),
// the accumulation is exact in int32 and the result is scaled once by f32(s) * f32(s_w) -- the same arithmetic as the
// fused fake-quant network (sesrq_fq.cu), so a per-node graph and the fused call agree bit for bit.
__global__ void op_conv_fq_kernel(const float* __restrict__ x, int Cin, int H, int W, const float* __restrict__ wcodes,
                                  int Cout, int k, size_t n_out, float s, float cscale, float* __restrict__ out) {
  const int pad = (k - 1) / 2;
  EW_LOOP(i, n_out) {
    int xx = (int)(i % W);
    size_t t = i / W;
    int yy = (int)(t % H);
    t /= H;
    int co = (int)(t % Cout);
    size_t b = t / Cout;
    int acc = 0;
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xp = x + (b * Cin + ci) * (size_t)H * W;
      const float* wp = wcodes + ((size_t)co * Cin + ci) * k * k;
      for (int dy = 0; dy < k; ++dy) {
        int sy = yy + dy - pad;
        if (sy < 0 || sy >= H) continue;
        for (int dx = 0; dx < k; ++dx) {
          int sx = xx + dx - pad;
          if (sx < 0 || sx >= W) continue;
          acc += __float2int_rn(__fdiv_rn(xp[(size_t)sy * W + sx], s)) * __float2int_rn(wp[dy * k + dx]);
        }
      }
    }
    out[i] = __fmul_rn((float)acc, cscale);
  }
}

// reshape_ouput_for_hardware_pe + PEs_and_bias_adder exe_mode 0 (:330-333, :386, :431-434, :459): float-domain clamps
__global__ void op_pe_merge_fq_kernel(const float* __restrict__ conv, int Cout, int HW, size_t n_out, float lo_acc,
                                      float hi_acc, float lo_add, float hi_add, const float* __restrict__ bias,
                                      float* __restrict__ out) {
  EW_LOOP(i, n_out) {
    size_t plane = i / HW;
    int hw = (int)(i - plane * HW);
    int co = (int)(plane % Cout);
    size_t b = plane / Cout;
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float v = conv[((b * 4 + p) * Cout + co) * (size_t)HW + hw];
      v = fminf(fmaxf(v, lo_acc), hi_acc);                         // :333
      s = __fadd_rn(s, v);                                         // :386
    }
    s = fminf(fmaxf(s, lo_add), hi_add);                           // :434
    out[i] = bias ? __fadd_rn(s, bias[co]) : s;                    // :459
  }
}

// ---- QAT fake-quant (models/quantize_utils_cuda.py:150-246): forward + straight-through backward ----------------------
// forward:  y = (clamp(round_half_away(x / s - zp), qmin, qmax) + zp) * s                       (:233-245, Round :164-165)
// backward: g_x = ((g * s) masked) / s, masked to zero where the rounded value leaves [qmin, qmax] (torch.clamp) or
//           the pre-round value t = x / s - zp leaves [lo, hi] (Round.backward :169-174), with
//           lo/hi = observer min/max mapped through the same affine map (and symmetrised for q_type 0, :154-158).
// s / zp / lo / hi are per tensor (rows = 1) or per output channel (row r covers `inner` consecutive elements).
__device__ __forceinline__ float round_half_away(float t) { return copysignf(floorf(__fadd_rn(fabsf(t), 0.5f)), t); }

__global__ void qat_fq_fwd_kernel(const float* __restrict__ x, size_t n, size_t inner, int rows,
                                  const float* __restrict__ scale, const float* __restrict__ zp, float qmin, float qmax,
                                  float* __restrict__ y) {
  EW_LOOP(i, n) {
    const int r = rows > 1 ? (int)((i / inner) % rows) : 0;
    const float s = scale[r], z = zp[r];
    const float t = __fsub_rn(__fdiv_rn(x[i], s), z);
    float q = round_half_away(t);
    q = fminf(fmaxf(q, qmin), qmax);
    y[i] = __fmul_rn(__fadd_rn(q, z), s);
  }
}

__global__ void qat_fq_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, size_t n, size_t inner, int rows,
                                  const float* __restrict__ scale, const float* __restrict__ zp,
                                  const float* __restrict__ obs_min, const float* __restrict__ obs_max, int q_type,
                                  float qmin, float qmax, float* __restrict__ gx) {
  EW_LOOP(i, n) {
    const int r = rows > 1 ? (int)((i / inner) % rows) : 0;
    const float s = scale[r], z = zp[r];
    // Round.forward's bounds (:154-162): the observer range through the same affine map, symmetrised for q_type 0
    float lo = __fsub_rn(__fdiv_rn(obs_min[r], s), z), hi = __fsub_rn(__fdiv_rn(obs_max[r], s), z);
    if (q_type == 0) { hi = fmaxf(fabsf(lo), fabsf(hi)); lo = -hi; }
    const float t = __fsub_rn(__fdiv_rn(x[i], s), z);
    const float q = round_half_away(t);
    const bool pass = (q >= qmin) && (q <= qmax) && !(t > hi) && !(t < lo);
    gx[i] = pass ? __fdiv_rn(__fmul_rn(g[i], s), s) : 0.f;
  }
}

// Quantizer.update_qparams (:299-327): (scale, zero_point) per row from the observer range
__global__ void qat_qparams_kernel(const float* __restrict__ mn, const float* __restrict__ mx, int rows, int q_type,
                                   float qmin, float qmax, float eps, float* __restrict__ scale, float* __restrict__ zp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  if (q_type == 0) {
    const float qr = __fdiv_rn(__fsub_rn(qmax, qmin), 2.f);                       // float(max - min) / 2
    const float fr = fmaxf(fabsf(mn[r]), fabsf(mx[r]));
    scale[r] = fmaxf(__fdiv_rn(fr, qr), eps);
    zp[r] = 0.f;
  } else {
    const float qr = __fsub_rn(qmax, qmin);
    const float sc = fmaxf(__fdiv_rn(__fsub_rn(mx[r], mn[r]), qr), eps);
    scale[r] = sc;
    const float m = mn[r];
    const float sg = (m > 0.f) ? 1.f : (m < 0.f ? -1.f : 0.f);
    zp[r] = __fmul_rn(sg, floorf(__fadd_rn(fabsf(__fdiv_rn(m, sc)), 0.5f)));
  }
}

// observer range update (:68-80 MinMaxObserver, :107-119 MovingAverageMinMaxObserver): mode 0 = take the current
// range, 1 = running min / max, 2 = exponential moving average with `momentum`
__global__ void qat_observer_combine_kernel(const float* __restrict__ cur_min, const float* __restrict__ cur_max, int rows,
                                            int mode, float momentum, float* __restrict__ mn, float* __restrict__ mx) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float a = cur_min[r], b = cur_max[r];
  if (mode == 0) { mn[r] = a; mx[r] = b; }
  else if (mode == 1) { mn[r] = fminf(a, mn[r]); mx[r] = fmaxf(b, mx[r]); }
  else {
    const float keep = __fsub_rn(1.f, momentum);
    mn[r] = __fadd_rn(__fmul_rn(keep, mn[r]), __fmul_rn(momentum, a));
    mx[r] = __fadd_rn(__fmul_rn(keep, mx[r]), __fmul_rn(momentum, b));
  }
}

__global__ void qat_fill2_kernel(float* mm) { mm[0] = INFINITY; mm[1] = -INFINITY; }

// k-th smallest |x| (1-based k, torch.kthvalue of HistogramObserver :136-138) by a two-level radix select on the float
// bit pattern (non-negative floats order like their bits): pass A histograms bits 30..15, pass B bits 14..0 inside the
// selected bucket.  Exact and deterministic.
__global__ void __launch_bounds__(256) kth_hist_kernel(const float* __restrict__ x, size_t n, int pass, const unsigned* __restrict__ sel,
                                                       unsigned* __restrict__ hist) {
  const unsigned want = pass ? sel[0] : 0u;
  EW_LOOP(i, n) {
    const unsigned b = __float_as_uint(fabsf(x[i]));
    if (pass == 0) atomicAdd(hist + (b >> 15), 1u);
    else if ((b >> 15) == want) atomicAdd(hist + (b & 0x7fffu), 1u);
  }
}
// one CTA: find the bucket holding rank k (sel[1] = rank inside it for the next pass); pass 1 writes the value
__global__ void __launch_bounds__(1024) kth_scan_kernel(const unsigned* __restrict__ hist, int nb, int pass, unsigned long long k,
                                                        unsigned* __restrict__ sel, float* __restrict__ out) {
  __shared__ unsigned long long s_part[1024];
  const int per = (nb + 1023) / 1024;
  const int b0 = threadIdx.x * per;
  unsigned long long sum = 0;
  for (int j = 0; j < per && b0 + j < nb; ++j) sum += hist[b0 + j];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long rank = pass ? (unsigned long long)sel[1] : k;         // 1-based rank still to skip into
    unsigned long long run = 0;
    int t = 0;
    while (t < 1023 && run + s_part[t] < rank) { run += s_part[t]; ++t; }
    int b = t * per;
    while (b < nb - 1 && run + hist[b] < rank) { run += hist[b]; ++b; }
    if (pass == 0) { sel[0] = (unsigned)b; sel[1] = (unsigned)(rank - run); }
    else out[0] = __uint_as_float((sel[0] << 15) | (unsigned)b);
  }
}

// per-row (min, max) of a [rows, cols] matrix: the "C" level observers (:27-30); one CTA per row
__global__ void __launch_bounds__(256) minmax_rows_kernel(const float* __restrict__ x, size_t cols, float* __restrict__ mn,
                                                          float* __restrict__ mx) {
  const float* row = x + (size_t)blockIdx.x * cols;
  float lo = INFINITY, hi = -INFINITY;
  for (size_t i = threadIdx.x; i < cols; i += blockDim.x) {
    const float v = row[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[8], s_hi[8];
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
    mn[blockIdx.x] = lo;
    mx[blockIdx.x] = hi;
  }
}

// Evaluation post-step of test.py:148-170 on device: out = clip(y + nearest_upsample(x), 0, 1) (the x2 model adds the
// upsampled input OUTSIDE the network, :149-155) and the squared-error sum of the PSNR the script prints:
//   metric 0: plain image values (compare_psnr, data_range 1),  metric 1: channel 0 * 255 (:165-166, x4),
//   metric 2: BT.601 luma of rgb_to_yuv (:133-136, :167-168, x2), clipped to [0, 255].
// One thread per HR pixel; fp64 partial sums, one atomicAdd per CTA.
__global__ void __launch_bounds__(256) sr_post_kernel(const float* __restrict__ y, const float* __restrict__ x,
                                                      const float* __restrict__ gt, int B, int C, int Hr, int Wr, int r,
                                                      int metric, float* __restrict__ out, double* __restrict__ sqerr) {
  const size_t plane = (size_t)Hr * Wr;
  const size_t n = (size_t)B * plane;
  double acc = 0.0;
  EW_LOOP(i, n) {
    const size_t b = i / plane, p = i - b * plane;
    const int hy = (int)(p / Wr), wx = (int)(p - (size_t)hy * Wr);
    const int H = Hr / r, W = Wr / r;
    double ypred = 0.0, ygt = 0.0;
    for (int c = 0; c < C; ++c) {
      const size_t o = (b * C + c) * plane + p;
      float v = y[o];
      if (x) v = __fadd_rn(v, x[((b * C + c) * H + hy / r) * (size_t)W + wx / r]);
      v = fminf(fmaxf(v, 0.f), 1.f);
      if (out) out[o] = v;
      if (gt) {
        const float g = gt[o];
        if (metric == 0) { const double d = (double)g - (double)v; acc += d * d; }
        else if (metric == 1) { if (c == 0) { const double d = (double)g * 255.0 - (double)v * 255.0; acc += d * d; } }
        else {
          const double wgt = c == 0 ? 65.481 : (c == 1 ? 128.553 : 24.966);
          ypred += (double)v * wgt;
          ygt += (double)g * wgt;
        }
      }
    }
    if (gt && metric == 2) {
      const double a = fmin(fmax(ygt + 16.0, 0.0), 255.0), bb = fmin(fmax(ypred + 16.0, 0.0), 255.0);
      acc += (a - bb) * (a - bb);
    }
  }
  if (!gt) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double s_acc[8];
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_acc[w];
    atomicAdd(sqerr, t);
  }
}


// Structural similarity as test.py:172-175 calls it (skimage.metrics.structural_similarity defaults: 7 x 7 uniform
// window, K1 = 0.01, K2 = 0.03, sample covariance, mean over the interior (H - 6) x (W - 6) pixels of every plane).
// One CTA = a 32 x 32 block of window centres of one plane; both images' (38 x 38) patches are staged in shared
// memory, every thread accumulates the five window sums of 4 centres in fp64 and the CTA adds its S total once.
constexpr int SSIM_WIN = 7, SSIM_T = 32, SSIM_P = SSIM_T + SSIM_WIN - 1;
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W,
                                                   double c1, double c2, double* __restrict__ total) {
  __shared__ float sa[SSIM_P][SSIM_P + 1], sb[SSIM_P][SSIM_P + 1];
  __shared__ double s_acc[8];
  const size_t plane = (size_t)blockIdx.z * H * W;
  const int x0 = blockIdx.x * SSIM_T, y0 = blockIdx.y * SSIM_T;        // top-left of the patch == first window origin
  for (int i = threadIdx.x; i < SSIM_P * SSIM_P; i += blockDim.x) {
    const int py = i / SSIM_P, px = i - py * SSIM_P;
    const int gy = y0 + py, gx = x0 + px;
    const bool in = gy < H && gx < W;
    sa[py][px] = in ? a[plane + (size_t)gy * W + gx] : 0.f;
    sb[py][px] = in ? b[plane + (size_t)gy * W + gx] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
  const double np_ = (double)(SSIM_WIN * SSIM_WIN), cov_norm = np_ / (np_ - 1.0);
  double acc = 0.0;
  for (int ty = ty0; ty < SSIM_T; ty += 8) {
    if (y0 + ty + SSIM_WIN > H || x0 + tx + SSIM_WIN > W) continue;     // window must lie inside the image
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
    for (int dy = 0; dy < SSIM_WIN; ++dy)
#pragma unroll
      for (int dx = 0; dx < SSIM_WIN; ++dx) {
        const double u = sa[ty + dy][tx + dx], v = sb[ty + dy][tx + dx];
        sx += u; sy += v; sxx += u * u; syy += v * v; sxy += u * v;
      }
    const double ux = sx / np_, uy = sy / np_;
    const double vx = cov_norm * (sxx / np_ - ux * ux), vy = cov_norm * (syy / np_ - uy * uy);
    const double vxy = cov_norm * (sxy / np_ - ux * uy);
    acc += ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_acc[w];
    atomicAdd(total, t);
  }
}

}  // namespace sesrq


using namespace sesrq;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int sesrq_op_quant_first(const float* d_x, size_t n, double scale, int zero, float* d_q, float* d_u, void* stream) {
  if (!d_x && n) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  op_quant_first_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, n, (float)scale, (float)zero,
                                                                   (float)std::max(zero, -128), d_q, d_u);
  SESRQ_LAUNCH_CHECK("op_quant_first_kernel");
  return SESRQ_OK;
}

int sesrq_op_quant_next(const float* d_x, size_t n, int zero, float* d_q, float* d_u, void* stream) {
  if (!d_x && n) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  op_quant_next_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, n, (float)zero, (float)std::max(zero, -128), d_q, d_u);
  SESRQ_LAUNCH_CHECK("op_quant_next_kernel");
  return SESRQ_OK;
}

int sesrq_op_quant_residual(const float* d_x, const float* d_shortcut, size_t n, int m16, int shift_n, int zero,
                            float* d_ic, float* d_q, float* d_u, void* stream) {
  if ((!d_x || !d_shortcut) && n) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  float mf = (float)((double)m16 * std::ldexp(1.0, -shift_n));
  op_quant_residual_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, d_shortcut, n, mf, (float)zero,
                                                                      (float)std::max(zero, -128), d_ic, d_q, d_u);
  SESRQ_LAUNCH_CHECK("op_quant_residual_kernel");
  return SESRQ_OK;
}

int sesrq_op_pe_split(const float* d_x, int B, int C, int HW, int pe_num, float* d_out, void* stream) {
  if (!d_x || !d_out || B < 0 || C <= 0 || HW <= 0 || pe_num != 4) return SESRQ_EINVAL;
  size_t n = (size_t)B * pe_num * C * HW;
  if (n == 0) return SESRQ_OK;
  op_pe_split_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, C, HW, pe_num, n, d_out);
  SESRQ_LAUNCH_CHECK("op_pe_split_kernel");
  return SESRQ_OK;
}

int sesrq_op_pe_merge(const float* d_conv, int B, int Cout, int HW, const float* d_pe_corr, const float* d_add_const,
                      int pe_acc_bits, int pe_add_bits, float* d_pe_out, float* d_pe_add, float* d_out,
                      int32_t* d_overflow, void* stream) {
  if (!d_conv || !d_pe_corr || !d_add_const || !d_out || B < 0 || Cout <= 0 || HW <= 0) return SESRQ_EINVAL;
  size_t n = (size_t)B * Cout * HW;
  if (n == 0) return SESRQ_OK;
  float lo_acc = -(float)(1 << (pe_acc_bits - 1)), hi_acc = (float)((1 << (pe_acc_bits - 1)) - 1);
  float lo_add = -(float)(1 << (pe_add_bits - 1)), hi_add = (float)((1 << (pe_add_bits - 1)) - 1);
  op_pe_merge_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_conv, Cout, HW, n, d_pe_corr, d_add_const, lo_acc,
                                                                hi_acc, lo_add, hi_add, d_pe_out, d_pe_add, d_out,
                                                                d_overflow);
  SESRQ_LAUNCH_CHECK("op_pe_merge_kernel");
  return SESRQ_OK;
}

int sesrq_op_requant(const float* d_x, size_t n, int m16, int shift_n, float* d_y, float* d_relu_copy, void* stream) {
  if ((!d_x || !d_y) && n) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  op_requant_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, n, (float)m16, (float)std::ldexp(1.0, -shift_n), d_y,
                                                               d_relu_copy);
  SESRQ_LAUNCH_CHECK("op_requant_kernel");
  return SESRQ_OK;
}

int sesrq_op_requant_final(const float* d_x, size_t n, int m16, int shift_n, int zero_out, double scale_out,
                           float* d_q5, float* d_y, void* stream) {
  if (!d_x && n) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  op_requant_final_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, n, (float)m16, (float)std::ldexp(1.0, -shift_n),
                                                                     (float)zero_out, (float)scale_out, d_q5, d_y);
  SESRQ_LAUNCH_CHECK("op_requant_final_kernel");
  return SESRQ_OK;
}

int sesrq_op_conv_int(const float* d_x, int B, int Cin, int H, int W, const float* d_w, int Cout, int k, float* d_out,
                      void* stream) {
  if (!d_x || !d_w || !d_out || B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || (k != 3 && k != 5 && k != 1))
    return SESRQ_EINVAL;
  size_t n = (size_t)B * Cout * H * W;
  if (n == 0) return SESRQ_OK;
  op_conv_int_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, Cin, H, W, d_w, Cout, k, n, d_out);
  SESRQ_LAUNCH_CHECK("op_conv_int_kernel");
  return SESRQ_OK;
}

int sesrq_op_fakequant(const float* d_x, size_t n, double scale, int zero, float* d_y, void* stream) {
  if ((!d_x || !d_y) && n) return SESRQ_EINVAL;
  if (!(scale > 0.0)) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  op_fakequant_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, n, (float)scale, (float)zero, d_y);
  SESRQ_LAUNCH_CHECK("op_fakequant_kernel");
  return SESRQ_OK;
}

int sesrq_op_conv_fq(const float* d_x, int B, int Cin, int H, int W, const float* d_wcodes, int Cout, int k,
                     double in_scale, double w_scale, float* d_out, void* stream) {
  if (!d_x || !d_wcodes || !d_out || B < 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || (k != 3 && k != 5 && k != 1))
    return SESRQ_EINVAL;
  if (!(in_scale > 0.0) || !(w_scale > 0.0)) return SESRQ_EINVAL;
  size_t n = (size_t)B * Cout * H * W;
  if (n == 0) return SESRQ_OK;
  const float sf = (float)in_scale;
  const float cscale = (float)((double)sf * (double)(float)w_scale);        // as sesrq_fq.cu
  op_conv_fq_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, Cin, H, W, d_wcodes, Cout, k, n, sf, cscale, d_out);
  SESRQ_LAUNCH_CHECK("op_conv_fq_kernel");
  return SESRQ_OK;
}

int sesrq_op_pe_merge_fq(const float* d_conv, int B, int Cout, int HW, float lo_acc, float hi_acc, float lo_add,
                         float hi_add, const float* d_bias, float* d_out, void* stream) {
  if (!d_conv || !d_out || B < 0 || Cout <= 0 || HW <= 0) return SESRQ_EINVAL;
  size_t n = (size_t)B * Cout * HW;
  if (n == 0) return SESRQ_OK;
  op_pe_merge_fq_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_conv, Cout, HW, n, lo_acc, hi_acc, lo_add, hi_add,
                                                                   d_bias, d_out);
  SESRQ_LAUNCH_CHECK("op_pe_merge_fq_kernel");
  return SESRQ_OK;
}

int sesrq_qat_fakequant_fwd(const float* d_x, size_t n, size_t inner, int rows, const float* d_scale, const float* d_zp,
                            float qmin, float qmax, float* d_y, void* stream) {
  if ((!d_x || !d_y || !d_scale || !d_zp) && n) return SESRQ_EINVAL;
  if (rows < 1 || inner < 1) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  qat_fq_fwd_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, n, inner, rows, d_scale, d_zp, qmin, qmax, d_y);
  SESRQ_LAUNCH_CHECK("qat_fq_fwd_kernel");
  return SESRQ_OK;
}

int sesrq_qat_fakequant_bwd(const float* d_x, const float* d_gy, size_t n, size_t inner, int rows, const float* d_scale,
                            const float* d_zp, const float* d_obs_min, const float* d_obs_max, int q_type, float qmin,
                            float qmax, float* d_gx, void* stream) {
  if ((!d_x || !d_gy || !d_gx || !d_scale || !d_zp || !d_obs_min || !d_obs_max) && n) return SESRQ_EINVAL;
  if (rows < 1 || inner < 1) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  qat_fq_bwd_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, d_gy, n, inner, rows, d_scale, d_zp, d_obs_min, d_obs_max,
                                                               q_type, qmin, qmax, d_gx);
  SESRQ_LAUNCH_CHECK("qat_fq_bwd_kernel");
  return SESRQ_OK;
}

int sesrq_qat_qparams(const float* d_min, const float* d_max, int rows, int q_type, float qmin, float qmax, float eps,
                      float* d_scale, float* d_zp, void* stream) {
  if (!d_min || !d_max || !d_scale || !d_zp || rows < 1) return SESRQ_EINVAL;
  qat_qparams_kernel<<<(rows + 127) / 128, 128, 0, ST(stream)>>>(d_min, d_max, rows, q_type, qmin, qmax, eps, d_scale, d_zp);
  SESRQ_LAUNCH_CHECK("qat_qparams_kernel");
  return SESRQ_OK;
}

int sesrq_qat_observe(const float* d_x, int rows, size_t cols, int mode, float momentum, float* d_min, float* d_max,
                      float* d_tmp, void* stream) {
  if (!d_x || !d_min || !d_max || !d_tmp || rows < 1 || cols < 1 || mode < 0 || mode > 2) return SESRQ_EINVAL;
  if (rows == 1) {                 // whole tensor: many CTAs + one atomic pair each (the activation observers)
    qat_fill2_kernel<<<1, 1, 0, ST(stream)>>>(d_tmp);
    SESRQ_LAUNCH_CHECK("qat_fill2_kernel");
    int rc = sesrq_minmax_update(d_x, cols, d_tmp, stream);
    if (rc) return rc;
    qat_observer_combine_kernel<<<1, 32, 0, ST(stream)>>>(d_tmp, d_tmp + 1, 1, mode, momentum, d_min, d_max);
  } else {                         // per output channel: one CTA per row (the weight observers)
    minmax_rows_kernel<<<rows, 256, 0, ST(stream)>>>(d_x, cols, d_tmp, d_tmp + rows);
    SESRQ_LAUNCH_CHECK("minmax_rows_kernel");
    qat_observer_combine_kernel<<<(rows + 127) / 128, 128, 0, ST(stream)>>>(d_tmp, d_tmp + rows, rows, mode, momentum, d_min, d_max);
  }
  SESRQ_LAUNCH_CHECK("qat_observer_combine_kernel");
  return SESRQ_OK;
}

int sesrq_kth_abs(const float* d_x, size_t n, unsigned long long k, float* d_out, void* d_ws, size_t ws_bytes, void* stream) {
  constexpr int NB_A = 1 << 16, NB_B = 1 << 15;
  if (!d_x || !d_out || !d_ws || n == 0 || k < 1 || k > n) return SESRQ_EINVAL;
  if (ws_bytes < (size_t)(NB_A + NB_B + 2) * sizeof(unsigned)) return SESRQ_ENOMEM;
  unsigned* hist_a = static_cast<unsigned*>(d_ws);
  unsigned* hist_b = hist_a + NB_A;
  unsigned* sel = hist_b + NB_B;
  SESRQ_CUDA(cudaMemsetAsync(d_ws, 0, (size_t)(NB_A + NB_B + 2) * sizeof(unsigned), ST(stream)));
  kth_hist_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, n, 0, sel, hist_a);
  SESRQ_LAUNCH_CHECK("kth_hist_kernel");
  kth_scan_kernel<<<1, 1024, 0, ST(stream)>>>(hist_a, NB_A, 0, k, sel, d_out);
  SESRQ_LAUNCH_CHECK("kth_scan_kernel");
  kth_hist_kernel<<<ew_grid(n), EW_THREADS, 0, ST(stream)>>>(d_x, n, 1, sel, hist_b);
  SESRQ_LAUNCH_CHECK("kth_hist_kernel");
  kth_scan_kernel<<<1, 1024, 0, ST(stream)>>>(hist_b, NB_B, 1, k, sel, d_out);
  SESRQ_LAUNCH_CHECK("kth_scan_kernel");
  return SESRQ_OK;
}

int sesrq_minmax_rows(const float* d_x, int rows, size_t cols, float* d_min, float* d_max, void* stream) {
  if (!d_x || !d_min || !d_max || rows < 1 || cols < 1) return SESRQ_EINVAL;
  minmax_rows_kernel<<<rows, 256, 0, ST(stream)>>>(d_x, cols, d_min, d_max);
  SESRQ_LAUNCH_CHECK("minmax_rows_kernel");
  return SESRQ_OK;
}

int sesrq_op_sr_post(const float* d_y, const float* d_x, const float* d_gt, int B, int C, int Hr, int Wr, int r, int metric,
                     float* d_out, double* d_sqerr, void* stream) {
  if (!d_y || B < 0 || C < 1 || Hr < 1 || Wr < 1 || r < 1 || Hr % r || Wr % r || metric < 0 || metric > 2) return SESRQ_EINVAL;
  if (d_gt && !d_sqerr) return SESRQ_EINVAL;
  if (metric == 2 && C != 3) return SESRQ_EINVAL;
  const size_t n = (size_t)B * Hr * Wr;
  if (n == 0) return SESRQ_OK;
  sr_post_kernel<<<ew_grid(n), EW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_y, d_x, d_gt, B, C, Hr, Wr, r, metric,
                                                                                       d_out, d_sqerr);
  SESRQ_LAUNCH_CHECK("sr_post_kernel");
  return SESRQ_OK;
}

int sesrq_op_ssim(const float* d_a, const float* d_b, int planes, int H, int W, double data_range, double* d_sum, void* stream) {
  if (!d_a || !d_b || !d_sum || planes < 1 || H < SSIM_WIN || W < SSIM_WIN || !(data_range > 0.0)) return SESRQ_EINVAL;
  const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
  dim3 grid((unsigned)((W - SSIM_WIN + 1 + SSIM_T - 1) / SSIM_T), (unsigned)((H - SSIM_WIN + 1 + SSIM_T - 1) / SSIM_T), (unsigned)planes);
  if (grid.y > 65535u || grid.z > 65535u) return SESRQ_EINVAL;
  ssim_kernel<<<grid, 256, 0, ST(stream)>>>(d_a, d_b, H, W, c1, c2, d_sum);
  SESRQ_LAUNCH_CHECK("ssim_kernel");
  return SESRQ_OK;
}

}  // extern "C"
