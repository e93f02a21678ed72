// Stand-alone myQL ops (exe_mode 1) on float32 NCHW device tensors that hold integer codes, exactly as
// the reference's un-fused fx graph passes them from node to node.  Used when a graph is not the full
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

}  // extern "C"
