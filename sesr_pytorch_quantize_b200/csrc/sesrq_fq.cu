// Fake-quant (exe_mode 0) forward: the test.py path (test.py:79-106,146) with per-frame dynamic
// quantisation of every domain, fused min/max observation and optional KL histograms.
//
// The reference computes conv((q - z) * s, W * s_w) in fp32 with library-defined summation order
// (quan_func.py:207-215 feeding nn.Conv2d).  Here the convolution is carried out exactly on the integer
// operands (q - z) and W and scaled once by f32(s) * f32(s_w); the float-domain emulation of the
// 18-/20-bit clamps (:330-333, :431-434) and the fake-quant bias (:402,:412,:459) follow the reference
// op for op.  Parity with the reference on this path is therefore tolerance-based:
// PSNR >= 60 dB at range 1.0 and max-abs <= 8 * s_5.
#include <algorithm>
#include <new>
#include <vector>
#include "sesrq_internal.h"

namespace sesrq {

int launch_hist_dev(const float* d_x, size_t n, const float* d_range, unsigned long long* d_hist, cudaStream_t st);

struct FqLayerDev {
  int k, cin, cout, relu;
  int32_t* d_w = nullptr;     // [tap][cin][16 co] int32 codes
  float* d_bias = nullptr;    // [cout]
  double w_scale;
};

// (scale, zero) of one domain from this frame's (min, max): quan_func.py:198-205 in Python doubles
struct DomQ { double s; double z; float sf; float zf; int degenerate; };
__device__ __forceinline__ DomQ dom_quant(const float* mm) {
  DomQ d;
  double mn = (double)mm[0], mx = (double)mm[1];
  d.degenerate = (mx == mn);
  d.s = d.degenerate ? 1.0 : (mx - mn) / 255.0;
  d.z = -128.0 - rint(mn / d.s);        // Python round(): half-to-even on the double
  d.sf = (float)d.s;
  d.zf = (float)d.z;
  return d;
}

__global__ void fq_reset_kernel(float* mm_frame) {
  int i = threadIdx.x;
  if (i < SESRQ_NUM_DOMAINS) { mm_frame[2 * i] = INFINITY; mm_frame[2 * i + 1] = -INFINITY; }
}

// X fp32 [C,H,W] -> (q - z) int16 NHWC16   (quan_func.py:207)
__global__ void __launch_bounds__(256) fq_quant_kernel(const float* __restrict__ x, int C, int HW,
                                                       const float* __restrict__ mm, uint4* __restrict__ d16) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  DomQ dq = dom_quant(mm);
  uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < SESRQ_MAX_CH; ++c) {
    if (c < C) {
      int q = quant_add_zero(__fdiv_rn(__ldg(x + (size_t)c * HW + i), dq.sf), dq.zf);
      int d = q - (int)dq.z;
      w[c >> 1] |= ((uint32_t)d & 0xffffu) << (16 * (c & 1));
    }
  }
  d16[2 * (size_t)i] = make_uint4(w[0], w[1], w[2], w[3]);
  d16[2 * (size_t)i + 1] = make_uint4(w[4], w[5], w[6], w[7]);
}

template <int K>
__global__ void __launch_bounds__(128)
fq_conv_kernel(const uint4* __restrict__ d16, const int32_t* __restrict__ wq, const float* __restrict__ bias,
               int cin, int cout, int relu, double w_scale, int acc_bits, int add_bits,
               const float* __restrict__ mm_in, const float* __restrict__ residual, float* __restrict__ out,
               float* __restrict__ mm_out, int H, int W) {
  constexpr int TAPS = K * K;
  constexpr int PAD = (K - 1) / 2;
  extern __shared__ int32_t w_s[];                 // [TAPS][cin][16]
  __shared__ float s_bias[SESRQ_MAX_CH];
  __shared__ float s_red[2][4];
  const int nw = TAPS * cin * 16;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) w_s[i] = wq[i];
  const DomQ dq = dom_quant(mm_in);
  const double bias_scale = dq.s * w_scale;                          // :448
  if (threadIdx.x < cout) {
    float bs = (float)bias_scale;
    float bq = fminf(fmaxf(rintf(__fdiv_rn(bias[threadIdx.x], bs)), -32768.f), 32767.f);   // :402
    s_bias[threadIdx.x] = __fmul_rn(bq, bs);                         // :412
  }
  __syncthreads();
  // float-domain clamp bounds, evaluated in double like the Python expressions (:331-332, :432-433)
  const double hi_a = (double)((1 << (acc_bits - 1)) - 1), lo_a = -(double)(1 << (acc_bits - 1));
  const double hi_s = (double)((1 << (add_bits - 1)) - 1), lo_s = -(double)(1 << (add_bits - 1));
  const float fmax_a = (float)(((hi_a - dq.z) * dq.s) * w_scale), fmin_a = (float)(((lo_a - dq.z) * dq.s) * w_scale);
  const float fmax_s = (float)(((hi_s - dq.z) * dq.s) * w_scale), fmin_s = (float)(((lo_s - dq.z) * dq.s) * w_scale);
  const float cscale = (float)((double)dq.sf * (double)(float)w_scale);

  const int HW = H * W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  float vmin = INFINITY, vmax = -INFINITY;
  if (pix < HW) {
    const int xx = pix % W, yy = pix / W;
    int acc[16][4];
#pragma unroll
    for (int co = 0; co < 16; ++co)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[co][p] = 0;
#pragma unroll 1
    for (int dy = 0; dy < K; ++dy) {
      const int sy = yy + dy - PAD;
      if (sy < 0 || sy >= H) continue;                               // zero padding of (q - z) * s
#pragma unroll 1
      for (int dx = 0; dx < K; ++dx) {
        const int sx = xx + dx - PAD;
        if (sx < 0 || sx >= W) continue;
        const uint4 a = __ldg(d16 + 2 * ((size_t)sy * W + sx));
        const uint4 b = __ldg(d16 + 2 * ((size_t)sy * W + sx) + 1);
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const int32_t* wt = w_s + (dy * K + dx) * cin * 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < cin) {
            const int d = (int)(int16_t)((v[c >> 1] >> (16 * (c & 1))) & 0xffffu);
            const int4* wc = reinterpret_cast<const int4*>(wt + c * 16);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int4 w4 = wc[g];
              acc[4 * g + 0][c & 3] += d * w4.x;
              acc[4 * g + 1][c & 3] += d * w4.y;
              acc[4 * g + 2][c & 3] += d * w4.z;
              acc[4 * g + 3][c & 3] += d * w4.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int co = 0; co < 16; ++co) {
      if (co < cout) {
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float yp = __fmul_rn((float)acc[co][p], cscale);
          yp = fminf(fmaxf(yp, fmin_a), fmax_a);                     // :333
          s = __fadd_rn(s, yp);                                      // :386
        }
        s = fminf(fmaxf(s, fmin_s), fmax_s);                         // :434
        s = __fadd_rn(s, s_bias[co]);                                // :459
        if (relu) s = fmaxf(s, 0.f);
        if (residual) s = __fadd_rn(s, residual[(size_t)co * HW + pix]);   // AddOp (models/sesr.py:45)
        out[(size_t)co * HW + pix] = s;
        vmin = fminf(vmin, s);
        vmax = fmaxf(vmax, s);
      }
    }
  }
  // block min/max -> frame min/max of the produced domain
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = vmin; s_red[1][threadIdx.x >> 5] = vmax; }
  __syncthreads();
  if (threadIdx.x == 0 && mm_out) {
    float lo = fminf(fminf(s_red[0][0], s_red[0][1]), fminf(s_red[0][2], s_red[0][3]));
    float hi = fmaxf(fmaxf(s_red[1][0], s_red[1][1]), fmaxf(s_red[1][2], s_red[1][3]));
    // float atomics via ordered-int CAS
    int* a = reinterpret_cast<int*>(mm_out);
    int old = a[0];
    while (lo < __int_as_float(old)) { int as = old; old = atomicCAS(a, as, __float_as_int(lo)); if (old == as) break; }
    old = a[1];
    while (hi > __int_as_float(old)) { int as = old; old = atomicCAS(a + 1, as, __float_as_int(hi)); if (old == as) break; }
  }
}

// domain-5 fake-quant + PixelShuffle (test.py:90-91 maps PixelShuffle to the quantiser; :207-215)
__global__ void __launch_bounds__(256) fq_final_kernel(const float* __restrict__ x5, int C, int H, int W, int r,
                                                       int do_quant, const float* __restrict__ mm, float* __restrict__ y) {
  const int HW = H * W;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)C * HW) return;
  float v = x5[i];
  if (do_quant) {
    DomQ dq = dom_quant(mm);
    float q = (float)quant_add_zero(__fdiv_rn(v, dq.sf), dq.zf);
    v = __fmul_rn(__fsub_rn(q, dq.zf), dq.sf);                      // :215
  }
  int c = (int)(i / HW);
  int p = (int)(i - (size_t)c * HW);
  int yy = p / W, xx = p % W;
  int r2 = r * r;
  int cc = c / r2, ii = (c % r2) / r, jj = c % r;
  y[((size_t)cc * (H * r) + (yy * r + ii)) * (size_t)(W * r) + (xx * r + jj)] = v;
}

__global__ void fq_commit_kernel(const float* __restrict__ mm_frame, float* __restrict__ mm_run,
                                 float* __restrict__ mm_frame_out) {
  int i = threadIdx.x;
  if (i < SESRQ_NUM_DOMAINS) {
    float lo = mm_frame[2 * i], hi = mm_frame[2 * i + 1];
    if (mm_run) {                                                    // quan_func.py:188-195
      if (lo < mm_run[2 * i]) mm_run[2 * i] = lo;
      if (hi > mm_run[2 * i + 1]) mm_run[2 * i + 1] = hi;
    }
    if (mm_frame_out) { mm_frame_out[2 * i] = lo; mm_frame_out[2 * i + 1] = hi; }
  }
}

}  // namespace sesrq

using namespace sesrq;

struct SesrqFq {
  SesrqFqDesc desc;
  FqLayerDev layers[SESRQ_NUM_LAYERS];
};

namespace {
struct FqPlan { size_t off_mm, off_d16, off_f[3], total; };
FqPlan fq_plan(int H, int W) {
  FqPlan p{};
  size_t hw = (size_t)H * W;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = 0;
  p.off_mm = off; off += 256;
  p.off_d16 = off; off += up(hw * 32);
  for (int i = 0; i < 3; ++i) { p.off_f[i] = off; off += up(hw * SESRQ_MAX_CH * 4); }
  p.total = off;
  return p;
}
}  // namespace

extern "C" {

int sesrq_fq_create(const SesrqFqDesc* d, SesrqFq** out) {
  if (!d || !out) return SESRQ_EINVAL;
  *out = nullptr;
  if (d->abi_version != SESRQ_ABI_VERSION || d->n_layers != SESRQ_NUM_LAYERS) return SESRQ_EINVAL;
  if (d->r_shuffle < 1 || d->r_shuffle > 4) return SESRQ_EINVAL;
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    if ((d->ksize[i] != 3 && d->ksize[i] != 5) || d->cin[i] < 1 || d->cin[i] > SESRQ_MAX_CH || d->cout[i] < 1 ||
        d->cout[i] > SESRQ_MAX_CH || !d->h_weight[i] || !d->h_bias[i])
      return SESRQ_EINVAL;
    if (i > 0 && d->cin[i] != d->cout[i - 1]) return SESRQ_EINVAL;
  }
  if (d->cout[4] % (d->r_shuffle * d->r_shuffle) != 0) return SESRQ_EINVAL;
  int dev = 0;
  SESRQ_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SESRQ_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return SESRQ_EARCH;
  SesrqFq* fq = new (std::nothrow) SesrqFq();
  if (!fq) return SESRQ_ENOMEM;
  fq->desc = *d;
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    FqLayerDev& L = fq->layers[i];
    L.k = d->ksize[i]; L.cin = d->cin[i]; L.cout = d->cout[i]; L.relu = d->relu[i]; L.w_scale = d->w_scale[i];
    const int taps = L.k * L.k;
    std::vector<int32_t> w((size_t)taps * L.cin * 16, 0);
    for (int co = 0; co < L.cout; ++co)
      for (int c = 0; c < L.cin; ++c)
        for (int t = 0; t < taps; ++t)
          w[((size_t)t * L.cin + c) * 16 + co] = d->h_weight[i][((size_t)co * L.cin + c) * taps + t];
    if (cudaMalloc(&L.d_w, w.size() * 4) != cudaSuccess || cudaMalloc(&L.d_bias, L.cout * 4) != cudaSuccess ||
        cudaMemcpy(L.d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(L.d_bias, d->h_bias[i], L.cout * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      set_cuda_error(cudaGetLastError(), "fq weight upload");
      sesrq_fq_destroy(fq);
      return SESRQ_ECUDA;
    }
  }
  *out = fq;
  return SESRQ_OK;
}

void sesrq_fq_destroy(SesrqFq* fq) {
  if (!fq) return;
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    if (fq->layers[i].d_w) cudaFree(fq->layers[i].d_w);
    if (fq->layers[i].d_bias) cudaFree(fq->layers[i].d_bias);
  }
  delete fq;
}

size_t sesrq_fq_ws_bytes(const SesrqFq* fq, int H, int W) {
  if (!fq || H <= 0 || W <= 0) return 0;
  return fq_plan(H, W).total;
}

int sesrq_fq_forward(const SesrqFq* fq, const float* d_x, int H, int W, float* d_y, float* d_minmax,
                     float* d_frame_minmax, unsigned long long* d_hist, const float* d_hist_range, void* d_ws,
                     size_t ws_bytes, void* stream) {
  if (!fq || !d_x || !d_y || !d_ws || H <= 0 || W <= 0) return SESRQ_EINVAL;
  if (d_hist && !d_hist_range) return SESRQ_EINVAL;
  if ((long long)H * W > (1LL << 27)) return SESRQ_EINVAL;
  FqPlan p = fq_plan(H, W);
  if (ws_bytes < p.total) return SESRQ_ENOMEM;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_ws);
  float* mm = reinterpret_cast<float*>(ws + p.off_mm);
  uint4* d16 = reinterpret_cast<uint4*>(ws + p.off_d16);
  float* F[3] = {reinterpret_cast<float*>(ws + p.off_f[0]), reinterpret_cast<float*>(ws + p.off_f[1]),
                 reinterpret_cast<float*>(ws + p.off_f[2])};
  const int HW = H * W;
  const SesrqFqDesc& D = fq->desc;

  fq_reset_kernel<<<1, 32, 0, st>>>(mm);
  SESRQ_LAUNCH_CHECK("fq_reset_kernel");
  int rc = sesrq_minmax_update(d_x, (size_t)D.cin[0] * HW, mm, stream);   // domain 0 (:177-178)
  if (rc) return rc;

  // tensor routing: X0=d_x, X1=F0 (kept for the AddOp), X2=F1, X3=F2, X4=F1, X5=F2
  const float* in_t[5] = {d_x, F[0], F[1], F[2], F[1]};
  float* out_t[5] = {F[0], F[1], F[2], F[1], F[2]};
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    const FqLayerDev& L = fq->layers[i];
    if (d_hist) {
      rc = launch_hist_dev(in_t[i], (size_t)L.cin * HW, d_hist_range + 2 * i, d_hist + (size_t)i * SESRQ_KL_BINS, st);
      if (rc) return rc;
    }
    fq_quant_kernel<<<cdiv(HW, 256), 256, 0, st>>>(in_t[i], L.cin, HW, mm + 2 * i, d16);
    SESRQ_LAUNCH_CHECK("fq_quant_kernel");
    const float* res = (i == 3 && D.residual) ? F[0] : nullptr;
    size_t smem = (size_t)L.k * L.k * L.cin * 16 * 4;
    if (L.k == 3)
      fq_conv_kernel<3><<<cdiv(HW, 128), 128, smem, st>>>(d16, L.d_w, L.d_bias, L.cin, L.cout, L.relu, L.w_scale,
                                                         D.pe_acc_bits, D.pe_add_bits, mm + 2 * i, res, out_t[i],
                                                         mm + 2 * (i + 1), H, W);
    else
      fq_conv_kernel<5><<<cdiv(HW, 128), 128, smem, st>>>(d16, L.d_w, L.d_bias, L.cin, L.cout, L.relu, L.w_scale,
                                                         D.pe_acc_bits, D.pe_add_bits, mm + 2 * i, res, out_t[i],
                                                         mm + 2 * (i + 1), H, W);
    SESRQ_LAUNCH_CHECK("fq_conv_kernel");
  }
  const int cout = D.cout[4];
  if (d_hist) {
    rc = launch_hist_dev(F[2], (size_t)cout * HW, d_hist_range + 10, d_hist + (size_t)5 * SESRQ_KL_BINS, st);
    if (rc) return rc;
  }
  // SESR: the PixelShuffle input is fake-quantised with this frame's range (test.py:90-91);
  // NRDM (observe_out): no PixelShuffle node, the conv output is only observed (quan_func.py:460-479)
  fq_final_kernel<<<cdiv(cout * HW, 256), 256, 0, st>>>(F[2], cout, H, W, D.r_shuffle, D.observe_out ? 0 : 1, mm + 10, d_y);
  SESRQ_LAUNCH_CHECK("fq_final_kernel");
  fq_commit_kernel<<<1, 32, 0, st>>>(mm, d_minmax, d_frame_minmax);
  SESRQ_LAUNCH_CHECK("fq_commit_kernel");
  return SESRQ_OK;
}

}  // extern "C"
