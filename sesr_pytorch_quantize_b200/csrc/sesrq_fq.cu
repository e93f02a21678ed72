// Fake-quant (exe_mode 0) forward: the test.py path (test.py:79-106,146) with per-frame dynamic
// quantisation of every domain, fused min/max observation and optional KL histograms -- BATCHED: one pass of
// ~20 launches processes all B frames (frames are a batch dimension; every frame keeps its own ranges).
//
// The reference computes conv((q - z) * s, W * s_w) in fp32 with library-defined summation order
// (quan_func.py:207-215 feeding nn.Conv2d).  Here the convolution is carried out exactly on the integer
// operands (q - z) and W and scaled once by f32(s) * f32(s_w); the float-domain emulation of the
// 18-/20-bit clamps (:330-333, :431-434) and the fake-quant bias (:402,:412,:459) follow the reference
// op for op.  Parity with the reference on this path is therefore tolerance-based:
// PSNR >= 60 dB at range 1.0 and max-abs <= 8 * s_5; parity with oracle/fq_oracle.py is bit-exact.
//
// Data flow per layer i (all frames at once):
//   fqb_params_kernel   (min, max) of domain i per frame -> FqFrameParams[B] (scale, zero, clamp bounds, bias; doubles)
//   fqb_quant_*_kernel  X_i fp32 -> operand bytes b = q + 128 (NHWC16, PE-major byte order); d = q - z = b + dmin
//   conv_umma_kernel<NS = 4, EPI_FQ>  tcgen05 kind::i8 on the bytes, four PE-masked accumulator sets; epilogue:
//                       per-PE float products, float clamps, + bias, ReLU, AddOp -> X_{i+1} fp32 NHWC16 + per-frame
//                       (min, max) of the produced domain (one atomic pair per warp and tile)
// then the domain-5 fake-quant + PixelShuffle and the fold of the per-frame ranges into the running observer.
#include <algorithm>
#include <cmath>
#include <new>
#include <vector>
#include "sesrq_internal.h"

namespace sesrq {

int launch_hist_dev(const float* d_x, size_t n, const float* d_range, unsigned long long* d_hist, cudaStream_t st);

// (scale, zero) of one domain from this frame's (min, max): quan_func.py:198-205 in Python doubles
struct DomQ { double s; double z; float sf; float zf; int degenerate; };
__device__ __forceinline__ DomQ dom_quant(const float* mm) {
  DomQ d;
  double mn = (double)mm[0], mx = (double)mm[1];
  d.degenerate = !(mx > mn);
  d.s = d.degenerate ? 1.0 : (mx - mn) / 255.0;
  d.z = -128.0 - rint(mn / d.s);        // Python round(): half-to-even on the double
  d.sf = (float)d.s;
  d.zf = (float)d.z;
  return d;
}

constexpr int MM_STRIDE = 2 * SESRQ_NUM_DOMAINS;      // floats per frame: (min, max) of the six domains

__global__ void fqb_init_mm_kernel(float* mm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mm[i] = (i & 1) ? -INFINITY : INFINITY;
}

// (min, max) of every frame of x [B, per_frame]: grid (chunks, B), one atomic pair per CTA
__global__ void __launch_bounds__(256) fqb_minmax_frames_kernel(const float* __restrict__ x, size_t per_frame,
                                                                float* __restrict__ mm) {
  const float* xf = x + (size_t)blockIdx.y * per_frame;
  float lo = INFINITY, hi = -INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_frame; i += (size_t)gridDim.x * blockDim.x) {
    const float v = __ldg(xf + i);
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
    atomic_min_float(mm + (size_t)blockIdx.y * MM_STRIDE, lo);
    atomic_max_float(mm + (size_t)blockIdx.y * MM_STRIDE + 1, hi);
  }
}

// per-frame constants of layer `dom` (dom = 5: only the quantiser of the output domain)
__global__ void fqb_params_kernel(const float* __restrict__ mm, int dom, int B, double w_scale, const float* __restrict__ bias,
                                  int cout, int acc_bits, int add_bits, double wpos, double wneg,
                                  FqFrameParams* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  const DomQ dq = dom_quant(mm + (size_t)f * MM_STRIDE + 2 * dom);
  FqFrameParams p;
  p.in_s = dq.sf;
  p.in_z = dq.zf;
  p.degenerate = dq.degenerate;
  p.pad0 = 0;
  p.dmin = (int)(-128.0 - dq.z);                                     // == round(min / s)
  p.cscale = (float)((double)dq.sf * (double)(float)w_scale);
  // float-domain clamp bounds, evaluated in double like the Python expressions (:331-332, :432-433)
  const double hi_a = (double)((1 << (acc_bits - 1)) - 1), lo_a = -(double)(1 << (acc_bits - 1));
  const double hi_s = (double)((1 << (add_bits - 1)) - 1), lo_s = -(double)(1 << (add_bits - 1));
  p.hi_acc = (float)(((hi_a - dq.z) * dq.s) * w_scale);
  p.lo_acc = (float)(((lo_a - dq.z) * dq.s) * w_scale);
  p.hi_add = (float)(((hi_s - dq.z) * dq.s) * w_scale);
  p.lo_add = (float)(((lo_s - dq.z) * dq.s) * w_scale);
  // can a clamp fire at all for operands d in [dmin, dmin + 255]?  |P_p| <= up / dn (worst PE of the layer)
  const double dlo = (double)p.dmin, dhi = dlo + 255.0;
  const double up = fmax(dhi, 0.0) * wpos + fmax(-dlo, 0.0) * wneg;
  const double dn = fmax(dhi, 0.0) * wneg + fmax(-dlo, 0.0) * wpos;
  p.clamp_acc = !(up * (1.0 + 1e-6) + 2.0 < (hi_a - dq.z) && -(dn * (1.0 + 1e-6) + 2.0) > (lo_a - dq.z));
  p.clamp_add = !(4.0 * up * (1.0 + 1e-6) + 8.0 < (hi_s - dq.z) && -(4.0 * dn * (1.0 + 1e-6) + 8.0) > (lo_s - dq.z));
  const float bs = (float)(dq.s * w_scale);                            // :448
#pragma unroll
  for (int c = 0; c < SESRQ_MAX_CH; ++c) {
    float v = 0.f;
    if (c < cout) {
      const float bq = fminf(fmaxf(rintf(__fdiv_rn(bias[c], bs)), -32768.f), 32767.f);   // :402
      v = __fmul_rn(bq, bs);                                                             // :412
    }
    p.bias[c] = v;
  }
  out[f] = p;
}

// operand bytes b = q + 128 in the PE-major byte order (byte j <-> channel (j & 3) * 4 + (j >> 2)), q from THIS frame's
// scale / zero (:207).  NCHW input (layer 0) and NHWC16 input (layers 1..4).
__global__ void __launch_bounds__(256) fqb_quant_nchw_kernel(const float* __restrict__ x, int C, int HW,
                                                             const FqFrameParams* __restrict__ fp, uint4* __restrict__ act_all) {
  const int f = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float zf = fp[f].in_z;
  const ExactDiv div_s(fp[f].in_s);                       // x / s, exactly rounded (sesrq_common.cuh)
  const float* xb = x + (size_t)f * C * HW + p;
  uint4* act = act_all + (size_t)f * HW;
  const int i = p;
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int c = 0; c < SESRQ_MAX_CH; ++c) {
    if (c < C) {
      const int q = quant_add_zero(div_s(__ldg(xb + (long long)c * HW)), zf);
      const int j = (c & 3) * 4 + (c >> 2);
      w[j >> 2] |= (uint32_t)(q + 128) << (8 * (j & 3));
    }
  }
  act[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// NHWC16 fp32 -> operand bytes, 4 lanes per pixel: lane q reads the float4 of channels 4q .. 4q + 3 (consecutive lanes
// read consecutive 16 B), quantises them, and a 4 x 4 byte transpose across the four lanes (two shuffles) leaves lane q
// with word q of the pixel's uint4 -- PE-major byte order: word e, byte q  <->  channel 4q + e.
constexpr int FQ_QUANT_UNROLL = 4;             // float4 per thread (independent loads in flight, frame constants amortised)
__global__ void __launch_bounds__(256) fqb_quant_nhwc_kernel(const float4* __restrict__ x4, int HW,
                                                             const FqFrameParams* __restrict__ fp, uint32_t* __restrict__ act4) {
  const int f = blockIdx.y;                                                    // grid.y = frame: no per-thread division
  const unsigned t0 = blockIdx.x * (blockDim.x * FQ_QUANT_UNROLL) + threadIdx.x;   // float4 index; whole warps stay together
  const unsigned n4 = (unsigned)HW * 4u;
  const float4* x = x4 + (size_t)f * HW * 4;
  uint32_t* act = act4 + (size_t)f * HW * 4;
  const float zf = fp[f].in_z;
  const ExactDiv div_s(fp[f].in_s);
  float4 v[FQ_QUANT_UNROLL];
#pragma unroll
  for (int u = 0; u < FQ_QUANT_UNROLL; ++u) {
    const unsigned t = t0 + u * blockDim.x;
    v[u] = (t < n4) ? __ldg(x + t) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int q = threadIdx.x & 3;
  // ExactDiv's operand range: tested once per float4 (activations of a real frame are always inside); a value outside
  // or NaN sends the whole warp through the clamps
  bool ok = true;
#pragma unroll
  for (int u = 0; u < FQ_QUANT_UNROLL; ++u) ok = ok && div_s.in_range4(v[u]);
  if (__any_sync(0xffffffffu, !ok)) {
#pragma unroll
    for (int u = 0; u < FQ_QUANT_UNROLL; ++u) {
      v[u].x = fminf(fmaxf(v[u].x, -div_s.lim), div_s.lim); v[u].y = fminf(fmaxf(v[u].y, -div_s.lim), div_s.lim);
      v[u].z = fminf(fmaxf(v[u].z, -div_s.lim), div_s.lim); v[u].w = fminf(fmaxf(v[u].w, -div_s.lim), div_s.lim);
    }
  }
#pragma unroll
  for (int u = 0; u < FQ_QUANT_UNROLL; ++u) {
    const unsigned t = t0 + u * blockDim.x;
    const uint32_t b0 = (uint32_t)(quant_add_zero(div_s.unclamped(v[u].x), zf) + 128);
    const uint32_t b1 = (uint32_t)(quant_add_zero(div_s.unclamped(v[u].y), zf) + 128);
    const uint32_t b2 = (uint32_t)(quant_add_zero(div_s.unclamped(v[u].z), zf) + 128);
    const uint32_t b3 = (uint32_t)(quant_add_zero(div_s.unclamped(v[u].w), zf) + 128);
    uint32_t w = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);                     // byte e = this lane's contribution to word e
    // transpose: lane q ends with (lane0.byte q, lane1.byte q, lane2.byte q, lane3.byte q)
    uint32_t p1 = __shfl_xor_sync(0xffffffffu, w, 1);
    w = (q & 1) ? __byte_perm(p1, w, 0x7351) : __byte_perm(w, p1, 0x6240);    // odd: (p.b1, own.b1, p.b3, own.b3); even: (own.b0, p.b0, own.b2, p.b2)
    uint32_t p2 = __shfl_xor_sync(0xffffffffu, w, 2);
    w = (q & 2) ? __byte_perm(p2, w, 0x7632) : __byte_perm(w, p2, 0x5410);    // upper pair: (p.h1, own.h1); lower pair: (own.h0, p.h0)
    if (t < n4) act[t] = w;
  }
}

// 2048-bin histogram of the first `nch` channels of an NHWC16 fp32 tensor over a fixed range (KL pass 2).  Per-CTA
// shared bins; float4 loads; ReLU outputs pile up in bin 0, which would serialise the shared atomics, so bin 0 is
// counted in a register and added once per warp.
__global__ void __launch_bounds__(256) fqb_hist_nhwc_kernel(const float4* __restrict__ x, size_t n4, int nch,
                                                            const float* __restrict__ d_range,
                                                            unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_hist[SESRQ_KL_BINS];
  const float lo = d_range[0], hi = d_range[1];
  if (!(hi > lo)) return;
  for (int i = threadIdx.x; i < SESRQ_KL_BINS; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const float scale = __fdiv_rn((float)SESRQ_KL_BINS, __fsub_rn(hi, lo));
  unsigned int zeros = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i & 3) * 4;                       // first channel of this float4
    if (c0 >= nch) continue;
    const float4 v4 = __ldg(x + i);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c0 + e < nch) {
        const float c = fminf(fmaxf(v[e], lo), hi);
        int b = (int)floorf(__fmul_rn(__fsub_rn(c, lo), scale));
        b = min(max(b, 0), SESRQ_KL_BINS - 1);
        if (b == 0) ++zeros;
        else atomicAdd(s_hist + b, 1u);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
  if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(s_hist, zeros);
  __syncthreads();
  for (int i = threadIdx.x; i < SESRQ_KL_BINS; i += blockDim.x) {
    const unsigned int c = s_hist[i];
    if (c) atomicAdd(hist + i, (unsigned long long)c);
  }
}

// domain-5 fake-quant with the frame's range (test.py:90-91 maps PixelShuffle to the quantiser; :207-215) + PixelShuffle:
// X_5 fp32 NHWC16 -> y fp32 NCHW [B, C / r^2, rH, rW].  4 lanes per pixel (lane q: channels 4q .. 4q + 3, one coalesced
// float4 load); for r = 2 / 4 the lane's channels are 2 x 2 / 1 x 4 neighbouring output pixels -> vector stores.
__global__ void __launch_bounds__(256) fqb_final_kernel(const float4* __restrict__ x5, int C, int H, int W, int r,
                                                        int do_quant, const FqFrameParams* __restrict__ fp, float* __restrict__ y) {
  const int f = blockIdx.y;
  const int HW = H * W;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = (int)(t >> 2);
  const int q = (int)(t & 3);
  if (p >= HW || 4 * q >= C) return;
  const int yy = p / W, xx = p - yy * W;
  const float4 v4 = __ldg(x5 + (size_t)f * HW * 4 + t);
  float v[4] = {v4.x, v4.y, v4.z, v4.w};
  if (do_quant) {
    const float sf = fp[f].in_s, zf = fp[f].in_z;
    const ExactDiv div_s(sf);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float qv = (float)quant_add_zero(div_s(v[e]), zf);
      v[e] = __fmul_rn(__fsub_rn(qv, zf), sf);                    // :215
    }
  }
  float* yb = y + (size_t)f * C * HW;
  const size_t Wr = (size_t)W * r;
  if (r == 2) {                                                    // channels 4q .. 4q + 3 = plane q, rows 2yy, 2yy + 1, columns 2xx, 2xx + 1
    float* d = yb + ((size_t)q * (H * 2) + 2 * yy) * Wr + 2 * xx;
    *reinterpret_cast<float2*>(d) = make_float2(v[0], v[1]);
    *reinterpret_cast<float2*>(d + Wr) = make_float2(v[2], v[3]);
  } else if (r == 4) {                                             // plane q / 4, row 4yy + q % 4, columns 4xx .. 4xx + 3
    float* d = yb + ((size_t)(q >> 2) * (H * 4) + 4 * yy + (q & 3)) * Wr + 4 * xx;
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    const int r2 = r * r;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e;
      if (c < C) {
        const int cc = c / r2, ii = (c % r2) / r, jj = c % r;
        yb[((size_t)cc * (H * r) + (yy * r + ii)) * Wr + (xx * r + jj)] = v[e];
      }
    }
  }
}

// fold the per-frame ranges into the running observer (quan_func.py:188-195) and hand them out
__global__ void fqb_commit_kernel(const float* __restrict__ mm, int B, float* __restrict__ mm_run,
                                  float* __restrict__ mm_frames_out) {
  const int d = threadIdx.x;
  if (d >= SESRQ_NUM_DOMAINS) return;
  float lo = INFINITY, hi = -INFINITY;
  for (int f = 0; f < B; ++f) {
    const float a = mm[(size_t)f * MM_STRIDE + 2 * d], b = mm[(size_t)f * MM_STRIDE + 2 * d + 1];
    lo = fminf(lo, a);
    hi = fmaxf(hi, b);
    if (mm_frames_out) { mm_frames_out[(size_t)f * MM_STRIDE + 2 * d] = a; mm_frames_out[(size_t)f * MM_STRIDE + 2 * d + 1] = b; }
  }
  if (mm_run) {
    if (lo < mm_run[2 * d]) mm_run[2 * d] = lo;
    if (hi > mm_run[2 * d + 1]) mm_run[2 * d + 1] = hi;
  }
}

}  // namespace sesrq

using namespace sesrq;

struct SesrqFq {
  SesrqFqDesc desc;
  LayerDev layers[SESRQ_NUM_LAYERS];          // tcgen05 weight images (4 PE-masked sets)
  int32_t* d_tapsum[SESRQ_NUM_LAYERS] = {};   // [k*k][64] per-tap, per-PE weight sums in accumulator column order
  float* d_bias[SESRQ_NUM_LAYERS] = {};
  double wpos[SESRQ_NUM_LAYERS] = {}, wneg[SESRQ_NUM_LAYERS] = {};
  size_t chunk_bytes = (size_t)4096 << 20;
};

namespace {
struct FqPlan { int chunk; size_t off_mm, off_fp, off_act, off_x[3], total; };
FqPlan fq_plan(const SesrqFq* fq, int B, int H, int W) {
  FqPlan p{};
  const size_t hw = (size_t)H * W;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t per_frame = hw * (16 + 3 * 64) + MM_STRIDE * 4 + sizeof(FqFrameParams);
  long long c = (long long)(fq->chunk_bytes / std::max<size_t>(per_frame, 1));
  c = std::max<long long>(1, std::min<long long>(c, std::max(B, 1)));
  const long long passes = (std::max(B, 1) + c - 1) / c;
  p.chunk = (int)((std::max(B, 1) + passes - 1) / passes);
  size_t off = 0;
  p.off_mm = off; off += up((size_t)p.chunk * MM_STRIDE * 4);
  p.off_fp = off; off += up((size_t)p.chunk * sizeof(FqFrameParams));
  p.off_act = off; off += up((size_t)p.chunk * hw * 16);
  for (int i = 0; i < 3; ++i) { p.off_x[i] = off; off += up((size_t)p.chunk * hw * 64); }
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
  if (d->pe_acc_bits < 9 || d->pe_acc_bits > 22 || d->pe_add_bits < 9 || d->pe_add_bits > 30) return SESRQ_EINVAL;
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    if ((d->ksize[i] != 3 && d->ksize[i] != 5) || d->cin[i] < 1 || d->cin[i] > SESRQ_MAX_CH || d->cout[i] < 1 ||
        d->cout[i] > SESRQ_MAX_CH || !d->h_weight[i] || !d->h_bias[i])
      return SESRQ_EINVAL;
    if (i > 0 && d->cin[i] != d->cout[i - 1]) return SESRQ_EINVAL;
  }
  if (d->residual && d->cout[0] != d->cout[3]) return SESRQ_EINVAL;
  if (d->cout[4] % (d->r_shuffle * d->r_shuffle) != 0) return SESRQ_EINVAL;
  int dev = 0;
  SESRQ_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SESRQ_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return SESRQ_EARCH;
  SesrqFq* fq = new (std::nothrow) SesrqFq();
  if (!fq) return SESRQ_ENOMEM;
  fq->desc = *d;
  if (const char* e = getenv("SESRQ_FQ_CHUNK_MB")) {
    long v = atol(e);
    if (v > 0) fq->chunk_bytes = (size_t)v << 20;
  }
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    LayerDev& L = fq->layers[i];
    LayerConst& c = L.c;
    memset(&c, 0, sizeof(c));
    c.k = d->ksize[i]; c.pad = (c.k - 1) / 2; c.cin = d->cin[i]; c.cout = d->cout[i];
    c.epi = EPI_FQ;
    c.pe_exact = 1;             // four PE-masked accumulator sets: the float-domain PE sums are rounded one by one
    c.add_inert = 0;
    c.relu = d->relu[i];
    c.r_shuffle = 1;
    const int8_t* w = d->h_weight[i];
    const int taps = c.k * c.k;
    int rc = umma_prepare_layer(L, w);
    // per-tap, per-PE weight sums (operand-offset correction) and the worst per-PE positive / negative weight mass
    std::vector<int32_t> ts((size_t)taps * 64, 0);
    double wpos = 0, wneg = 0;
    for (int co = 0; co < c.cout; ++co) {
      double pp[4] = {0, 0, 0, 0}, nn[4] = {0, 0, 0, 0};
      for (int ci = 0; ci < c.cin; ++ci)
        for (int t = 0; t < taps; ++t) {
          const int v = w[((size_t)co * c.cin + ci) * taps + t];
          ts[(size_t)t * 64 + (ci & 3) * 16 + byte_of_chan(co)] += v;
          if (v > 0) pp[ci & 3] += v; else nn[ci & 3] -= v;
        }
      for (int p = 0; p < 4; ++p) { wpos = std::max(wpos, pp[p]); wneg = std::max(wneg, nn[p]); }
    }
    fq->wpos[i] = wpos;
    fq->wneg[i] = wneg;
    if (rc == SESRQ_OK &&
        (cudaMalloc(&fq->d_tapsum[i], ts.size() * 4) != cudaSuccess || cudaMalloc(&fq->d_bias[i], SESRQ_MAX_CH * 4) != cudaSuccess ||
         cudaMemcpy(fq->d_tapsum[i], ts.data(), ts.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
         cudaMemset(fq->d_bias[i], 0, SESRQ_MAX_CH * 4) != cudaSuccess ||
         cudaMemcpy(fq->d_bias[i], d->h_bias[i], c.cout * 4, cudaMemcpyHostToDevice) != cudaSuccess)) {
      set_cuda_error(cudaGetLastError(), "fq weight upload");
      rc = SESRQ_ECUDA;
    }
    if (rc != SESRQ_OK) {
      sesrq_fq_destroy(fq);
      return rc;
    }
  }
  *out = fq;
  return SESRQ_OK;
}

void sesrq_fq_destroy(SesrqFq* fq) {
  if (!fq) return;
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    if (fq->layers[i].d_wumma) cudaFree(fq->layers[i].d_wumma);
    if (fq->d_tapsum[i]) cudaFree(fq->d_tapsum[i]);
    if (fq->d_bias[i]) cudaFree(fq->d_bias[i]);
  }
  delete fq;
}

size_t sesrq_fq_ws_bytes_batch(const SesrqFq* fq, int B, int H, int W) {
  if (!fq || B < 0 || H <= 0 || W <= 0) return 0;
  return fq_plan(fq, B, H, W).total;
}

size_t sesrq_fq_ws_bytes(const SesrqFq* fq, int H, int W) { return sesrq_fq_ws_bytes_batch(fq, 1, H, W); }

int sesrq_fq_forward_batch(const SesrqFq* fq, const float* d_x, int B, int H, int W, float* d_y, float* d_minmax,
                           float* d_frame_minmax, unsigned long long* d_hist, const float* d_hist_range, void* d_ws,
                           size_t ws_bytes, void* stream) {
  if (!fq || B < 0 || H <= 0 || W <= 0) return SESRQ_EINVAL;
  if (B == 0) return SESRQ_OK;
  if (!d_x || !d_y || !d_ws) return SESRQ_EINVAL;
  if (d_hist && !d_hist_range) return SESRQ_EINVAL;
  if ((long long)H * W > (1LL << 27)) return SESRQ_EINVAL;
  const FqPlan p = fq_plan(fq, B, H, W);
  if (ws_bytes < p.total) return SESRQ_ENOMEM;
  if ((reinterpret_cast<uintptr_t>(d_ws) & 15) != 0) return SESRQ_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_ws);
  float* mm = reinterpret_cast<float*>(ws + p.off_mm);
  FqFrameParams* fp = reinterpret_cast<FqFrameParams*>(ws + p.off_fp);
  uint8_t* act = ws + p.off_act;
  float* X[3] = {reinterpret_cast<float*>(ws + p.off_x[0]), reinterpret_cast<float*>(ws + p.off_x[1]),
                 reinterpret_cast<float*>(ws + p.off_x[2])};
  const SesrqFqDesc& D = fq->desc;
  const int HW = H * W;
  const int cin = D.cin[0], cout = D.cout[4];
  const size_t y_frame = (size_t)cout * HW;                 // Cout / r^2 * rH * rW
  for (int b0 = 0; b0 < B; b0 += p.chunk) {
    const int nb = std::min(p.chunk, B - b0);
    const float* x = d_x + (size_t)b0 * cin * HW;
    const long long npix = (long long)nb * HW;
    fqb_init_mm_kernel<<<(nb * MM_STRIDE + 255) / 256, 256, 0, st>>>(mm, nb * MM_STRIDE);
    SESRQ_LAUNCH_CHECK("fqb_init_mm_kernel");
    {
      const size_t per_frame = (size_t)cin * HW;                                  // domain 0 (:177-178)
      dim3 g((unsigned)std::min<size_t>(std::max<size_t>(per_frame / 4096, 1), 64), (unsigned)nb);
      fqb_minmax_frames_kernel<<<g, 256, 0, st>>>(x, per_frame, mm);
      SESRQ_LAUNCH_CHECK("fqb_minmax_frames_kernel");
    }
    // tensor routing: X_1 = X[0] (kept for the AddOp), X_2 = X[1], X_3 = X[2], X_4 = X[1], X_5 = X[2]
    const float* in_t[5] = {nullptr, X[0], X[1], X[2], X[1]};
    float* out_t[5] = {X[0], X[1], X[2], X[1], X[2]};
    for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
      const LayerDev& L = fq->layers[i];
      fqb_params_kernel<<<(nb + 127) / 128, 128, 0, st>>>(mm, i, nb, D.w_scale[i], fq->d_bias[i], L.c.cout, D.pe_acc_bits,
                                                          D.pe_add_bits, fq->wpos[i], fq->wneg[i], fp);
      SESRQ_LAUNCH_CHECK("fqb_params_kernel");
      if (i == 0) {
        if (d_hist) {
          int rc = launch_hist_dev(x, (size_t)nb * cin * HW, d_hist_range, d_hist, st);
          if (rc) return rc;
        }
        fqb_quant_nchw_kernel<<<dim3((unsigned)((HW + 255) / 256), (unsigned)nb), 256, 0, st>>>(x, cin, HW, fp,
                                                                                                reinterpret_cast<uint4*>(act));
        SESRQ_LAUNCH_CHECK("fqb_quant_nchw_kernel");
      } else {
        if (d_hist) {
          const size_t n4 = (size_t)npix * 4;
          const unsigned hg = (unsigned)std::min<size_t>(std::max<size_t>((n4 + 2047) / 2048, 1), 148 * 8);
          fqb_hist_nhwc_kernel<<<hg, 256, 0, st>>>(reinterpret_cast<const float4*>(in_t[i]), n4, L.c.cin, d_hist_range + 2 * i,
                                                   d_hist + (size_t)i * SESRQ_KL_BINS);
          SESRQ_LAUNCH_CHECK("fqb_hist_nhwc_kernel");
        }
        fqb_quant_nhwc_kernel<<<dim3((unsigned)((HW * 4 + 256 * FQ_QUANT_UNROLL - 1) / (256 * FQ_QUANT_UNROLL)), (unsigned)nb), 256, 0, st>>>(
            reinterpret_cast<const float4*>(in_t[i]), HW, fp, reinterpret_cast<uint32_t*>(act));
        SESRQ_LAUNCH_CHECK("fqb_quant_nhwc_kernel");
      }
      FqArgs fa{};
      fa.fp = fp;
      fa.xout = out_t[i];
      fa.xres = (i == 3 && D.residual) ? X[0] : nullptr;
      fa.mm = mm + 2 * (i + 1);
      fa.tapsum = fq->d_tapsum[i];
      fa.mm_stride = MM_STRIDE;
      fa.relu = D.relu[i];
      fa.cout = L.c.cout;
      fa.one = 1.0f;
      int rc = launch_conv_umma_fq(L, act, fa, nb, H, W, st);
      if (rc) return rc;
    }
    if (d_hist) {
      const size_t n4 = (size_t)npix * 4;
      const unsigned hg = (unsigned)std::min<size_t>(std::max<size_t>((n4 + 2047) / 2048, 1), 148 * 8);
      fqb_hist_nhwc_kernel<<<hg, 256, 0, st>>>(reinterpret_cast<const float4*>(X[2]), n4, cout, d_hist_range + 10,
                                               d_hist + (size_t)5 * SESRQ_KL_BINS);
      SESRQ_LAUNCH_CHECK("fqb_hist_nhwc_kernel");
    }
    // SESR: the PixelShuffle input is fake-quantised with this frame's range (test.py:90-91);
    // NRDM (observe_out): no PixelShuffle node, the conv output is only observed (quan_func.py:460-479)
    fqb_params_kernel<<<(nb + 127) / 128, 128, 0, st>>>(mm, 5, nb, 1.0, fq->d_bias[4], 0, D.pe_acc_bits, D.pe_add_bits, 0.0, 0.0, fp);
    SESRQ_LAUNCH_CHECK("fqb_params_kernel");
    fqb_final_kernel<<<dim3((unsigned)((HW * 4 + 255) / 256), (unsigned)nb), 256, 0, st>>>(
        reinterpret_cast<const float4*>(X[2]), cout, H, W, D.r_shuffle, D.observe_out ? 0 : 1, fp, d_y + (size_t)b0 * y_frame);
    SESRQ_LAUNCH_CHECK("fqb_final_kernel");
    fqb_commit_kernel<<<1, 32, 0, st>>>(mm, nb, d_minmax, d_frame_minmax ? d_frame_minmax + (size_t)b0 * MM_STRIDE : nullptr);
    SESRQ_LAUNCH_CHECK("fqb_commit_kernel");
  }
  return SESRQ_OK;
}

int sesrq_fq_forward(const SesrqFq* fq, const float* d_x, int H, int W, float* d_y, float* d_minmax,
                     float* d_frame_minmax, unsigned long long* d_hist, const float* d_hist_range, void* d_ws,
                     size_t ws_bytes, void* stream) {
  return sesrq_fq_forward_batch(fq, d_x, 1, H, W, d_y, d_minmax, d_frame_minmax, d_hist, d_hist_range, d_ws, ws_bytes, stream);
}

}  // extern "C"
