// Generic conv path: CUDA-core dp4a, four per-PE accumulators per output, any zero point.
// Always exact; it is the fallback for configurations the tcgen05 kernel does not take (z' != -128)
// and the on-device cross-check for it.  One thread = one output pixel x all output channels.
#include "sesrq_internal.h"

namespace sesrq {

__device__ __forceinline__ int dp4a_u8s8(uint32_t a, int32_t b, int32_t c) {
  int32_t d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// ---- layer-0 input quantiser: fp32 NCHW -> NHWC16 bytes (quan_func.py:222-225) -------------------
__global__ void __launch_bounds__(256) quant_input_kernel(const float* __restrict__ x, int C, int HW,
                                                          long long npix, float s0, float z0,
                                                          uint4* __restrict__ act) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  long long b = i / HW;
  int p = (int)(i - b * HW);
  const float* xb = x + b * (long long)C * HW + p;
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int c = 0; c < SESRQ_MAX_CH; ++c) {
    if (c < C) {
      int q = quant_input(__ldg(xb + (long long)c * HW), s0, z0);
      const int j = (c & 3) * 4 + (c >> 2);   // byte_of_chan(c)
      w[j >> 2] |= (uint32_t)(q + 128) << (8 * (j & 3));
    }
  }
  act[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

int launch_quant_input(const float* d_x, int B, int C, int H, int W, float s0, float z0, uint8_t* d_act,
                       cudaStream_t st) {
  long long npix = (long long)B * H * W;
  if (npix == 0) return SESRQ_OK;
  unsigned grid = (unsigned)((npix + 255) / 256);
  quant_input_kernel<<<grid, 256, 0, st>>>(d_x, C, H * W, npix, s0, z0, reinterpret_cast<uint4*>(d_act));
  SESRQ_LAUNCH_CHECK("quant_input_kernel");
  return SESRQ_OK;
}

// ---- conv + PE tree + requant epilogue ------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128)
conv_generic_kernel(const __grid_constant__ LayerConst L, const uint4* __restrict__ in,
                    const int4* __restrict__ wpk, const uint4* __restrict__ shortcut,
                    uint4* __restrict__ out, uint4* __restrict__ out_rc, int8_t* __restrict__ q5,
                    float* __restrict__ y, int B, int H, int W) {
  constexpr int TAPS = K * K;
  constexpr int PAD = (K - 1) / 2;
  __shared__ int4 w_s[TAPS * 16];
  for (int i = threadIdx.x; i < TAPS * 16; i += blockDim.x) w_s[i] = wpk[i];
  __syncthreads();

  long long npix = (long long)B * H * W;
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  int xx = (int)(pix % W);
  long long t = pix / W;
  int yy = (int)(t % H);
  long long b = t / H;
  const uint4* frame = in + b * (long long)H * W;

  int acc[16][4];
#pragma unroll
  for (int co = 0; co < 16; ++co)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[co][p] = 0;

  const uint32_t padw = (uint32_t)L.pad_byte * 0x01010101u;
#pragma unroll 1
  for (int dy = 0; dy < K; ++dy) {
    int sy = yy + dy - PAD;
    bool rowok = (sy >= 0) && (sy < H);
#pragma unroll 1
    for (int dx = 0; dx < K; ++dx) {
      int sx = xx + dx - PAD;
      uint4 v = make_uint4(padw, padw, padw, padw);
      if (rowok && sx >= 0 && sx < W) v = __ldg(frame + (long long)sy * W + sx);
      const int4* wt = w_s + (dy * K + dx) * 16;
#pragma unroll
      for (int co = 0; co < 16; ++co) {
        int4 w = wt[co];
        acc[co][0] = dp4a_u8s8(v.x, w.x, acc[co][0]);
        acc[co][1] = dp4a_u8s8(v.y, w.y, acc[co][1]);
        acc[co][2] = dp4a_u8s8(v.z, w.z, acc[co][2]);
        acc[co][3] = dp4a_u8s8(v.w, w.w, acc[co][3]);
      }
    }
  }

  // epilogue
  uint32_t ow[4] = {0u, 0u, 0u, 0u};
  uint32_t rw[4] = {0u, 0u, 0u, 0u};
  uint4 sc = make_uint4(0u, 0u, 0u, 0u);
  if (L.epi == EPI_RESIDUAL) sc = __ldg(shortcut + pix);
  const uint32_t scw[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int co = (j & 3) * 4 + (j >> 2);   // chan_of_byte(j)
    if (co >= L.cout) continue;
    int A = pe_tree(acc[co], L, co);
    float yv = requant(A, L.mf);
    if (L.epi == EPI_FINAL) {
      int q = quant_add_zero(yv, L.z_next);                         // :585 (no activation before it)
      long long hw = (long long)H * W;
      if (q5) q5[(b * L.cout + co) * hw + (long long)yy * W + xx] = (int8_t)q;
      if (y) {
        float o = __fmul_rn(__fsub_rn((float)q, L.z_next), L.s_out);   // :594
        int r = L.r_shuffle, r2 = r * r;
        int cc = co / r2, ii = (co % r2) / r, jj = co % r;
        int Co = L.cout / r2;
        y[((b * Co + cc) * (long long)(H * r) + (yy * r + ii)) * (long long)(W * r) + (xx * r + jj)] = o;
      }
    } else {
      float a = L.relu ? fmaxf(yv, 0.f) : yv;
      int v;
      if (L.epi == EPI_RESIDUAL) {
        int ic = quant_add_zero(a, -128.f);                          // :251
        int vrc = (scw[j >> 2] >> (8 * (j & 3))) & 0xff;             // rc + 128   (:250)
        float tt = __fmul_rn((float)(vrc + ic + 128), L.res_mf);     // (rc + ic + 256) * Mr * 2^-nr (:252,:269)
        v = quant_add_zero(tt, L.z_res) + 128;                       // :270
      } else {
        v = quant_add_zero(a, L.z_next) + 128;                       // :280
        if (L.save_rc) rw[j >> 2] |= (uint32_t)(quant_add_zero(a, -128.f) + 128) << (8 * (j & 3));
      }
      ow[j >> 2] |= (uint32_t)v << (8 * (j & 3));
    }
  }
  if (L.epi != EPI_FINAL) {
    out[pix] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    if (L.save_rc && out_rc) out_rc[pix] = make_uint4(rw[0], rw[1], rw[2], rw[3]);
  }
}

int launch_conv_generic(const LayerDev& L, const uint8_t* d_in, const uint8_t* d_shortcut, uint8_t* d_out,
                        uint8_t* d_rc, int8_t* d_q5, float* d_y, int B, int H, int W, cudaStream_t st) {
  long long npix = (long long)B * H * W;
  if (npix == 0) return SESRQ_OK;
  unsigned grid = (unsigned)((npix + 127) / 128);
  auto in = reinterpret_cast<const uint4*>(d_in);
  auto sc = reinterpret_cast<const uint4*>(d_shortcut);
  auto out = reinterpret_cast<uint4*>(d_out);
  auto rc = reinterpret_cast<uint4*>(d_rc);
  auto w = reinterpret_cast<const int4*>(L.d_wpk);
  if (L.c.k == 3)
    conv_generic_kernel<3><<<grid, 128, 0, st>>>(L.c, in, w, sc, out, rc, d_q5, d_y, B, H, W);
  else if (L.c.k == 5)
    conv_generic_kernel<5><<<grid, 128, 0, st>>>(L.c, in, w, sc, out, rc, d_q5, d_y, B, H, W);
  else
    return SESRQ_EINVAL;
  SESRQ_LAUNCH_CHECK("conv_generic_kernel");
  return SESRQ_OK;
}

// ---- NHWC16 bytes -> int8 NCHW codes (layer-by-layer parity reads) -------------------------------
__global__ void __launch_bounds__(256) act_to_nchw_kernel(const uint4* __restrict__ act, int C, int HW,
                                                          long long npix, int8_t* __restrict__ q) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  uint4 v = __ldg(act + i);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  long long b = i / HW;
  int p = (int)(i - b * HW);
  for (int c = 0; c < C; ++c) {
    int j = byte_of_chan(c);
    int val = (int)((w[j >> 2] >> (8 * (j & 3))) & 0xff) - 128;
    q[(b * C + c) * (long long)HW + p] = (int8_t)val;
  }
}

int launch_act_to_nchw(const uint8_t* d_act, int B, int C, int H, int W, int8_t* d_q, cudaStream_t st) {
  long long npix = (long long)B * H * W;
  if (npix == 0) return SESRQ_OK;
  unsigned grid = (unsigned)((npix + 255) / 256);
  act_to_nchw_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(d_act), C, H * W, npix, d_q);
  SESRQ_LAUNCH_CHECK("act_to_nchw_kernel");
  return SESRQ_OK;
}

}  // namespace sesrq
