// Activation calibration kernels: running min/max, 2048-bin histogram, on-device KL threshold search.
//   min/max  : the observer inside quantize_asymmetrical_by_tensor exe_mode 0 (quan_func.py:175-196)
//              and myMinMaxObserver (quan_modules.py:14-40)
//   histogram + KL : the reference abandoned KL calibration (define.py:16-20, BINS_NUM=2048,
//              TGT_BINS_NUM=128 survive as comments); the search implemented here is the published NVIDIA
//              pytorch-quantization / TensorRT entropy calibrator, restated on the CPU in oracle/kl_oracle.py.
#include <algorithm>
#include "sesrq_internal.h"

namespace sesrq {

// ---- running min / max ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, size_t n, float* __restrict__ mm) {
  float lo = INFINITY, hi = -INFINITY;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  size_t n4 = aligned ? n / 4 : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (size_t i = tid; i < n4; i += nth) {
    float4 v = __ldg(x4 + i);
    lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
    hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  for (size_t i = n4 * 4 + tid; i < n; i += nth) {
    float v = __ldg(x + i);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[8], s_hi[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_lo[w] = lo; s_hi[w] = hi; }
  __syncthreads();
  if (w == 0) {
    lo = l < 8 ? s_lo[l] : INFINITY;
    hi = l < 8 ? s_hi[l] : -INFINITY;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (l == 0) {
      atomic_min_float(mm + 0, lo);
      atomic_max_float(mm + 1, hi);
    }
  }
}

// ---- histogram: per-CTA shared-memory bins, warp-aggregated atomics ---------------------------------
__device__ __forceinline__ int hist_bin(float v, float lo, float hi, float scale, int nbins) {
  float c = fminf(fmaxf(v, lo), hi);
  int b = (int)floorf(__fmul_rn(__fsub_rn(c, lo), scale));
  return min(max(b, 0), nbins - 1);
}

__device__ __forceinline__ void hist_add_aggregated(unsigned int* s_hist, int bin, bool valid) {
  // lanes of the warp that hit the same bin elect a leader that adds their count once: ReLU outputs
  // pile up in bin 0, which would otherwise serialise 32-way on one shared-memory word.
  unsigned active = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  unsigned peers = __match_any_sync(active, bin);
  if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(s_hist + bin, __popc(peers));
}

__global__ void __launch_bounds__(256) hist_kernel(const float* __restrict__ x, size_t n, float lo, float hi,
                                                   const float* __restrict__ d_range,
                                                   unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_hist[SESRQ_KL_BINS];
  if (d_range) { lo = d_range[0]; hi = d_range[1]; }
  if (!(hi > lo)) return;
  for (int i = threadIdx.x; i < SESRQ_KL_BINS; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const float scale = __fdiv_rn((float)SESRQ_KL_BINS, __fsub_rn(hi, lo));
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  // whole warps iterate together so the aggregated atomics see a full mask
  const size_t n_round = ((n + 31) / 32) * 32;
  for (size_t i = tid; i < n_round; i += nth) {
    bool valid = i < n;
    float v = valid ? __ldg(x + i) : 0.f;
    hist_add_aggregated(s_hist, hist_bin(v, lo, hi, scale, SESRQ_KL_BINS), valid);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SESRQ_KL_BINS; i += blockDim.x) {
    unsigned int c = s_hist[i];
    if (c) atomicAdd(hist + i, (unsigned long long)c);
  }
}

// ---- KL threshold search ----------------------------------------------------------------------------
// The published entropy calibrator (NVIDIA pytorch-quantization `_compute_amax_entropy` = TensorRT entropy calibration
// 2; restated in oracle/kl_oracle.py), sized by the reference's surviving constants (define.py:19-20).
// One CTA per candidate bin count i in [tgt, nbins].  h[0] := h[1]; P = h[0:i] with the outlier mass folded into
// bin i-1; Q = h[0:i] merged into `tgt` groups -- source bin k belongs to group floor(k * tgt / i), i.e. group g covers
// [ceil(g i / tgt), ceil((g + 1) i / tgt)) -- and spread uniformly over the non-empty bins of each group; KL(P||Q) in
// fp64.  Every CTA builds the two prefix arrays in shared memory with a block scan (8 bins per thread).
constexpr int KL_THREADS = 256;

__global__ void __launch_bounds__(KL_THREADS) kl_kernel(const unsigned long long* __restrict__ hist, int nbins,
                                                        int tgt, double* __restrict__ kl_out) {
  extern __shared__ unsigned long long s_mem[];
  unsigned long long* s_pref = s_mem;                       // nbins + 1 prefix sums of counts
  unsigned int* s_nz = reinterpret_cast<unsigned int*>(s_pref + nbins + 1);   // nbins + 1 prefix of non-empty flags
  __shared__ double s_red[KL_THREADS / 32];
  __shared__ unsigned long long s_wsum[KL_THREADS / 32];
  __shared__ unsigned int s_wnz[KL_THREADS / 32];
  const int i = tgt + blockIdx.x;
  // ---- block scan: thread t owns bins [t * per, (t + 1) * per) ----
  {
    const int per = (nbins + KL_THREADS - 1) / KL_THREADS;
    const int k0 = threadIdx.x * per;
    unsigned long long a = 0;
    unsigned int z = 0;
    for (int k = k0; k < min(k0 + per, nbins); ++k) {
      const unsigned long long h = hist[(k == 0 && nbins > 1) ? 1 : k];      // h[0] := h[1]
      a += h;
      z += (h != 0);
    }
    unsigned long long ia = a;                              // inclusive scan over the warp
    unsigned int iz = z;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long ta = __shfl_up_sync(0xffffffffu, ia, o);
      const unsigned int tz = __shfl_up_sync(0xffffffffu, iz, o);
      if ((threadIdx.x & 31) >= o) { ia += ta; iz += tz; }
    }
    if ((threadIdx.x & 31) == 31) { s_wsum[threadIdx.x >> 5] = ia; s_wnz[threadIdx.x >> 5] = iz; }
    __syncthreads();
    unsigned long long base = ia - a;
    unsigned int basez = iz - z;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) { base += s_wsum[w]; basez += s_wnz[w]; }
    if (threadIdx.x == 0) { s_pref[0] = 0; s_nz[0] = 0; }
    for (int k = k0; k < min(k0 + per, nbins); ++k) {
      const unsigned long long h = hist[(k == 0 && nbins > 1) ? 1 : k];
      base += h;
      basez += (h != 0);
      s_pref[k + 1] = base;
      s_nz[k + 1] = basez;
    }
  }
  __syncthreads();
  const unsigned long long total = s_pref[nbins];
  const unsigned long long inside = s_pref[i];
  const unsigned long long outliers = total - inside;
  double acc = 0.0;
  bool infinite = false;
  if (total != 0 && inside != 0) {
    const double inv_total = 1.0 / (double)total;
    const double inv_inside = 1.0 / (double)inside;
    for (int k = threadIdx.x; k < i; k += KL_THREADS) {
      unsigned long long h = s_pref[k + 1] - s_pref[k];
      unsigned long long pk = h + ((k == i - 1) ? outliers : 0ull);
      if (pk == 0) continue;
      if (h == 0) { infinite = true; continue; }            // P > 0 where Q == 0
      const long long g = ((long long)k * tgt) / i;
      const int start = (int)((g * i + tgt - 1) / tgt);
      const int stop = (int)(((g + 1) * i + tgt - 1) / tgt);
      unsigned long long gs = s_pref[stop] - s_pref[start];
      unsigned int gn = s_nz[stop] - s_nz[start];
      double p = (double)pk * inv_total;
      double q = ((double)gs / (double)gn) * inv_inside;
      acc += p * log(p / q);
    }
  } else {
    infinite = true;
  }
  // deterministic tree reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  int any_inf = __syncthreads_or(infinite ? 1 : 0);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < KL_THREADS / 32; ++w) t += s_red[w];
    kl_out[i] = any_inf ? INFINITY : t;
  }
}

// the published tie-break: the LAST (largest i) minimum; all-infinite -> keep the full range
__global__ void __launch_bounds__(256) kl_argmin_kernel(const double* __restrict__ kl, int nbins, int tgt,
                                                        int32_t* __restrict__ idx) {
  __shared__ double s_v[256];
  __shared__ int s_i[256];
  double best = INFINITY;
  int bi = nbins;
  for (int i = tgt + threadIdx.x; i <= nbins; i += blockDim.x) {
    double v = kl[i];
    if (v < best || (v == best && v != INFINITY)) { best = v; bi = i; }   // ascending scan: ties keep the larger i
  }
  s_v[threadIdx.x] = best;
  s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      double v = s_v[threadIdx.x + o];
      int j = s_i[threadIdx.x + o];
      if (v < s_v[threadIdx.x] || (v == s_v[threadIdx.x] && v != INFINITY && j > s_i[threadIdx.x])) {
        s_v[threadIdx.x] = v;
        s_i[threadIdx.x] = j;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) idx[0] = s_i[0];
}

int launch_hist_dev(const float* d_x, size_t n, const float* d_range, unsigned long long* d_hist, cudaStream_t st) {
  if (n == 0) return SESRQ_OK;
  size_t blocks = (n + 256 * 16 - 1) / (256 * 16);
  unsigned grid = (unsigned)std::min<size_t>(std::max<size_t>(blocks, 1), 148 * 4);
  hist_kernel<<<grid, 256, 0, st>>>(d_x, n, 0.f, 0.f, d_range, d_hist);
  SESRQ_LAUNCH_CHECK("hist_kernel");
  return SESRQ_OK;
}

}  // namespace sesrq

using namespace sesrq;

extern "C" {

int sesrq_minmax_update(const float* d_x, size_t n, float* d_mm, void* stream) {
  if (!d_mm || (!d_x && n)) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  size_t blocks = (n / 4 + 255) / 256;
  unsigned grid = (unsigned)std::min<size_t>(std::max<size_t>(blocks, 1), 148 * 8);
  minmax_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_x, n, d_mm);
  SESRQ_LAUNCH_CHECK("minmax_kernel");
  return SESRQ_OK;
}

int sesrq_hist_update(const float* d_x, size_t n, float lo, float hi, unsigned long long* d_hist, void* stream) {
  if (!d_hist || (!d_x && n) || !(hi > lo)) return SESRQ_EINVAL;
  if (n == 0) return SESRQ_OK;
  size_t blocks = (n + 256 * 16 - 1) / (256 * 16);
  unsigned grid = (unsigned)std::min<size_t>(std::max<size_t>(blocks, 1), 148 * 4);
  hist_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_x, n, lo, hi, nullptr, d_hist);
  SESRQ_LAUNCH_CHECK("hist_kernel");
  return SESRQ_OK;
}

int sesrq_kl_threshold(const unsigned long long* d_hist, int nbins, int tgt_bins, int32_t* d_idx, double* d_kl,
                       void* stream) {
  if (!d_hist || !d_idx || !d_kl || tgt_bins < 1 || nbins < tgt_bins || nbins > 8192) return SESRQ_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  size_t smem = (size_t)(nbins + 1) * (sizeof(unsigned long long) + sizeof(unsigned int));
  kl_kernel<<<nbins - tgt_bins + 1, KL_THREADS, smem, st>>>(d_hist, nbins, tgt_bins, d_kl);
  SESRQ_LAUNCH_CHECK("kl_kernel");
  kl_argmin_kernel<<<1, 256, 0, st>>>(d_kl, nbins, tgt_bins, d_idx);
  SESRQ_LAUNCH_CHECK("kl_argmin_kernel");
  return SESRQ_OK;
}

}  // extern "C"
