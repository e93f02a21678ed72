// Kernel-developer micro-benchmark #2 (not part of the product): tensor-pipe throughput of tcgen05.mma.kind::i8
// (M = 128, K = 32) when W warps issue concurrently with a minimal (unrolled, constant-descriptor) instruction stream.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define MMA(d, ad, bd, idesc) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory")
__global__ void __launch_bounds__(256, 1) probe(int n_cols, int nwarps, int iters, int shift, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const int warp = threadIdx.x >> 5;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps && (threadIdx.x & 31) == 0) {
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a0 = ((smem_u32(smem) + warp * 8192) & 0x3FFFFu) >> 4;
    const uint32_t b0 = ((smem_u32(smem) + 80 * 1024) & 0x3FFFFu) >> 4;
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a0 | (1u << 16));
    const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b0 | ((uint32_t)n_cols << 16));
    const uint32_t d0 = tmem_base + warp * 128;           // disjoint 128-column regions per warp
    const uint32_t d1 = d0 + shift;
    t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
      MMA(d0, ad, bd, idesc); MMA(d1, ad, bd, idesc); MMA(d0, ad, bd, idesc); MMA(d1, ad, bd, idesc);
      MMA(d0, ad, bd, idesc); MMA(d1, ad, bd, idesc); MMA(d0, ad, bd, idesc); MMA(d1, ad, bd, idesc);
    }
    t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(0u) : "memory");
    long long t2 = clock64();
    out[2 * warp] = t1 - t0;
    out[2 * warp + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}
int main() {
  long long* d_out;
  cudaMalloc(&d_out, 128);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int iters = 4000;
  const int ns[] = {16, 32, 48, 64, 80, 96, 128};
  for (int shift = 0; shift <= 16; shift += 16)
  for (int n : ns)
    for (int w = 1; w <= 4; ++w) {
      if (n + shift > 128) continue;
      long long h[16] = {0};
      for (int rep = 0; rep < 2; ++rep) {
        probe<<<1, 256, 96 * 1024>>>(n, w, iters, shift, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d W=%d: %s\n", n, w, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d_out, 128, cudaMemcpyDeviceToHost);
      }
      long long mx = 0, mi = 0;
      for (int i = 0; i < w; ++i) { if (h[2 * i + 1] > mx) mx = h[2 * i + 1]; if (h[2 * i] > mi) mi = h[2 * i]; }
      printf("N=%3d shift=%2d warps=%d  issue %.1f clk/mma/warp  pipe %.1f clk/mma (all warps)\n", n, shift, w, (double)mi / iters,
             (double)mx / (iters * w));
      fflush(stdout);
    }
  return 0;
}
