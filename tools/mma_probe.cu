// Kernel-developer micro-benchmark (not part of the product): cycles per tcgen05.mma.kind::i8 (M = 128, K = 32,
// no-swizzle K-major operands in shared memory) as a function of N, of how the accumulator address D moves between
// consecutive instructions, and of how often tcgen05.commit is issued.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_probe tools/mma_probe.cu && /tmp/mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(int n_cols, int d_step, int d_wrap, int iters, int per_commit, int a_step,
                                                int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1));
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
  if (threadIdx.x == 0) {
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t b0 = ((smem_u32(smem) + 48 * 1024) & 0x3FFFFu) >> 4;
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t d_off = 0, a_off = 0;
    const int inner = per_commit > 0 ? per_commit : iters;
    const int outer = iters / inner;
    const uint32_t bar_a = smem_u32(&bar);
    long long t0 = clock64();
    for (int o = 0; o < outer; ++o) {
      for (int i = 0; i < inner; ++i) {
        const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)((a0 + a_off) | (1u << 16));
        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b0 | ((uint32_t)n_cols << 16));
        if (!(mode & 1))
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
              ::"r"(tmem_base + d_off), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        d_off += d_step;
        if (d_off + n_cols > (uint32_t)d_wrap) d_off = 0;
        a_off += a_step;
        if (a_off > 2048) a_off = 0;
      }
      if (per_commit > 0 && !(mode & 2))
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    }
    long long t1 = clock64();
    // drain: one final commit to a barrier nobody else arrives on (intermediate commits are never waited for)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  struct Cfg { int n, dstep, dwrap, per_commit, a_step, mode; const char* what; };
  const Cfg cfgs[] = {
      {16, 0, 512, 0, 0, 1, "loop only (no mma)"},
      {16, 0, 512, 4, 0, 1, "loop + commit/4 (no mma)"},
      {16, 0, 512, 1, 0, 1, "loop + commit/1 (no mma)"},
      {16, 0, 512, 0, 0, 0, "N=16 same D"},
      {16, 16, 512, 0, 0, 0, "N=16 step 16"},
      {48, 16, 512, 0, 0, 0, "N=48 shifted by 16"},
      {48, 16, 512, 0, 8, 0, "N=48 shifted, A moves"},
      {48, 16, 512, 1, 8, 0, "N=48 shifted, commit/1"},
      {48, 16, 512, 2, 8, 0, "N=48 shifted, commit/2"},
      {48, 16, 512, 4, 8, 0, "N=48 shifted, commit/4"},
      {48, 16, 512, 8, 8, 0, "N=48 shifted, commit/8"},
      {48, 16, 512, 20, 8, 0, "N=48 shifted, commit/20"},
      {80, 16, 512, 10, 8, 0, "N=80 shifted, commit/10"},
      {160, 32, 512, 6, 8, 0, "N=160 shifted 32, commit/6"},
      {128, 0, 512, 0, 0, 0, "N=128 same D"},
      {256, 0, 512, 0, 0, 0, "N=256 same D"},
  };
  for (const Cfg& c : cfgs) {
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<1, 128, 64 * 1024>>>(c.n, c.dstep, c.dwrap, iters, c.per_commit, c.a_step, c.mode, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.what, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    }
    printf("%-36s issue %.1f clk/mma   complete %.1f clk/mma\n", c.what, (double)h[0] / iters, (double)h[1] / iters);
    fflush(stdout);
  }
  return 0;
}
