// Internal host-side declarations shared by the translation units of libsesrq.
#pragma once
#include <atomic>
#include <cstdio>
#include "sesrq_common.cuh"

namespace sesrq {

extern std::atomic<uint64_t> g_launches;     // kernels launched by this library (sesrq_launch_count)
void set_cuda_error(cudaError_t e, const char* where);
const char* get_cuda_error();

#define SESRQ_CUDA(expr)                                   \
  do {                                                     \
    cudaError_t e__ = (expr);                              \
    if (e__ != cudaSuccess) {                              \
      ::sesrq::set_cuda_error(e__, #expr);                 \
      return SESRQ_ECUDA;                                  \
    }                                                      \
  } while (0)

#define SESRQ_LAUNCH_CHECK(name)                           \
  do {                                                     \
    ::sesrq::g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) {                              \
      ::sesrq::set_cuda_error(e__, name);                  \
      return SESRQ_ECUDA;                                  \
    }                                                      \
  } while (0)

// Device-side resources of one layer.
struct LayerDev {
  LayerConst c;
  int32_t* d_wpk = nullptr;      // generic path: [k*k][16 co][4 pe] packed s8x4 (PE-major bytes)
  uint8_t* d_wumma = nullptr;    // tcgen05 path: B operand image (see sesrq_umma.cu)
  int impl = SESRQ_IMPL_GENERIC;
  int umma_n = 16;               // accumulator columns per 128-pixel block (16 fast / 64 pe-exact)
  int layout = 0;                // SESRQ_LAYOUT_*: which accumulator layout the tcgen05 path chose (0 = generic path)
};

// ---- generic (CUDA-core) path: sesrq_generic.cu -------------------------------------------------
int launch_quant_input(const float* d_x, int B, int C, int H, int W, float s0, float z0, uint8_t* d_act,
                       cudaStream_t st);
int launch_conv_generic(const LayerDev& L, const uint8_t* d_in, const uint8_t* d_shortcut, uint8_t* d_out,
                        uint8_t* d_rc, int8_t* d_q5, float* d_y, int B, int H, int W, cudaStream_t st);
int launch_act_to_nchw(const uint8_t* d_act, int B, int C, int H, int W, int8_t* d_q, cudaStream_t st);

// ---- tcgen05 path: sesrq_umma.cu -----------------------------------------------------------------
bool umma_supported(const LayerConst& c);
int umma_prepare_layer(LayerDev& L, const int8_t* h_w);                 // builds d_wumma
int launch_conv_umma(const LayerDev& L, const float* d_x /*layer 0*/, const uint8_t* d_in,
                     const uint8_t* d_shortcut, uint8_t* d_out, int8_t* d_q5, float* d_y, int B, int H,
                     int W, cudaStream_t st, int rev = 0);   // rev: walk the tiles last to first

// ---- two 3x3 layers per launch: sesrq_chain.cu ---------------------------------------------------
bool chain2_supported(const LayerDev& A, const LayerDev& B);            // both plain 3x3, 16 -> 16 channels, tcgen05 layout NS = 1
int launch_conv_chain2(const LayerDev& A, const LayerDev& B, const uint8_t* d_in, uint8_t* d_out, int Bn, int H, int W,
                       cudaStream_t st, int rev = 0);

// exe_mode 0 conv on the tcgen05 path (4 PE-masked accumulator sets, float-domain epilogue): L.c.epi == EPI_FQ
int launch_conv_umma_fq(const LayerDev& L, const uint8_t* d_in, const FqArgs& fa, int B, int H, int W, cudaStream_t st);

}  // namespace sesrq
