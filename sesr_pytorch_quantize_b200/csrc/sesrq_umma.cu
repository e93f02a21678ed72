// tcgen05 implicit-GEMM conv path (placeholder until the kernel lands in this round).
#include "sesrq_internal.h"
namespace sesrq {
bool umma_supported(const LayerConst&) { return false; }
int umma_prepare_layer(LayerDev&, const int8_t*) { return SESRQ_EUNSUPPORTED; }
int launch_conv_umma(const LayerDev&, const float*, const uint8_t*, const uint8_t*, uint8_t*, int8_t*, float*,
                     int, int, int, cudaStream_t) { return SESRQ_EUNSUPPORTED; }
}
