/* sesrq.h -- C ABI of libsesrq.so: the B200-native (sm_100a) SESR / NRDM int8 hardware-sim hot path.
 *
 * The reference (gui-yupeng/sesr-pytorch-quantize) is 100 % Python and has no FFI; its boundary is
 * the set of Python callables that torch.fx inserts around every nn.Conv2d (myQL/graph_modify.py:33-154,
 * sim.py:82-114).  Each entry point below names the reference interface it replaces.  The Python
 * mirror of that interface (package sesr_pytorch_quantize_b200: define.py, myQL/, models/) binds these
 * symbols with ctypes -- see INTEGRATION.md for the stub.
 *
 * Conventions: plain pointers and sizes only; every `d_` pointer is DEVICE memory owned by the caller,
 * every `h_` pointer HOST memory; work is enqueued on `stream` (a cudaStream_t passed as void*; NULL =
 * the legacy default stream) with no hidden synchronisation unless stated; functions return 0 or a
 * negative SESRQ_E* code (sesrq_strerror).  A SesrqNet is immutable after creation, so one handle may
 * be used from several host threads on different streams with different workspaces.
 */
#ifndef SESRQ_H_
#define SESRQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SESRQ_ABI_VERSION 1
#define SESRQ_NUM_LAYERS 5      /* conv_first, residual_block.{0,1,2}, conv_last (models/sesr_sim.py:14-28) */
#define SESRQ_MAX_CH 16
#define SESRQ_NUM_DOMAINS 6     /* quantisation domains input.{0..5} (test.py:186-217) */
#define SESRQ_KL_BINS 2048      /* define.py:19 (commented remnant BINS_NUM) */
#define SESRQ_KL_TARGET_BINS 128/* define.py:20 (TGT_BINS_NUM) */

enum {
  SESRQ_OK = 0,
  SESRQ_EINVAL = -1,     /* bad argument / unsupported shape */
  SESRQ_ENOMEM = -2,     /* workspace too small or allocation failure */
  SESRQ_ECUDA = -3,      /* CUDA runtime error (see sesrq_last_cuda_error) */
  SESRQ_EARCH = -4,      /* device is not sm_100 (B200) */
  SESRQ_EUNSUPPORTED = -5
};

/* kernel selection for the conv stack */
enum {
  SESRQ_IMPL_AUTO = 0,    /* tcgen05 where its preconditions hold, else generic */
  SESRQ_IMPL_GENERIC = 1, /* CUDA-core dp4a, 4 per-PE accumulators, any zero point (always exact) */
  SESRQ_IMPL_UMMA = 2     /* tcgen05.mma kind::i8 implicit GEMM + TMA halo tiles */
};

/* One collapsed conv layer after the host-side "compile" step.
 * Replaces the per-layer state the reference keeps in ./output_pt (quan_func.py:77-78,408,489,536-537). */
typedef struct SesrqLayerDesc {
  int32_t ksize;                      /* 3 or 5, stride 1, pad (k-1)/2 (models/model_utils_pt.py:9-10) */
  int32_t cin;                        /* 1..16 */
  int32_t cout;                       /* 1..16 */
  int32_t relu;                       /* activation after requant (ReLU for layers 0-3, Identity for 4) */
  int32_t z_in;                       /* input zero point z_i, unclamped (quan_func.py:420) */
  int32_t m16;                        /* requant multiplier M16 (quan_func.py:495-515) */
  int32_t shift_n;                    /* requant shift n: y = acc * M16 * 2^-n */
  int32_t reserved0;
  const int8_t* h_weight;             /* HOST [cout][cin][k][k] int8 codes (conv.weight.{i}.pt) */
  int32_t add_const[SESRQ_MAX_CH];    /* clamp16(bias_q - sum(W)*z_i) (conv.bias.quan{i}.pt, :481-489) */
} SesrqLayerDesc;

/* Whole network.  Topology is the reference's: layer 0 output is the shortcut, re-added in the
 * input quantiser of layer 4 (quan_func.py:244-270), layer 3 requantises into domain 1 (:554-560). */
typedef struct SesrqNetDesc {
  int32_t abi_version;                /* SESRQ_ABI_VERSION */
  int32_t n_layers;                   /* must be SESRQ_NUM_LAYERS */
  int32_t r_shuffle;                  /* PixelShuffle factor: 4 (x4), 2 (x2), 1 (NRDM: none) */
  int32_t pe_num;                     /* define.PE, must be 4 */
  int32_t pe_acc_bits;                /* define.PE_ACC_BIT (18): per-PE accumulator clamp */
  int32_t pe_add_bits;                /* define.PE_ADD_BIT (20): PE adder-tree clamp */
  int32_t quan_bits;                  /* define.QUAN_BIT, must be 8 */
  int32_t impl;                       /* SESRQ_IMPL_* */
  double scale_in;                    /* s_0  (input.0.scale.pt) */
  double scale_out;                   /* s_5  (input.5.scale.pt) */
  int32_t z_out;                      /* z_5  (input.5.zero.pt) */
  int32_t res_m16;                    /* residual add requant M16 = qbc(s_1/s_4) (:259-260) */
  int32_t res_shift_n;
  int32_t reserved0;
  SesrqLayerDesc layers[SESRQ_NUM_LAYERS];
} SesrqNetDesc;

typedef struct SesrqNet SesrqNet;

/* ---- library ------------------------------------------------------------------------------- */
int sesrq_abi_version(void);
const char* sesrq_strerror(int code);
const char* sesrq_last_cuda_error(void);         /* text of the last CUDA failure on this thread */
/* Number of kernels this library has launched since load (all streams, this process); bench.py
 * reports the delta over the timed region as "gpu_launches". */
uint64_t sesrq_launch_count(void);

/* ---- network: replaces GraphModule.forward over the 5 x (quantize_asymmetrical_by_tensor ->
 *      reshape_input_for_hardware_pe -> Conv2d -> PEs_and_bias_adder -> requan_conv2d_output ->
 *      activation) chain + PixelShuffle built by sim.py:82-114 and run at sim.py:205 ------------ */
int sesrq_net_create(const SesrqNetDesc* desc, SesrqNet** out_net);
void sesrq_net_destroy(SesrqNet* net);
/* which kernel family each layer resolved to (SESRQ_IMPL_GENERIC / _UMMA), and whether the static
 * weight-norm bound proved the PE clamps inert for it */
int sesrq_net_layer_info(const SesrqNet* net, int layer, int32_t* impl, int32_t* pe_exact);

/* Accumulator layout the tcgen05 path chose for a layer (csrc/sesrq_umma.cu header): how the reference's per-PE 18-bit
 * clamp (quan_func.py:338-370) is honoured.  NS1: provably inert, one 16-column set; SPARE: the few (channel, PE) sums that
 * can clamp ride in the unused columns 12..15 of a <= 12-channel final layer; NS2: in a second set of 4 correction
 * columns; NS4: four PE-masked sets and the literal clamp tree.  0 = the layer runs on the generic dp4a path. */
#define SESRQ_LAYOUT_GENERIC 0
#define SESRQ_LAYOUT_NS1 1
#define SESRQ_LAYOUT_SPARE 2
#define SESRQ_LAYOUT_NS2 3
#define SESRQ_LAYOUT_NS4 4
int sesrq_net_layer_layout(const SesrqNet* net, int layer);   /* SESRQ_LAYOUT_* or a negative error code */

/* Layer fusion.  When residual_block.0 and residual_block.1 (layers 1, 2: plain 3x3, 16 -> 16 channels, no PE clamp that
 * can fire) both run on the tcgen05 path, sesrq_net_forward_i8 executes them in ONE launch and hands the intermediate
 * codes (the reference's input.2.pt) from the first to the second through shared memory -- they never reach HBM
 * (csrc/sesrq_chain.cu).  Results are bit-identical either way.  Enabled by default (SESRQ_FUSE=0 in the environment
 * at create time disables it); sesrq_net_set_fusion switches it per handle and returns the previous setting (0 / 1)
 * or a negative error code.  With fusion on, sesrq_ws_read_act_i8(domain 2) returns SESRQ_EUNSUPPORTED. */
int sesrq_net_set_fusion(SesrqNet* net, int enable);
int sesrq_net_fused_layers(const SesrqNet* net);   /* 1 if the next forward will fuse layers 1 + 2, else 0 */

/* Bytes of device workspace needed by sesrq_net_forward_i8 for a [B,Cin,H,W] batch. */
size_t sesrq_ws_bytes(const SesrqNet* net, int B, int H, int W);

/* Integer-sim forward.  d_x: float32 NCHW [B,Cin,H,W].  d_q5 (nullable): int8 NCHW [B,Cout,H,W],
 * the reference's input.5.pt (quan_func.py:585,592).  d_y (nullable): float32 NCHW
 * [B,Cout/r^2,rH,rW] = PixelShuffle((q5 - z5) * s5), the tensor sim.py:205 returns (:594).
 * At least one of the two must be given; with only one the last layer runs a single-output kernel
 * (1080p x2, 16 frames: y only 0.38 ms, q5 only 0.41 ms, both 0.55 ms). */
int sesrq_net_forward_i8(const SesrqNet* net, const float* d_x, int B, int H, int W,
                         int8_t* d_q5, float* d_y, void* d_ws, size_t ws_bytes, void* stream);

/* Same call with HOST buffers: copies in, runs, copies out and returns when the results are in host memory.
 * The batch is cut into chunks that flow through a 3-stage pipeline (H2D / kernels on `stream` / D2H on two
 * internal copy streams, double-buffered device staging), so with pinned host memory the PCIe transfers in both
 * directions overlap the kernels; pageable memory also works.  The workspace and staging buffers are owned by the
 * net (grown on demand; NOT thread-safe on one handle). */
int sesrq_net_forward_i8_host(SesrqNet* net, const float* h_x, int B, int H, int W,
                              int8_t* h_q5, float* h_y, void* stream);

/* Per-kernel device timing: when enabled, every kernel sesrq_net_forward_i8 launches is bracketed by
 * CUDA events on the launching stream.  sesrq_net_profile_read synchronises on them, adds the elapsed
 * milliseconds and launch counts per layer into ms[] / launches[] (index SESRQ_NUM_LAYERS = the separate
 * input-quantise pass of the generic path) and clears the record.  Not thread-safe on one handle. */
int sesrq_net_profile_enable(SesrqNet* net, int enable);
int sesrq_net_profile_read(SesrqNet* net, double* ms, uint64_t* launches);

/* Intermediate activations of the LAST sesrq_net_forward_i8 on this workspace, for layer-by-layer
 * parity: writes q_i (the reference's input.{i}.pt, i = 0..4) as int8 NCHW [B,C_i,H,W]. */
int sesrq_ws_read_act_i8(const SesrqNet* net, const void* d_ws, size_t ws_bytes, int B, int H, int W,
                         int domain, int8_t* d_q_nchw, void* stream);

/* ---- calibration: replaces the running min/max observer of quantize_asymmetrical_by_tensor
 *      exe_mode 0 (quan_func.py:175-196), myMinMaxObserver (quan_modules.py:14-40) and the abandoned
 *      KL calibrator (define.py:16-20) ------------------------------------------------------------ */
/* d_mm[0] = min(d_mm[0], min x), d_mm[1] = max(d_mm[1], max x) (running, order independent). */
int sesrq_minmax_update(const float* d_x, size_t n, float* d_mm, void* stream);
/* d_hist[SESRQ_KL_BINS] (uint64) += histogram of clamp(x, lo, hi), bin = min(B-1, floor((x-lo)*B/(hi-lo))). */
int sesrq_hist_update(const float* d_x, size_t n, float lo, float hi, unsigned long long* d_hist, void* stream);
/* KL-entropy threshold search on device: the published NVIDIA pytorch-quantization / TensorRT entropy calibrator
 * (`_compute_amax_entropy`: h[0] := h[1], source bin k -> target group floor(k * tgt / i), outliers folded into the
 * last kept bin, last minimum wins; restated in oracle/kl_oracle.py).  Writes the chosen bin count i in
 * [tgt_bins, nbins] to d_idx[0]; d_kl (required, nbins+1 doubles) receives KL(P||Q) per candidate i (entries below
 * tgt_bins are untouched).  Clipping threshold = lo + i * (hi - lo) / nbins. */
int sesrq_kl_threshold(const unsigned long long* d_hist, int nbins, int tgt_bins, int32_t* d_idx,
                       double* d_kl, void* stream);

/* ---- fake-quant forward: replaces the exe_mode 0 chain of test.py:79-106,146 ------------------ */
typedef struct SesrqFqDesc {
  int32_t abi_version;
  int32_t n_layers;                   /* SESRQ_NUM_LAYERS */
  int32_t r_shuffle;
  int32_t residual;                   /* 1: AddOp(residual_features, initial_features) before conv_last */
  int32_t observe_out;                /* 1: also observe the last conv's output as domain 5 (MFLAG 3, :460-479) */
  int32_t pe_acc_bits, pe_add_bits, reserved0;
  int32_t ksize[SESRQ_NUM_LAYERS], cin[SESRQ_NUM_LAYERS], cout[SESRQ_NUM_LAYERS], relu[SESRQ_NUM_LAYERS];
  const int8_t* h_weight[SESRQ_NUM_LAYERS];   /* int8 codes */
  double w_scale[SESRQ_NUM_LAYERS];           /* conv.weight.{i}.scale.pt */
  const float* h_bias[SESRQ_NUM_LAYERS];      /* float biases */
} SesrqFqDesc;
typedef struct SesrqFq SesrqFq;
int sesrq_fq_create(const SesrqFqDesc* desc, SesrqFq** out);
void sesrq_fq_destroy(SesrqFq* fq);
size_t sesrq_fq_ws_bytes(const SesrqFq* fq, int H, int W);                 /* == sesrq_fq_ws_bytes_batch(fq, 1, H, W) */
size_t sesrq_fq_ws_bytes_batch(const SesrqFq* fq, int B, int H, int W);
/* B frames [B,Cin,H,W] in one pass (~20 launches for the whole batch).  The reference path is per-frame dynamic:
 * every domain of every frame is quantised with THAT frame's min/max (quan_func.py:177-207), so frames stay
 * independent; they are only processed together.  d_y: float32 [B,Cout/r^2,rH,rW].
 * d_minmax: float[SESRQ_NUM_DOMAINS][2] running (min,max) per domain, updated in place on device (:188-195).
 * d_frame_minmax (nullable): float[B][SESRQ_NUM_DOMAINS][2], every frame's (min,max) per domain.
 * d_hist (nullable) with d_hist_range float[SESRQ_NUM_DOMAINS][2] = (lo,hi): second calibration pass,
 * adds every domain tensor of every frame to its SESRQ_KL_BINS-bin histogram (uint64[6][2048]).
 * Kernels: per layer a per-frame constants kernel, an operand quantiser and the tcgen05 convolution with four
 * PE-masked accumulator sets and the float-domain epilogue (sesrq_umma.cu, EPI_FQ). */
int sesrq_fq_forward_batch(const SesrqFq* fq, const float* d_x, int B, int H, int W, float* d_y,
                           float* d_minmax, float* d_frame_minmax, unsigned long long* d_hist,
                           const float* d_hist_range, void* d_ws, size_t ws_bytes, void* stream);
/* One frame: sesrq_fq_forward_batch with B = 1. */
int sesrq_fq_forward(const SesrqFq* fq, const float* d_x, int H, int W, float* d_y,
                     float* d_minmax, float* d_frame_minmax, unsigned long long* d_hist,
                     const float* d_hist_range, void* d_ws, size_t ws_bytes, void* stream);

/* ---- stand-alone myQL ops on float32 NCHW device tensors (exe_mode 1), for graphs that are not
 *      fused (partial passes, dump flags set): each replaces one reference callable -------------------------------------------------- */
/* quantize_asymmetrical_by_tensor, func_id 0 (:222-225) and 1..3 (:275-280): returns q - max(z,-128) */
int sesrq_op_quant_first(const float* d_x, size_t n, double scale, int zero, float* d_q, float* d_u, void* stream);
int sesrq_op_quant_next(const float* d_x, size_t n, int zero, float* d_q, float* d_u, void* stream);
/* quantize_asymmetrical_by_tensor, func_id 4 (:244-270): residual add + requant */
int sesrq_op_quant_residual(const float* d_x, const float* d_shortcut, size_t n, int m16, int shift_n,
                            int zero, float* d_ic, float* d_q, float* d_u, void* stream);
/* reshape_input_for_hardware_pe (:298-318): [B,C,H,W] -> [4B,C,H,W] */
int sesrq_op_pe_split(const float* d_x, int B, int C, int HW, int pe_num, float* d_out, void* stream);
/* PEs_and_bias_adder exe_mode 1 (:418-493) on the PE-split conv output [4B,Cout,H,W]:
 * d_pe_corr[cout*4] = sum(W_pe)*z', add_const[cout]; optional dumps d_pe_out (4B,..), d_pe_add. */
int sesrq_op_pe_merge(const float* d_conv, int B, int Cout, int HW, const float* d_pe_corr,
                      const float* d_add_const, int pe_acc_bits, int pe_add_bits,
                      float* d_pe_out, float* d_pe_add, float* d_out, int32_t* d_overflow, void* stream);
/* requan_conv2d_output exe_mode 1 (:517-611): y = x * M16 * 2^-n, optional relu copy (shortcut),
 * final-layer variant rounds into domain 5 and dequantises. */
int sesrq_op_requant(const float* d_x, size_t n, int m16, int shift_n, float* d_y, float* d_relu_copy, void* stream);
int sesrq_op_requant_final(const float* d_x, size_t n, int m16, int shift_n, int zero_out, double scale_out,
                           float* d_q5, float* d_y, void* stream);
/* exact integer conv of float-held codes (the nn.Conv2d node of the un-fused graph) */
int sesrq_op_conv_int(const float* d_x, int B, int Cin, int H, int W, const float* d_w, int Cout, int k,
                      float* d_out, void* stream);

/* ---- the same callables in exe_mode 0 (fake-quant "staged" float, the un-fused test.py graph) ---------------------- */
/* quantize_asymmetrical_by_tensor exe_mode 0 (:207-215): y = (clamp(rint(x / s + z)) - z) * s; the caller derives
 * (s, z) from this tensor's min/max (sesrq_minmax_update) as quan_func.py:198-205 does. */
int sesrq_op_fakequant(const float* d_x, size_t n, double scale, int zero, float* d_y, void* stream);
/* the nn.Conv2d node of the mode-0 graph: d_x holds (q - z) * s (PE-split), d_wcodes the integer weight codes as
 * float32 [Cout,Cin,k,k]; exact integer accumulation scaled once by f32(s) * f32(s_w) (bit-identical to sesrq_fq_forward) */
int sesrq_op_conv_fq(const float* d_x, int B, int Cin, int H, int W, const float* d_wcodes, int Cout, int k,
                     double in_scale, double w_scale, float* d_out, void* stream);
/* reshape_ouput_for_hardware_pe + PEs_and_bias_adder exe_mode 0 (:330-333, :386, :431-434, :459) on the PE-split conv
 * output [4B,Cout,H,W]: clamp each PE to [lo_acc, hi_acc], sum, clamp to [lo_add, hi_add], + d_bias[Cout] (nullable). */
int sesrq_op_pe_merge_fq(const float* d_conv, int B, int Cout, int HW, float lo_acc, float hi_acc, float lo_add,
                         float hi_add, const float* d_bias, float* d_out, void* stream);

/* ---- QAT library (models/quantize_utils_cuda.py:150-246 Round + Quantizer.forward): fake-quant forward and the
 *      straight-through backward.  d_scale / d_zp / d_lo / d_hi hold `rows` entries: 1 (per tensor, q_level "L") or the
 *      channel count (q_level "C": element i belongs to row (i / inner) % rows).
 *      fwd: y = (clamp(round_half_away(x / s - zp), qmin, qmax) + zp) * s
 *      bwd: g_x = (g * s) / s where the clamp and the Round mask (t = x / s - zp inside [lo, hi]) pass, else 0 ------- */
int sesrq_qat_fakequant_fwd(const float* d_x, size_t n, size_t inner, int rows, const float* d_scale, const float* d_zp,
                            float qmin, float qmax, float* d_y, void* stream);
int sesrq_qat_fakequant_bwd(const float* d_x, const float* d_gy, size_t n, size_t inner, int rows, const float* d_scale,
                            const float* d_zp, const float* d_obs_min, const float* d_obs_max, int q_type, float qmin,
                            float qmax, float* d_gx, void* stream);
/* Quantizer.update_qparams (:299-327): q_type 0 symmetric (scale = max(|min|,|max|) / ((qmax - qmin) / 2), zp = 0),
 * 1 asymmetric (scale = (max - min) / (qmax - qmin), zp = sign(min) * floor(|min / scale| + 0.5)); scale >= eps */
int sesrq_qat_qparams(const float* d_min, const float* d_max, int rows, int q_type, float qmin, float qmax, float eps,
                      float* d_scale, float* d_zp, void* stream);
/* ObserverBase.forward + update_range (:22-35, :68-80, :107-119): (min, max) of each of the `rows` rows of the
 * [rows, cols] view of d_x, folded into d_min / d_max: mode 0 = set, 1 = running min / max, 2 = moving average
 * (1 - momentum) * old + momentum * cur.  d_tmp: 2 * rows floats of scratch. */
int sesrq_qat_observe(const float* d_x, int rows, size_t cols, int mode, float momentum, float* d_min, float* d_max,
                      float* d_tmp, void* stream);
/* per-row (min, max) of a [rows, cols] matrix: ObserverBase.forward q_level "C" (:27-30) */
int sesrq_minmax_rows(const float* d_x, int rows, size_t cols, float* d_min, float* d_max, void* stream);
/* k-th smallest |x| (1-based; torch.kthvalue(input.abs().view(-1), k) of HistogramObserver, :136-138), exact two-level
 * radix select on the bit pattern.  d_ws: (65536 + 32768 + 2) * 4 bytes of scratch. */
int sesrq_kth_abs(const float* d_x, size_t n, unsigned long long k, float* d_out, void* d_ws, size_t ws_bytes, void* stream);

/* ---- evaluation post-step of test.py:148-170: out = clip(y + nearest_upsample_r(x), 0, 1) (d_x nullable: no add; the
 *      x2 model adds the upsampled input outside the network) and, with d_gt, *d_sqerr += sum of squared errors of the
 *      PSNR the script prints: metric 0 plain values, 1 channel 0 * 255 (:165-166), 2 BT.601 luma of rgb_to_yuv
 *      (:133-136, C must be 3).  d_y / d_gt / d_out: float32 [B,C,Hr,Wr]; d_x: [B,C,Hr/r,Wr/r]. ------------------ */
int sesrq_op_sr_post(const float* d_y, const float* d_x, const float* d_gt, int B, int C, int Hr, int Wr, int r, int metric,
                     float* d_out, double* d_sqerr, void* stream);

/* SSIM of test.py:172-175 (skimage.metrics.structural_similarity with its defaults: 7 x 7 uniform window, K1 = 0.01,
 * K2 = 0.03, sample covariance; channel_axis = 2 averages the per-plane values): *d_sum += sum of the SSIM map over the
 * interior (H - 6) x (W - 6) window centres of `planes` planes [planes, H, W]; the caller divides by
 * planes * (H - 6) * (W - 6). */
int sesrq_op_ssim(const float* d_a, const float* d_b, int planes, int H, int W, double data_range, double* d_sum, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SESRQ_H_ */
