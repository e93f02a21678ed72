// tcgen05 implicit-GEMM convolution for the 16-channel layers (sm_100a only).
//
// Data flow per CTA (persistent, one CTA per SM, warp-specialised):
//   warp 0      TMA producer: one cp.async.bulk.tensor.3d per tile brings a (TH+2p)-row x 128-pixel halo tile of
//               NHWC16 bytes into shared memory.  The tensor map views a frame row as W*2 8-byte elements so
//               the box's inner extent is a contiguous 2048-byte run (a 16-byte inner box makes TMA fetch one
//               32-byte sector per pixel: measured 2x traffic, profiles/r1a).  Out-of-frame pixels are
//               zero-filled by TMA, and byte 0 == code -128 == the reference's pad value z'
//               (quan_func.py:351-352), so padding is free.
//   warps 1-3   MMA issuers (accumulator blocks dealt round-robin): one elected thread per warp issues
//               tcgen05.mma.kind::i8 (u8 activations x s8 weights -> s32 in TMEM).  No im2col: the tile is
//               pixel-major (16 B per pixel), so the A operand of tap (dy,dx) for 128 consecutive output
//               positions is the same tile shifted by (dy*128+dx) pixels.  A K=32 instruction covers two taps:
//               its canonical no-swizzle K-major descriptor takes the two 16-byte K chunks LBO bytes apart, and
//               LBO is simply the distance between the two taps.  The tile pitch is 128 pixels, so accumulator
//               block b is output row b of the tile and every 128-row operand is one uniform-stride run
//               (SBO = 128 B); the 2p rightmost columns of a block are garbage and the epilogue drops them.
//               Descriptors live in registers: per MMA the only arithmetic is one add on the low word.
//               Warp 1 also owns the TMEM allocation (512 columns).
//   warps 4..   epilogue: tcgen05.ld (lane = pixel, column = output channel), exact requant arithmetic of
//               SURVEY App. A, one 16-byte NHWC16 store per pixel (or int8 NCHW + pixel-shuffled fp32 vector
//               stores for the last layer).
//
// PE semantics (quan_func.py:338-370: each of the 4 PEs clamps its partial sum to PE_ACC_BIT before the adder
// tree).  The host proves per (output channel, PE) whether that clamp can ever fire (exact range of the partial
// sum over all int8 codes).  Three accumulator layouts, chosen per layer:
//   NS = 1  no clamp can fire: one 128x16 accumulator, the clamps are skipped.
//   NS = 2  at most one PE per channel can clamp (x2 conv_last: 2 of 48 pairs): 128x32.  Columns 0-15 hold the
//           full sums, column 16+n the partial sum of channel n's one risky PE (B rows masked to that PE's
//           input channels, zero rows elsewhere); the epilogue adds clamp(v) - v.
//   NS = 4  anything else: 128x64, four PE-masked copies of the weights, the literal clamp tree.
// All three issue the same number of MMAs; the tensor pipe has the headroom.
#include <cuda.h>
#include <algorithm>
#include <cstring>
#include <vector>
#include "sesrq_internal.h"

namespace sesrq {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// bytes [a,b,c,d] (a lowest) of the four ints saturated to int8: two I2IP instructions.
// PTX: cvt.pack.sat.s8.s32.b32 d, x, y, z  ->  d = (z << 16) | (sat8(x) << 8) | sat8(y)
__device__ __forceinline__ uint32_t pack4_sat_s8(int a, int b, int c, int d) {
  uint32_t t, r;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(t));
  return r;
}

// ------------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------------
constexpr int UMMA_M = 128;
constexpr int MAX_PAIRS = 13;            // ceil(25 / 2)
constexpr int TMEM_COLS = 512;
constexpr int CTRL_WARPS = 4;            // warp 0: TMA producer; warps 1-3: MMA issuers (keeps epilogue warp % 4 == lane quarter)
constexpr int MMA_WARPS = 3;
constexpr int MAX_STAGES = 4;

struct UmmaParams {
  // per-column epilogue constants (column n of the accumulator <-> output channel col_chan(n))
  int32_t c_fused[16];                   // NS=1: sum_p pe_corr + K
  int32_t c_sum[16];                     // NS=2: sum_p pe_corr
  int32_t c_k[16];                       // NS=2/4: K (add const)
  int32_t c_x[16];                       // NS=2: pe_corr of the channel's risky PE
  int32_t c_pe[4][16];                   // NS=4: pe_corr per PE
  uint32_t xmask;                        // NS=2: bit n set <=> channel of column n has a risky PE
  int32_t lo_acc, hi_acc, lo_add, hi_add;
  float mf, z_next, res_mf, z_res, s_out;
  int32_t relu, cout;
  // geometry
  int32_t B, H, W, TH, TW, tiles_x, tiles_y, num_tiles, pad;
  uint32_t box_bytes, stage_bytes, w_bytes;
  int32_t stages;                        // tile buffers in flight (2..MAX_STAGES)
  int32_t pair_off[MAX_PAIRS];           // pixel offset of the first tap of the pair inside the tile
  int32_t pair_lbo[MAX_PAIRS];           // pixel distance to the second tap
  uint32_t idesc;
  // layer-0 kernel only
  int32_t cin;
  uint32_t pad_word;                     // out-of-frame pixel: byte c = z' + 128 for c < cin
  float in_scale, in_zero;
};

// ------------------------------------------------------------------------------------------------
// epilogue of one output pixel: 16 (x NS) accumulator columns -> codes.  Shared by both kernels.
// ------------------------------------------------------------------------------------------------
template <int NS, int EPI, int RS>
__device__ __forceinline__ void epilogue_pixel(const UmmaParams& P, const int (&acc)[NS][16], const uint4 sc,
                                               uint4* __restrict__ out_px, int8_t* __restrict__ qb,
                                               float* __restrict__ yb, const long long HW, const float relu_floor) {
  // ---- A = clampAdd(sum_p clampAcc(P_p)) + K  (quan_func.py:370,437,491) ----
  int A[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    if (NS == 1) {
      A[n] = acc[0][n] + P.c_fused[n];
    } else if (NS == 2) {
      int s = acc[0][n] + P.c_sum[n];
      if (P.xmask & (1u << n)) {
        const int v = acc[NS - 1][n] + P.c_x[n];
        s += clampi(v, P.lo_acc, P.hi_acc) - v;
      }
      A[n] = clampi(s, P.lo_add, P.hi_add) + P.c_k[n];
    } else {
      int s = 0;
#pragma unroll
      for (int p = 0; p < NS; ++p) s += clampi(acc[p][n] + P.c_pe[p][n], P.lo_acc, P.hi_acc);
      A[n] = clampi(s, P.lo_add, P.hi_add) + P.c_k[n];
    }
  }
  if (EPI == EPI_FINAL) {
    // q5 = clamp(rint(y + z5)); out = (q5 - z5) * s5 (:585,:594); columns are in natural channel order
    constexpr int R2 = RS * RS;
    const long long HrWr = HW * R2;
    const int Wr = P.W * RS;
#pragma unroll
    for (int n0 = 0; n0 < 16; n0 += RS) {
      if (n0 < P.cout) {
        float o[RS];
#pragma unroll
        for (int j = 0; j < RS; ++j) {
          const float yv = requant(A[n0 + j], P.mf);
          const int q = min(max(__float2int_rn(__fadd_rn(yv, P.z_next)), -128), 127);
          if (qb) qb[(long long)(n0 + j) * HW] = (int8_t)q;
          o[j] = __fmul_rn(__fsub_rn((float)q, P.z_next), P.s_out);
        }
        if (yb) {
          float* dst = yb + (long long)(n0 / R2) * HrWr + (long long)((n0 % R2) / RS) * Wr;
          if constexpr (RS == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
          else if constexpr (RS == 2) *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
          else *dst = o[0];
        }
      }
    }
  } else {
    uint32_t ow[4];
    const uint32_t scw[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      int q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = fmaxf(requant(A[4 * w + j], P.mf), relu_floor);
        if (EPI == EPI_RESIDUAL) {
          const int ic = min(__float2int_rn(__fadd_rn(a, -128.f)), 127);         // :251 (a >= 0)
          const int vrc = (int)__byte_perm(scw[w], 0u, 0x4440u + j);              // rc + 128 (:250)
          const float tt = __fmul_rn((float)(vrc + ic + 128), P.res_mf);          // :252,:269
          q[j] = __float2int_rn(__fadd_rn(tt, P.z_res));                          // :270, clamped by the pack
        } else {
          q[j] = __float2int_rn(__fadd_rn(a, P.z_next));                          // :280, clamped by the pack
        }
      }
      ow[w] = pack4_sat_s8(q[0], q[1], q[2], q[3]) ^ 0x80808080u;                 // v = q + 128
    }
    *out_px = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel.  KS: kernel size; NS: accumulator sets (1/2/4, see header); EPI: EpilogueKind; RS: PixelShuffle
// factor of the final epilogue (1 otherwise); EPI_WARPS: number of epilogue warps.
// ------------------------------------------------------------------------------------------------
template <int KS, int NS, int EPI, int RS, int EPI_WARPS>
__global__ void __launch_bounds__((CTRL_WARPS + EPI_WARPS) * 32, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ UmmaParams P,
                 const uint8_t* __restrict__ w_img, const uint4* __restrict__ shortcut, uint4* __restrict__ out,
                 int8_t* __restrict__ q5, float* __restrict__ y) {
  constexpr int NCOLS = 16 * NS;
  constexpr int NSLOT = TMEM_COLS / NCOLS;
  constexpr int NP = (KS * KS + 1) / 2;   // tap pairs = MMAs per accumulator block
  constexpr int NWG = EPI_WARPS / 4;
  static_assert(NSLOT % NWG == 0, "slot ownership must be stable per warpgroup");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const tile0 = smem;
  uint8_t* const w_s = smem + (size_t)P.stages * P.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + ((P.w_bytes + 127) & ~127u));
  uint64_t* tile_full = bars;                            // [MAX_STAGES]
  uint64_t* tile_empty = bars + MAX_STAGES;              // [MAX_STAGES]
  uint64_t* acc_full = bars + 2 * MAX_STAGES;            // [NSLOT]
  uint64_t* acc_empty = bars + 2 * MAX_STAGES + NSLOT;   // [NSLOT]
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 2 * NSLOT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup ----
  for (uint32_t i = threadIdx.x; i < P.w_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(w_s)[i] = __ldg(reinterpret_cast<const uint4*>(w_img) + i);
  fence_proxy_async();                      // weights were written through the generic proxy, UMMA reads via async proxy
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&tile_full[i], 1);
      mbar_init(&tile_empty[i], MMA_WARPS);
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);          // one elected arrival per epilogue warp of the owning warpgroup
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  const int tiles_per_frame = P.tiles_x * P.tiles_y;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;                       // phase parity of the current pass over the stage ring
      for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
        mbar_wait(&tile_empty[st], ph ^ 1);
        const int b = t / tiles_per_frame;
        const int r = t - b * tiles_per_frame;
        const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
        mbar_arrive_expect_tx(&tile_full[st], P.box_bytes);
        tma_load_3d(tile0 + (size_t)st * P.stage_bytes, &tmap, (tx * P.TW - P.pad) * 2, ty * P.TH - P.pad, b, &tile_full[st]);
        if (++st == P.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp < CTRL_WARPS) {
    // ================= MMA issuers (warps 1..3) =================
    const int mw = warp - 1;
    const bool leader = elect_one();
    uint32_t a_lo[NP], b_lo[NP];
    const uint32_t w_addr = smem_u32(w_s);
#pragma unroll
    for (int pr = 0; pr < NP; ++pr) {
      a_lo[pr] = (uint32_t)P.pair_off[pr] | ((uint32_t)P.pair_lbo[pr] << 16);                 // 16-byte units
      b_lo[pr] = (((w_addr + (uint32_t)pr * (2u * NCOLS * 16u)) & 0x3FFFFu) >> 4) | ((uint32_t)NCOLS << 16);
    }
    // cute::UMMA::SmemDescriptor high word: SBO = 128 B (bits 32-45), version 1 (bit 46), no swizzle
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t idesc = P.idesc;
    const uint32_t tile_addr0 = (smem_u32(tile0) & 0x3FFFFu) >> 4;
    int st = 0;
    uint32_t ph = 0;
    uint32_t g0 = 0;
    for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
      const int r = t % tiles_per_frame;
      const int nblk = min(P.TH, P.H - (r / P.tiles_x) * P.TH);     // rows of this tile that exist
      mbar_wait(&tile_full[st], ph);
      tc_fence_after();
      const uint32_t a_base = tile_addr0 + (uint32_t)st * (P.stage_bytes >> 4);
      for (int mb = mw; mb < nblk; mb += MMA_WARPS) {
        const uint32_t g = g0 + mb;
        const uint32_t slot = g % NSLOT;
        mbar_wait(&acc_empty[slot], ((g / NSLOT) & 1) ^ 1);
        tc_fence_after();
        if (leader) {
          const uint32_t d = tmem_base + slot * NCOLS;
          const uint32_t row = a_base + (uint32_t)mb * (UMMA_M * 16 / 16);
#pragma unroll
          for (int pr = 0; pr < NP; ++pr) {
            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo[pr] + row);
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)b_lo[pr];
            umma_i8(d, ad, bd, idesc, pr > 0 ? 1u : 0u);
          }
          umma_commit(&acc_full[slot]);       // implies tcgen05.fence::before_thread_sync
        }
        __syncwarp();
      }
      if (leader) umma_commit(&tile_empty[st]);   // fires when this warp's MMAs on the tile have retired
      __syncwarp();
      g0 += nblk;
      if (++st == P.stages) { st = 0; ph ^= 1; }
    }
  } else {
    // ================= epilogue =================
    const int wg = (warp - CTRL_WARPS) >> 2;      // warpgroup
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int ox = quarter * 32 + lane;           // pitch == UMMA_M: block b is output row b, lane is the column
    const long long HW = (long long)P.H * P.W;
    const float relu_floor = P.relu ? 0.f : -INFINITY;
    uint32_t g0 = 0;
    for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
      const int b = t / tiles_per_frame;
      const int r = t - b * tiles_per_frame;
      const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
      const int y0 = ty * P.TH, gx = tx * P.TW + ox;
      const int nblk = min(P.TH, P.H - y0);
      const bool valid = (ox < P.TW) && (gx < P.W);
      const long long pix0 = (long long)b * HW + (long long)y0 * P.W + gx;
      const int first = (int)((wg + NWG - (g0 % NWG)) % NWG);
      for (int mb = first; mb < nblk; mb += NWG) {
        const uint32_t g = g0 + mb;
        const uint32_t slot = g % NSLOT;
        const long long pix = pix0 + (long long)mb * P.W;
        uint4 sc = make_uint4(0u, 0u, 0u, 0u);
        if (EPI == EPI_RESIDUAL && valid) sc = __ldg(shortcut + pix);

        __syncwarp();                        // the .sync.aligned TMEM loads below need a converged warp
        mbar_wait(&acc_full[slot], (g / NSLOT) & 1);
        tc_fence_after();
        int acc[NS][16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * NCOLS;
#pragma unroll
        for (int p = 0; p < NS; ++p) tmem_ld16(taddr + p * 16, acc[p]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);

        if (valid) {
          float* yb = nullptr;
          int8_t* qb = nullptr;
          if (EPI == EPI_FINAL) {
            constexpr int R2 = RS * RS;
            if (y) yb = y + (long long)b * (P.cout / R2) * (HW * R2) + (long long)(y0 + mb) * RS * (P.W * RS) + (long long)gx * RS;
            if (q5) qb = q5 + (long long)b * P.cout * HW + (long long)(y0 + mb) * P.W + gx;
          }
          epilogue_pixel<NS, EPI, RS>(P, acc, sc, out + pix, qb, yb, HW, relu_floor);
        }
      }
      g0 += nblk;
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Layer-0 kernel: fp32 NCHW frames with cin <= 4 channels.  Builder warps quantise the halo tile on the fly
// (clamp(rint(x / s0 + z0)), true fp32 division, quan_func.py:225) into a 4-byte-per-pixel plane S, then
// expand it into the K-major operand plane  E[row][X] = S[row][X .. X+3]  (16 B per entry).  For output
// column X and tap row dy, the two 16-byte K chunks of one MMA are E[y+dy][X] (taps dx = 0..3) and
// E[y+dy][X+4] (dx = 4..7, weights zero beyond the kernel width): LBO = 64 B, SBO = 128 B, one MMA per tap
// row.  Everything downstream (TMEM accumulators, PE layouts, epilogue) is shared with the 16-channel kernel.
// ------------------------------------------------------------------------------------------------
constexpr int L0_EP = 132;               // E row pitch in entries (X up to 127 + 4)
constexpr int L0_SP = 136;               // S row pitch in words  (X + 3 up to 134)
constexpr int L0_BUILD_WARPS = 8;

// x / s with a hoisted reciprocal.  `rcp` is the Newton-refined reciprocal of s; the three operations are
// exactly the quotient steps of div.rn.f32's fast path (q0 = x*rcp; e = fma(-s, q0, x); q = fma(rcp, e, q0)),
// which is correctly rounded whenever neither the operands nor the quotient are near the exponent limits.
// Tiny |x| (< 2^-100): the code is decided by the zero point, not by the last bit of the quotient.  Huge |x|
// would overflow the quotient (Inf - Inf = NaN in the residual), so x is first clamped to +-lim with
// lim / s = 2^40: beyond it the code is saturated whatever the int32 zero point is.  The quantised result is
// therefore identical to true division for every finite input (tests: rounding-boundary / extreme inputs).
struct ExactDiv {
  float s, rcp, lim;
  __device__ __forceinline__ explicit ExactDiv(float s_) : s(s_) {
    float y0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(s_));
    rcp = __fmaf_rn(y0, __fmaf_rn(-s_, y0, 1.0f), y0);
    lim = fminf(__fmul_rn(s_, 1099511627776.f), 1.2676506e30f);      // min(s * 2^40, 2^100)
  }
  __device__ __forceinline__ float operator()(float x) const {
    x = fminf(fmaxf(x, -lim), lim);
    const float q0 = __fmul_rn(x, rcp);
    const float e = __fmaf_rn(-s, q0, x);
    return __fmaf_rn(rcp, e, q0);
  }
};

template <int KS, int NS, int CIN, int EPI_WARPS>
__global__ void __launch_bounds__((CTRL_WARPS + EPI_WARPS + L0_BUILD_WARPS) * 32, 1)
conv_umma_l0_kernel(const __grid_constant__ UmmaParams P, const uint8_t* __restrict__ w_img,
                    const float* __restrict__ x, uint4* __restrict__ out) {
  constexpr int NCOLS = 16 * NS;
  constexpr int NSLOT = TMEM_COLS / NCOLS;
  constexpr int NWG = EPI_WARPS / 4;
  constexpr int PAD = (KS - 1) / 2;
  static_assert(NSLOT % NWG == 0, "slot ownership must be stable per warpgroup");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const e0 = smem;                                         // E planes: P.stages buffers of stage_bytes
  uint32_t* const s_plane = reinterpret_cast<uint32_t*>(smem + (size_t)P.stages * P.stage_bytes);
  const uint32_t s_bytes = (uint32_t)(((P.TH + 2 * PAD) * L0_SP * 4 + 127) & ~127);
  uint8_t* const w_s = smem + (size_t)P.stages * P.stage_bytes + s_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + ((P.w_bytes + 127) & ~127u));
  uint64_t* e_full = bars;                               // [MAX_STAGES] count = builder warps
  uint64_t* e_empty = bars + MAX_STAGES;                 // [MAX_STAGES] count = MMA warps (tcgen05.commit)
  uint64_t* acc_full = bars + 2 * MAX_STAGES;
  uint64_t* acc_empty = bars + 2 * MAX_STAGES + NSLOT;
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 2 * NSLOT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (uint32_t i = threadIdx.x; i < P.w_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(w_s)[i] = __ldg(reinterpret_cast<const uint4*>(w_img) + i);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&e_full[i], L0_BUILD_WARPS);
      mbar_init(&e_empty[i], MMA_WARPS);
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const int tiles_per_frame = P.tiles_x * P.tiles_y;
  const long long HW = (long long)P.H * P.W;

  if (warp == 0) {
    // idle (no TMA in this kernel: the operand plane is computed, not copied)
  } else if (warp < CTRL_WARPS) {
    // ================= MMA issuers =================
    const int mw = warp - 1;
    const bool leader = elect_one();
    uint32_t b_lo[KS];
    const uint32_t w_addr = smem_u32(w_s);
#pragma unroll
    for (int dy = 0; dy < KS; ++dy)
      b_lo[dy] = (((w_addr + (uint32_t)dy * (2u * NCOLS * 16u)) & 0x3FFFFu) >> 4) | ((uint32_t)NCOLS << 16);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t idesc = P.idesc;
    const uint32_t e_addr0 = (smem_u32(e0) & 0x3FFFFu) >> 4;
    int st = 0;
    uint32_t ph = 0;
    uint32_t g0 = 0;
    for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
      const int r = t % tiles_per_frame;
      const int nblk = min(P.TH, P.H - (r / P.tiles_x) * P.TH);
      mbar_wait(&e_full[st], ph);
      tc_fence_after();
      const uint32_t a_base = e_addr0 + (uint32_t)st * (P.stage_bytes >> 4);
      for (int mb = mw; mb < nblk; mb += MMA_WARPS) {
        const uint32_t g = g0 + mb;
        const uint32_t slot = g % NSLOT;
        mbar_wait(&acc_empty[slot], ((g / NSLOT) & 1) ^ 1);
        tc_fence_after();
        if (leader) {
          const uint32_t d = tmem_base + slot * NCOLS;
#pragma unroll
          for (int dy = 0; dy < KS; ++dy) {
            const uint32_t a_lo = (a_base + (uint32_t)(mb + dy) * L0_EP) | (4u << 16);     // LBO = 4 entries = 64 B
            umma_i8(d, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo[dy], idesc, dy > 0 ? 1u : 0u);
          }
          umma_commit(&acc_full[slot]);
        }
        __syncwarp();
      }
      if (leader) umma_commit(&e_empty[st]);
      __syncwarp();
      g0 += nblk;
      if (++st == P.stages) { st = 0; ph ^= 1; }
    }
  } else if (warp < CTRL_WARPS + EPI_WARPS) {
    // ================= epilogue (identical to the 16-channel kernel, EPI_PLAIN) =================
    const int wg = (warp - CTRL_WARPS) >> 2;
    const int quarter = warp & 3;
    const int ox = quarter * 32 + lane;
    const float relu_floor = P.relu ? 0.f : -INFINITY;
    uint32_t g0 = 0;
    for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
      const int b = t / tiles_per_frame;
      const int r = t - b * tiles_per_frame;
      const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
      const int y0 = ty * P.TH, gx = tx * P.TW + ox;
      const int nblk = min(P.TH, P.H - y0);
      const bool valid = (ox < P.TW) && (gx < P.W);
      const long long pix0 = (long long)b * HW + (long long)y0 * P.W + gx;
      const int first = (int)((wg + NWG - (g0 % NWG)) % NWG);
      for (int mb = first; mb < nblk; mb += NWG) {
        const uint32_t g = g0 + mb;
        const uint32_t slot = g % NSLOT;
        __syncwarp();
        mbar_wait(&acc_full[slot], (g / NSLOT) & 1);
        tc_fence_after();
        int acc[NS][16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * NCOLS;
#pragma unroll
        for (int p = 0; p < NS; ++p) tmem_ld16(taddr + p * 16, acc[p]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
        if (valid)
          epilogue_pixel<NS, EPI_PLAIN, 1>(P, acc, make_uint4(0u, 0u, 0u, 0u), out + pix0 + (long long)mb * P.W, nullptr,
                                           nullptr, HW, relu_floor);
      }
      g0 += nblk;
    }
  } else {
    // ================= builders =================
    const int bt = threadIdx.x - (CTRL_WARPS + EPI_WARPS) * 32;       // 0 .. 255
    constexpr int NB = L0_BUILD_WARPS * 32;
    const ExactDiv div_s0(P.in_scale);
    const float z0 = P.in_zero;
    int st = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
      const int b = t / tiles_per_frame;
      const int r = t - b * tiles_per_frame;
      const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
      const int y0 = ty * P.TH - PAD, x0 = tx * P.TW - PAD;
      const int rows = min(P.TH, P.H - ty * P.TH) + 2 * PAD;
      const float* xb = x + (long long)b * CIN * HW;
      // phase A: quantise (rows x 128 pixels) into S; columns 128..135 only feed discarded outputs
      asm volatile("bar.sync 1, %0;" ::"n"(NB) : "memory");           // previous tile's phase B has drained S
      // (U pixels x cin loads are issued back to back so the DRAM latency is paid once per batch, not per pixel)
      constexpr int U = 6;
      const int npx = rows * UMMA_M;
      for (int i0 = bt; i0 < npx; i0 += NB * U) {
        float v[U][CIN];
        bool inside[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * NB;
          const int gy = y0 + (i >> 7), gx = x0 + (i & (UMMA_M - 1));
          inside[u] = (i < npx) && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
          const float* px = xb + (long long)gy * P.W + gx;
#pragma unroll
          for (int c = 0; c < CIN; ++c) v[u][c] = inside[u] ? __ldg(px + (long long)c * HW) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * NB;
          if (i < npx) {
            int q[4] = {-128, -128, -128, -128};            // unused channels: byte 0 (their weights are zero)
#pragma unroll
            for (int c = 0; c < CIN; ++c) q[c] = __float2int_rn(__fadd_rn(div_s0(v[u][c]), z0));   // :225, clamped by the pack
            const uint32_t word = inside[u] ? (pack4_sat_s8(q[0], q[1], q[2], q[3]) ^ 0x80808080u) : P.pad_word;
            s_plane[(i >> 7) * L0_SP + (i & (UMMA_M - 1))] = word;
          }
        }
      }
      if (bt < rows * (L0_SP - UMMA_M)) {                              // tail columns: keep them defined
        const int rr = bt / (L0_SP - UMMA_M), cc = UMMA_M + bt % (L0_SP - UMMA_M);
        s_plane[rr * L0_SP + cc] = P.pad_word;
      }
      mbar_wait(&e_empty[st], ph ^ 1);                                 // the MMAs that read this E stage have retired
      asm volatile("bar.sync 1, %0;" ::"n"(NB) : "memory");           // S complete
      // phase B: E[row][X] = S[row][X .. X+3]
      uint4* e = reinterpret_cast<uint4*>(e0 + (size_t)st * P.stage_bytes);
      for (int i = bt; i < rows * L0_EP; i += NB) {
        const int rr = i / L0_EP, xx = i - rr * L0_EP;
        const uint32_t* sp = s_plane + rr * L0_SP + xx;
        e[i] = make_uint4(sp[0], sp[1], sp[2], sp[3]);
      }
      fence_proxy_async();                                             // generic-proxy stores -> visible to UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&e_full[st]);
      if (++st == P.stages) { st = 0; ph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static bool is_l0_layout(const LayerConst& c) { return c.cin <= 4; }

bool umma_supported(const LayerConst& c) {
  // layer 0: fp32 frames with cin <= 4 (any input zero point: the builder writes the pad value itself);
  // other layers: 16 input channels from an NHWC16 tensor whose out-of-frame value is byte 0 (z' == -128).
  // Anything else runs on the generic path.
  if (is_l0_layout(c)) return c.epi == EPI_PLAIN && c.save_rc == 0 && (c.k == 3 || c.k == 5);
  if (!(c.cin == 16 && c.pad_byte == 0 && c.save_rc == 0 && (c.k == 3 || c.k == 5))) return false;
  if (c.epi == EPI_FINAL && !(c.r_shuffle == 1 || c.r_shuffle == 2 || c.r_shuffle == 4)) return false;
  return true;
}

// Accumulator layout for this layer (see header): 1, 2 or 4 sets of 16 columns.
static int accumulator_sets(const LayerConst& c) {
  if (!c.pe_exact) return 1;
  for (int co = 0; co < c.cout; ++co)
    if (c.pe_risky[co] > 3) return 4;          // more than one PE of a channel can clamp
  return 2;
}

// B operand image: for every tap pair, [2 K-halves][N/8 row groups][8 rows][16 B]; row n of set s holds the
// weights of output channel col_chan(n) for the 16 input-channel bytes, masked per set as the layout requires.
// Layer-0 image: one K=32 chunk per tap row dy; byte k of half h <-> tap dx = 4h + k/4, input channel k%4
// (PE p == channel p for cin <= 4).
static int umma_prepare_layer0(LayerDev& L, const int8_t* w) {
  const LayerConst& c = L.c;
  const int ns = accumulator_sets(c);
  const int N = 16 * ns;
  L.umma_n = N;
  std::vector<uint8_t> img((size_t)c.k * 2 * N * 16, 0);
  for (int dy = 0; dy < c.k; ++dy)
    for (int h = 0; h < 2; ++h)
      for (int n = 0; n < N; ++n) {
        const int set = n / 16, nn = n % 16;
        const int co = chan_of_byte(nn);
        if (co >= c.cout) continue;
        int only_pe = -1;
        if (ns == 4) only_pe = set;
        if (ns == 2 && set == 1) {
          if (c.pe_risky[co] < 0) continue;
          only_pe = c.pe_risky[co];
        }
        uint8_t* row = img.data() + (((size_t)dy * 2 + h) * N + n) * 16;
        for (int k = 0; k < 16; ++k) {
          const int dx = 4 * h + k / 4, ci = k % 4;
          if (dx >= c.k || ci >= c.cin) continue;
          if (only_pe >= 0 && ci != only_pe) continue;
          row[k] = (uint8_t)w[(((size_t)co * c.cin + ci) * c.k + dy) * c.k + dx];
        }
      }
  if (cudaMalloc(&L.d_wumma, img.size()) != cudaSuccess ||
      cudaMemcpy(L.d_wumma, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "umma layer-0 weight upload");
    return SESRQ_ECUDA;
  }
  return SESRQ_OK;
}

int umma_prepare_layer(LayerDev& L, const int8_t* w) {
  const LayerConst& c = L.c;
  if (is_l0_layout(c)) return umma_prepare_layer0(L, w);
  const int taps = c.k * c.k, pairs = (taps + 1) / 2;
  const int ns = accumulator_sets(c);
  const int N = 16 * ns;
  L.umma_n = N;
  std::vector<uint8_t> img((size_t)pairs * 2 * N * 16, 0);
  const bool natural = (c.epi == EPI_FINAL);
  for (int pr = 0; pr < pairs; ++pr)
    for (int h = 0; h < 2; ++h) {
      const int t = 2 * pr + h;
      if (t >= taps) continue;
      for (int n = 0; n < N; ++n) {
        const int set = n / 16, nn = n % 16;
        const int co = natural ? nn : chan_of_byte(nn);
        if (co >= c.cout) continue;
        int only_pe = -1;                      // -1: all PEs
        if (ns == 4) only_pe = set;
        if (ns == 2 && set == 1) {
          if (c.pe_risky[co] < 0) continue;    // no risky PE: zero rows
          only_pe = c.pe_risky[co];
        }
        uint8_t* row = img.data() + (((size_t)pr * 2 + h) * N + n) * 16;
        for (int j = 0; j < 16; ++j) {
          const int ci = chan_of_byte(j);
          if (ci >= c.cin) continue;
          if (only_pe >= 0 && (j >> 2) != only_pe) continue;
          row[j] = (uint8_t)w[((size_t)co * c.cin + ci) * taps + t];
        }
      }
    }
  if (cudaMalloc(&L.d_wumma, img.size()) != cudaSuccess ||
      cudaMemcpy(L.d_wumma, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), "umma weight upload");
    return SESRQ_ECUDA;
  }
  return SESRQ_OK;
}

template <int KS, int NS, int EPI, int RS, int EPI_WARPS>
static int launch_one(const CUtensorMap& tmap, const UmmaParams& P, const LayerDev& L, const uint8_t* d_shortcut,
                      uint8_t* d_out, int8_t* d_q5, float* d_y, int grid, size_t smem, cudaStream_t st) {
  auto kern = conv_umma_kernel<KS, NS, EPI, RS, EPI_WARPS>;
  static bool attr_set = false;
  if (!attr_set) {
    SESRQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  kern<<<grid, (CTRL_WARPS + EPI_WARPS) * 32, smem, st>>>(tmap, P, L.d_wumma, reinterpret_cast<const uint4*>(d_shortcut),
                                                          reinterpret_cast<uint4*>(d_out), d_q5, d_y);
  SESRQ_LAUNCH_CHECK("conv_umma_kernel");
  return SESRQ_OK;
}

template <int KS, int NS, int EPI_WARPS>
static int launch_epi(const LayerConst& c, const CUtensorMap& tmap, const UmmaParams& P, const LayerDev& L,
                      const uint8_t* d_shortcut, uint8_t* d_out, int8_t* d_q5, float* d_y, int grid, size_t smem,
                      cudaStream_t st) {
  if (c.epi == EPI_PLAIN) return launch_one<KS, NS, EPI_PLAIN, 1, EPI_WARPS>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
  if (c.epi == EPI_RESIDUAL) return launch_one<KS, NS, EPI_RESIDUAL, 1, EPI_WARPS>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
  if (c.r_shuffle == 4) return launch_one<KS, NS, EPI_FINAL, 4, EPI_WARPS>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
  if (c.r_shuffle == 2) return launch_one<KS, NS, EPI_FINAL, 2, EPI_WARPS>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
  return launch_one<KS, NS, EPI_FINAL, 1, EPI_WARPS>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
}

template <int KS>
static int launch_ns(int ns, const LayerConst& c, const CUtensorMap& tmap, const UmmaParams& P, const LayerDev& L,
                     const uint8_t* d_shortcut, uint8_t* d_out, int8_t* d_q5, float* d_y, int grid, size_t smem,
                     cudaStream_t st) {
  if (ns == 1) return launch_epi<KS, 1, 16>(c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
  if (ns == 2) return launch_epi<KS, 2, 16>(c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
  return launch_epi<KS, 4, 8>(c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
}

// Tile height and stage count: maximise (SM load balance of the static round-robin schedule) x (halo
// efficiency), subject to >= 2 (preferably >= 3) tile buffers fitting next to the fixed shared memory.
struct TilePlan { int TH, tiles_y, stages; };
static TilePlan choose_tiling(int H, int B, int tiles_x, int pad, size_t row_bytes, size_t slack_bytes,
                              size_t per_row_extra, size_t fixed, int sms) {
  TilePlan best{0, 0, 0};
  double best_score = -1.0;
  const size_t total = 227 * 1024;
  for (int th = std::min(H, 32); th >= 1; --th) {
    const int ty = cdiv(H, th);
    const int TH = cdiv(H, ty);                       // balanced rows per tile
    const size_t stage = ((size_t)(TH + 2 * pad) * row_bytes + slack_bytes + 1023) & ~(size_t)1023;
    const size_t extra = (size_t)(TH + 2 * pad) * per_row_extra + 256;
    if (fixed + extra + 2 * stage > total) continue;
    const int stages = (int)std::min<size_t>(MAX_STAGES, (total - fixed - extra) / stage);
    const long long tiles = (long long)tiles_x * ty * B;
    const double balance = (double)tiles / ((double)sms * (double)((tiles + sms - 1) / sms));
    const double halo = 1.0 - 0.3 * (2.0 * pad) / (TH + 2.0 * pad);
    const double depth = stages >= 3 ? 1.0 : 0.9;
    const double score = balance * halo * depth;
    if (score > best_score + 1e-9) { best_score = score; best = {TH, ty, stages}; }
  }
  return best;
}

static void fill_epilogue_params(UmmaParams& P, const LayerConst& c, int ns) {
  const bool natural = (c.epi == EPI_FINAL);
  for (int n = 0; n < 16; ++n) {
    const int co = natural ? n : chan_of_byte(n);
    if (co >= c.cout) continue;
    P.c_fused[n] = c.fused_const[co];
    P.c_k[n] = c.add_const[co];
    int sum = 0;
    for (int p = 0; p < 4; ++p) {
      P.c_pe[p][n] = c.pe_corr[p][co];
      sum += c.pe_corr[p][co];
    }
    P.c_sum[n] = sum;
    if (ns == 2 && c.pe_risky[co] >= 0) {
      P.c_x[n] = c.pe_corr[c.pe_risky[co]][co];
      P.xmask |= 1u << n;
    }
  }
  P.lo_acc = c.lo_acc; P.hi_acc = c.hi_acc; P.lo_add = c.lo_add; P.hi_add = c.hi_add;
  P.mf = c.mf; P.z_next = c.z_next; P.res_mf = c.res_mf; P.z_res = c.z_res; P.s_out = c.s_out;
  P.relu = c.relu; P.cout = c.cout;
}

template <int KS, int NS, int CIN, int EPI_WARPS>
static int launch_l0_one(const UmmaParams& P, const LayerDev& L, const float* d_x, uint8_t* d_out, int grid, size_t smem,
                         cudaStream_t st) {
  auto kern = conv_umma_l0_kernel<KS, NS, CIN, EPI_WARPS>;
  static bool attr_set = false;
  if (!attr_set) {
    SESRQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  kern<<<grid, (CTRL_WARPS + EPI_WARPS + L0_BUILD_WARPS) * 32, smem, st>>>(P, L.d_wumma, d_x, reinterpret_cast<uint4*>(d_out));
  SESRQ_LAUNCH_CHECK("conv_umma_l0_kernel");
  return SESRQ_OK;
}

static int launch_conv_umma_l0(const LayerDev& L, const float* d_x, uint8_t* d_out, int B, int H, int W, cudaStream_t st) {
  const LayerConst& c = L.c;
  if (!d_x || !d_out) return SESRQ_EINVAL;
  const int pad = c.pad, ns = L.umma_n / 16, N = L.umma_n;
  UmmaParams P;
  memset(&P, 0, sizeof(P));
  P.tiles_x = cdiv(W, UMMA_M - 2 * pad);
  P.TW = cdiv(W, P.tiles_x);
  const uint32_t w_bytes = (uint32_t)c.k * 2 * N * 16;
  const size_t fixed = ((w_bytes + 127) & ~127u) + 8 * (2 * MAX_STAGES + 2 * (TMEM_COLS / N)) + 64 + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const TilePlan tp = choose_tiling(H, B, P.tiles_x, pad, (size_t)L0_EP * 16, 8 * 16, (size_t)L0_SP * 4, fixed, sms);
  if (tp.stages < 2) return SESRQ_EUNSUPPORTED;
  P.TH = tp.TH;
  P.tiles_y = tp.tiles_y;
  P.stages = tp.stages;
  P.stage_bytes = (uint32_t)((((size_t)(P.TH + 2 * pad) * L0_EP + 8) * 16 + 1023) & ~(size_t)1023);
  const size_t smem = (size_t)P.stages * P.stage_bytes + (((size_t)(P.TH + 2 * pad) * L0_SP * 4 + 127) & ~(size_t)127) + fixed;
  if (smem > 227 * 1024) return SESRQ_EUNSUPPORTED;
  P.w_bytes = w_bytes;
  P.num_tiles = P.tiles_x * P.tiles_y * B;
  P.B = B; P.H = H; P.W = W; P.pad = pad;
  P.idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24);
  fill_epilogue_params(P, c, ns);
  P.cin = c.cin;
  P.pad_word = 0;
  for (int ci = 0; ci < c.cin; ++ci) P.pad_word |= (uint32_t)c.pad_byte << (8 * ci);
  P.in_scale = c.in_scale;
  P.in_zero = c.in_zero;
  const int grid = std::min(P.num_tiles, sms);
#define SESRQ_L0_NS(KS_, CIN_)                                                                   \
  do {                                                                                           \
    if (ns == 1) return launch_l0_one<KS_, 1, CIN_, 16>(P, L, d_x, d_out, grid, smem, st);       \
    if (ns == 2) return launch_l0_one<KS_, 2, CIN_, 16>(P, L, d_x, d_out, grid, smem, st);       \
    return launch_l0_one<KS_, 4, CIN_, 8>(P, L, d_x, d_out, grid, smem, st);                     \
  } while (0)
#define SESRQ_L0_CASE(KS_)                                                                       \
  do {                                                                                           \
    if (c.cin == 1) SESRQ_L0_NS(KS_, 1);                                                         \
    if (c.cin == 2) SESRQ_L0_NS(KS_, 2);                                                         \
    if (c.cin == 3) SESRQ_L0_NS(KS_, 3);                                                         \
    SESRQ_L0_NS(KS_, 4);                                                                         \
  } while (0)
  if (c.k == 3) SESRQ_L0_CASE(3);
  SESRQ_L0_CASE(5);
#undef SESRQ_L0_NS
#undef SESRQ_L0_CASE
}

int launch_conv_umma(const LayerDev& L, const float* d_x, const uint8_t* d_in, const uint8_t* d_shortcut,
                     uint8_t* d_out, int8_t* d_q5, float* d_y, int B, int H, int W, cudaStream_t st) {
  const LayerConst& c = L.c;
  if (is_l0_layout(c)) return launch_conv_umma_l0(L, d_x, d_out, B, H, W, st);
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return SESRQ_EUNSUPPORTED;
  if (!d_in) return SESRQ_EINVAL;
  const int pad = c.pad;
  const int ns = L.umma_n / 16;
  const int N = L.umma_n;

  // ---- tiling: 128-pixel pitch (one TMA box row = 2048 B), TH output rows per tile ----
  UmmaParams P;
  memset(&P, 0, sizeof(P));
  P.tiles_x = cdiv(W, UMMA_M - 2 * pad);
  P.TW = cdiv(W, P.tiles_x);
  const int taps = c.k * c.k, pairs = (taps + 1) / 2;
  const uint32_t w_bytes = (uint32_t)pairs * 2 * N * 16;
  const size_t fixed = ((w_bytes + 127) & ~127u) + 8 * (2 * MAX_STAGES + 2 * (TMEM_COLS / N)) + 64 + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const TilePlan tp = choose_tiling(H, B, P.tiles_x, pad, (size_t)UMMA_M * 16, (size_t)2 * pad * 16, 0, fixed, sms);
  if (tp.stages < 2) return SESRQ_EUNSUPPORTED;
  P.TH = tp.TH;
  P.tiles_y = tp.tiles_y;
  P.stages = tp.stages;
  P.stage_bytes = (uint32_t)((((size_t)(P.TH + 2 * pad) * UMMA_M + 2 * pad) * 16 + 1023) & ~(size_t)1023);
  P.box_bytes = (uint32_t)((P.TH + 2 * pad) * UMMA_M * 16);
  P.w_bytes = w_bytes;
  P.num_tiles = P.tiles_x * P.tiles_y * B;
  P.B = B; P.H = H; P.W = W; P.pad = pad;
  for (int pr = 0; pr < pairs; ++pr) {
    const int t0 = 2 * pr, t1 = 2 * pr + 1;
    const int o0 = (t0 / c.k) * UMMA_M + (t0 % c.k);
    const int o1 = (t1 < taps) ? (t1 / c.k) * UMMA_M + (t1 % c.k) : o0 + 1;   // dummy half: weights are zero
    P.pair_off[pr] = o0;
    P.pair_lbo[pr] = o1 - o0;
  }
  // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32, A = u8, B = s8, K-major both, N, M = 128
  P.idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24);
  fill_epilogue_params(P, c, ns);

  // ---- tensor map over the NHWC16 input viewed as (W*2 x 8-byte, H, B) ----
  CUtensorMap tmap;
  cuuint64_t gdim[3] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  cuuint32_t box[3] = {(cuuint32_t)UMMA_M * 2, (cuuint32_t)(P.TH + 2 * pad), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(d_in), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
    return SESRQ_ECUDA;
  }

  const int grid = std::min(P.num_tiles, sms);
  const size_t smem = (size_t)P.stage_bytes * P.stages + fixed;
  if (c.k == 3) return launch_ns<3>(ns, c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
  return launch_ns<5>(ns, c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st);
}

}  // namespace sesrq
