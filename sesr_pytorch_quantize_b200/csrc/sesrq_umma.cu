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
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "sesrq_internal.h"

#include "sesrq_umma_dev.cuh"

namespace sesrq {

// ------------------------------------------------------------------------------------------------
// exe_mode 0 (fake-quant) epilogue of one output pixel.  The four accumulator sets hold the per-PE integer partial sums
// of the operand bytes (pre-armed with MAGIC_BITS, so each word is the float 12582912 + A_p): y_p = f32(A_p) * cscale,
// float-domain PE clamp, sum of the four PEs, adder clamp, + fake-quantised bias, ReLU, AddOp -- the reference's op order
// (quan_func.py:330-333, :386, :431-434, :459; models/sesr.py:45), restated in oracle/fq_oracle.py.
//   SLOW: per-frame operand offset dmin != 0 (the frame's minimum is not code 0: zero padding is then NOT byte 0, so the
//         in-frame tap sums times dmin are added per pixel) and / or clamps that can fire.  Rare, kept out of line.
// ------------------------------------------------------------------------------------------------
// General form (SLOW): scalar, with the operand-offset correction and the clamps.
// Out of line, on a COPY of the accumulators in local memory (the caller's stay in registers for the fast form).
__device__ __noinline__ void epilogue_pixel_fq_slow(const FqArgs& fa, const FqFrameParams& f, const int* acc /*[4][16]*/,
                                                    float* o /*[16]*/, const int gy, const int gx, const int H, const int W,
                                                    const int ks) {
#pragma unroll 1
  for (int n = 0; n < 16; ++n) {
    const int co = chan_of_byte(n);                    // accumulator column n <-> output channel (PE-major byte order)
    float s = 0.f;
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
      float a = __fadd_rn(__int_as_float(acc[p * 16 + n]), -MAGIC_F);                        // exact: A_p as a float
      if (f.dmin != 0) {
        // d = byte + dmin and out-of-frame d = 0: add dmin * (sum of the PE's weights over the IN-FRAME taps); integer
        // arithmetic like the oracle's int64 conv, converted to float once
        const int pad = (ks - 1) / 2;
        long long corr = 0;
        for (int dy = 0; dy < ks; ++dy) {
          const int sy = gy + dy - pad;
          if (sy < 0 || sy >= H) continue;
          for (int dx = 0; dx < ks; ++dx) {
            const int sx = gx + dx - pad;
            if (sx < 0 || sx >= W) continue;
            corr += fa.tapsum[(dy * ks + dx) * 64 + p * 16 + n];
          }
        }
        a = (float)((long long)__float2int_rn(a) + corr * (long long)f.dmin);
      }
      float yp = __fmul_rn(a, f.cscale);
      yp = fminf(fmaxf(yp, f.lo_acc), f.hi_acc);                                             // :333
      s = __fadd_rn(s, yp);                                                                  // :386
    }
    s = fminf(fmaxf(s, f.lo_add), f.hi_add);                                                 // :434
    s = __fadd_rn(s, f.bias[co]);                                                            // :459
    if (fa.relu) s = fmaxf(s, 0.f);
    o[co] = s;
  }
}

// Fast form: no operand offset (dmin == 0).  Channel pairs (columns 2k, 2k + 1 = channels c, c + 4) on packed fp32
// pairs: exact A_p (FADD2), product (FMUL2), bias (FADD2); every lane is the scalar IEEE operation, so the result is
// bit-identical to the general form.  CLAMP: the float-domain PE / adder clamps can fire for this frame's operand range
// (decided per frame by fqb_params_kernel from the layer's weight mass), otherwise they are skipped as provably inert.
template <bool CLAMP>
__device__ __forceinline__ void epilogue_half_fq_fast(const FqArgs& fa, const float cscale, const float4 cl, const float2 (&bias2)[8],
                                                      const int h, const int (&a4)[4][8], float (&o)[16]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    // NOTE: ptxas contracts `mul.rn.f32x2` feeding `add.rn.f32x2` into one FFMA2 (seen in the SASS: a single rounding,
    // not the reference's two), and it also folds an FMA with a -0.0 addend back into that pattern.  The PE sum is
    // therefore written as an FMA with a multiplier 1.0 that only the host knows, RN(y * 1 + s) = RN(y + s).
    float2 s = make_float2(0.f, 0.f);
    const float2 one2 = make_float2(fa.one, fa.one);      // 1.0 as a run-time kernel parameter: ptxas cannot turn the FMA back into an add
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float2 a = f2_add(make_float2(__int_as_float(a4[p][2 * kk]), __int_as_float(a4[p][2 * kk + 1])), -MAGIC_F);
      float2 yp = f2_mul(a, cscale);
      if (CLAMP) {
        yp.x = fminf(fmaxf(yp.x, cl.x), cl.y);                                               // :333
        yp.y = fminf(fmaxf(yp.y, cl.x), cl.y);
      }
      s = f2_fma3(yp, one2, s);                                                              // :386 (first add: 0 + y_0)
    }
    if (CLAMP) {
      s.x = fminf(fmaxf(s.x, cl.z), cl.w);                                                   // :434
      s.y = fminf(fmaxf(s.y, cl.z), cl.w);
    }
    // bias pair k = 4h + kk
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
      if (hh == h) s = f2_add2(s, bias2[4 * hh + kk]);                                       // :459
    if (fa.relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
      if (hh == h) { o[chan_of_byte(8 * hh + 2 * kk)] = s.x; o[chan_of_byte(8 * hh + 2 * kk + 1)] = s.y; }
  }
}

// One output row of the exe_mode 0 path (the accumulator group is one row, GroupRows<4>): the 64 accumulator columns are
// read in two halves of 8 columns x 4 PEs (32 registers live instead of 64, so that 16 epilogue warps fit the register
// file), the slot is re-armed with the MAGIC constant, and the 16 channels go out as four float4.
__device__ __forceinline__ void epilogue_row_fq(const FqArgs& fa, const FqFrameParams* __restrict__ fpb, const float cscale,
                                                const float4 cl, const float2 (&bias2)[8], const int mode, const uint32_t taddr,
                                                float* __restrict__ xo, const float* __restrict__ xr, const bool valid,
                                                const int gy, const int gx, const int H, const int W, const int ks, float& vmin,
                                                float& vmax) {
  float o[16];
  if (mode == 2) {                          // operand offset: general form, out of line, on a copy in local memory
    int acc_l[64];
    float o_l[16];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int t16[16];
      tmem_ld16(taddr + p * 16, t16);
      tmem_ld_wait();
      reg_fence16(t16);
#pragma unroll
      for (int n = 0; n < 16; ++n) acc_l[p * 16 + n] = t16[n];
    }
    const FqFrameParams f = *fpb;
    epilogue_pixel_fq_slow(fa, f, acc_l, o_l, gy, gx, H, W, ks);
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = o_l[c];
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int a4[4][8];
#pragma unroll
      for (int p = 0; p < 4; ++p) tmem_ld8(taddr + p * 16 + 8 * h, a4[p]);
      tmem_ld_wait();
#pragma unroll
      for (int p = 0; p < 4; ++p)
        asm volatile("" : "+r"(a4[p][0]), "+r"(a4[p][1]), "+r"(a4[p][2]), "+r"(a4[p][3]), "+r"(a4[p][4]), "+r"(a4[p][5]),
                     "+r"(a4[p][6]), "+r"(a4[p][7]));
      if (mode == 1) epilogue_half_fq_fast<true>(fa, cscale, cl, bias2, h, a4, o);
      else epilogue_half_fq_fast<false>(fa, cscale, cl, bias2, h, a4, o);
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) tmem_st16_const(taddr + p * 16, MAGIC_BITS);     // re-arm; awaited by the caller
  if (xr != nullptr && valid) {                                                              // AddOp (models/sesr.py:45)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(xr) + q);
      const float2 lo = f2_add2(make_float2(o[4 * q], o[4 * q + 1]), make_float2(r4.x, r4.y));
      const float2 hi = f2_add2(make_float2(o[4 * q + 2], o[4 * q + 3]), make_float2(r4.z, r4.w));
      o[4 * q] = lo.x; o[4 * q + 1] = lo.y; o[4 * q + 2] = hi.x; o[4 * q + 3] = hi.y;
    }
  }
  if (valid) {
    if (fa.cout == 16) {
      float lo = o[0], hi = o[0];
#pragma unroll
      for (int c = 1; c < 16; ++c) { lo = fminf(lo, o[c]); hi = fmaxf(hi, o[c]); }           // 3-input FMNMX3 chains
      vmin = fminf(vmin, lo);
      vmax = fmaxf(vmax, hi);
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (c < fa.cout) { vmin = fminf(vmin, o[c]); vmax = fmaxf(vmax, o[c]); }
        else o[c] = 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      reinterpret_cast<float4*>(xo)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  }
}

// Rows of one completion group.  All RG rows are always processed: rows a frame's short bottom tile does not have
// (i >= cnt) still own pre-armed slots that no MMA touched, so loading and re-arming them is harmless and their stores
// are predicated off.  The loop therefore carries no row-count branches between the .sync.aligned TMEM instructions and
// exists once in the code (the final-layer bodies are large; code size matters).
template <int NS, int EPI, int RS, int COUT, int OUTM, bool FLOOR>
__device__ __forceinline__ void epilogue_group(const UmmaParams& P, const int cnt, const uint32_t taddr0, const int (&cinit)[NS][16],
                                               const uint4* __restrict__ sc_row, uint4* __restrict__ orow,
                                               int8_t* __restrict__ qb, float* __restrict__ yb, const long long HW,
                                               const float relu_floor, const bool valid, uint64_t* empty_bar, const int lane,
                                               const int wg, const FqArgs& fa, const FqFrameParams* __restrict__ fpb, const float fq_cscale,
                                               const float4 fq_cl, const long long fq_px,
                                               const int gy0, const int gx, float& vmin, float& vmax,
                                               const uint4* __restrict__ sc_smem, const float2 (&fq_bias2)[8], const int fq_slow) {
  constexpr int NCOLS = 16 * NS;
  constexpr int RG = GroupRows<NS>::value;
  if constexpr (EPI == EPI_FQ && NS == 4) {                   // exe_mode 0: one row per group, own TMEM access pattern
    static_assert(RG == 1, "the fake-quant epilogue handles one row per group");
    epilogue_row_fq(fa, fpb, fq_cscale, fq_cl, fq_bias2, fq_slow, taddr0, fa.xout + fq_px * 16,
                    fa.xres ? fa.xres + fq_px * 16 : nullptr, valid && cnt > 0, gy0, gx, P.H, P.W, 2 * P.pad + 1, vmin, vmax);
    return;
  }
  uint4 sc[RG];
#pragma unroll
  for (int i = 0; i < RG; ++i) {
    sc[i] = make_uint4(0u, 0u, 0u, 0u);
    if (EPI == EPI_RESIDUAL) {
      // TMA-staged tile (row pitch 128 pixels; out-of-frame pixels are zero-filled and their results never stored),
      // else straight from global memory
      if (sc_smem) sc[i] = sc_smem[i * UMMA_M];
      else if (valid && i < cnt) sc[i] = __ldg(sc_row + (long long)i * P.W);
    }
  }
  // software pipelining of the TMEM loads (next row's load in flight during this row's arithmetic).  The x2 last layer
  // writing only y (8 epilogue warps) is faster WITHOUT it: 0.378 -> 0.367 ms (both outputs: 0.523 -> 0.528, so it keeps it)
  constexpr bool PIPE = !(EPI == EPI_FINAL && (X_NOPIPE || (OUTM == 2 && RS == 2)));
  constexpr bool BRANCH = X_BRANCH && EPI == EPI_FINAL;
  int acc[PIPE ? 2 : 1][NS][16];
  if (PIPE) slot_load<NS>(taddr0, acc[0]);
#pragma unroll
  for (int i = 0; i < RG; ++i) {
    constexpr int CUR = 0;
    const int cur = PIPE ? (i & 1) : CUR;
    if (!PIPE) slot_load<NS>(taddr0 + (uint32_t)(i * NCOLS), acc[0]);
    tmem_ld_wait();
    slot_fence<NS>(acc[cur]);
    if (PIPE && i + 1 < RG) slot_load<NS>(taddr0 + (uint32_t)((i + 1) * NCOLS), acc[(i + 1) & (PIPE ? 1 : 0)]);   // next row in flight
    slot_store<NS>(taddr0 + (uint32_t)(i * NCOLS), cinit);      // re-arm the slot; awaited once per group
    const bool v_i = valid && i < cnt;
    if (SESRQ_X_ROWSYNC == 2 || (SESRQ_X_ROWSYNC == 1 && EPI == EPI_FINAL)) asm volatile("bar.sync %0, 128;" ::"r"(2 + wg) : "memory");
    if (BRANCH) {
      if (v_i) epilogue_pixel<NS, EPI, RS, COUT, OUTM, FLOOR>(P, acc[cur], sc[i], orow, qb, yb, HW, relu_floor, true);
      __syncwarp();
    } else {
      epilogue_pixel<NS, EPI, RS, COUT, OUTM, FLOOR>(P, acc[cur], sc[i], orow, qb, yb, HW, relu_floor, v_i);
    }
    if (EPI == EPI_FINAL) {
      if (OUTM != 2) qb += P.W;
      if (OUTM != 1) yb += (long long)(RS * RS) * P.W;
    } else {
      orow += P.W;
    }
    if (X_ROWARRIVE) {
      tmem_st_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(empty_bar);
    }
  }
}

// Epilogue warps, shared by both kernels.  A row slot cycles:  pre-loaded with c_init (tcgen05.st)  ->  the issuer
// accumulates one 128-pixel output row on top  ->  tcgen05.ld, immediate re-load of c_init.  Completion (acc_full ring,
// indexed by the running group counter) and re-arming (acc_empty, one barrier per group of RG slots, 4 arrivals = the
// warps of the consuming warpgroup) are signalled per group of RG consecutive rows of a tile; warpgroup wg consumes the
// groups G == wg (mod NWG) and software-pipelines the TMEM loads of a group's rows.
template <int NS, int EPI, int RS, int COUT, int OUTM, bool FLOOR, int EPI_WARPS, bool PIN>
__device__ __forceinline__ void epilogue_warps(const UmmaParams& P, const int ewarp, const int lane, const uint32_t tmem_base,
                                               uint64_t* acc_full, uint64_t* acc_empty,
                                               const uint4* __restrict__ shortcut, uint4* __restrict__ out,
                                               int8_t* __restrict__ q5, float* __restrict__ y, const FqArgs& fa,
                                               const uint8_t* sc_buf = nullptr, uint64_t* sc_full = nullptr,
                                               uint64_t* sc_empty = nullptr) {
  constexpr int NCOLS = 16 * NS;
  constexpr int NSLOT = TMEM_COLS / NCOLS;
  constexpr int NWG = EPI_WARPS / 4;
  constexpr int RG = GroupRows<NS>::value;
  constexpr int NGS = NSLOT / RG;               // group slots
  int sst = 0;                                  // EPI_RESIDUAL: shortcut stage ring position / phase parity
  uint32_t sph = 0;
  const bool sc_tma = (EPI == EPI_RESIDUAL) && P.sc_stages > 0;
  const int wg = ewarp >> 2;                    // warpgroup
  const int quarter = ewarp & 3;                // TMEM lane quarter this warp may access (== warp id % 4)
  const int ox = quarter * 32 + lane;           // pitch == UMMA_M: a slot is one output row of the tile, lane is the column
  const long long HW = (long long)P.H * P.W;
  const float relu_floor = P.relu ? 0.f : -INFINITY;
  const int tiles_per_frame = P.tiles_x * P.tiles_y;
  const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16);

  int cinit[NS][16];
#pragma unroll
  for (int p = 0; p < NS; ++p)
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      cinit[p][n] = (NS == 2 && p == 1 && n >= 4) ? 0 : P.c_init[p][n];
      // PIN keeps the constants in registers: without it the compiler re-reads them from the constant bank before
      // the re-arms of every group (8 LDC.64); it is only set where the register budget has the room (no spills)
      if (PIN) asm volatile("" : "+r"(cinit[p][n]));
    }
  for (int g = wg; g < NGS; g += NWG)
#pragma unroll
    for (int i = 0; i < RG; ++i) slot_store<NS>(tq + (uint32_t)((g * RG + i) * NCOLS), cinit);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0)
    for (int g = wg; g < NGS; g += NWG)
      for (int i = 0; i < (X_ROWARRIVE ? RG : 1); ++i) mbar_arrive(&acc_empty[g]);

  constexpr int GPB = TB / RG;
  const uint32_t nring = (uint32_t)ring_slots(NSLOT, P.TH);
  RingPos ring;
  uint32_t G0 = 0;                              // groups of all earlier tiles
  for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x, ring.advance((uint32_t)P.TH, nring)) {
    const int tt = P.dbg < 0 ? P.num_tiles - 1 - t : t;      // tile order (launch_conv_umma: alternates from layer to layer)
    const int b = tt / tiles_per_frame;
    const int r = tt - b * tiles_per_frame;
    const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
    const int y0 = ty * P.TH, gx = tx * P.TW + ox;
    const int nblk = min(P.TH, P.H - y0);
    const int ngrp = tile_groups<RG>(nblk);
    const bool valid = (ox < P.TW) && (gx < P.W);
    const long long pix0 = (long long)b * HW + (long long)y0 * P.W + gx;
    // final layer: per-tile output bases (row mb adds mb * W bytes / mb * RS^2 * W floats)
    int8_t* qb_tile = nullptr;
    float* yb_tile = nullptr;
    if (EPI == EPI_FINAL) {
      constexpr int R2 = RS * RS;
      if (OUTM != 2) qb_tile = q5 + (long long)b * P.cout * HW + (long long)y0 * P.W + gx;
      if (OUTM != 1) yb_tile = y + (long long)b * (P.cout / R2) * (HW * R2) + (long long)y0 * RS * (P.W * RS) + (long long)gx * RS;
    }
    // exe_mode 0: this tile's frame constants, and the running (min, max) of what this warp produces for the frame
    const FqFrameParams* fpb = nullptr;
    float vmin = INFINITY, vmax = -INFINITY;
    float2 fq_bias2[8];                     // this frame's fake-quantised bias in accumulator-column pairs
    float fq_cscale = 0.f;
    float4 fq_cl = make_float4(0.f, 0.f, 0.f, 0.f);
    int fq_slow = 0;                        // 0: fast, 1: fast with clamps, 2: general form (operand offset)
    if (EPI == EPI_FQ) {
      fpb = fa.fp + b;
      fq_slow = (fpb->dmin != 0) ? 2 : ((fpb->clamp_acc || fpb->clamp_add) ? 1 : 0);   // per frame: warp-uniform
      fq_cscale = fpb->cscale;
      fq_cl = make_float4(fpb->lo_acc, fpb->hi_acc, fpb->lo_add, fpb->hi_add);
#pragma unroll
      for (int k = 0; k < 8; ++k) fq_bias2[k] = make_float2(fpb->bias[chan_of_byte(2 * k)], fpb->bias[chan_of_byte(2 * k + 1)]);
    }
    const uint4* sc_tile = nullptr;
    if (sc_tma) {                                 // this tile's shortcut rows have landed in shared memory
      mbar_wait(&sc_full[sst], sph);
      sc_tile = reinterpret_cast<const uint4*>(sc_buf + (size_t)sst * P.sc_stage_bytes) + ox;
    }
    const int first = (int)((wg + NWG - (G0 % NWG)) % NWG);
    for (int j = first; j < ngrp; j += NWG) {
      const uint32_t G = G0 + j;
      const int mb0 = (j / GPB) * TB + (j % GPB) * RG;          // tile-local group j = band * GPB + group within the band
      const int cnt = min(RG, nblk - mb0);
      const uint32_t slot0 = ring.base + (uint32_t)mb0;         // a multiple of RG
      const uint32_t taddr0 = tq + slot0 * NCOLS;
      const long long prow = pix0 + (long long)mb0 * P.W;
      int8_t* const qb = (EPI == EPI_FINAL && OUTM != 2) ? qb_tile + (long long)mb0 * P.W : nullptr;
      float* const yb = (EPI == EPI_FINAL && OUTM != 1) ? yb_tile + (long long)(mb0 * RS * RS) * P.W : nullptr;
      __syncwarp();                        // the .sync.aligned TMEM accesses below need a converged warp
      mbar_wait(&acc_full[G % NSLOT], (G / NSLOT) & 1);
      tc_fence_after();
      epilogue_group<NS, EPI, RS, COUT, OUTM, FLOOR>(P, cnt, taddr0, cinit, shortcut + prow, out + prow, qb, yb, HW, relu_floor, valid,
                                                     &acc_empty[slot0 / RG], lane, wg, fa, fpb, fq_cscale, fq_cl, prow, y0 + mb0, gx, vmin, vmax,
                                                     sc_tile ? sc_tile + mb0 * UMMA_M : nullptr, fq_bias2, fq_slow);
      if (!X_ROWARRIVE) {
        tmem_st_wait();                    // warp-collective: every lane's re-arm stores have landed
        tc_fence_before();
        if (lane == 0) mbar_arrive(&acc_empty[slot0 / RG]);
      }
    }
    G0 += ngrp;
    if (sc_tma) {                           // every epilogue warp releases the shortcut buffer, with or without groups in the tile
      __syncwarp();
      if (lane == 0) mbar_arrive(&sc_empty[sst]);
      if (++sst == P.sc_stages) { sst = 0; sph ^= 1; }
    }
    if (EPI == EPI_FQ) {                    // one atomic pair per warp and tile (a tile never spans frames)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      }
      if (lane == 0 && vmin <= vmax) {
        atomic_min_float(fa.mm + (long long)b * fa.mm_stride, vmin);
        atomic_max_float(fa.mm + (long long)b * fa.mm_stride + 1, vmax);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// MMA issuer (one thread per CTA), shared by both kernels.
//
// An M128 x K32 tcgen05.mma streams 4 KB of A operand from shared memory whatever N is (measured: ~30 cycles
// per instruction at N = 16, profiles/r1 pipeline experiments), so the cost of a layer is the number of MMAs.
// The vertical taps are therefore batched into N: input row r of the tile contributes to the KS output rows
// o = r - dy, whose accumulators are KS CONSECUTIVE TMEM slots, so one MMA with N = KS * NCOLS per
// (input row, horizontal tap pair) updates all of them at once:
//     D[slot(o)][pixel X][co] += sum_{dx in pair, c} in[r][X + dx][c] * W[dy = r - o][dx][co]
// The B image holds, per tap pair and K half, the row blocks in order dy = KS-1 .. 0 so that ascending output
// rows see ascending D columns; tile edges simply use a sub-range of blocks (smaller N, offset B start).
// 3x3: 2 MMAs per row instead of 5; 5x5: 3 instead of 13; layer 0 (K = all 8 dx of one row): 1 instead of 5.
// Horizontal shifts stay in the A descriptor (start address = pixel offset, LBO = distance to the second tap).
// ------------------------------------------------------------------------------------------------
// What limits this kernel is the instruction stream of the issuing THREAD, not the tensor pipe (tools/mma_probe*.cu
// on B200: an M128 x K32 kind::i8 MMA occupies the pipe for ~32 + N/4 cycles -- both operands stream from shared
// memory at 128 B/clk -- for any accumulator address pattern, but a lone thread issues at most one MMA per ~51
// cycles, pays ~5 cycles per additional dependent instruction, ~185 per tcgen05.commit and ~200 per
// mbarrier.try_wait + tcgen05.fence).  Hence:
//   * a tile's output rows are split into `nbw` bands of TB rows, each issued by its own warp (warps 1..nbw).  A
//     band's MMAs touch only that band's accumulator slots (edge rows use a sub-range of the B row blocks), so the
//     warps are independent: no ordering between them, each tracks its own completions;
//   * slot readiness (acc_empty, one barrier per group of RG row slots) is polled once per band by TB / RG lanes at once;
//   * completion is signalled per GROUP of RG consecutive rows (one tcgen05.commit per group; barrier ring acc_full
//     indexed by a running group counter); an epilogue warpgroup consumes one group at a time;
//   * the code for a full band is completely unrolled: every operand is `band base + compile-time offset`.
// MMAs of input row r (band-relative, 0 .. rows + KS - 2) of a band with `rows` output rows.
//   output rows o = r - dy, clipped to [0, rows); B row block jblk = o - (r - KS + 1)  <->  dy = KS - 1 - jblk.
template <int KS, int NS, int NPAIR, int ROW_UNITS, int PAIR_STEP>
__device__ __forceinline__ void issue_row(const int r, const int rows, const uint32_t d_band, const uint32_t a_band_lo,
                                          const uint32_t w_units, const uint32_t idesc0) {
  constexpr int NCOLS = 16 * NS;
  constexpr int NTOT = KS * NCOLS;                           // rows of the B image per K half
  constexpr int JMAX = (256 / NCOLS) < KS ? (256 / NCOLS) : KS;   // row blocks per instruction (N <= 256)
  // cute::UMMA::SmemDescriptor high word: SBO = 128 B (bits 32-45), version 1 (bit 46), no swizzle
  const uint32_t desc_hi = (128u >> 4) | (1u << 14);
  const uint32_t b_lbo = (uint32_t)NTOT << 16;
  const int o_first = r - (KS - 1);
  const int o_lo = o_first < 0 ? 0 : o_first;
  const int o_hi = r < rows - 1 ? r : rows - 1;
#pragma unroll
  for (int o = o_lo; o <= o_hi; o += JMAX) {
    const int cnt = (o_hi - o + 1) < JMAX ? (o_hi - o + 1) : JMAX;
    const int jblk = o - o_first;
    const uint32_t idesc = idesc0 | ((uint32_t)((cnt * NCOLS) >> 3) << 17);
#pragma unroll
    for (int q = 0; q < NPAIR; ++q)
      umma_i8(d_band + (uint32_t)(o * NCOLS), make_desc(a_band_lo + (uint32_t)(r * ROW_UNITS + q * PAIR_STEP), desc_hi),
              make_desc((w_units + (uint32_t)(q * 2 * NTOT + jblk * NCOLS)) | b_lbo, desc_hi), idesc, 1u);
  }
}

// One band of one tile.  Called by the whole issuer warp (converged); lane 0 issues the tcgen05 instructions.
// nblk: rows of the tile that exist; ring: slot base / parity of the tile; Gt: groups of all earlier tiles.
template <int KS, int NS, int NPAIR, int ROW_UNITS, int PAIR_STEP>
__device__ __forceinline__ void mma_issue_band(const int bw, const int lane, const uint32_t tmem_base, const uint32_t a_tile,
                                               const uint32_t a_lbo, const uint32_t w_units, const uint32_t idesc0,
                                               const int nblk, const RingPos ring, const uint32_t Gt,
                                               const uint32_t acc_full_addr, const uint32_t acc_empty_addr) {
  constexpr int NCOLS = 16 * NS;
  constexpr int NSLOT = TMEM_COLS / NCOLS;
  constexpr int RG = GroupRows<NS>::value;
  constexpr int GPB = TB / RG;
  const int rows = min(TB, nblk - bw * TB);                  // output rows of this band that exist (may be <= 0)
  const uint32_t slot0 = ring.base + (uint32_t)(bw * TB);
  // ---- NS = 4 (one row per group, the whole accumulator ring is ONE tile): waiting for all 8 slots of the tile before the
  // first MMA would serialise this tile's MMAs behind the previous tile's whole epilogue.  Wait two slots at a time, right
  // before the source rows that touch them first (source row r touches the output rows <= r), so that the first rows of
  // a tile are issued while the epilogue is still draining the last rows of the previous one. ----
  if (RG == 1 && rows == TB) {
    const uint32_t d_band = tmem_base + slot0 * NCOLS;
    const uint32_t a_band_lo = (a_tile + (uint32_t)(bw * TB * ROW_UNITS)) | (a_lbo << 16);
    const uint32_t Gb = Gt + (uint32_t)(bw * GPB);
#pragma unroll
    for (int q = 0; q < TB / 2; ++q) {
      {
        const uint32_t bar = acc_empty_addr + (slot0 + (uint32_t)(2 * q + (lane & 1))) * 8u;
        uint32_t ok;
        do {
          ok = 1u;
          if (lane < 2)
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(ok) : "r"(bar), "r"(ring.par), "r"(MBAR_SUSPEND_NS) : "memory");
        } while (!__all_sync(0xffffffffu, ok != 0u));
        tc_fence_after();
      }
      if (elect_one()) {
#pragma unroll
        for (int r = 2 * q; r < (q == TB / 2 - 1 ? TB + KS - 1 : 2 * q + 2); ++r) {
          issue_row<KS, NS, NPAIR, ROW_UNITS, PAIR_STEP>(r, TB, d_band, a_band_lo, w_units, idesc0);
          const int o_done = r - (KS - 1);
          if (o_done >= 0) umma_commit_addr(acc_full_addr + ((Gb + (uint32_t)o_done) % NSLOT) * 8u);
        }
      }
      __syncwarp();
    }
    return;
  }
  // ---- all groups of the band pre-loaded?  (lane l < GPB <-> group l of the band; acc_empty is per group of RG slots) ----
  {
    const int gl = lane & (GPB - 1);
    const uint32_t bar = acc_empty_addr + (slot0 / RG + (uint32_t)gl) * 8u;
    uint32_t ok;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(ok) : "r"(bar), "r"(ring.par), "r"(MBAR_SUSPEND_NS) : "memory");
    } while (!__all_sync(0xffffffffu, ok != 0u));
    tc_fence_after();
    // groups the tile does not have: nobody will consume their slots, so complete the phase here (4 = arrivals of a
    // consuming warpgroup); the slots stay pre-loaded
    if (lane < GPB && lane * RG >= max(rows, 0))
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(X_ROWARRIVE ? 4 * RG : 4) : "memory");
  }
  if (rows > 0 && elect_one()) {          // elect.sync: the compiler keeps the block on the uniform datapath
    const uint32_t d_band = tmem_base + slot0 * NCOLS;
    const uint32_t a_band_lo = (a_tile + (uint32_t)(bw * TB * ROW_UNITS)) | (a_lbo << 16);
    const uint32_t Gb = Gt + (uint32_t)(bw * GPB);
    if (rows == TB) {
#pragma unroll
      for (int r = 0; r < TB + KS - 1; ++r) {
        issue_row<KS, NS, NPAIR, ROW_UNITS, PAIR_STEP>(r, TB, d_band, a_band_lo, w_units, idesc0);
        const int o_done = r - (KS - 1);
        if (o_done >= 0 && (o_done % RG) == RG - 1) umma_commit_addr(acc_full_addr + ((Gb + (uint32_t)(o_done / RG)) % NSLOT) * 8u);
      }
    } else {
      for (int r = 0; r < rows + KS - 1; ++r) {
        issue_row<KS, NS, NPAIR, ROW_UNITS, PAIR_STEP>(r, rows, d_band, a_band_lo, w_units, idesc0);
        const int o_done = r - (KS - 1);
        if (o_done >= 0 && ((o_done % RG) == RG - 1 || o_done == rows - 1))
          umma_commit_addr(acc_full_addr + ((Gb + (uint32_t)(o_done / RG)) % NSLOT) * 8u);
      }
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// the kernel.  KS: kernel size; NS: accumulator sets (1/2/4, see header); EPI: EpilogueKind; RS: PixelShuffle
// factor of the final epilogue (1 otherwise); EPI_WARPS: number of epilogue warps.
// ------------------------------------------------------------------------------------------------
template <int KS, int NS, int EPI, int RS, int COUT, int OUTM, bool FLOOR, int EPI_WARPS>
__global__ void __launch_bounds__((CTRL_WARPS + EPI_WARPS) * 32, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ UmmaParams P,
                 const uint8_t* __restrict__ w_img, const uint4* __restrict__ shortcut, uint4* __restrict__ out,
                 int8_t* __restrict__ q5, float* __restrict__ y, const __grid_constant__ FqArgs fa,
                 const __grid_constant__ CUtensorMap smap) {
  constexpr int NCOLS = 16 * NS;
  constexpr int NSLOT = TMEM_COLS / NCOLS;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const tile0 = smem;
  uint8_t* const sc0 = smem + (size_t)P.stages * P.stage_bytes;              // EPI_RESIDUAL: shortcut tile buffers
  uint8_t* const w_s = sc0 + (size_t)P.sc_stages * P.sc_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + ((P.w_bytes + 127) & ~127u));
  uint64_t* tile_full = bars;                            // [MAX_STAGES]
  uint64_t* tile_empty = bars + MAX_STAGES;              // [MAX_STAGES]
  uint64_t* acc_full = bars + 2 * MAX_STAGES;            // [NSLOT]
  uint64_t* acc_empty = bars + 2 * MAX_STAGES + NSLOT;   // [NSLOT]
  uint64_t* sc_full = bars + 2 * MAX_STAGES + 2 * NSLOT; // [MAX_STAGES]
  uint64_t* sc_empty = sc_full + MAX_STAGES;             // [MAX_STAGES]
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(sc_empty + MAX_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup (overlaps the tail of the previous kernel of the stream: nothing here depends on it) ----
  pdl_launch_dependents();
  for (uint32_t i = threadIdx.x; i < P.w_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(w_s)[i] = __ldg(reinterpret_cast<const uint4*>(w_img) + i);
  fence_proxy_async();                      // weights were written through the generic proxy, UMMA reads via async proxy
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&tile_full[i], 1);
      mbar_init(&tile_empty[i], P.nbw);
      mbar_init(&sc_full[i], 1);
      mbar_init(&sc_empty[i], EPI_WARPS);
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], X_ROWARRIVE ? 4 * GroupRows<NS>::value : 4);   // one elected arrival per epilogue warp of the consuming warpgroup
    }
    if (EPI == EPI_RESIDUAL && P.sc_stages > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&smap)) : "memory");
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  pdl_wait();                               // from here on the activations the previous kernel wrote are read

  const int tiles_per_frame = P.tiles_x * P.tiles_y;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int st = 0, sst = 0;
      uint32_t ph = 0, sph = 0;              // phase parity of the current pass over the stage rings
      for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
        const int tt = P.dbg < 0 ? P.num_tiles - 1 - t : t;      // tile order (launch_conv_umma: alternates from layer to layer)
        mbar_wait(&tile_empty[st], ph ^ 1);
        const int b = tt / tiles_per_frame;
        const int r = tt - b * tiles_per_frame;
        const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
        mbar_arrive_expect_tx(&tile_full[st], P.box_bytes);
        tma_load_3d(tile0 + (size_t)st * P.stage_bytes, &tmap, (tx * P.TW - P.pad) * 2, ty * P.TH - P.pad, b, &tile_full[st]);
        if (++st == P.stages) { st = 0; ph ^= 1; }
        if (EPI == EPI_RESIDUAL && P.sc_stages > 0) {
          // the shortcut pixels of the tile's OUTPUT rows (no halo): consumed by the epilogue, so the 0.5 GB re-read
          // overlaps the MMAs instead of sitting in front of the residual arithmetic
          mbar_wait(&sc_empty[sst], sph ^ 1);
          mbar_arrive_expect_tx(&sc_full[sst], P.sc_stage_bytes);
          tma_load_3d(sc0 + (size_t)sst * P.sc_stage_bytes, &smap, tx * P.TW * 2, ty * P.TH, b, &sc_full[sst]);
          if (++sst == P.sc_stages) { sst = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp < CTRL_WARPS) {
    // ================= MMA issuer warps 1 .. nbw (one band of every tile each) =================
    const int bw = warp - 1;
    if (bw < P.nbw) {
      constexpr int NPAIR = (KS + 1) / 2;
      constexpr int RG = GroupRows<NS>::value;
      const uint32_t w_units = (smem_u32(w_s) & 0x3FFFFu) >> 4;
      const uint32_t tile_addr0 = (smem_u32(tile0) & 0x3FFFFu) >> 4;
      const uint32_t nring = (uint32_t)ring_slots(NSLOT, P.TH);
      int st = 0;
      uint32_t ph = 0;
      uint32_t Gt = 0;
      RingPos ring;
      for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
        const int tt = P.dbg < 0 ? P.num_tiles - 1 - t : t;      // tile order (launch_conv_umma: alternates from layer to layer)
        const int r = tt % tiles_per_frame;
        const int nblk = min(P.TH, P.H - (r / P.tiles_x) * P.TH);     // rows of this tile that exist
        mbar_wait(&tile_full[st], ph);
        tc_fence_after();
        // A: row pitch 128 pixels; pair q = taps (2q, 2q+1) of the row: start 2q pixels in, LBO = one pixel
        mma_issue_band<KS, NS, NPAIR, UMMA_M, 2>(bw, lane, tmem_base, tile_addr0 + (uint32_t)st * (P.stage_bytes >> 4), 1u,
                                                 w_units, P.idesc, nblk, ring, Gt, smem_u32(acc_full), smem_u32(acc_empty));
        if (lane == 0) umma_commit(&tile_empty[st]);     // fires when this warp's MMAs on the tile have retired
        __syncwarp();
        Gt += tile_groups<RG>(nblk);
        ring.advance((uint32_t)P.TH, nring);
        if (++st == P.stages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    constexpr bool PIN = !X_NOPIN && NS == 1 && (EPI == EPI_PLAIN || (EPI == EPI_FINAL && OUTM != 0));
    epilogue_warps<NS, EPI, RS, COUT, OUTM, FLOOR, EPI_WARPS, PIN>(P, warp - CTRL_WARPS, lane, tmem_base, acc_full, acc_empty,
                                                                   shortcut, out, q5, y, fa, sc0, sc_full, sc_empty);
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
#ifndef SESRQ_L0_BUILD_WARPS
#define SESRQ_L0_BUILD_WARPS 8
#endif
#ifndef SESRQ_L0_EPI_WARPS
#define SESRQ_L0_EPI_WARPS 12        // epilogue warps of the NS = 1 layer-0 kernel
#endif
constexpr int L0_BUILD_WARPS = SESRQ_L0_BUILD_WARPS;
constexpr int L0_STATIC_SMEM = ((L0_BUILD_WARPS * (128 + 4) * 4 + 1023) / 1024) * 1024;   // stage_w (static shared memory), rounded up
constexpr int L0_XP = 136;               // fp32 tile row pitch in floats = TMA box width: 132 pixels + up to 3 of alignment
                                         // (the box must start on a 16-byte boundary: an unaligned inner coordinate faults)
constexpr int L0_XSTAGES = 4;            // fp32 tile buffers (upper bound; P.xstages are used)

// ExactDiv (sesrq_common.cuh) on two quotients per instruction (FMUL2 / FFMA2): lane-wise identical to the scalar sequence
// max of three, NaN if any operand is NaN (max.NaN.f32, 3-input form of sm_100)
__device__ __forceinline__ float max3_nan(float a, float b, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// The operands must already lie in [-lim, lim] (ExactDiv's range clamp): the row builder checks a whole row's worth of
// values with a handful of 3-input maxima and clamps only in the (never seen in image data) case that one is outside.
__device__ __forceinline__ float2 exact_div2(const ExactDiv& d, float2 x) {
  const float2 q0 = f2_mul(x, d.rcp);
  const float2 e = f2_fma3(make_float2(-d.s, -d.s), q0, x);
  return f2_fma3(make_float2(d.rcp, d.rcp), e, q0);
}

template <int KS, int NS, int CIN, bool FLOOR, int EPI_WARPS>
__global__ void __launch_bounds__((CTRL_WARPS + EPI_WARPS + L0_BUILD_WARPS) * 32, 1)
conv_umma_l0_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ UmmaParams P,
                    const uint8_t* __restrict__ w_img, const float* __restrict__ x, uint4* __restrict__ out) {
  constexpr int NCOLS = 16 * NS;
  constexpr int NSLOT = TMEM_COLS / NCOLS;
  constexpr int PAD = (KS - 1) / 2;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const e0 = smem;                                         // E planes: P.stages buffers of stage_bytes
  // second region: the fp32 tile buffers (TMA path) or the S plane (fallback path)
  uint8_t* const region2 = smem + (size_t)P.stages * P.stage_bytes;
  uint32_t* const s_plane = reinterpret_cast<uint32_t*>(region2);
  const uint32_t s_bytes = (uint32_t)(((P.TH + 2 * PAD) * L0_SP * 4 + 127) & ~127);
  const uint32_t r2_bytes = P.x_tma ? (uint32_t)P.xstages * P.xstage_bytes : s_bytes;
  uint8_t* const w_s = region2 + r2_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + ((P.w_bytes + 127) & ~127u));
  uint64_t* e_full = bars;                               // [MAX_STAGES] count = builder warps
  uint64_t* e_empty = bars + MAX_STAGES;                 // [MAX_STAGES] count = MMA warps (tcgen05.commit)
  uint64_t* acc_full = bars + 2 * MAX_STAGES;
  uint64_t* acc_empty = bars + 2 * MAX_STAGES + NSLOT;
  uint64_t* x_full = bars + 2 * MAX_STAGES + 2 * NSLOT;  // [L0_XSTAGES] TMA transaction barriers
  uint64_t* x_empty = x_full + L0_XSTAGES;               // [L0_XSTAGES] count = builder warps
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(x_empty + L0_XSTAGES);

  __shared__ uint32_t stage_w[L0_BUILD_WARPS][UMMA_M + 4];   // builders: the pixel words of the tile row in flight
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();                               // the next layer may start its prologue now
  for (uint32_t i = threadIdx.x; i < P.w_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(w_s)[i] = __ldg(reinterpret_cast<const uint4*>(w_img) + i);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    if (P.x_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&e_full[i], L0_BUILD_WARPS);
      mbar_init(&e_empty[i], P.nbw);
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], X_ROWARRIVE ? 4 * GroupRows<NS>::value : 4);
    }
    for (int i = 0; i < L0_XSTAGES; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], L0_BUILD_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  pdl_wait();                               // from here on the activations the previous kernel wrote are read
  const int tiles_per_frame = P.tiles_x * P.tiles_y;
  const long long HW = (long long)P.H * P.W;

  if (warp == 0) {
    // ================= TMA producer: fp32 halo tiles, one box of (L0_XP, TH + 2 pad, CIN) per tile.  Out-of-frame
    // elements are zero-filled, and x = 0 quantises to max(z0, -128), the reference's pad code (:351-352) =========
    if (P.x_tma && lane == 0) {
      int xs = 0;
      uint32_t xph = 0;
      for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
        const int tt = P.dbg < 0 ? P.num_tiles - 1 - t : t;      // tile order (launch_conv_umma: alternates from layer to layer)
#ifdef SESRQ_DEBUG_KNOBS
        if (P.dbg & 512) break;
#endif
        mbar_wait(&x_empty[xs], xph ^ 1);
        const int b = tt / tiles_per_frame;
        const int r = tt - b * tiles_per_frame;
        const int ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
        mbar_arrive_expect_tx(&x_full[xs], (uint32_t)(CIN * (P.TH + 2 * PAD) * L0_XP * 4));
        tma_load_4d(region2 + (size_t)xs * P.xstage_bytes, &xmap, (tx * P.TW - PAD) & ~3, ty * P.TH - PAD, 0, b, &x_full[xs]);
        if (++xs == P.xstages) { xs = 0; xph ^= 1; }
      }
    }
  } else if (warp < CTRL_WARPS) {
    // ================= MMA issuer warps 1 .. nbw =================
    const int bw = warp - 1;
    if (bw < P.nbw) {
      constexpr int RG = GroupRows<NS>::value;
      const uint32_t w_units = (smem_u32(w_s) & 0x3FFFFu) >> 4;
      const uint32_t e_addr0 = (smem_u32(e0) & 0x3FFFFu) >> 4;
      const uint32_t nring = (uint32_t)ring_slots(NSLOT, P.TH);
      int st = 0;
      uint32_t ph = 0;
      uint32_t Gt = 0;
      RingPos ring;
      for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
        const int tt = P.dbg < 0 ? P.num_tiles - 1 - t : t;      // tile order (launch_conv_umma: alternates from layer to layer)
        const int r = tt % tiles_per_frame;
        const int nblk = min(P.TH, P.H - (r / P.tiles_x) * P.TH);
        mbar_wait(&e_full[st], ph);
        tc_fence_after();
        // A: E plane, row pitch L0_EP entries; one K = 32 chunk covers taps dx = 0..7 of a row: LBO = 4 entries
        mma_issue_band<KS, NS, 1, L0_EP, 0>(bw, lane, tmem_base, e_addr0 + (uint32_t)st * (P.stage_bytes >> 4), 4u, w_units,
                                            P.idesc, nblk, ring, Gt, smem_u32(acc_full), smem_u32(acc_empty));
        if (lane == 0) umma_commit(&e_empty[st]);
        __syncwarp();
        Gt += tile_groups<RG>(nblk);
        ring.advance((uint32_t)P.TH, nring);
        if (++st == P.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp < CTRL_WARPS + EPI_WARPS) {
    // ================= epilogue (shared with the 16-channel kernel, EPI_PLAIN) =================
    const FqArgs no_fq{};
    epilogue_warps<NS, EPI_PLAIN, 1, 0, 0, FLOOR, EPI_WARPS, false>(P, warp - CTRL_WARPS, lane, tmem_base, acc_full, acc_empty, nullptr,
                                                                    out, nullptr, nullptr, no_fq);
  } else {
    // ================= builders =================
    const int bt = threadIdx.x - (CTRL_WARPS + EPI_WARPS) * 32;       // 0 .. 255
    constexpr int NB = L0_BUILD_WARPS * 32;
    const ExactDiv div_s0(P.in_scale);
    const float z0 = P.in_zero;
    int st = 0;
    uint32_t ph = 0;
    if (P.x_tma) {
      // ---- TMA path: one warp turns one tile row at a time into operand entries.  The row's 128 pixels (4 passes of 32
      // lanes, every address = row base + immediate) are quantised on packed pairs and staged as 4-byte words in shared
      // memory; entry E[row][X] = (w(X), w(X+1), w(X+2), w(X+3)) is then one own word + three staged neighbours (warp
      // shuffles were measured ~5x slower next to the UMMA traffic).  Words 128..131 only meet zero weights. ----
      const int bwid = bt >> 5;
      const int rp = P.TH + 2 * PAD;                                   // rows per fp32 plane
      uint32_t* const sw = stage_w[bwid];
      if (lane < 4) sw[UMMA_M + lane] = 0u;
      int xs = 0;
      uint32_t xph = 0;
      for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
        const int tt = P.dbg < 0 ? P.num_tiles - 1 - t : t;      // tile order (launch_conv_umma: alternates from layer to layer)
        const int r = tt % tiles_per_frame;
#ifdef SESRQ_DEBUG_KNOBS
        const int rows = (P.dbg & 256) ? 0 : min(P.TH, P.H - (r / P.tiles_x) * P.TH) + 2 * PAD;
#else
        const int rows = min(P.TH, P.H - (r / P.tiles_x) * P.TH) + 2 * PAD;
#endif
        const int delta = ((r % P.tiles_x) * P.TW - PAD) & 3;          // pixels the aligned box starts early
        if (lane == 0) {
          mbar_wait(&x_full[xs], xph);
          mbar_wait(&e_empty[st], ph ^ 1);                             // the MMAs that read this E stage have retired
        }
        __syncwarp();
        const float* xsrc = reinterpret_cast<const float*>(region2 + (size_t)xs * P.xstage_bytes) + delta + lane;
        uint4* e = reinterpret_cast<uint4*>(e0 + (size_t)st * P.stage_bytes) + lane;
        // row pointers advance by constants (the compiler re-derived them from `row` with ~10 IMADs per row)
        const float* xr = xsrc + bwid * L0_XP;
        uint4* er = e + bwid * L0_EP;
        for (int row = bwid; row < rows; row += L0_BUILD_WARPS, xr += L0_BUILD_WARPS * L0_XP, er += L0_BUILD_WARPS * L0_EP) {
          float v[4][CIN];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CIN; ++c) v[u][c] = xr[c * rp * L0_XP + 32 * u];
          {
            // ExactDiv's operand range [-lim, lim] (lim / s = 2^40: beyond it the code saturates): test all of the
            // lane's values at once; !(m <= lim) also catches NaN.  The clamp itself runs only if some lane needs it.
            float m = 0.f;
#pragma unroll
            for (int u = 0; u < 4; u += 2)
#pragma unroll
              for (int c = 0; c < CIN; ++c) m = max3_nan(m, fabsf(v[u][c]), fabsf(v[u + 1][c]));
            const bool wild = !(m <= div_s0.lim);
            if (__any_sync(0xffffffffu, wild)) {
#pragma unroll
              for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int c = 0; c < CIN; ++c) v[u][c] = fminf(fmaxf(v[u][c], -div_s0.lim), div_s0.lim);
            }
          }
          uint32_t w0[4];
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            int q[2][4] = {{-128, -128, -128, -128}, {-128, -128, -128, -128}};   // unused channels: byte 0 (zero weights)
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
              const float2 tq = f2_add(exact_div2(div_s0, make_float2(v[u][c], v[u + 1][c])), z0);   // :225, clamped by the pack
              q[0][c] = __float2int_rn(tq.x);
              q[1][c] = __float2int_rn(tq.y);
            }
            w0[u] = pack4_sat_s8(q[0][0], q[0][1], q[0][2], q[0][3]) ^ 0x80808080u;
            w0[u + 1] = pack4_sat_s8(q[1][0], q[1][1], q[1][2], q[1][3]) ^ 0x80808080u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) sw[32 * u + lane] = w0[u];
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t* s1 = sw + 32 * u + lane + 1;
            er[32 * u] = make_uint4(w0[u], s1[0], s1[1], s1[2]);
          }
          __syncwarp();
        }
        fence_proxy_async();                                           // generic-proxy stores -> visible to UMMA
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&e_full[st]);
          mbar_arrive(&x_empty[xs]);
        }
        if (++st == P.stages) { st = 0; ph ^= 1; }
        if (++xs == P.xstages) { xs = 0; xph ^= 1; }
      }
    } else
    for (int t = blockIdx.x; t < P.num_tiles; t += gridDim.x) {
      const int tt = P.dbg < 0 ? P.num_tiles - 1 - t : t;      // tile order (launch_conv_umma: alternates from layer to layer)
      const int b = tt / tiles_per_frame;
      const int r = tt - b * tiles_per_frame;
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
EncodeTiledFn get_encode_fn() {
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

// Tuning knobs from the environment, read once per process (first use).  They select among correct variants only; the
// knock-out masks that produce wrong results exist only in -DSESRQ_DEBUG_KNOBS builds.
struct Tuning {
  int nbw = 0, nbw_l0 = 0, nbw_final = 0, l0_no_tma = 0, l0_xstages = 0, sc_no_tma = 0, dbg = 0;
  Tuning() {
    auto geti = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
    nbw = geti("SESRQ_UMMA_NBW");
    nbw_l0 = geti("SESRQ_UMMA_NBW_L0");
    nbw_final = geti("SESRQ_UMMA_NBW_FINAL");
    l0_no_tma = getenv("SESRQ_L0_NO_TMA") ? 1 : 0;
    sc_no_tma = getenv("SESRQ_SC_NO_TMA") ? 1 : 0;
    l0_xstages = geti("SESRQ_L0_XSTAGES");
#ifdef SESRQ_DEBUG_KNOBS
    dbg = geti("SESRQ_UMMA_DBG");
#endif
  }
};
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SESRQ_PDL"); return !(e && atoi(e) == 0); }();
  return on;
}

static const Tuning& tuning() {
  static const Tuning t;
  return t;
}

int sm_count() {
  static std::atomic<int> cached[MAX_DEVICES];
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < MAX_DEVICES) {
    const int c = cached[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < MAX_DEVICES) cached[dev].store(sms, std::memory_order_relaxed);
  return sms;
}
// layer-0 layout: fp32 NCHW frames with cin <= 4 -- only the FIRST layer reads fp32 (a hidden layer with <= 4 channels
// reads NHWC16 bytes and runs on the generic path)
static bool is_l0_layout(const LayerConst& c) { return c.first_layer != 0 && c.cin <= 4; }

// explicit max(y, 0) is only needed for a ReLU layer whose consumer's zero point is above -128 (see epilogue_pixel)
bool needs_floor(const LayerConst& c) { return c.epi == EPI_PLAIN && c.relu && c.z_next > -128.f; }

bool umma_supported(const LayerConst& c) {
  // layer 0: fp32 frames with cin <= 4 (any input zero point: the builder writes the pad value itself);
  // other layers: 16 input channels from an NHWC16 tensor whose out-of-frame value is byte 0 (z' == -128).
  // Anything else runs on the generic path.
  // The float-bits accumulator (MAGIC_BITS + A) needs |A| < 2^22: A is bounded by the adder clamp plus the
  // 16-bit constant, so pe_add_bits <= 22 suffices.  The final epilogue needs 12582912 + z5 exact in fp32.
  if (c.hi_add >= (1 << 21) || c.hi_acc >= (1 << 21)) return false;
  if (c.epi == EPI_FINAL && (c.relu || !(fabsf(c.z_next) <= 1048576.f))) return false;
  if (is_l0_layout(c)) return c.epi == EPI_PLAIN && c.save_rc == 0 && (c.k == 3 || c.k == 5);
  if (!(c.cin == 16 && c.pad_byte == 0 && c.save_rc == 0 && (c.k == 3 || c.k == 5))) return false;
  if (c.epi == EPI_FINAL && !(c.r_shuffle == 1 || c.r_shuffle == 2 || c.r_shuffle == 4)) return false;
  return true;
}

// Final layers with <= 12 output channels leave accumulator columns 12..15 unused: up to 4 risky (channel, PE)
// partial sums ride there and the layer keeps the single 16-column set (x2 conv_last: 12 channels, 2 risky pairs).
static bool spare_column_layout(const LayerConst& c) {
  if (!(c.pe_exact && c.add_inert && c.epi == EPI_FINAL && c.cout <= 12)) return false;
  int pairs = 0;
  for (int co = 0; co < c.cout; ++co) {
    if (c.pe_risky[co] > 3) return false;
    if (c.pe_risky[co] >= 0) ++pairs;
  }
  return pairs <= 4;
}

// Accumulator layout for this layer (see header): 1, 2 or 4 sets of 16 columns.
static int accumulator_sets(const LayerConst& c) {
  if (!c.pe_exact) return 1;
  if (spare_column_layout(c)) return 1;
  if (!c.add_inert) return 4;                  // the adder-tree clamp can fire: literal tree
  int pairs = 0;
  for (int co = 0; co < c.cout; ++co) {
    if (c.pe_risky[co] > 3) return 4;          // more than one PE of a channel can clamp
    if (c.pe_risky[co] >= 0) ++pairs;
  }
  return pairs <= 4 ? 2 : 4;                   // NS = 2 keeps 4 correction columns live
}

// NS = 2: the (channel, PE) pairs whose 18-bit clamp can fire, in set-0 column order.  Set-1 column j carries the
// partial sum of pair j; cols[j] is the set-0 column it corrects, chans[j] / pes[j] the channel and PE.
static int risky_pairs(const LayerConst& c, bool natural, int* cols, int* chans, int* pes) {
  int cnt = 0;
  for (int n = 0; n < 16; ++n) {
    const int co = natural ? n : chan_of_byte(n);
    if (co >= c.cout || c.pe_risky[co] < 0) continue;
    cols[cnt] = n; chans[cnt] = co; pes[cnt] = c.pe_risky[co];
    ++cnt;
  }
  return cnt;
}

// (output channel, PE mask) of accumulator column n (set = n / 16) for layout ns; co < 0: zero row
static void column_role(const LayerConst& c, int ns, bool natural, int n, int* co_out, int* only_pe) {
  const int set = n / 16, nn = n % 16;
  *only_pe = -1;
  *co_out = -1;
  if (ns == 2 && set == 1) {
    int cols[16], chans[16], pes[16];
    const int cnt = risky_pairs(c, natural, cols, chans, pes);
    if (nn < cnt) { *co_out = chans[nn]; *only_pe = pes[nn]; }
    return;
  }
  const int co = natural ? nn : chan_of_byte(nn);
  if (ns == 1 && nn >= 12 && spare_column_layout(c)) {       // spare column 12 + j carries risky pair j
    int cols[16], chans[16], pes[16];
    const int cnt = risky_pairs(c, natural, cols, chans, pes);
    if (nn - 12 < cnt) { *co_out = chans[nn - 12]; *only_pe = pes[nn - 12]; }
    return;
  }
  if (co >= c.cout) return;
  *co_out = co;
  if (ns == 4) *only_pe = set;
}

// B operand image (see mma_issue_tile): [tap pair q][K half h][row block jblk <-> dy = k-1-jblk][N columns][16 B].
// Row (jblk, n) of pair q, half h holds the weights of tap (dy, dx = 2q + h) for the output channel / PE mask that
// accumulator column n stands for, over the 16 input-channel bytes (zero rows for dx >= k).
// Layer-0 image: one pair; byte k of half h <-> tap dx = 4h + k/4, input channel k%4 (PE p == channel p for cin <= 4).
static int umma_upload(LayerDev& L, const std::vector<uint8_t>& img, const char* what) {
  if (cudaMalloc(&L.d_wumma, img.size()) != cudaSuccess ||
      cudaMemcpy(L.d_wumma, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_cuda_error(cudaGetLastError(), what);
    return SESRQ_ECUDA;
  }
  return SESRQ_OK;
}

static int umma_prepare_layer0(LayerDev& L, const int8_t* w) {
  const LayerConst& c = L.c;
  const int ns = accumulator_sets(c);
  const int N = 16 * ns, NTOT = c.k * N;
  L.umma_n = N;
  L.layout = ns == 4 ? SESRQ_LAYOUT_NS4 : (ns == 2 ? SESRQ_LAYOUT_NS2 : (spare_column_layout(c) ? SESRQ_LAYOUT_SPARE : SESRQ_LAYOUT_NS1));
  std::vector<uint8_t> img((size_t)2 * NTOT * 16, 0);
  for (int h = 0; h < 2; ++h)
    for (int jblk = 0; jblk < c.k; ++jblk) {
      const int dy = c.k - 1 - jblk;
      for (int n = 0; n < N; ++n) {
        int co, only_pe;
        column_role(c, ns, false, n, &co, &only_pe);
        if (co < 0) continue;
        uint8_t* row = img.data() + (((size_t)h * NTOT) + (size_t)jblk * N + n) * 16;
        for (int k = 0; k < 16; ++k) {
          const int dx = 4 * h + k / 4, ci = k % 4;
          if (dx >= c.k || ci >= c.cin) continue;
          if (only_pe >= 0 && ci != only_pe) continue;
          row[k] = (uint8_t)w[(((size_t)co * c.cin + ci) * c.k + dy) * c.k + dx];
        }
      }
    }
  return umma_upload(L, img, "umma layer-0 weight upload");
}

int umma_prepare_layer(LayerDev& L, const int8_t* w) {
  const LayerConst& c = L.c;
  if (is_l0_layout(c)) return umma_prepare_layer0(L, w);
  const int taps = c.k * c.k, pairs = (c.k + 1) / 2;
  const int ns = accumulator_sets(c);
  const int N = 16 * ns, NTOT = c.k * N;
  L.umma_n = N;
  L.layout = ns == 4 ? SESRQ_LAYOUT_NS4 : (ns == 2 ? SESRQ_LAYOUT_NS2 : (spare_column_layout(c) ? SESRQ_LAYOUT_SPARE : SESRQ_LAYOUT_NS1));
  std::vector<uint8_t> img((size_t)pairs * 2 * NTOT * 16, 0);
  const bool natural = (c.epi == EPI_FINAL);
  for (int q = 0; q < pairs; ++q)
    for (int h = 0; h < 2; ++h) {
      const int dx = 2 * q + h;
      if (dx >= c.k) continue;
      for (int jblk = 0; jblk < c.k; ++jblk) {
        const int dy = c.k - 1 - jblk;
        for (int n = 0; n < N; ++n) {
          int co, only_pe;
          column_role(c, ns, natural, n, &co, &only_pe);
          if (co < 0) continue;
          uint8_t* row = img.data() + ((((size_t)q * 2 + h) * NTOT) + (size_t)jblk * N + n) * 16;
          for (int j = 0; j < 16; ++j) {
            const int ci = chan_of_byte(j);
            if (ci >= c.cin) continue;
            if (only_pe >= 0 && (j >> 2) != only_pe) continue;
            row[j] = (uint8_t)w[((size_t)co * c.cin + ci) * taps + dy * c.k + dx];
          }
        }
      }
    }
  return umma_upload(L, img, "umma weight upload");
}

template <int KS, int NS, int EPI, int RS, int COUT, int OUTM, bool FLOOR, int EPI_WARPS>
static int launch_one(const CUtensorMap& tmap, const UmmaParams& P, const LayerDev& L, const uint8_t* d_shortcut,
                      uint8_t* d_out, int8_t* d_q5, float* d_y, int grid, size_t smem, cudaStream_t st,
                      const FqArgs& fa = FqArgs{}, const CUtensorMap* smap = nullptr) {
  auto kern = conv_umma_kernel<KS, NS, EPI, RS, COUT, OUTM, FLOOR, EPI_WARPS>;
  static std::atomic<uint64_t> optin_done{0};
  if (int rc = ensure_smem_optin(kern, 227 * 1024, optin_done)) return rc;
  SESRQ_CUDA(launch_pdl(kern, grid, (CTRL_WARPS + EPI_WARPS) * 32, smem, st, tmap, P, (const uint8_t*)L.d_wumma,
                        reinterpret_cast<const uint4*>(d_shortcut), reinterpret_cast<uint4*>(d_out), d_q5, d_y, fa, smap ? *smap : tmap));
  SESRQ_LAUNCH_CHECK("conv_umma_kernel");
  return SESRQ_OK;
}

template <int KS, int NS, int EPI_WARPS>
static int launch_epi(const LayerConst& c, const CUtensorMap& tmap, const UmmaParams& P, const LayerDev& L,
                      const uint8_t* d_shortcut, uint8_t* d_out, int8_t* d_q5, float* d_y, int grid, size_t smem,
                      cudaStream_t st, const CUtensorMap* smap) {
#define SESRQ_GO(EPI_, RS_, COUT_, OUTM_, FLOOR_) \
  return launch_one<KS, NS, EPI_, RS_, COUT_, OUTM_, FLOOR_, EPI_WARPS>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, \
                                                                        FqArgs{}, smap)
#define SESRQ_GO_FINAL(RS_, COUT_)                              \
  do {                                                          \
    if (d_q5 && d_y) SESRQ_GO(EPI_FINAL, RS_, COUT_, 0, false); \
    if (d_y) SESRQ_GO(EPI_FINAL, RS_, COUT_, 2, false);         \
    SESRQ_GO(EPI_FINAL, RS_, COUT_, 1, false);                  \
  } while (0)
  if (c.epi == EPI_PLAIN) {
    if (needs_floor(c)) SESRQ_GO(EPI_PLAIN, 1, 0, 0, true);
    SESRQ_GO(EPI_PLAIN, 1, 0, 0, false);
  }
  if (c.epi == EPI_RESIDUAL) SESRQ_GO(EPI_RESIDUAL, 1, 0, 0, false);
  if (!d_q5 && !d_y) return SESRQ_EINVAL;
  // the reference's three networks get kernels specialised on their channel count; anything else the run-time loop
  if (c.r_shuffle == 4) { if (c.cout == 16) SESRQ_GO_FINAL(4, 16); SESRQ_GO_FINAL(4, 0); }
  if (c.r_shuffle == 2) { if (c.cout == 12) SESRQ_GO_FINAL(2, 12); SESRQ_GO_FINAL(2, 0); }
  if (c.cout == 3) SESRQ_GO_FINAL(1, 3);
  SESRQ_GO_FINAL(1, 0);
#undef SESRQ_GO_FINAL
#undef SESRQ_GO
}

template <int KS>
static int launch_ns(int ns, const LayerConst& c, const CUtensorMap& tmap, const UmmaParams& P, const LayerDev& L,
                     const uint8_t* d_shortcut, uint8_t* d_out, int8_t* d_q5, float* d_y, int grid, size_t smem,
                     cudaStream_t st, const CUtensorMap* smap) {
  // x2 final layer (PixelShuffle(2)) with the fp32 output materialised: 8 epilogue warps instead of 16.  Fewer warps = fewer
  // concurrent write streams per SM; measured on 16 x 1080p (ms, 4 / 8 / 12 / 16 / 20 warps): y only 0.416 / 0.375 / 0.409 /
  // 0.390 / 0.375, both outputs 0.731 / 0.525 / 0.578 / 0.607 / 0.533, q5 only 0.707 / 0.449 / - / 0.345 / 0.384 (so q5-only
  // keeps 16).  The x4 and NRDM final layers show no gain (1.176 vs 1.176 ms, 1.125 vs 1.139 ms per step) and keep 16.
  if (ns == 1 && c.epi == EPI_FINAL && c.r_shuffle == 2 && d_y) {
    const FqArgs nofq{};
    if (c.cout == 12) {
      if (d_q5) return launch_one<KS, 1, EPI_FINAL, 2, 12, 0, false, 8>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, nofq, smap);
      return launch_one<KS, 1, EPI_FINAL, 2, 12, 2, false, 8>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, nofq, smap);
    }
    if (d_q5) return launch_one<KS, 1, EPI_FINAL, 2, 0, 0, false, 8>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, nofq, smap);
    return launch_one<KS, 1, EPI_FINAL, 2, 0, 2, false, 8>(tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, nofq, smap);
  }
  if (ns == 1) return launch_epi<KS, 1, 16>(c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, smap);
  if (ns == 2) return launch_epi<KS, 2, 16>(c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, smap);
  return launch_epi<KS, 4, 8>(c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, smap);
}

// Tile height = TB * nbw rows (nbw issuer warps / row bands).  More bands = more MMA issue bandwidth but (k - 1)
// extra edge rows per band; the accumulator ring (512 / (16 ns) row slots) must hold at least one tile.  Pick the
// largest nbw <= the preferred one whose tile buffers fit at least twice next to the fixed shared memory.
struct TilePlan { int TH, tiles_y, stages, nbw; size_t stage_bytes; };
static TilePlan choose_tiling(int H, int ns, int prefer, int pad, size_t row_bytes, size_t slack_bytes, size_t per_row_extra,
                              size_t fixed) {
  const int nslot = TMEM_COLS / (16 * ns);
  // `prefer` = the caller's measured band count for NS = 1 (see launch_conv_umma / launch_conv_umma_l0)
  int want = (ns == 1) ? prefer : (ns == 2 ? 2 : 1);
  if (tuning().nbw > 0) want = tuning().nbw;                                     // tuning experiments
  want = std::max(1, std::min(want, std::min(3, nslot / TB)));
  const size_t total = 227 * 1024;
  TilePlan best{0, 0, 0, 0, 0};
  for (int nbw = want; nbw >= 1; --nbw) {
    const int TH = TB * nbw;
    const size_t stage = ((size_t)(TH + 2 * pad) * row_bytes + slack_bytes + 1023) & ~(size_t)1023;
    const size_t extra = (size_t)(TH + 2 * pad) * per_row_extra + 256;
    if (fixed + extra + 2 * stage > total) continue;
    best = {TH, cdiv(H, TH), (int)std::min<size_t>(MAX_STAGES, (total - fixed - extra) / stage), nbw, stage};
    break;
  }
  return best;
}

void fill_epilogue_params(UmmaParams& P, const LayerConst& c, int ns) {
  const bool natural = (c.epi == EPI_FINAL);
  for (int n = 0; n < 16; ++n) {
    const int co = natural ? n : chan_of_byte(n);
    if (co >= c.cout) {                       // unused column: keep the float-bits form valid
      if (ns <= 2) P.c_init[0][n] = MAGIC_BITS;
      P.c_kmagic[n] = MAGIC_BITS;
      continue;
    }
    if (ns <= 2) {
      P.c_init[0][n] = c.fused_const[co] + MAGIC_BITS;        // sum_p pe_corr + K, as float bits
    } else {
      for (int p = 0; p < 4; ++p) P.c_init[p][n] = c.pe_corr[p][co];
      P.c_kmagic[n] = c.add_const[co] + MAGIC_BITS;
    }
  }
  if (ns == 2 || (ns == 1 && spare_column_layout(c))) {
    int cols[16], chans[16], pes[16];
    P.n_risky = risky_pairs(c, natural, cols, chans, pes);
    for (int j = 0; j < P.n_risky; ++j) {
      P.risky_col[j] = cols[j];
      if (ns == 2) P.c_init[1][j] = c.pe_corr[pes[j]][chans[j]];
      else P.c_init[0][12 + j] = c.pe_corr[pes[j]][chans[j]];          // spare column: plain integer partial sum
    }
  }
  P.lo_acc = c.lo_acc; P.hi_acc = c.hi_acc; P.lo_add = c.lo_add; P.hi_add = c.hi_add;
  P.mf = c.mf;
  P.nbias = (float)(-12582912.0 * (double)c.mf);              // exact: 3 * M16 * 2^(22 - n)
  P.z_next = c.z_next;
  P.res_mf = c.res_mf;
  P.res_nbias = (float)(-12582912.0 * (double)c.res_mf);
  P.z_res = c.z_res;
  P.s_out = c.s_out;
  P.out_off = -(12582912.f + c.z_next);
  P.relu = c.relu; P.cout = c.cout;
}

template <int KS, int NS, int CIN, bool FLOOR, int EPI_WARPS>
static int launch_l0_one(const CUtensorMap& xmap, const UmmaParams& P, const LayerDev& L, const float* d_x, uint8_t* d_out,
                         int grid, size_t smem, cudaStream_t st) {
  auto kern = conv_umma_l0_kernel<KS, NS, CIN, FLOOR, EPI_WARPS>;
  // the kernel also has ~4.2 KB of static shared memory; static + dynamic <= 227 KB
  static std::atomic<uint64_t> optin_done{0};
  if (int rc = ensure_smem_optin(kern, 227 * 1024 - L0_STATIC_SMEM, optin_done)) return rc;
  // `smem` is the CTA's whole budget (the plan counts the static part in `fixed`): the dynamic request excludes it
  SESRQ_CUDA(launch_pdl(kern, grid, (CTRL_WARPS + EPI_WARPS + L0_BUILD_WARPS) * 32, smem - L0_STATIC_SMEM, st, xmap, P, (const uint8_t*)L.d_wumma, d_x,
                        reinterpret_cast<uint4*>(d_out)));
  SESRQ_LAUNCH_CHECK("conv_umma_l0_kernel");
  return SESRQ_OK;
}

static int launch_conv_umma_l0(const LayerDev& L, const float* d_x, uint8_t* d_out, int B, int H, int W, cudaStream_t st, int rev) {
  const LayerConst& c = L.c;
  if (!d_x || !d_out) return SESRQ_EINVAL;
  const int pad = c.pad, ns = L.umma_n / 16, N = L.umma_n;
  UmmaParams P;
  memset(&P, 0, sizeof(P));
  P.tiles_x = cdiv(W, UMMA_M - 2 * pad);
  P.TW = cdiv(W, P.tiles_x);
  const uint32_t w_bytes = (uint32_t)2 * c.k * N * 16;
  // fixed part: weights, barriers, slack for alignment, and the kernel's 4 KB of static shared memory (stage_w)
  const size_t fixed = ((w_bytes + 127) & ~127u) + 8 * (2 * MAX_STAGES + 2 * (TMEM_COLS / N) + 2 * L0_XSTAGES) + 64 + 1024 + L0_STATIC_SMEM;
  const int sms = sm_count();
  // fp32 tiles by TMA need 16-byte aligned rows (W % 4 == 0), a 16-byte aligned base, and x = 0 -> pad code (z0 <= 127)
  EncodeTiledFn encode = get_encode_fn();
  P.x_tma = (encode && (W % 4) == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && c.in_zero <= 127.f) ? 1 : 0;
  if (tuning().l0_no_tma) P.x_tma = 0;
  const size_t per_row_extra = P.x_tma ? (size_t)2 * c.cin * L0_XP * 4 : (size_t)L0_SP * 4;
  // Bands of 8 rows per tile, 2 or 3: the cost is (rows incl. the padding of the last tile) x (cost per row), and per
  // row 3 bands are ~5 % cheaper for RGB input (fewer halo rows to quantise: 1080p x2 0.278 vs 0.292 ms) but ~4 % dearer
  // for single-channel input (x4 540x960: 0.271 vs 0.256 ms); the padding decides for short frames (H = 80: 96 vs 80
  // rows, 0.197 vs 0.178 ms).
  const double rows2 = (double)cdiv(H, 2 * TB) * 2 * TB, rows3 = (double)cdiv(H, 3 * TB) * 3 * TB;
  int prefer = (c.cin == 1 ? rows3 * 1.04 < rows2 : rows3 * 0.95 <= rows2) ? 3 : 2;
  if (tuning().nbw_l0 > 0) prefer = tuning().nbw_l0;                            // tuning experiments
  const TilePlan tp = choose_tiling(H, ns, prefer, pad, (size_t)L0_EP * 16, 8 * 16, per_row_extra, fixed + 256);
  if (tp.stages < 2) return SESRQ_EUNSUPPORTED;
  P.TH = tp.TH;
  P.tiles_y = tp.tiles_y;
  P.nbw = tp.nbw;
  P.stage_bytes = (uint32_t)tp.stage_bytes;
  P.xstage_bytes = (uint32_t)(((size_t)c.cin * (P.TH + 2 * pad) * L0_XP * 4 + 127) & ~(size_t)127);
  const size_t s_region = (((size_t)(P.TH + 2 * pad) * L0_SP * 4 + 127) & ~(size_t)127);
  if (P.x_tma) {
    // two operand-plane buffers suffice (built while the other is consumed); the fp32 tiles need depth to hide
    // the HBM latency + transfer time of a tile behind the build of the previous ones
    P.stages = 2;
    const size_t room = 227 * 1024 - fixed - (size_t)P.stages * P.stage_bytes;
    P.xstages = (int)std::min<size_t>(L0_XSTAGES, room / P.xstage_bytes);
    if (tuning().l0_xstages > 0) P.xstages = std::min(P.xstages, std::max(2, tuning().l0_xstages));
    if (P.xstages < 2) return SESRQ_EUNSUPPORTED;
  } else {
    P.stages = std::min(tp.stages, 3);
  }
  const size_t region2 = P.x_tma ? (size_t)P.xstages * P.xstage_bytes : s_region;
  const size_t smem = (size_t)P.stages * P.stage_bytes + region2 + fixed;
  if (smem > 227 * 1024) return SESRQ_EUNSUPPORTED;
  CUtensorMap xmap;
  memset(&xmap, 0, sizeof(xmap));
  if (P.x_tma) {
    cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)c.cin, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)c.cin * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)L0_XP, (cuuint32_t)(P.TH + 2 * pad), (cuuint32_t)c.cin, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d_x), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (layer-0 input)");
      return SESRQ_ECUDA;
    }
  }
  P.w_bytes = w_bytes;
  P.num_tiles = P.tiles_x * P.tiles_y * B;
  P.B = B; P.H = H; P.W = W; P.pad = pad;
  P.idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(UMMA_M >> 4) << 24);      // N is OR-ed in per instruction
  fill_epilogue_params(P, c, ns);
  P.cin = c.cin;
  P.pad_word = 0;
  for (int ci = 0; ci < c.cin; ++ci) P.pad_word |= (uint32_t)c.pad_byte << (8 * ci);
  P.in_scale = c.in_scale;
  P.in_zero = c.in_zero;
  P.dbg = tuning().dbg | (rev ? UMMA_REV : 0);
  const int grid = std::min(P.num_tiles, sms);
#define SESRQ_L0_NS(KS_, CIN_)                                                                              \
  do {                                                                                                      \
    if (needs_floor(c)) {                                                                                   \
      if (ns == 1) return launch_l0_one<KS_, 1, CIN_, true, SESRQ_L0_EPI_WARPS>(xmap, P, L, d_x, d_out, grid, smem, st);          \
      if (ns == 2) return launch_l0_one<KS_, 2, CIN_, true, 16>(xmap, P, L, d_x, d_out, grid, smem, st);          \
      return launch_l0_one<KS_, 4, CIN_, true, 8>(xmap, P, L, d_x, d_out, grid, smem, st);                        \
    }                                                                                                       \
    if (ns == 1) return launch_l0_one<KS_, 1, CIN_, false, SESRQ_L0_EPI_WARPS>(xmap, P, L, d_x, d_out, grid, smem, st);           \
    if (ns == 2) return launch_l0_one<KS_, 2, CIN_, false, 16>(xmap, P, L, d_x, d_out, grid, smem, st);           \
    return launch_l0_one<KS_, 4, CIN_, false, 8>(xmap, P, L, d_x, d_out, grid, smem, st);                         \
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

// Tiling, tensor map and kernel parameters of one 16-channel-layout launch (shared by the integer and fake-quant paths).
static int plan_conv_launch(const LayerDev& L, const uint8_t* d_in, int B, int H, int W, UmmaParams& P, CUtensorMap& tmap,
                            int& grid, size_t& smem, const uint8_t* d_shortcut = nullptr, CUtensorMap* smap = nullptr) {
  const LayerConst& c = L.c;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return SESRQ_EUNSUPPORTED;
  if (!d_in) return SESRQ_EINVAL;
  const int pad = c.pad;
  const int ns = L.umma_n / 16;
  const int N = L.umma_n;

  // ---- tiling: 128-pixel pitch (one TMA box row = 2048 B), TH output rows per tile ----
  memset(&P, 0, sizeof(P));
  P.tiles_x = cdiv(W, UMMA_M - 2 * pad);
  P.TW = cdiv(W, P.tiles_x);
  const int pairs = (c.k + 1) / 2;
  const uint32_t w_bytes = (uint32_t)pairs * 2 * c.k * N * 16;
  size_t fixed = ((w_bytes + 127) & ~127u) + 8 * (4 * MAX_STAGES + 2 * (TMEM_COLS / N)) + 64 + 1024;
  const int sms = sm_count();
  // residual layer: the shortcut rows of a tile are staged by TMA as well (16 rows x 2 KB per buffer, 3 buffers: a
  // buffer lives until the epilogue is through with the tile, i.e. longer than the input tile the MMAs consume)
  const bool sc_tma = c.epi == EPI_RESIDUAL && d_shortcut && smap && ns == 1 && !tuning().sc_no_tma &&
                      (reinterpret_cast<uintptr_t>(d_shortcut) & 15) == 0;
  const int sc_stages = sc_tma ? 3 : 0;
  const size_t sc_reserved = (size_t)sc_stages * (2 * TB) * UMMA_M * 16;
  fixed += sc_reserved;
  // slack: the last row's rightmost K chunk reaches 2*pad + 1 pixels past the row
  // measured: 16-channel layers are fastest with 2 bands (TH = 16: the 32-slot ring then holds two tiles; 1080p x2 3x3
  // 0.178 vs 0.185 ms with 3 bands); the 16-output x4 final layer with 3 (540x960: 0.909 vs 0.955 ms)
  // (round 2: the 3-output NRDM final layer, 512 frames of 80x960: 1.254 / 1.170 / 1.133 ms per step with 1 / 2 / 3 bands)
  int prefer = (c.epi == EPI_FINAL && (c.r_shuffle == 4 || c.cout <= 4)) ? 3 : 2;
  if (c.epi == EPI_FINAL && tuning().nbw_final > 0) prefer = tuning().nbw_final;   // tuning experiments
  const TilePlan tp = choose_tiling(H, ns, prefer, pad, (size_t)UMMA_M * 16, (size_t)(2 * pad + 1) * 16, 0, fixed);
  if (tp.stages < 2) return SESRQ_EUNSUPPORTED;
  P.TH = tp.TH;
  P.tiles_y = tp.tiles_y;
  P.stages = tp.stages;
  P.nbw = tp.nbw;
  P.stage_bytes = (uint32_t)tp.stage_bytes;
  P.box_bytes = (uint32_t)((P.TH + 2 * pad) * UMMA_M * 16);
  P.w_bytes = w_bytes;
  P.num_tiles = P.tiles_x * P.tiles_y * B;
  P.B = B; P.H = H; P.W = W; P.pad = pad;
  P.sc_stages = sc_stages;
  P.sc_stage_bytes = (uint32_t)(P.TH * UMMA_M * 16);
  // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32, A = u8, B = s8, K-major both, M = 128;
  // N is OR-ed in per instruction (mma_issue_tile)
  P.idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(UMMA_M >> 4) << 24);
  fill_epilogue_params(P, c, ns);
  P.dbg = tuning().dbg;

  if (sc_tma) {                         // the shortcut tensor has the input's layout; box = the tile's output rows, no halo
    cuuint64_t sdim[3] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t sstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    cuuint32_t sbox[3] = {(cuuint32_t)UMMA_M * 2, (cuuint32_t)P.TH, 1};
    cuuint32_t sest[3] = {1, 1, 1};
    CUresult sr = encode(smap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(d_shortcut), sdim, sstr, sbox, sest,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (sr != CUDA_SUCCESS) {
      set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (shortcut)");
      return SESRQ_ECUDA;
    }
  }

  // ---- tensor map over the NHWC16 input viewed as (W*2 x 8-byte, H, B) ----
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
  grid = std::min(P.num_tiles, sms);
  // `fixed` reserved 3 shortcut buffers of 2 * TB rows; the tile may have come out shorter
  smem = (size_t)P.stage_bytes * P.stages + fixed - sc_reserved + (size_t)P.sc_stages * P.sc_stage_bytes;
  return SESRQ_OK;
}

int launch_conv_umma(const LayerDev& L, const float* d_x, const uint8_t* d_in, const uint8_t* d_shortcut,
                     uint8_t* d_out, int8_t* d_q5, float* d_y, int B, int H, int W, cudaStream_t st, int rev) {
  const LayerConst& c = L.c;
  if (is_l0_layout(c)) return launch_conv_umma_l0(L, d_x, d_out, B, H, W, st, rev);
  UmmaParams P;
  CUtensorMap tmap, smap;
  int grid = 0;
  size_t smem = 0;
  if (int rc = plan_conv_launch(L, d_in, B, H, W, P, tmap, grid, smem, d_shortcut, &smap)) return rc;
  // Tile order.  A layer finishes with its last tiles' output still in the 126 MB L2; its consumer starts with those tiles
  // when it walks the tiles in the opposite direction (sesrq_net_forward_i8 alternates `rev` from launch to launch).
  if (rev) P.dbg |= UMMA_REV;
  const int ns = L.umma_n / 16;
  const CUtensorMap* sp = P.sc_stages > 0 ? &smap : nullptr;
  if (c.k == 3) return launch_ns<3>(ns, c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, sp);
  return launch_ns<5>(ns, c, tmap, P, L, d_shortcut, d_out, d_q5, d_y, grid, smem, st, sp);
}

// exe_mode 0: four PE-masked accumulator sets (the float-domain PE sums are rounded separately, quan_func.py:333,:386),
// slots pre-armed with MAGIC_BITS (no integer constants: bias and clamps live in the float domain), fp32 output.
int launch_conv_umma_fq(const LayerDev& L, const uint8_t* d_in, const FqArgs& fa, int B, int H, int W, cudaStream_t st) {
  const LayerConst& c = L.c;
  if (c.epi != EPI_FQ || L.umma_n != 64 || !fa.fp || !fa.xout || !fa.mm || !fa.tapsum) return SESRQ_EINVAL;
  UmmaParams P;
  CUtensorMap tmap;
  int grid = 0;
  size_t smem = 0;
  if (int rc = plan_conv_launch(L, d_in, B, H, W, P, tmap, grid, smem)) return rc;
  for (int p = 0; p < 4; ++p)
    for (int n = 0; n < 16; ++n) P.c_init[p][n] = MAGIC_BITS;
  if (c.k == 3) return launch_one<3, 4, EPI_FQ, 1, 0, 0, false, SESRQ_FQ_EPI_WARPS>(tmap, P, L, nullptr, nullptr, nullptr, nullptr, grid, smem, st, fa);
  return launch_one<5, 4, EPI_FQ, 1, 0, 0, false, SESRQ_FQ_EPI_WARPS>(tmap, P, L, nullptr, nullptr, nullptr, nullptr, grid, smem, st, fa);
}

}  // namespace sesrq
