// Device-side building blocks of the tcgen05 kernels (sm_100a only): PTX wrappers, the kernel parameter block, the exact
// requantisation epilogue of one pixel and the TMEM <-> register helpers.  Included by sesrq_umma.cu (one layer per
// launch) and sesrq_chain.cu (two 3x3 layers per launch, intermediate rows through shared memory).
#pragma once
#include <cuda.h>
#include <atomic>
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
// Waiting warps share their scheduler with working ones and a spinning warp is always "ready", so a plain try_wait
// loop steals issue slots (ncu: 150 M of 510 M instructions of the layer-0 kernel were such spins and the operand
// builders ran at 0.1 IPC).  The suspend-time hint parks the warp in hardware until the phase completes (or the hint
// expires), with no blind sleep on the critical path.
constexpr uint32_t MBAR_SUSPEND_NS = 200000;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(a), "r"(parity), "r"(MBAR_SUSPEND_NS)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// Programmatic dependent launch: the next kernel of the stream may start its prologue (weights to shared memory, barrier
// init, TMEM allocation) while this one is still running; it must not touch memory the predecessor writes before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const int (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const int (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
// 16 columns, every word the same value (one source register)
__device__ __forceinline__ void tmem_st16_const(uint32_t taddr, int v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2): two IEEE round-to-nearest operations per issue slot, each lane
// bit-identical to the scalar instruction.  The scalar operands are broadcast by the hardware.
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(uint64_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float2 f2_fma(float2 a, float b, float c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b, b)), "l"(f2_pack(c, c)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 f2_fma3(float2 a, float2 b, float2 c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)), "l"(f2_pack(c.x, c.y)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 f2_add2(float2 a, float2 b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 f2_add(float2 a, float b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b, b)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b, b)));
  return f2_unpack(r);
}

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
constexpr int TMEM_COLS = 512;
#ifndef SESRQ_FQ_EPI_WARPS
#define SESRQ_FQ_EPI_WARPS 12           // epilogue warps of the exe_mode 0 kernels (calibration leg: 8 warps 4.90 ms, 12: 4.78 ms, 16: 5.1 ms -- spills)
#endif
constexpr int CTRL_WARPS = 4;            // warp 0: TMA producer; warps 1-3: MMA issuers (keeps epilogue warp % 4 == lane quarter)
constexpr int MAX_STAGES = 4;

// The accumulator slots in TMEM are not zeroed: the epilogue pre-loads every slot (tcgen05.st) with a per-column
// constant and every MMA accumulates on top of it.  For the single-sum layouts that constant is
//     c_fused + 0x4B400000   (c_fused = sum_p pe_corr_p + K, 0x4B400000 = bits of 1.5 * 2^23),
// so the 32-bit word the epilogue reads back IS the float 12582912 + A (exact for |A| < 2^22), and one FFMA
//     y = fma(m, mf, -12582912 * mf)
// yields RN(A * mf): m * mf is exact inside the FMA and 12582912 * mf = 3 * M16 * 2^(22-n) is exactly
// representable, so there is a single rounding, of the exact product A * M16 * 2^-n -- the reference's
// RN_f32(A * M16) * 2^-n (quan_func.py:529).  No integer add, no int->float conversion, no separate multiply.
// output rows per completion group (see mma_issue_tile / epilogue_warps)
template <int NS> struct GroupRows { static constexpr int value = NS == 1 ? 4 : (NS == 2 ? 2 : 1); };

// Accumulator-slot ring.  A tile of TH = TB * nbw output rows owns TH consecutive row slots starting at a multiple
// of TH (a short bottom tile leaves its trailing slots unused), so slot addresses inside a band are compile-time
// offsets and a tile never wraps around the ring.  Band w (rows [w*TB, w*TB+TB) of the tile) is issued by warp 1+w.
// A band is cut into groups of RG rows, the unit of completion signalling.
constexpr int TB = 8;
struct RingPos {                              // slot base and phase parity of the current tile
  uint32_t base = 0, par = 0;
  __device__ __forceinline__ void advance(uint32_t th, uint32_t nring) {
    base += th;
    if (base == nring) { base = 0; par ^= 1u; }
  }
};
__host__ __device__ __forceinline__ int ring_slots(int nslot, int th) { return (nslot / th) * th; }
template <int RG>
__device__ __forceinline__ int tile_groups(int nblk) {         // groups of a tile with nblk rows
  return (nblk / TB) * (TB / RG) + ((nblk % TB) + RG - 1) / RG;
}

// experiment switches for the final-layer epilogue (tools/variants.sh builds one library per combination)
#ifdef SESRQ_X_NOPIPE
constexpr bool X_NOPIPE = true;      // no software pipelining of the TMEM loads (one accumulator buffer)
#else
constexpr bool X_NOPIPE = false;
#endif
#ifdef SESRQ_X_BRANCH
constexpr bool X_BRANCH = true;      // out-of-frame lanes branch around the row instead of predicating every store
#else
constexpr bool X_BRANCH = false;
#endif
#ifdef SESRQ_X_ROWARRIVE
constexpr bool X_ROWARRIVE = true;   // re-arm completion signalled per row (RG arrivals per warp and group) instead of once per group
#else
constexpr bool X_ROWARRIVE = false;
#endif
// The 4 warps of a warpgroup meet at a named barrier before every row, so that their stores to the same output rows
// (q5: 4 x 32 misaligned bytes of one plane row = whole sectors only together; y: 4 x 256 B of one HR row) leave the SM
// within a few hundred cycles of each other.  Measured on 16 x 1080p: final layer both 0.745 -> 0.626 ms, y only
// 0.435 -> 0.411 (profiles/final_layer_study_r2.txt); 0 = off, 1 = final layer (default), 2 = every layer.
#ifndef SESRQ_X_ROWSYNC
#define SESRQ_X_ROWSYNC 1
#endif
#ifndef SESRQ_X_STCS
#define SESRQ_X_STCS 2               // 1: streaming (evict-first) stores for y, 2: for y and q5 (default: L4 y only 0.412 -> 0.404 ms)
#endif
#ifdef SESRQ_X_NOPIN
constexpr bool X_NOPIN = true;
#else
constexpr bool X_NOPIN = false;
#endif

constexpr int32_t MAGIC_BITS = 0x4B400000;
constexpr float MAGIC_F = 12582912.f;

struct UmmaParams {
  int32_t c_init[4][16];                 // TMEM initial value per accumulator set and column
  int32_t c_kmagic[16];                  // NS=4: K + MAGIC_BITS (added after the clamp tree)
  int32_t n_risky;                       // NS=2: number of (channel, PE) pairs whose 18-bit clamp can fire
  int32_t risky_col[16];                 // NS=2: set-1 column j holds the partial sum that corrects set-0 column risky_col[j]
  int32_t lo_acc, hi_acc, lo_add, hi_add;
  float mf, nbias;                       // M16 * 2^-n and -12582912 * mf
  float z_next;
  float res_mf, res_nbias, z_res;        // EPI_RESIDUAL: Mres * 2^-nres, -12582912 * res_mf, z_4
  float s_out, out_off;                  // EPI_FINAL: f32(s_5), -(12582912 + z_5)
  int32_t relu, cout;
  int32_t dbg;                           // bits 0..30: SESRQ_UMMA_DBG mask (pipeline experiments only; results are wrong when set);
                                         // bit 31 (UMMA_REV): walk the tiles last to first (launch_conv_umma).  It lives here because
                                         // growing this struct by one word took conv_chain2_kernel from 74 to 96 registers + spills
  // geometry
  int32_t B, H, W, TH, TW, tiles_x, tiles_y, num_tiles, pad;
  uint32_t box_bytes, stage_bytes, w_bytes;
  int32_t stages;                        // tile buffers in flight (2..MAX_STAGES)
  int32_t sc_stages;                     // EPI_RESIDUAL: shortcut tile buffers staged by TMA (0 = the epilogue reads global memory)
  uint32_t sc_stage_bytes;               //   one buffer: TH rows x 128 pixels x 16 B
  int32_t nbw;                           // issuer warps = row bands per tile (1..3)
  uint32_t idesc;
  // layer-0 kernel only
  int32_t cin;
  uint32_t pad_word;                     // out-of-frame pixel: byte c = z' + 128 for c < cin
  float in_scale, in_zero;
  int32_t x_tma;                         // 1: fp32 input tiles arrive by TMA (needs W % 4 == 0), 0: builders read global memory
  int32_t xstages;                       // fp32 tile buffers in flight (2..L0_XSTAGES)
  uint32_t xstage_bytes;                 // one fp32 tile buffer: cin planes of (TH + 2 pad) rows x L0_XP floats
};

constexpr int32_t UMMA_REV = (int32_t)0x80000000;

__device__ __forceinline__ void st_y2(float* p, float2 v) {
  if (SESRQ_X_STCS >= 1) __stcs(reinterpret_cast<float2*>(p), v);
  else *reinterpret_cast<float2*>(p) = v;
}
__device__ __forceinline__ void st_q(int8_t* p, int v) {
  if (SESRQ_X_STCS >= 2) __stcs(reinterpret_cast<char*>(p), (char)v);
  else *p = (int8_t)v;
}

// Stores of the final layer for one pixel: COUT int8 codes to the NCHW q5 planes (stride HW bytes) and the
// de-quantised values to the pixel-shuffled fp32 output (PixelShuffle(RS): channel n -> plane n / RS^2, row offset
// (n % RS^2) / RS, column offset n % RS; the RS values of one row are contiguous -> one vector store).
template <int COUT, int RS, bool Q, bool Y>
__device__ __forceinline__ void final_store(const int (&code)[16], const float (&o)[16], int8_t* __restrict__ qb,
                                            float* __restrict__ yb, const long long HW, const long long plane, const int Wr,
                                            const bool valid) {
  constexpr int R2 = RS * RS;
  if (Q) {
    int8_t* p = qb;
#pragma unroll
    for (int n = 0; n < COUT; ++n, p += HW)
      if (valid) *p = (int8_t)code[n];
  }
  if (Y) {
#pragma unroll
    for (int c = 0; c < COUT / R2; ++c) {
      float* pc = yb + c * plane;
#pragma unroll
      for (int i = 0; i < RS; ++i) {
        float* dst = pc + (long long)i * Wr;
        const int n0 = c * R2 + i * RS;
        if (valid) {
          if constexpr (RS == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[n0], o[n0 + 1], o[n0 + 2], o[n0 + 3]);
          else if constexpr (RS == 2) *reinterpret_cast<float2*>(dst) = make_float2(o[n0], o[n0 + 1]);
          else *dst = o[n0];
        }
      }
    }
  }
}

// Final-layer arithmetic for the first NP channel pairs: q5 = clamp(rint(y + z5)); out = (q5 - z5) * s5
// (quan_func.py:585,:594).  rint via the 1.5*2^23 trick: rm = 12582912 + q5 exactly, its low byte is the int8 code.
template <int NP>
__device__ __forceinline__ void final_math(const UmmaParams& P, const float2 (&yv)[8], int (&code)[16], float (&o)[16]) {
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    float2 t = f2_add(yv[k], P.z_next);
    t.x = fminf(fmaxf(t.x, -128.f), 127.f);
    t.y = fminf(fmaxf(t.y, -128.f), 127.f);
    const float2 rm = f2_add(t, MAGIC_F);
    code[2 * k] = __float_as_int(rm.x);
    code[2 * k + 1] = __float_as_int(rm.y);
    const float2 ov = f2_mul(f2_add(rm, P.out_off), P.s_out);
    o[2 * k] = ov.x;
    o[2 * k + 1] = ov.y;
  }
}

// Final layer of the x2 network (12 channels, r = 2), math and stores interleaved pair by pair: the packed result of
// channels (2k, 2k+1) IS the float2 that PixelShuffle(2) places at output row i = k & 1 of plane k >> 1, so it is stored
// straight from its register pair (no staging arrays, no register moves).
template <bool Q, bool Y>
__device__ __forceinline__ void final_x2(const UmmaParams& P, const float2 (&yv)[8], int8_t* __restrict__ qb,
                                         float* __restrict__ yb, const long long HW, const long long plane, const int Wr,
                                         const bool valid) {
  int8_t* qp = qb;
  float* yp = yb;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float2 t = f2_add(yv[k], P.z_next);
    t.x = fminf(fmaxf(t.x, -128.f), 127.f);
    t.y = fminf(fmaxf(t.y, -128.f), 127.f);
    const float2 rm = f2_add(t, MAGIC_F);
    if (Q) {
      if (valid) {
        st_q(qp, __float_as_int(rm.x));
        st_q(qp + HW, __float_as_int(rm.y));
      }
      qp += 2 * HW;
    }
    if (Y) {
      const float2 ov = f2_mul(f2_add(rm, P.out_off), P.s_out);
      if (valid) st_y2(yp, ov);
      yp += (k & 1) ? plane - Wr : Wr;                       // row 0 -> row 1 of the plane -> row 0 of the next plane
    }
  }
}

// ------------------------------------------------------------------------------------------------
// epilogue of one output pixel: 16 (x NS) accumulator columns -> codes.  Shared by both kernels.
// FLOOR: apply max(y, relu_floor) explicitly.  It is only needed for a ReLU layer whose consumer has z > -128:
// with z <= -128, clamp(rint(max(y,0) + z)) == clamp(rint(y + z)) because every y < 0 saturates at -128 either way.
// ------------------------------------------------------------------------------------------------
// OUTM (final layer): 0 = both outputs may be wanted, decided by the pointers at run time; 1 = q5 only, 2 = y only
// (compile time: the single-output kernels carry only their own store sequence, which makes them markedly faster).
// COUT (final layer): the output channel count when it is one of the reference's networks (x4: 16 with r = 4, x2: 12 with
// r = 2, NRDM: 3 with r = 1) -- those kernels carry only their own store sequence; 0 = any count, decided at run time.
template <int NS, int EPI, int RS, int COUT, int OUTM, bool FLOOR>
__device__ __forceinline__ void epilogue_pixel(const UmmaParams& P, int (&acc)[NS][16], const uint4 sc,
                                               uint4* __restrict__ out_px, int8_t* __restrict__ qb,
                                               float* __restrict__ yb, const long long HW, const float relu_floor,
                                               const bool valid) {
  // ---- m = bits of the float 12582912 + A,  A = clampAdd(sum_p clampAcc(P_p)) + K  (quan_func.py:370,437,491) ----
  // NS = 2 (correction columns in set 1) and the spare-column layout of NS = 1 final layers (columns 12 + j): the
  // adder-tree clamp is proved inert; the few PE clamps that can fire are corrected here by d = clamp(v) - v on the
  // column they belong to.  Firing is rare, so the hot path is the test only and the fix-up is kept compact (it indexes
  // the accumulator by a run-time column: 12 predicated adds per pair instead of an unrolled decision tree, which used
  // to make up 90 % of this kernel's code).
  if (NS == 2 || (NS == 1 && EPI == EPI_FINAL)) {
    if (P.n_risky > 0) {
      int d[4];
      bool fire = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {              // both layouts carry at most 4 risky pairs (accumulator_sets)
        const int v = (NS == 2) ? acc[NS - 1][j] : acc[0][12 + j];
        d[j] = (j < P.n_risky) ? clampi(v, P.lo_acc, P.hi_acc) - v : 0;
        fire |= (d[j] != 0);
      }
      if (fire) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = P.risky_col[j];
#pragma unroll
          for (int n = 0; n < (NS == 2 ? 16 : 12); ++n) acc[0][n] += (n == col) ? d[j] : 0;
        }
      }
    }
  }
  float2 yv[8];                                        // channel pairs (2k, 2k+1)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int m[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = 2 * k + h;
      if (NS <= 2) {
        m[h] = acc[0][n];
      } else {
        int s = 0;
#pragma unroll
        for (int p = 0; p < NS; ++p) s += clampi(acc[p][n], P.lo_acc, P.hi_acc);
        m[h] = clampi(s, P.lo_add, P.hi_add) + P.c_kmagic[n];
      }
    }
    yv[k] = f2_fma(make_float2(__int_as_float(m[0]), __int_as_float(m[1])), P.mf, P.nbias);   // RN(A * M16 * 2^-n)  (:529)
    if (FLOOR) { yv[k].x = fmaxf(yv[k].x, relu_floor); yv[k].y = fmaxf(yv[k].y, relu_floor); }
  }
  if (EPI == EPI_FINAL) {
    // q5 = clamp(rint(y + z5)); out = (q5 - z5) * s5 (:585,:594); columns are in natural channel order.
    // rint via the 1.5*2^23 trick: rm = 12582912 + q5 exactly, its low byte is the int8 code.
    constexpr int R2 = RS * RS;
    constexpr bool Q = OUTM != 2, Y = OUTM != 1;           // OUTM 0 is only launched with both pointers (launch_epi)
    int code[16];
    float o[16];
    const long long plane = HW * R2;                       // one pixel-shuffled output plane
    const int Wr = P.W * RS;
    if (RS == 4 && COUT == 16) { final_math<8>(P, yv, code, o); final_store<16, RS, Q, Y>(code, o, qb, yb, HW, plane, Wr, valid); }
    else if (RS == 2 && COUT == 12) final_x2<Q, Y>(P, yv, qb, yb, HW, plane, Wr, valid);
    else if (RS == 1 && COUT == 3) { final_math<2>(P, yv, code, o); final_store<3, RS, Q, Y>(code, o, qb, yb, HW, plane, Wr, valid); }
    else {
      final_math<8>(P, yv, code, o);
      if (Q) {
#pragma unroll
        for (int n = 0; n < 16; ++n)
          if (n < P.cout && valid) qb[(long long)n * HW] = (int8_t)code[n];
      }
      if (Y) {
#pragma unroll
        for (int n0 = 0; n0 < 16; n0 += RS) {
          if (n0 < P.cout && valid) {
            float* dst = yb + (long long)(n0 / R2) * plane + (long long)((n0 % R2) / RS) * Wr;
            if constexpr (RS == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[n0], o[n0 + 1], o[n0 + 2], o[n0 + 3]);
            else if constexpr (RS == 2) *reinterpret_cast<float2*>(dst) = make_float2(o[n0], o[n0 + 1]);
            else *dst = o[n0];
          }
        }
      }
    }
  } else if (EPI == EPI_RESIDUAL) {
    // ic = clamp(rint(relu(y) - 128)) (:251; the low clamp subsumes the ReLU), rc from the shortcut (:250),
    // t = (rc + ic + 256) * Mres * 2^-nres (:252,:269), q4 = clamp(rint(t + z4)) (:270).
    // Bytes: shortcut byte = rc + 128, (ic ^ 0x80) = ic + 128, so their sum IS rc + ic + 256.
    uint32_t ow[4];
    const uint32_t scw[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 u0 = f2_add(yv[2 * w], -128.f), u1 = f2_add(yv[2 * w + 1], -128.f);
      const uint32_t vic = pack4_sat_s8(__float2int_rn(u0.x), __float2int_rn(u0.y), __float2int_rn(u1.x), __float2int_rn(u1.y)) ^
                           0x80808080u;
      // 16-bit lanes: even bytes and odd bytes
      const uint32_t s_even = (scw[w] & 0x00ff00ffu) + (vic & 0x00ff00ffu);
      const uint32_t s_odd = __byte_perm(scw[w], 0u, 0x4341u) + __byte_perm(vic, 0u, 0x4341u);
      // bits of the float 12582912 + S: low two bytes from the 16-bit lane, high two from the magic constant
      const float2 m01 = make_float2(__uint_as_float(__byte_perm(s_even, (uint32_t)MAGIC_BITS, 0x7610u)),
                                     __uint_as_float(__byte_perm(s_odd, (uint32_t)MAGIC_BITS, 0x7610u)));
      const float2 m23 = make_float2(__uint_as_float(__byte_perm(s_even, (uint32_t)MAGIC_BITS, 0x7632u)),
                                     __uint_as_float(__byte_perm(s_odd, (uint32_t)MAGIC_BITS, 0x7632u)));
      const float2 t01 = f2_add(f2_fma(m01, P.res_mf, P.res_nbias), P.z_res);
      const float2 t23 = f2_add(f2_fma(m23, P.res_mf, P.res_nbias), P.z_res);
      ow[w] = pack4_sat_s8(__float2int_rn(t01.x), __float2int_rn(t01.y), __float2int_rn(t23.x), __float2int_rn(t23.y)) ^
              0x80808080u;
    }
    if (valid) *out_px = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  } else {
    uint32_t ow[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 t0 = f2_add(yv[2 * w], P.z_next), t1 = f2_add(yv[2 * w + 1], P.z_next);     // :280, clamped by the pack
      ow[w] = pack4_sat_s8(__float2int_rn(t0.x), __float2int_rn(t0.y), __float2int_rn(t1.x), __float2int_rn(t1.y)) ^
              0x80808080u;                                                                      // v = q + 128
    }
    if (valid) *out_px = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// TMEM <-> register helpers of the epilogue
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void reg_fence16(int (&r)[16]) {
  // the values were produced asynchronously (tcgen05.ld): pin every later use behind the preceding wait::ld
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}

// TMEM <-> registers for one row slot.  NS = 2: only the first 4 columns of set 1 are live.
template <int NS>
__device__ __forceinline__ void slot_load(uint32_t taddr, int (&a)[NS][16]) {
  tmem_ld16(taddr, a[0]);
  if (NS == 2) tmem_ld4(taddr + 16, a[1]);
  if (NS == 4) {
#pragma unroll
    for (int p = 1; p < NS; ++p) tmem_ld16(taddr + p * 16, a[p]);
  }
}
template <int NS>
__device__ __forceinline__ void slot_store(uint32_t taddr, const int (&a)[NS][16]) {
  tmem_st16(taddr, a[0]);
  if (NS == 2) tmem_st4(taddr + 16, a[1]);
  if (NS == 4) {
#pragma unroll
    for (int p = 1; p < NS; ++p) tmem_st16(taddr + p * 16, a[p]);
  }
}
template <int NS>
__device__ __forceinline__ void slot_fence(int (&a)[NS][16]) {
  reg_fence16(a[0]);
  if (NS == 2) asm volatile("" : "+r"(a[1][0]), "+r"(a[1][1]), "+r"(a[1][2]), "+r"(a[1][3]));
  if (NS == 4) {
#pragma unroll
    for (int p = 1; p < NS; ++p) reg_fence16(a[p]);
  }
}


__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}


// ------------------------------------------------------------------------------------------------
// host helpers shared by the tcgen05 translation units (defined in sesrq_umma.cu)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn();
constexpr int MAX_DEVICES = 64;
int sm_count();
bool needs_floor(const LayerConst& c);
void fill_epilogue_params(UmmaParams& P, const LayerConst& c, int ns);

// The dynamic shared memory opt-in is a per-device (per-context) attribute of the kernel: set it once per
// (kernel instantiation, device).  cudaFuncSetAttribute is idempotent, so a race between two threads is benign.
// `done`: one bit per device; the caller keeps one variable per kernel instantiation (a function-pointer type is shared
// by all instantiations with the same signature, so it cannot live in here).
// Launch with the programmatic-stream-serialization attribute (see pdl_wait): every kernel launched this way calls
// pdl_wait() before its first access to activations.  SESRQ_PDL=0 (read once) launches plainly.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <typename Kern>
static int ensure_smem_optin(Kern kern, int bytes, std::atomic<uint64_t>& done) {
  int dev = 0;
  SESRQ_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < MAX_DEVICES && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return SESRQ_OK;
  SESRQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (dev >= 0 && dev < MAX_DEVICES) done.fetch_or(1ull << dev, std::memory_order_release);
  return SESRQ_OK;
}

}  // namespace sesrq
