/* CPU ORACLE (plain C, OpenMP) for the SESR / NRDM int8 hardware-sim hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Same restatement of the reference algorithm as oracle/sesr_oracle.py (which is pinned bit-for-bit against
 * the unmodified reference by oracle/gen_golden.py); tests/test_c_oracle.py checks this file against the
 * numpy oracle and the committed golden vectors.  It exists so that (a) parity can be checked at the full
 * BASELINE sizes in seconds and (b) bench.py has a CPU baseline that uses all host cores.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Citations are into the reference's myQL/quan_func.py.  Build: oracle/Makefile (-ffp-contract=off is
 * essential: the fp32 requant product must not be fused).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t k, cin, cout, relu;
  int32_t z;          /* input zero point (unclamped) */
  int32_t m16, n;     /* requant */
  const int8_t* w;    /* [cout][cin][k][k] */
  const int32_t* K;   /* [cout] add const */
} OracleLayer;

typedef struct {
  OracleLayer layers[5];
  double s0, s5;
  int32_t z5, mr, nr, r, acc_bits, add_bits;
} OracleNet;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int quant_add_zero(float t, float z) {        /* :280 clamp(round(t + z)) half-even */
  float s = t + z;
  s = s < -128.f ? -128.f : (s > 127.f ? 127.f : s);
  return (int)nearbyintf(s);
}
static inline float requant(int a, int m16, int n) {          /* :529 (acc * M16) * 2^-n in fp32 */
  float p = (float)a * (float)m16;
  return p * ldexpf(1.0f, -n);
}

/* one layer on one frame: q_in [cin][H][W] int8 codes -> A [cout][H][W] int32 (PE tree + add const) */
static void conv_pe(const OracleNet* net, const OracleLayer* L, const int8_t* q, int H, int W, int32_t* A) {
  const int k = L->k, pad = (k - 1) / 2, cin = L->cin, cout = L->cout;
  const int zp = L->z < -128 ? -128 : L->z;                   /* :351-352 */
  const int lo18 = -(1 << (net->acc_bits - 1)), hi18 = (1 << (net->acc_bits - 1)) - 1;
  const int lo20 = -(1 << (net->add_bits - 1)), hi20 = (1 << (net->add_bits - 1)) - 1;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  /* padded copy, channel-last per PE is not needed: keep planar int16 */
  int16_t* qp = (int16_t*)malloc((size_t)cin * Hp * Wp * sizeof(int16_t));
  for (int c = 0; c < cin; ++c)
    for (int y = 0; y < Hp; ++y)
      for (int x = 0; x < Wp; ++x) {
        int sy = y - pad, sx = x - pad;
        qp[((size_t)c * Hp + y) * Wp + x] = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? q[((size_t)c * H + sy) * W + sx] : (int16_t)zp;
      }
#pragma omp parallel
  {
    int32_t* acc = (int32_t*)malloc((size_t)4 * W * sizeof(int32_t));
#pragma omp for collapse(2) schedule(static)
    for (int co = 0; co < cout; ++co)
      for (int y = 0; y < H; ++y) {
        memset(acc, 0, (size_t)4 * W * sizeof(int32_t));
        for (int c = 0; c < cin; ++c) {
          int32_t* ap = acc + (size_t)(c & 3) * W;            /* PE p takes channels c == p (mod 4), :315 */
          for (int dy = 0; dy < k; ++dy) {
            const int16_t* row = qp + ((size_t)c * Hp + y + dy) * Wp;
            for (int dx = 0; dx < k; ++dx) {
              const int wv = L->w[(((size_t)co * cin + c) * k + dy) * k + dx];
              if (wv == 0) continue;
              const int16_t* r = row + dx;
              for (int x = 0; x < W; ++x) ap[x] += wv * r[x];
            }
          }
        }
        int32_t* out = A + ((size_t)co * H + y) * W;
        for (int x = 0; x < W; ++x) {
          int s = 0;
          for (int p = 0; p < 4; ++p) s += clampi(acc[(size_t)p * W + x], lo18, hi18);   /* :370 */
          out[x] = clampi(s, lo20, hi20) + L->K[co];                                      /* :437, :491 */
        }
      }
    free(acc);
  }
  free(qp);
}

/* x float32 [B][cin][H][W] -> q5 int8 [B][cout][H][W] (nullable), y float32 pixel-shuffled (nullable),
 * trace_q (nullable): int8 [5][16][H][W] codes q0..q4 of frame 0 */
int sesr_oracle_forward(const OracleNet* net, const float* x, int B, int H, int W, int8_t* q5, float* y, int8_t* trace_q) {
  const size_t hw = (size_t)H * W;
  const int cin0 = net->layers[0].cin, coutL = net->layers[4].cout, r = net->r;
  int8_t* qa = (int8_t*)malloc(16 * hw);
  int8_t* q1 = (int8_t*)malloc(16 * hw);       /* rc codes of the shortcut */
  int32_t* A = (int32_t*)malloc(16 * hw * sizeof(int32_t));
  const float s0 = (float)net->s0, s5 = (float)net->s5;
  for (int b = 0; b < B; ++b) {
    const float* xb = x + (size_t)b * cin0 * hw;
    const float z0 = (float)net->layers[0].z;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)(cin0 * hw); ++i) qa[i] = (int8_t)quant_add_zero(xb[i] / s0, z0);   /* :225 */
    for (int li = 0; li < 5; ++li) {
      const OracleLayer* L = &net->layers[li];
      if (trace_q && b == 0) memcpy(trace_q + (size_t)li * 16 * hw, qa, (size_t)L->cin * hw);
      conv_pe(net, L, qa, H, W, A);
      const long long n = (long long)L->cout * hw;
      if (li < 3) {
        const float zn = (float)net->layers[li + 1].z;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; ++i) {
          float v = requant(A[i], L->m16, L->n);
          v = v > 0.f ? v : 0.f;
          if (li == 0) q1[i] = (int8_t)quant_add_zero(v, -128.f);      /* rc, :250 (shortcut = relu(y), :530) */
          qa[i] = (int8_t)quant_add_zero(v, zn);                        /* :280 */
        }
      } else if (li == 3) {
        const float z4 = (float)net->layers[4].z;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; ++i) {
          float v = requant(A[i], L->m16, L->n);
          v = v > 0.f ? v : 0.f;
          int ic = quant_add_zero(v, -128.f);                           /* :251 */
          float t = (float)(q1[i] + ic + 256);                          /* :252 */
          t = (t * (float)net->mr) * ldexpf(1.0f, -net->nr);            /* :269 */
          qa[i] = (int8_t)quant_add_zero(t, z4);                        /* :270 */
        }
      } else {
        const float z5 = (float)net->z5;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; ++i) {
          float v = requant(A[i], L->m16, L->n);
          int q = quant_add_zero(v, z5);                                /* :585 */
          if (q5) q5[(size_t)b * coutL * hw + i] = (int8_t)q;
          if (y) {
            float o = ((float)q - z5) * s5;                             /* :594 */
            int c = (int)(i / (long long)hw);
            long long p = i - (long long)c * hw;
            int yy = (int)(p / W), xx = (int)(p % W);
            int r2 = r * r, cc = c / r2, ii = (c % r2) / r, jj = c % r;
            y[(((size_t)b * (coutL / r2) + cc) * ((size_t)H * r) + ((size_t)yy * r + ii)) * ((size_t)W * r) + ((size_t)xx * r + jj)] = o;
          }
        }
      }
    }
  }
  free(qa); free(q1); free(A);
  return 0;
}

/* Explicit thread count: a launcher's OMP_NUM_THREADS (torchrun exports 1) must not decide how many host cores the
 * CPU baseline uses; bench.py passes the size of its affinity mask. */
void sesr_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int sesr_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
