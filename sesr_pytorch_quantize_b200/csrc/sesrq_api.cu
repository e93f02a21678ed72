// C ABI of libsesrq: network handle, host-side constant folding, forward orchestration.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>
#include "sesrq_internal.h"

namespace sesrq {

std::atomic<uint64_t> g_launches{0};
static thread_local std::string t_cuda_error;

void set_cuda_error(cudaError_t e, const char* where) {
  t_cuda_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + where;
}
const char* get_cuda_error() { return t_cuda_error.c_str(); }

static const int kNextDomain[SESRQ_NUM_LAYERS] = {1, 2, 3, 1, 5};   // quan_func.py:525,601,556,579

// exact range of the per-PE and total partial sums over all codes q in [-128,127]
static void clamp_bounds(const int8_t* w, int cout, int cin, int k, int acc_bits, int add_bits,
                         bool* inert_acc, bool* inert_add, int32_t* risky) {
  const long long lo_acc = -(1LL << (acc_bits - 1)), hi_acc = (1LL << (acc_bits - 1)) - 1;
  const long long lo_add = -(1LL << (add_bits - 1)), hi_add = (1LL << (add_bits - 1)) - 1;
  *inert_acc = true;
  *inert_add = true;
  for (int co = 0; co < cout; ++co) {
    long long hi_p[4] = {0, 0, 0, 0}, lo_p[4] = {0, 0, 0, 0};
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < k * k; ++t) {
        long long wv = w[((size_t)co * cin + c) * k * k + t];
        long long a = wv * -128, b = wv * 127;
        hi_p[c & 3] += std::max(a, b);
        lo_p[c & 3] += std::min(a, b);
      }
    long long hi = 0, lo = 0;
    risky[co] = -1;
    for (int p = 0; p < 4; ++p) {
      if (hi_p[p] > hi_acc || lo_p[p] < lo_acc) {
        *inert_acc = false;
        risky[co] = (risky[co] < 0) ? p : 4;
      }
      hi += hi_p[p];
      lo += lo_p[p];
    }
    if (hi > hi_add || lo < lo_add) *inert_add = false;
  }
}

}  // namespace sesrq

using namespace sesrq;

struct SesrqNet {
  SesrqNetDesc desc;
  LayerDev layers[SESRQ_NUM_LAYERS];
  std::vector<int8_t> h_w[SESRQ_NUM_LAYERS];
  bool need_rc = false;       // z_1 != -128: shortcut codes differ from q_1 (quan_func.py:250 vs :280)
  bool need_a0 = false;       // layer 0 runs on the generic path -> separate input-quantise pass
  int alternate = 1;          // 1: consecutive launches walk their tiles in opposite directions (SESRQ_ALTERNATE=0 disables)
  bool fuse12 = true;         // layers 1 + 2 in one launch (sesrq_chain.cu); domain 2 is then never materialised
  int device = 0;
  size_t chunk_bytes = (size_t)2048 << 20;   // activation workspace per pass (SESRQ_CHUNK_MB, read at create)
  int host_chunk = 0;                        // frames per host-pipeline chunk (SESRQ_HOST_CHUNK, read at create; 0 = auto)
  // host-call staging (sesrq_net_forward_i8_host): two sets of device buffers, copy-in / copy-out streams
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_start = nullptr;
  // per-kernel event timing (sesrq_net_profile_*)
  mutable bool profiling = false;
  struct EvRec { cudaEvent_t a, b; int slot; };
  mutable std::vector<EvRec> ev_used;
  mutable std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_free;
};

namespace {
// brackets one launch with events when profiling is on
struct ProfScope {
  const SesrqNet* net; cudaStream_t st; int slot; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(const SesrqNet* n, cudaStream_t s, int sl) : net(n), st(s), slot(sl) {
    if (!net->profiling) return;
    if (!net->ev_free.empty()) { a = net->ev_free.back().first; b = net->ev_free.back().second; net->ev_free.pop_back(); }
    else if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    net->ev_used.push_back({a, b, slot});
  }
};
}  // namespace

namespace {

struct WsPlan {
  int chunk;                 // frames per pass
  size_t act_bytes;          // one NHWC16 activation buffer for `chunk` frames
  size_t off_a[5];           // A0..A4
  size_t off_rc;
  size_t total;
};

WsPlan plan_ws(const SesrqNet* net, int B, int H, int W) {
  WsPlan p{};
  size_t per_frame = (size_t)H * W * 16;
  int nbuf = 4 + (net->need_a0 ? 1 : 0) + (net->need_rc ? 1 : 0);
  size_t per_frame_all = per_frame * nbuf;
  // Frames per pass: as many as fit the workspace budget, then balanced over the passes.  Large passes win:
  // each launch pays a ramp-up/tail and the static tile schedule quantises to the SM count (measured:
  // 1080p x2 233 -> 167 us/frame going from 1 to 16 frames per pass), which outweighs L2 residency.
  long long c = per_frame_all ? (long long)(net->chunk_bytes / per_frame_all) : 1;
  c = std::max<long long>(1, std::min<long long>(c, std::max(B, 1)));
  const long long passes = (std::max(B, 1) + c - 1) / c;
  p.chunk = (int)((std::max(B, 1) + passes - 1) / passes);
  p.act_bytes = (per_frame * p.chunk + 255) & ~(size_t)255;
  size_t off = 0;
  for (int i = 0; i < 5; ++i) {
    p.off_a[i] = off;
    if (i > 0 || net->need_a0) off += p.act_bytes;
  }
  p.off_rc = off;
  if (net->need_rc) off += p.act_bytes;
  p.total = off;
  return p;
}

}  // namespace

extern "C" {

int sesrq_abi_version(void) { return SESRQ_ABI_VERSION; }

const char* sesrq_strerror(int code) {
  switch (code) {
    case SESRQ_OK: return "ok";
    case SESRQ_EINVAL: return "invalid argument or unsupported shape";
    case SESRQ_ENOMEM: return "workspace too small or out of memory";
    case SESRQ_ECUDA: return "CUDA runtime error";
    case SESRQ_EARCH: return "device is not an sm_100 (B200) GPU";
    case SESRQ_EUNSUPPORTED: return "operation not supported in this configuration";
    default: return "unknown error";
  }
}

const char* sesrq_last_cuda_error(void) { return get_cuda_error(); }
uint64_t sesrq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int sesrq_net_create(const SesrqNetDesc* d, SesrqNet** out_net) {
  if (!d || !out_net) return SESRQ_EINVAL;
  *out_net = nullptr;
  if (d->abi_version != SESRQ_ABI_VERSION || d->n_layers != SESRQ_NUM_LAYERS) return SESRQ_EINVAL;
  if (d->pe_num != 4 || d->quan_bits != 8) return SESRQ_EINVAL;              // define.py:5 "only 4 PEs"
  if (d->pe_acc_bits < 9 || d->pe_acc_bits > 30 || d->pe_add_bits < 9 || d->pe_add_bits > 30) return SESRQ_EINVAL;
  if (d->r_shuffle < 1 || d->r_shuffle > 4) return SESRQ_EINVAL;
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    const SesrqLayerDesc& l = d->layers[i];
    if ((l.ksize != 3 && l.ksize != 5) || l.cin < 1 || l.cin > SESRQ_MAX_CH || l.cout < 1 ||
        l.cout > SESRQ_MAX_CH || !l.h_weight)
      return SESRQ_EINVAL;
    if (l.m16 < 0 || l.m16 > 65535 || l.shift_n < -16 || l.shift_n > 32) return SESRQ_EINVAL;
    if (i > 0 && l.cin != d->layers[i - 1].cout) return SESRQ_EINVAL;
  }
  if (d->layers[0].cout != d->layers[3].cout) return SESRQ_EINVAL;           // shortcut add needs equal channels
  if (d->layers[4].cout % (d->r_shuffle * d->r_shuffle) != 0) return SESRQ_EINVAL;

  int dev = 0;
  SESRQ_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SESRQ_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return SESRQ_EARCH;

  SesrqNet* net = new (std::nothrow) SesrqNet();
  if (!net) return SESRQ_ENOMEM;
  net->desc = *d;
  net->device = dev;
  if (const char* e = getenv("SESRQ_CHUNK_MB")) {
    long v = atol(e);
    if (v > 0) net->chunk_bytes = (size_t)v << 20;
  }
  if (const char* e = getenv("SESRQ_HOST_CHUNK")) net->host_chunk = std::max(0, atoi(e));
  if (const char* e = getenv("SESRQ_FUSE")) net->fuse12 = atoi(e) != 0;
  if (const char* e = getenv("SESRQ_ALTERNATE")) net->alternate = atoi(e) != 0 ? 1 : 0;
  int zs[SESRQ_NUM_DOMAINS];
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) zs[i] = d->layers[i].z_in;
  zs[5] = d->z_out;
  net->need_rc = (zs[1] != -128);

  int rc = SESRQ_OK;
  for (int i = 0; i < SESRQ_NUM_LAYERS && rc == SESRQ_OK; ++i) {
    const SesrqLayerDesc& l = d->layers[i];
    LayerDev& L = net->layers[i];
    LayerConst& c = L.c;
    memset(&c, 0, sizeof(c));
    const int taps = l.ksize * l.ksize;
    net->h_w[i].assign(l.h_weight, l.h_weight + (size_t)l.cout * l.cin * taps);
    net->desc.layers[i].h_weight = net->h_w[i].data();
    const int8_t* w = net->h_w[i].data();
    c.k = l.ksize;
    c.pad = (l.ksize - 1) / 2;
    c.cin = l.cin;
    c.cout = l.cout;
    c.pad_byte = std::max(l.z_in, -128) + 128;                                 // quan_func.py:351-352
    if (c.pad_byte > 255) { rc = SESRQ_EINVAL; break; }
    c.lo_acc = -(1 << (d->pe_acc_bits - 1));
    c.hi_acc = (1 << (d->pe_acc_bits - 1)) - 1;
    c.lo_add = -(1 << (d->pe_add_bits - 1));
    c.hi_add = (1 << (d->pe_add_bits - 1)) - 1;
    bool inert_acc, inert_add;
    clamp_bounds(w, l.cout, l.cin, l.ksize, d->pe_acc_bits, d->pe_add_bits, &inert_acc, &inert_add, c.pe_risky);
    c.pe_exact = (inert_acc && inert_add) ? 0 : 1;
    c.add_inert = inert_add ? 1 : 0;
    c.epi = (i == 3) ? EPI_RESIDUAL : (i == 4 ? EPI_FINAL : EPI_PLAIN);
    c.relu = l.relu;
    c.save_rc = (i == 0 && net->need_rc) ? 1 : 0;
    c.r_shuffle = d->r_shuffle;
    c.first_layer = (i == 0) ? 1 : 0;
    c.mf = (float)((double)l.m16 * std::ldexp(1.0, -l.shift_n));               // exact: m16 < 2^16
    // layer 3 requantises into domain 1 (:554-560) but its codes are consumed by the residual
    // quantiser of layer 4, which adds z_4 (:270); z_next is unused by EPI_RESIDUAL.
    c.z_next = (float)zs[i == 3 ? 4 : kNextDomain[i]];
    c.res_mf = (float)((double)d->res_m16 * std::ldexp(1.0, -d->res_shift_n));
    c.z_res = (float)zs[4];
    c.s_out = (float)d->scale_out;
    c.in_scale = (float)d->scale_in;
    c.in_zero = (float)zs[0];
    for (int co = 0; co < l.cout; ++co) {
      long long sum_p[4] = {0, 0, 0, 0};
      for (int ci = 0; ci < l.cin; ++ci)
        for (int t = 0; t < taps; ++t) sum_p[ci & 3] += w[((size_t)co * l.cin + ci) * taps + t];
      long long fused = l.add_const[co];
      for (int p = 0; p < 4; ++p) {
        c.pe_corr[p][co] = (int32_t)(-128 * sum_p[p]);
        fused += c.pe_corr[p][co];
      }
      c.add_const[co] = l.add_const[co];
      c.fused_const[co] = (int32_t)fused;
    }
    // generic-path weights: [tap][co][pe] words, byte m of word p = W[co][4m + p][tap]
    std::vector<int32_t> wpk((size_t)taps * 16 * 4, 0);
    for (int t = 0; t < taps; ++t)
      for (int co = 0; co < l.cout; ++co)
        for (int p = 0; p < 4; ++p) {
          uint32_t word = 0;
          for (int m = 0; m < 4; ++m) {
            int ci = 4 * m + p;
            int8_t v = (ci < l.cin) ? w[((size_t)co * l.cin + ci) * taps + t] : (int8_t)0;
            word |= (uint32_t)(uint8_t)v << (8 * m);
          }
          wpk[((size_t)t * 16 + co) * 4 + p] = (int32_t)word;
        }
    if (cudaMalloc(&L.d_wpk, wpk.size() * 4) != cudaSuccess ||
        cudaMemcpy(L.d_wpk, wpk.data(), wpk.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      set_cuda_error(cudaGetLastError(), "weight upload");
      rc = SESRQ_ECUDA;
      break;
    }
    // kernel family
    L.impl = SESRQ_IMPL_GENERIC;
    if (d->impl == SESRQ_IMPL_UMMA || d->impl == SESRQ_IMPL_AUTO) {
      if (umma_supported(c)) {
        rc = umma_prepare_layer(L, w);
        if (rc == SESRQ_OK) L.impl = SESRQ_IMPL_UMMA;
      } else if (d->impl == SESRQ_IMPL_UMMA) {
        rc = SESRQ_EUNSUPPORTED;
      }
    }
  }
  if (rc != SESRQ_OK) {
    sesrq_net_destroy(net);
    return rc;
  }
  net->need_a0 = (net->layers[0].impl == SESRQ_IMPL_GENERIC);
  *out_net = net;
  return SESRQ_OK;
}

void sesrq_net_destroy(SesrqNet* net) {
  if (!net) return;
  for (int i = 0; i < SESRQ_NUM_LAYERS; ++i) {
    if (net->layers[i].d_wpk) cudaFree(net->layers[i].d_wpk);
    if (net->layers[i].d_wumma) cudaFree(net->layers[i].d_wumma);
  }
  if (net->d_stage) cudaFree(net->d_stage);
  if (net->s_in) cudaStreamDestroy(net->s_in);
  if (net->s_out) cudaStreamDestroy(net->s_out);
  for (int i = 0; i < 2; ++i) {
    if (net->ev_in[i]) cudaEventDestroy(net->ev_in[i]);
    if (net->ev_comp[i]) cudaEventDestroy(net->ev_comp[i]);
    if (net->ev_out[i]) cudaEventDestroy(net->ev_out[i]);
  }
  if (net->ev_start) cudaEventDestroy(net->ev_start);
  for (auto& r : net->ev_used) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto& r : net->ev_free) { cudaEventDestroy(r.first); cudaEventDestroy(r.second); }
  delete net;
}

int sesrq_net_layer_info(const SesrqNet* net, int layer, int32_t* impl, int32_t* pe_exact) {
  if (!net || layer < 0 || layer >= SESRQ_NUM_LAYERS) return SESRQ_EINVAL;
  if (impl) *impl = net->layers[layer].impl;
  if (pe_exact) *pe_exact = net->layers[layer].c.pe_exact;
  return SESRQ_OK;
}

int sesrq_net_set_fusion(SesrqNet* net, int enable) {
  if (!net) return SESRQ_EINVAL;
  const int prev = net->fuse12 ? 1 : 0;
  net->fuse12 = enable != 0;
  return prev;
}

int sesrq_net_fused_layers(const SesrqNet* net) {
  return (net && net->fuse12 && chain2_supported(net->layers[1], net->layers[2])) ? 1 : 0;
}

int sesrq_net_layer_layout(const SesrqNet* net, int layer) {
  if (!net || layer < 0 || layer >= SESRQ_NUM_LAYERS) return SESRQ_EINVAL;
  return net->layers[layer].impl == SESRQ_IMPL_UMMA ? net->layers[layer].layout : SESRQ_LAYOUT_GENERIC;
}

size_t sesrq_ws_bytes(const SesrqNet* net, int B, int H, int W) {
  if (!net || B < 0 || H <= 0 || W <= 0) return 0;
  return plan_ws(net, B, H, W).total;
}

int sesrq_net_forward_i8(const SesrqNet* net, const float* d_x, int B, int H, int W, int8_t* d_q5,
                         float* d_y, void* d_ws, size_t ws_bytes, void* stream) {
  if (!net || B < 0 || H <= 0 || W <= 0) return SESRQ_EINVAL;
  if (B == 0) return SESRQ_OK;
  if (!d_x || !d_ws) return SESRQ_EINVAL;
  if (!d_q5 && !d_y) return SESRQ_EINVAL;                       // nothing to materialise
  if ((long long)H * W > (1LL << 30)) return SESRQ_EINVAL;
  WsPlan p = plan_ws(net, B, H, W);
  if (ws_bytes < p.total) return SESRQ_ENOMEM;
  if ((reinterpret_cast<uintptr_t>(d_ws) & 15) != 0) return SESRQ_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(d_ws);
  uint8_t* A[5];
  for (int i = 0; i < 5; ++i) A[i] = ws + p.off_a[i];
  uint8_t* RC = net->need_rc ? ws + p.off_rc : nullptr;
  const int cin = net->layers[0].c.cin, cout = net->layers[4].c.cout;
  const int r = net->desc.r_shuffle;
  const size_t hw = (size_t)H * W;

  for (int b0 = 0; b0 < B; b0 += p.chunk) {
    const int nb = std::min(p.chunk, B - b0);
    const float* x = d_x + (size_t)b0 * cin * hw;
    int8_t* q5 = d_q5 ? d_q5 + (size_t)b0 * cout * hw : nullptr;
    float* y = d_y ? d_y + (size_t)b0 * cout * hw : nullptr;     // Cout/r^2 * rH * rW == Cout * H * W
    (void)r;
    int rc;
    int rev = 0;                                  // tile order alternates from launch to launch (L2 reuse, launch_conv_umma)
    // layer 0
    if (net->layers[0].impl == SESRQ_IMPL_UMMA) {
      ProfScope ps(net, st, 0);
      rc = launch_conv_umma(net->layers[0], x, nullptr, nullptr, A[1], nullptr, nullptr, nb, H, W, st, rev);
    } else {
      {
        ProfScope ps(net, st, SESRQ_NUM_LAYERS);
        rc = launch_quant_input(x, nb, cin, H, W, net->layers[0].c.in_scale, net->layers[0].c.in_zero, A[0], st);
      }
      if (rc) return rc;
      ProfScope ps(net, st, 0);
      rc = launch_conv_generic(net->layers[0], A[0], nullptr, A[1], RC, nullptr, nullptr, nb, H, W, st);
    }
    if (rc) return rc;
    // layers 1..4
    for (int i = 1; i < SESRQ_NUM_LAYERS; ++i) {
      const LayerDev& L = net->layers[i];
      if (i == 1 && net->fuse12 && chain2_supported(net->layers[1], net->layers[2])) {
        ProfScope ps(net, st, 1);                 // both layers in one launch, A[2] is skipped
        rc = launch_conv_chain2(net->layers[1], net->layers[2], A[1], A[3], nb, H, W, st, rev ^= net->alternate);
        if (rc) return rc;
        i = 2;
        continue;
      }
      const uint8_t* sc = (i == 3) ? (net->need_rc ? RC : A[1]) : nullptr;
      uint8_t* out = (i < 4) ? A[i + 1] : nullptr;
      ProfScope ps(net, st, i);
      if (L.impl == SESRQ_IMPL_UMMA)
        rc = launch_conv_umma(L, nullptr, A[i], sc, out, i == 4 ? q5 : nullptr, i == 4 ? y : nullptr, nb, H, W, st, rev ^= net->alternate);
      else
        rc = launch_conv_generic(L, A[i], sc, out, nullptr, i == 4 ? q5 : nullptr, i == 4 ? y : nullptr, nb, H, W, st);
      if (rc) return rc;
    }
  }
  return SESRQ_OK;
}

int sesrq_net_profile_enable(SesrqNet* net, int enable) {
  if (!net) return SESRQ_EINVAL;
  net->profiling = enable != 0;
  return SESRQ_OK;
}

int sesrq_net_profile_read(SesrqNet* net, double* ms, uint64_t* launches) {
  if (!net || !ms || !launches) return SESRQ_EINVAL;
  for (auto& r : net->ev_used) {
    SESRQ_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    SESRQ_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.slot] += t;
    launches[r.slot] += 1;
    net->ev_free.emplace_back(r.a, r.b);
  }
  net->ev_used.clear();
  return SESRQ_OK;
}

int sesrq_ws_read_act_i8(const SesrqNet* net, const void* d_ws, size_t ws_bytes, int B, int H, int W,
                         int domain, int8_t* d_q_nchw, void* stream) {
  if (!net || !d_ws || !d_q_nchw || domain < 0 || domain > 4 || B <= 0) return SESRQ_EINVAL;
  WsPlan p = plan_ws(net, B, H, W);
  if (ws_bytes < p.total) return SESRQ_ENOMEM;
  if (B > p.chunk) return SESRQ_EUNSUPPORTED;        // chunked forward reuses the buffers
  if (domain == 2 && sesrq_net_fused_layers(net)) return SESRQ_EUNSUPPORTED;   // never materialised (sesrq_chain.cu)
  if (domain == 0 && !net->need_a0) return SESRQ_EUNSUPPORTED;   // fused into the layer-0 kernel
  const int C = (domain == 0) ? net->layers[0].c.cin : net->layers[domain - 1].c.cout;
  const uint8_t* act = static_cast<const uint8_t*>(d_ws) + p.off_a[domain];
  return launch_act_to_nchw(act, B, C, H, W, d_q_nchw, reinterpret_cast<cudaStream_t>(stream));
}

int sesrq_net_forward_i8_host(SesrqNet* net, const float* h_x, int B, int H, int W, int8_t* h_q5,
                              float* h_y, void* stream) {
  if (!net || B < 0 || H <= 0 || W <= 0) return SESRQ_EINVAL;
  if (B == 0) return SESRQ_OK;
  if (!h_x || (!h_q5 && !h_y)) return SESRQ_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t hw = (size_t)H * W;
  const int cin = net->layers[0].c.cin, cout = net->layers[4].c.cout;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  // The batch is cut into chunks that flow through a 3-stage pipeline (H2D on s_in, the five layers on the caller's
  // stream, D2H on s_out): PCIe is full duplex, so the input copy of chunk i+1 and the output copy of chunk i-1
  // overlap the kernels of chunk i.  Device staging is double-buffered.
  const size_t frame_io = (size_t)cin * hw * 4 + (h_q5 ? (size_t)cout * hw : 0) + (h_y ? (size_t)cout * hw * 4 : 0);
  // chunk size: about B/8 frames, at least 2 frames and at least ~32 MB of PCIe traffic (measured on 16 x 1080p:
  // 1 frame/chunk 54.5 ms, 2: 38.5, 4: 39.4, 8: 40.2, unpipelined 45.1; the D2H volume alone is 36 ms at 55 GB/s)
  int chunk = (int)std::max<size_t>((B + 7) / 8, (((size_t)32 << 20) + frame_io - 1) / std::max<size_t>(frame_io, 1));
  chunk = std::min(std::max(chunk, 2), B);
  if (net->host_chunk > 0) chunk = std::min(B, net->host_chunk);             // tuning experiments
  const int nchunks = (B + chunk - 1) / chunk;
  const size_t x_bytes = up((size_t)chunk * cin * hw * 4);
  const size_t q_bytes = h_q5 ? up((size_t)chunk * cout * hw) : 0;
  const size_t y_bytes = h_y ? up((size_t)chunk * cout * hw * 4) : 0;
  // the workspace plan balances frames over passes, so it is not monotonic in the frame count: size it for the
  // full chunks AND for the tail chunk
  const int tail = B % chunk;
  const size_t ws_bytes = std::max(sesrq_ws_bytes(net, chunk, H, W), tail ? sesrq_ws_bytes(net, tail, H, W) : (size_t)0);
  const size_t set_bytes = x_bytes + q_bytes + y_bytes;
  const size_t need = 2 * set_bytes + ws_bytes;
  if (need > net->stage_bytes) {
    if (net->d_stage) SESRQ_CUDA(cudaFree(net->d_stage));
    net->d_stage = nullptr;
    net->stage_bytes = 0;
    SESRQ_CUDA(cudaMalloc(&net->d_stage, need));
    net->stage_bytes = need;
  }
  if (!net->s_in) {
    SESRQ_CUDA(cudaStreamCreateWithFlags(&net->s_in, cudaStreamNonBlocking));
    SESRQ_CUDA(cudaStreamCreateWithFlags(&net->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      SESRQ_CUDA(cudaEventCreateWithFlags(&net->ev_in[i], cudaEventDisableTiming));
      SESRQ_CUDA(cudaEventCreateWithFlags(&net->ev_comp[i], cudaEventDisableTiming));
      SESRQ_CUDA(cudaEventCreateWithFlags(&net->ev_out[i], cudaEventDisableTiming));
    }
    SESRQ_CUDA(cudaEventCreateWithFlags(&net->ev_start, cudaEventDisableTiming));
  }
  uint8_t* base = static_cast<uint8_t*>(net->d_stage);
  void* d_ws = base + 2 * set_bytes;
  // everything already enqueued on the caller's stream comes first
  SESRQ_CUDA(cudaEventRecord(net->ev_start, st));
  SESRQ_CUDA(cudaStreamWaitEvent(net->s_in, net->ev_start, 0));
  SESRQ_CUDA(cudaStreamWaitEvent(net->s_out, net->ev_start, 0));
  for (int k = 0; k < nchunks; ++k) {
    const int s = k & 1;
    const int f0 = k * chunk, nf = std::min(chunk, B - f0);
    uint8_t* set = base + (size_t)s * set_bytes;
    float* d_x = reinterpret_cast<float*>(set);
    int8_t* d_q5 = h_q5 ? reinterpret_cast<int8_t*>(set + x_bytes) : nullptr;
    float* d_y = h_y ? reinterpret_cast<float*>(set + x_bytes + q_bytes) : nullptr;
    // input buffer s is free once the layer-0 kernel of chunk k-2 has read it (ev_comp covers the whole chunk)
    if (k >= 2) SESRQ_CUDA(cudaStreamWaitEvent(net->s_in, net->ev_comp[s], 0));
    SESRQ_CUDA(cudaMemcpyAsync(d_x, h_x + (size_t)f0 * cin * hw, (size_t)nf * cin * hw * 4, cudaMemcpyHostToDevice, net->s_in));
    SESRQ_CUDA(cudaEventRecord(net->ev_in[s], net->s_in));
    SESRQ_CUDA(cudaStreamWaitEvent(st, net->ev_in[s], 0));
    if (k >= 2) SESRQ_CUDA(cudaStreamWaitEvent(st, net->ev_out[s], 0));      // output buffers s drained by chunk k-2's D2H
    int rc = sesrq_net_forward_i8(net, d_x, nf, H, W, d_q5, d_y, d_ws, ws_bytes, stream);
    if (rc) {
      // copies to / from the caller's buffers may still be in flight: drain them before handing control back
      cudaStreamSynchronize(net->s_in);
      cudaStreamSynchronize(net->s_out);
      cudaStreamSynchronize(st);
      return rc;
    }
    SESRQ_CUDA(cudaEventRecord(net->ev_comp[s], st));
    SESRQ_CUDA(cudaStreamWaitEvent(net->s_out, net->ev_comp[s], 0));
    if (h_q5)
      SESRQ_CUDA(cudaMemcpyAsync(h_q5 + (size_t)f0 * cout * hw, d_q5, (size_t)nf * cout * hw, cudaMemcpyDeviceToHost, net->s_out));
    if (h_y)
      SESRQ_CUDA(cudaMemcpyAsync(h_y + (size_t)f0 * cout * hw, d_y, (size_t)nf * cout * hw * 4, cudaMemcpyDeviceToHost, net->s_out));
    SESRQ_CUDA(cudaEventRecord(net->ev_out[s], net->s_out));
  }
  SESRQ_CUDA(cudaStreamSynchronize(net->s_out));
  SESRQ_CUDA(cudaStreamSynchronize(st));
  return SESRQ_OK;
}

}  // extern "C"
