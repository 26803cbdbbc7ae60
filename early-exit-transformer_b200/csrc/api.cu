// api.cu -- C-ABI glue: error state, device check, GEMM backend dispatch.
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace eec {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// active-item limits are keyed by the stream the kernels are launched on: two models (or two streams) in one process never see each
// other's limit, and the handle a caller passes is the stream it already passes to every entry point
static std::mutex g_active_mu;
static std::vector<std::pair<cudaStream_t, ActiveItems>> g_active;
ActiveItems active_items(cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_active_mu);
  for (const auto& e : g_active)
    if (e.first == stream) return e.second;
  return ActiveItems{nullptr, 0, 0};
}

static bool force_simt_gemm() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EEC_FORCE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EEC_PDL"); v = (e && e[0] == '1') ? 1 : 0; }   // measured: see common.cuh
  return v == 1;
}

#ifdef EEC_EXPERIMENTS   // `make experiments` (eec/libeec_exp.so): the first-generation tcgen05 GEMM kept for A/B timing, EEC_GEMM_V1=1
static bool gemm_v1() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EEC_GEMM_V1"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
#else
static bool gemm_v1() { return false; }
#endif
static int tc_dispatch(const eec_gemm_desc* d, cudaStream_t st, int32_t* argmax, float* entropy, int logsoftmax) {
#ifdef EEC_EXPERIMENTS
  if (gemm_v1()) return gemm_tc(d, st, argmax, entropy, logsoftmax);
#endif
  return gemm_tc2(d, st, argmax, entropy, logsoftmax);
}

// LayerNorm tails for the FFMA path (the tcgen05 path fuses them into the GEMM epilogue)
static int simt_with_tails(const eec_gemm_desc* d, cudaStream_t st) {
  if (!d->ln_out) return gemm_simt(d, st);
  EEC_CHECK_ARG(d->N == 256 && d->out_dtype == EEC_F32, "gemm: LN tail needs N == 256 and fp32 C");
  eec_gemm_desc g = *d;
  g.ln_out = nullptr;
  if (d->ln2_gamma) {
    EEC_CHECK_ARG(d->x_pre != nullptr, "gemm: ln2 needs x_pre");
    g.C = d->x_pre;
  }
  if (int r = gemm_simt(&g, st)) return r;
  eec_stream_t es = reinterpret_cast<eec_stream_t>(st);
  if (d->ln2_gamma) {
    if (int r = eec_layernorm_fwd(d->x_pre, d->ln_gamma, d->ln_beta, d->C, EEC_F32, d->ln_mean, d->ln_rstd, d->M, 256, es)) return r;
    return eec_layernorm_fwd((const float*)d->C, d->ln2_gamma, d->ln2_beta, d->ln_out, d->ln_dtype, d->ln2_mean, d->ln2_rstd, d->M, 256, es);
  }
  return eec_layernorm_fwd((const float*)d->C, d->ln_gamma, d->ln_beta, d->ln_out, d->ln_dtype, d->ln_mean, d->ln_rstd, d->M, 256, es);
}

}  // namespace eec

using namespace eec;

extern "C" const char* eec_last_error(void) { return g_err; }
extern "C" int eec_version(void) { return 100; }
extern "C" long long eec_launch_count(void) { return g_launches.load(); }

extern "C" int eec_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

extern "C" int eec_set_active_items(const int32_t* n_items_dev, int rows_per_item, int pad_items, eec_stream_t stream) {
  EEC_CHECK_ARG(n_items_dev == nullptr || (rows_per_item > 0 && pad_items >= 0), "set_active_items: rows_per_item must be positive and pad_items >= 0 (got %d, %d)", rows_per_item, pad_items);
  std::lock_guard<std::mutex> lk(g_active_mu);
  for (size_t i = 0; i < g_active.size(); ++i)
    if (g_active[i].first == S(stream)) {
      if (n_items_dev) g_active[i].second = ActiveItems{n_items_dev, rows_per_item, pad_items};
      else g_active.erase(g_active.begin() + i);
      return 0;
    }
  if (n_items_dev) g_active.emplace_back(S(stream), ActiveItems{n_items_dev, rows_per_item, pad_items});
  return 0;
}

extern "C" int eec_gemm(const eec_gemm_desc* d, eec_stream_t stream) {
  EEC_CHECK_ARG(d != nullptr, "gemm: NULL descriptor");
  if (d->M == 0 || d->N == 0) return 0;
  EEC_CHECK_ARG(!d->a_colsum || (d->in_dtype == EEC_BF16 && !force_simt_gemm() && !gemm_v1()),
                "gemm: a_colsum is implemented by the bf16 tcgen05 (v3) kernel only; use eec_colsum on the other paths");
  if (d->in_dtype == EEC_F32 || force_simt_gemm()) return simt_with_tails(d, S(stream));
  return tc_dispatch(d, S(stream), nullptr, nullptr, 0);
}

extern "C" int eec_head_logsoftmax(const void* x, int dtype, const void* w, const float* bias, float* out,
                                   int32_t* argmax, float* entropy, float* logits_ws, int rows, int d, int V,
                                   eec_stream_t stream) {
  EEC_CHECK_ARG(d == 256 && V == 256, "head: d and V must be 256 (got %d, %d)", d, V);
  if (rows == 0) return 0;
  eec_gemm_desc g;
  memset(&g, 0, sizeof(g));
  g.M = rows; g.N = V; g.K = d;
  g.A = x; g.lda = d; g.a_kmajor = 1;
  g.B = w; g.ldb = d; g.b_kmajor = 1;
  g.in_dtype = dtype; g.bias = bias; g.alpha = 1.0f; g.out_dtype = EEC_F32;
  if (dtype == EEC_BF16 && !force_simt_gemm()) {
    g.C = out; g.ldc = V;
    return tc_dispatch(&g, S(stream), argmax, entropy, 1);
  }
  EEC_CHECK_ARG(logits_ws != nullptr, "head: fp32 path needs a logits workspace");
  g.C = logits_ws; g.ldc = V;
  if (int r = gemm_simt(&g, S(stream))) return r;
  return eec_logsoftmax_fwd(logits_ws, out, argmax, entropy, rows, V, stream);
}
