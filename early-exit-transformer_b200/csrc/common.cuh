// common.cuh -- shared helpers for libeec.so (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "eec.h"

namespace eec {

void set_error(const char* fmt, ...);

#define EEC_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      eec::set_error(__VA_ARGS__);          \
      return 1;                             \
    }                                       \
  } while (0)

#define EEC_CUDA(call)                                                                    \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      eec::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

void count_launch();
#define EEC_LAUNCH_CHECK()         \
  do {                             \
    eec::count_launch();           \
    EEC_CUDA(cudaGetLastError());  \
  } while (0)

static inline cudaStream_t S(eec_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// A training step is a chain of ~600 dependent kernels, most of them 10-50 us long; launch latency, block scheduling and
// the per-kernel prologue (barrier init, TMEM allocation, parameter staging) of kernel k+1 overlap the tail of kernel k
// when k+1 is launched with programmatic stream serialisation and both follow this protocol:
//   pdl_trigger()  first statement of every kernel: "my dependents may start launching" (they still WAIT below)
//   pdl_wait()     before the first access to global memory another kernel may have written (or may still read):
//                  returns when the preceding kernel has completed and its writes are visible
// Every kernel launched through launch_pdl() MUST call pdl_wait(); both are no-ops under a plain launch.
// Measured on B200 inside the CUDA graphs (bench.py, same box, 20 steps): training step 15.52 ms with vs 15.34-15.57 ms without,
// inference forward 4.63 ms with vs 4.43 ms without -- graph replay already removes the launch gaps and the early-resident
// blocks of kernel k+1 compete with kernel k's tail.  So the attribute is OFF by default; EEC_PDL=1 turns it on.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers ---------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float dsiluf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.0f + x * (1.0f - s));
}
// accurate variants for the fp32 parity path
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// 8-wide vector load/store of T as floats (T = float: 2x float4; T = bf16: one uint4)
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---- dropout: counter-based keep masks (Philox4x32-10; include/eec.h "dropout") -----------------------------------
// One Philox call yields 128 random bits = the 16-bit draws of 8 consecutive elements (a "group"); every kernel that touches
// a dropout site (forward epilogue, backward re-generation, the standalone eec_dropout used by the tests) derives the same
// bits from (seed, offset, site, element index / 8) and nothing else, so no mask is ever stored.
struct DropArgs {          // built on the host (make_drop), passed to kernels by value
  const uint64_t* state;   // device {seed, offset}; NULL = dropout off
  uint32_t site;
  uint32_t thr;            // element dropped iff its 16-bit draw < thr   (thr = round(p * 65536))
  float scale;             // 65536 / (65536 - thr)
  const void* bits;        // tensor-core kernels: keep-mask words written by eec_dropout_bits (they do not run Philox themselves)
};
static inline DropArgs make_drop(const uint64_t* state, float p, uint32_t site) {
  DropArgs a{};
  if (!state || !(p > 0.f)) return a;
  double t = (double)p * 65536.0 + 0.5;
  uint32_t thr = t >= 65535.0 ? 65535u : (uint32_t)t;
  a.state = state; a.site = site; a.thr = thr; a.scale = (float)(65536.0 / (65536.0 - (double)thr));
  return a;
}
struct DropKey { uint32_t k0, k1, c2, c3; };
__device__ __forceinline__ DropKey drop_key(const DropArgs& a) {
  const uint64_t seed = a.state[0], off = a.state[1];
  DropKey k;
  k.k0 = (uint32_t)seed; k.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(off >> 32);
  k.c2 = a.site; k.c3 = (uint32_t)off;
  return k;
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// f[i] = scale (kept) or 0 (dropped) for elements 8*group .. 8*group+7
__device__ __forceinline__ void drop_factors8(const DropKey& k, const DropArgs& a, uint64_t group, float (&f)[8]) {
  const uint4 r = philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), k.c2, k.c3, k.k0, k.k1);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = ((w[i] & 0xffffu) >= a.thr) ? a.scale : 0.f;
    f[2 * i + 1] = ((w[i] >> 16) >= a.thr) ? a.scale : 0.f;
  }
}
// single element (the CUDA-core parity kernels touch one element per thread)
__device__ __forceinline__ float drop_factor1(const DropKey& k, const DropArgs& a, uint64_t elem) {
  const uint64_t group = elem >> 3;
  const uint4 r = philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), k.c2, k.c3, k.k0, k.k1);
  const uint32_t i = (uint32_t)elem & 7u;
  const uint32_t w = (i >> 1) == 0 ? r.x : (i >> 1) == 1 ? r.y : (i >> 1) == 2 ? r.z : r.w;
  const uint32_t u = (i & 1u) ? (w >> 16) : (w & 0xffffu);
  return (u >= a.thr) ? a.scale : 0.f;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- active-item limit (early-exit inference with batch compaction; include/eec.h eec_set_active_items) ----------------
// When set, the inference kernels process only the first *n_dev utterances (= *n_dev * rows_per_item leading rows of every
// frame-major tensor): tiles / rows / utterance blocks past that limit return immediately.  The count lives on the DEVICE
// (eec_exit_select updates it), so no host synchronisation and no re-capture of a CUDA graph is needed when it changes.
// pad_items extra utterances past the count are still processed: the tcgen05 attention loads 128-row K/V tiles that overhang into the
// next utterance's rows (masked in the softmax, but 0 x NaN = NaN in P.V), so the rows right behind the last survivor must hold
// finite numbers -- one padding utterance (T' >= 128 rows) of stale-but-finite data guarantees that.
struct ActiveItems { const int32_t* n_dev; int rows_per_item; int pad_items; };
ActiveItems active_items(cudaStream_t stream);   // the limit set for THIS stream (eec_set_active_items), or {nullptr, 0, 0}
__device__ __forceinline__ int active_count(const ActiveItems& a) { return *a.n_dev + a.pad_items; }
__device__ __forceinline__ int active_rows(const ActiveItems& a, int rows) {
  return a.n_dev ? min(rows, active_count(a) * a.rows_per_item) : rows;
}

// backends implemented in separate translation units
int gemm_simt(const eec_gemm_desc* d, cudaStream_t st);
int gemm_tc(const eec_gemm_desc* d, cudaStream_t st, int32_t* argmax, float* entropy, int logsoftmax);   // v1 (EEC_GEMM_V1=1)
int gemm_tc2(const eec_gemm_desc* d, cudaStream_t st, int32_t* argmax, float* entropy, int logsoftmax);  // persistent v2
int gemm_tc3(const eec_gemm_desc* d, cudaStream_t st);   // v3 epilogue (GENERIC / GLU modes), called by gemm_tc2
int gemm_ln3(const eec_gemm_desc* d, cudaStream_t st);   // v3 LayerNorm-tail epilogue (N == 256), called by gemm_tc2
bool gemm_ws_ok(const eec_gemm_desc* d);                 // weight-stationary K = 256 kernel (gemm_ws.cu): eligible?
int gemm_ws(const eec_gemm_desc* d, cudaStream_t st);
int gemm_ws2(const eec_gemm_desc* d, cudaStream_t st);   // CTA-pair (cta_group::2) version, gemm_ws2.cu
bool gemm_pair_ok(const eec_gemm_desc* d, cudaStream_t st);   // fp32-output streaming GEMMs (weight / long-K data gradients), CTA pairs: gemm_pair.cu
int gemm_pair(const eec_gemm_desc* d, cudaStream_t st);
bool gemm_lnp_ok(const eec_gemm_desc* d, cudaStream_t st);    // LayerNorm-tail GEMM, CTA pairs: gemm_lnp.cu
int gemm_lnp(const eec_gemm_desc* d, cudaStream_t st);
// x[j] = bit j of `word` ? x[j] * scale : 0   (j < n <= 32; the epilogue-side half of the dropout)
template <int N>
__device__ __forceinline__ void drop_apply_bits(float (&x)[N], uint32_t word, float scale) {
#pragma unroll
  for (int j = 0; j < N; ++j) x[j] = (word & (1u << j)) ? x[j] * scale : 0.f;
}

// general (decoder) attention on the tensor cores; attn_general_tc_ok: can this geometry run there (else the CUDA-core kernels do it)
bool attn_general_tc_ok(const eec_attn_desc* d);
int attn_general_fwd_tc(const eec_attn_desc* d, void* ctx, int ldo, float* lse, const DropArgs& drop, cudaStream_t st);
int attn_general_bwd_tc(const eec_attn_desc* d, const void* ctx, const void* dctx, int ldo, const float* lse, void* dq, int lddq, void* dk,
                        int lddk, void* dv, int lddv, float* dvec, float* dq32, const DropArgs& drop, cudaStream_t st);
int attn_bwd_tc(const void* qkv, const void* ctx, const void* dctx, const float* lse, const int32_t* key_len, void* dqkv,
                float* dvec, float* dq32, int B, int T, int H, int dh, const DropArgs& drop, cudaStream_t st);
int attn_fwd_tc(const void* qkv, const int32_t* key_len, void* ctx, float* lse, int B, int T, int H, int dh,
                const DropArgs& drop, cudaStream_t st);

}  // namespace eec
