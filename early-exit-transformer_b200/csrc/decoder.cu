// decoder.cu -- glue kernels of the AED decoder stacks (reference early_exit.py:701-717, :742-800, train.py:36-51):
// token embedding + positional encoding (and its scatter-add backward), the padding-mask bit words of a token matrix, and
// nn.CrossEntropyLoss (mean over all rows, no ignore_index) with its gradient.  The decoder's contractions run on the GEMM and
// attention kernels shared with the encoder (eec_gemm with RELU / DRELU epilogues, eec_attn_general_fwd / _bwd).
#include "common.cuh"

namespace eec {
namespace {

// one warp per token row: x[row, :] = emb[tok, :] + pe[t, :]     (D % 128 == 0: float4 per lane)
__global__ void __launch_bounds__(256) embed_pe_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ emb,
                                                       const float* __restrict__ pe, float* __restrict__ x, int rows, int L, int D, int V) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int t = row % L;
  long tok = tokens[row];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  for (int c = lane * 4; c < D; c += 128) {
    const float4 e = *reinterpret_cast<const float4*>(emb + tok * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pe + (long)t * D + c);
    *reinterpret_cast<float4*>(x + (long)row * D + c) = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
}

__global__ void __launch_bounds__(256) embed_bwd_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ dx,
                                                        float* __restrict__ demb, int rows, int D, int V) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  long tok = tokens[row];
  tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  for (int c = lane; c < D; c += 32) atomicAdd(demb + tok * D + c, dx[(long)row * D + c]);
}

__global__ void key_bits_kernel(const int64_t* __restrict__ tokens, int B, int L, int64_t pad, uint32_t* __restrict__ bits) {
  pdl_trigger();
  pdl_wait();
  const int W = (L + 31) >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (b, word)
  if (i >= B * W) return;
  const int b = i / W, w = i % W;
  uint32_t v = 0;
  for (int k = 0; k < 32; ++k) {
    const int t = w * 32 + k;
    if (t < L && tokens[(long)b * L + t] != pad) v |= 1u << k;
  }
  bits[i] = v;
}

// one warp per row (V % 32 == 0, V <= 1024 handled as a loop): loss += (lse - logit[target]) / rows; dlogits = (softmax - onehot) / rows
__global__ void __launch_bounds__(256) cross_entropy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets, int rows,
                                                            int V, float* __restrict__ loss_out, float* __restrict__ dlogits) {
  pdl_trigger();
  pdl_wait();
  __shared__ float part[8];
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float contrib = 0.f;
  if (row < rows) {
    const float* x = logits + (long)row * V;
    float mx = -INFINITY;
    for (int c = lane; c < V; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(x[c] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    long tg = targets[row];
    tg = tg < 0 ? 0 : (tg >= V ? V - 1 : tg);
    const float inv = 1.0f / (float)rows;
    contrib = (lse - x[tg]) * inv;
    if (dlogits) {
      float* g = dlogits + (long)row * V;
      for (int c = lane; c < V; c += 32) g[c] = (expf(x[c] - lse) - (c == tg ? 1.f : 0.f)) * inv;
    }
  }
  if (lane == 0) part[warp] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    atomicAdd(loss_out, t);
  }
}

}  // namespace
}  // namespace eec

using namespace eec;

extern "C" int eec_embed_pe(const int64_t* tokens, const float* emb, const float* pe, float* x, int B, int L, int D, int V,
                            eec_stream_t stream) {
  EEC_CHECK_ARG(D % 128 == 0, "embed_pe: D %% 128 (got %d)", D);
  const int rows = B * L;
  if (rows == 0) return 0;
  launch_pdl(embed_pe_kernel, dim3(cdiv(rows * 32, 256)), dim3(256), 0, S(stream), tokens, emb, pe, x, rows, L, D, V);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_embed_bwd(const int64_t* tokens, const float* dx, float* demb, int B, int L, int D, int V, eec_stream_t stream) {
  const int rows = B * L;
  if (rows == 0) return 0;
  launch_pdl(embed_bwd_kernel, dim3(cdiv(rows * 32, 256)), dim3(256), 0, S(stream), tokens, dx, demb, rows, D, V);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_key_bits_from_tokens(const int64_t* tokens, int B, int L, int64_t pad, uint32_t* bits, eec_stream_t stream) {
  const int n = B * cdiv(L, 32);
  if (n == 0) return 0;
  launch_pdl(key_bits_kernel, dim3(cdiv(n, 128)), dim3(128), 0, S(stream), tokens, B, L, pad, bits);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_cross_entropy(const float* logits, const int64_t* targets, int rows, int V, float* loss_out, float* dlogits,
                                 eec_stream_t stream) {
  EEC_CHECK_ARG(V >= 1, "cross_entropy: V");
  if (rows == 0) return 0;
  launch_pdl(cross_entropy_kernel, dim3(cdiv(rows * 32, 256)), dim3(256), 0, S(stream), logits, targets, rows, V, loss_out, dlogits);
  EEC_LAUNCH_CHECK();
  return 0;
}
