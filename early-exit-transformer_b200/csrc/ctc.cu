// ctc.cu -- multi-exit CTC forward+backward in one kernel (train.py:57-65 ->
// nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True)), greedy decode
// (util/beam_infer.py:21-23) and the on-device early-exit selection (north star; no reference).
//
// CTC: one CTA per (exit, utterance); thread s owns extended-target state s (2U+1 states).
// alpha rows are kept in a shared-memory ping-pong and spilled to a workspace; the beta sweep
// fuses the occupancy reduction and writes d(loss)/d(logits) = scale*(softmax - occupancy)
// directly, so log-softmax backward never runs as a separate kernel.
#include "common.cuh"

namespace eec {

__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// grid (B, E).  lp/grad: [E][B][T][V]; nll: [E][B]; loss_out: [E]
template <int NT>
__global__ void __launch_bounds__(NT) ctc_kernel(const float* __restrict__ lp_all, const int64_t* __restrict__ targets,
                                                 const int64_t* __restrict__ target_len, int B, int T, int V, int Lmax,
                                                 int blank, float gscale, float* __restrict__ nll_all,
                                                 float* __restrict__ loss_out, float* __restrict__ grad_all,
                                                 float* __restrict__ alpha_ws, int Smax) {
  extern __shared__ float sm[];
  float* buf0 = sm;              // [Smax + 2] with 2 leading -inf pads
  float* buf1 = sm + (Smax + 2);
  float* occ = sm + 2 * (Smax + 2);            // [V]
  int* ext = reinterpret_cast<int*>(occ + V);  // [Smax]
  __shared__ float s_ll;
  const int b = blockIdx.x, e = blockIdx.y, tid = threadIdx.x;
  const long eb = (long)e * B + b;
  const float* lp = lp_all + eb * T * V;
  float* grad = grad_all ? grad_all + eb * T * V : nullptr;
  float* aw = alpha_ws + eb * (long)T * Smax;
  int U = (int)target_len[b];
  if (U > Lmax) U = Lmax;
  const int S = 2 * U + 1;
  for (int s = tid; s < Smax; s += NT) ext[s] = (s < S && (s & 1)) ? (int)targets[(long)b * Lmax + (s >> 1)] : blank;
  if (tid < 2) { buf0[tid] = -INFINITY; buf1[tid] = -INFINITY; }
  __syncthreads();
  const int s = tid;
  const bool active = s < S;
  const int my = active ? ext[s] : blank;
  const bool skip = active && s >= 2 && my != blank && my != ext[s - 2];
  float* prev = buf0 + 2;
  float* cur = buf1 + 2;
  // ---- alpha
  if (s < Smax) {
    float a = -INFINITY;
    if (active && s < 2) a = lp[my];
    prev[s] = a;
    aw[s] = a;
  }
  __syncthreads();
  for (int t = 1; t < T; ++t) {
    if (s < Smax) {
      float a = -INFINITY;
      if (active) a = lse3(prev[s], prev[s - 1], skip ? prev[s - 2] : -INFINITY) + lp[(long)t * V + my];
      cur[s] = a;
      aw[(long)t * Smax + s] = a;
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  if (tid == 0) s_ll = (S >= 2) ? lse2(prev[S - 1], prev[S - 2]) : prev[S - 1];
  __syncthreads();
  const float ll = s_ll;
  const bool feasible = (ll != -INFINITY);
  const float denom = (float)B * (float)max(U, 1);
  if (tid == 0) {
    nll_all[eb] = feasible ? -ll : 0.f;
    if (feasible && loss_out) atomicAdd(loss_out + e, -ll / denom);
  }
  if (!grad) return;
  if (!feasible) {  // zero_infinity: zero gradient rows
    for (long i = tid; i < (long)T * V; i += NT) grad[i] = 0.f;
    return;
  }
  // ---- beta sweep + occupancy + gradient.  Reuse the ping-pong with 2 TRAILING pads.
  __syncthreads();
  float* bprev = buf0;  // [Smax] + 2 trailing
  float* bcur = buf1;
  if (tid < 2) { buf0[Smax + tid] = -INFINITY; buf1[Smax + tid] = -INFINITY; }
  const bool skip_fwd = active && (s + 2 < S) && ext[s + 2] != blank && ext[s + 2] != my;
  const float sc = gscale / denom;
  for (int t = T - 1; t >= 0; --t) {
    float bv = -INFINITY;
    if (active) {
      const float l = lp[(long)t * V + my];
      if (t == T - 1) bv = (s >= S - 2) ? l : -INFINITY;
      else bv = lse3(bprev[s], bprev[s + 1], skip_fwd ? bprev[s + 2] : -INFINITY) + l;
    }
    if (s < Smax) bcur[s] = bv;
    for (int c = tid; c < V; c += NT) occ[c] = 0.f;
    __syncthreads();
    if (active) {
      const float a = aw[(long)t * Smax + s];
      if (a != -INFINITY && bv != -INFINITY) atomicAdd(&occ[my], expf(a + bv - lp[(long)t * V + my] - ll));
    }
    __syncthreads();
    for (int c = tid; c < V; c += NT) grad[(long)t * V + c] = sc * (expf(lp[(long)t * V + c]) - occ[c]);
    float* tmp = bprev; bprev = bcur; bcur = tmp;
    // next iteration's writes to bcur/occ are ordered by the two barriers above
  }
}

// greedy collapse: one warp per utterance
__global__ void greedy_collapse_kernel(const int32_t* __restrict__ argmax, int32_t* __restrict__ tokens,
                                       int32_t* __restrict__ n_tokens, int B, int T, int blank) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int32_t* a = argmax + (long)b * T;
  int32_t* out = tokens + (long)b * T;
  int count = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    int v = (t < T) ? a[t] : blank;
    int pv = (t > 0 && t < T) ? a[t - 1] : -1;
    bool keep = (t < T) && (v != pv) && (v != blank);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) out[count + __popc(m & ((1u << lane) - 1u))] = v;
    count += __popc(m);
  }
  for (int t = count + lane; t < T; t += 32) out[t] = -1;
  if (lane == 0) n_tokens[b] = count;
}

// early-exit selection + stable compaction bookkeeping; ONE block of 1024 threads.
__global__ void __launch_bounds__(1024) exit_select_kernel(const float* __restrict__ entropy, const int32_t* __restrict__ argmax,
                                                           const int32_t* __restrict__ key_len_alive,
                                                           const int32_t* __restrict__ row_map, int32_t* __restrict__ n_alive,
                                                           int exit_idx, int is_last, float threshold,
                                                           int32_t* __restrict__ exit_index, int32_t* __restrict__ tokens,
                                                           int32_t* __restrict__ n_tokens, int32_t* __restrict__ new_row_map,
                                                           int32_t* __restrict__ new_key_len, int32_t* __restrict__ gather_idx,
                                                           float* __restrict__ mean_entropy_out, int B, int T, int blank) {
  __shared__ int done[1024];
  __shared__ int pos[1024];
  __shared__ int s_new;
  const int n = *n_alive;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = w; i < n; i += 32) {
    const int kl = min(key_len_alive[i], T);
    float h = 0.f;
    for (int t = lane; t < kl; t += 32) h += entropy[(long)i * T + t];
    h = warp_sum(h) / (float)max(kl, 1);
    const int fin = (is_last || h < threshold) ? 1 : 0;
    if (lane == 0) {
      done[i] = fin;
      if (mean_entropy_out) mean_entropy_out[(long)exit_idx * B + row_map[i]] = h;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0;
    for (int i = 0; i < n; ++i) {
      pos[i] = k;
      if (!done[i]) ++k;
    }
    s_new = k;
  }
  __syncthreads();
  // survivors: compacted bookkeeping (read old maps before anyone overwrites: separate out arrays)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (!done[i]) {
      const int j = pos[i];
      new_row_map[j] = row_map[i];
      new_key_len[j] = key_len_alive[i];
      gather_idx[j] = i;
    }
  }
  // finalised rows: greedy tokens of this exit into the ORIGINAL row
  for (int i = w; i < n; i += 32) {
    if (!done[i]) continue;
    const int orig = row_map[i];
    const int32_t* a = argmax + (long)i * T;
    int32_t* out = tokens + (long)orig * T;
    int count = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      int v = (t < T) ? a[t] : blank;
      int pv = (t > 0 && t < T) ? a[t - 1] : -1;
      bool keep = (t < T) && (v != pv) && (v != blank);
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) out[count + __popc(m & ((1u << lane) - 1u))] = v;
      count += __popc(m);
    }
    for (int t = count + lane; t < T; t += 32) out[t] = -1;
    if (lane == 0) { n_tokens[orig] = count; exit_index[orig] = exit_idx; }
  }
  __syncthreads();
  if (threadIdx.x == 0) *n_alive = s_new;
}

__global__ void gather_rows_kernel(const float4* __restrict__ x, float4* __restrict__ y, const int32_t* __restrict__ gather_idx,
                                   const int32_t* __restrict__ n_alive, long row_vec) {
  const int j = blockIdx.y;
  if (j >= *n_alive) return;
  const long src = (long)gather_idx[j] * row_vec, dst = (long)j * row_vec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < row_vec; i += (long)gridDim.x * blockDim.x) y[dst + i] = x[src + i];
}

// ---- front-end im2col / col2im for Conv1d(k=3, s=2)
template <typename TI, typename TO>
__global__ void im2col_k3s2_kernel(const TI* __restrict__ in, long sb, long sc, long st, TO* __restrict__ out, int ldo, int B,
                                   int C, int T_out) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)B * T_out * ldo;
  if (i >= total) return;
  int k = (int)(i % ldo);
  long r = i / ldo;
  int t = (int)(r % T_out), b = (int)(r / T_out);
  float v = 0.f;
  if (k < 3 * C) {
    int c = k / 3, j = k % 3;
    v = ld_as_float<TI>(in + (long)b * sb + (long)c * sc + (long)(2 * t + j) * st);
  }
  st_from_float<TO>(out + i, v);
}

__global__ void col2im_k3s2_kernel(const float* __restrict__ dcols, int ldc, float* __restrict__ dx, int B, int C, int T_in,
                                   int T_out) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)B * T_in * C;
  if (i >= total) return;
  int c = (int)(i % C);
  long r = i / C;
  int f = (int)(r % T_in), b = (int)(r / T_in);
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    int ft = f - j;
    if (ft >= 0 && (ft & 1) == 0 && (ft >> 1) < T_out) v += dcols[((long)b * T_out + (ft >> 1)) * ldc + c * 3 + j];
  }
  dx[i] = v;
}

}  // namespace eec

using namespace eec;

static int ctc_smax(int Lmax) { return ((2 * Lmax + 1 + 31) / 32) * 32; }

extern "C" int64_t eec_ctc_workspace_bytes(int E, int B, int T, int Lmax) {
  return (int64_t)E * B * T * ctc_smax(Lmax) * (int64_t)sizeof(float);
}

extern "C" int eec_ctc_fwd_bwd(const float* lp, const int64_t* targets, const int64_t* target_len, int E, int B, int T,
                               int V, int Lmax, int blank, float gscale, float* nll, float* loss_out, float* grad,
                               void* workspace, eec_stream_t stream) {
  if (E == 0 || B == 0) return 0;
  EEC_CHECK_ARG(T >= 1, "ctc: T must be >= 1");
  const int Smax = ctc_smax(Lmax);
  EEC_CHECK_ARG(Smax <= 1024, "ctc: target length %d too long (2L+1 must be <= 1024)", Lmax);
  EEC_CHECK_ARG(workspace != nullptr, "ctc: workspace is NULL");
  size_t smem = (size_t)(2 * (Smax + 2) + V) * sizeof(float) + (size_t)Smax * sizeof(int);
  dim3 grid(B, E);
  float* ws = reinterpret_cast<float*>(workspace);
  if (Smax <= 256)
    ctc_kernel<256><<<grid, 256, smem, S(stream)>>>(lp, targets, target_len, B, T, V, Lmax, blank, gscale, nll, loss_out, grad, ws, Smax);
  else if (Smax <= 512)
    ctc_kernel<512><<<grid, 512, smem, S(stream)>>>(lp, targets, target_len, B, T, V, Lmax, blank, gscale, nll, loss_out, grad, ws, Smax);
  else
    ctc_kernel<1024><<<grid, 1024, smem, S(stream)>>>(lp, targets, target_len, B, T, V, Lmax, blank, gscale, nll, loss_out, grad, ws, Smax);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_greedy_collapse(const int32_t* argmax, int32_t* tokens, int32_t* n_tokens, int B, int T, int blank,
                                   eec_stream_t stream) {
  if (B == 0) return 0;
  greedy_collapse_kernel<<<cdiv(B * 32, 128), 128, 0, S(stream)>>>(argmax, tokens, n_tokens, B, T, blank);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_exit_select(const float* entropy, const int32_t* argmax, const int32_t* key_len_alive,
                               const int32_t* row_map, int32_t* n_alive, int exit_idx, int is_last, float threshold,
                               int32_t* exit_index, int32_t* tokens, int32_t* n_tokens, int32_t* new_row_map,
                               int32_t* new_key_len, int32_t* gather_idx, float* mean_entropy_out, int B, int T, int blank,
                               eec_stream_t stream) {
  EEC_CHECK_ARG(B <= 1024, "exit_select: batch must be <= 1024 (got %d)", B);
  if (B == 0) return 0;
  exit_select_kernel<<<1, 1024, 0, S(stream)>>>(entropy, argmax, key_len_alive, row_map, n_alive, exit_idx, is_last, threshold,
                                               exit_index, tokens, n_tokens, new_row_map, new_key_len, gather_idx,
                                               mean_entropy_out, B, T, blank);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_gather_rows(const float* x, float* y, const int32_t* gather_idx, const int32_t* n_alive, int B,
                               int64_t row_elems, eec_stream_t stream) {
  EEC_CHECK_ARG(row_elems % 4 == 0, "gather_rows: row_elems %% 4");
  if (B == 0) return 0;
  long rv = row_elems / 4;
  dim3 grid((unsigned)min((long)32, cdiv64(rv, 256)), B);
  gather_rows_kernel<<<grid, 256, 0, S(stream)>>>((const float4*)x, (float4*)y, gather_idx, n_alive, rv);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_im2col_k3s2(const void* in, int in_dtype, int64_t sb, int64_t sc, int64_t st, void* out, int out_dtype,
                               int ldo, int B, int C, int T_out, eec_stream_t stream) {
  EEC_CHECK_ARG(ldo >= 3 * C, "im2col: ldo < 3*C");
  EEC_CHECK_ARG(in_dtype == EEC_F32, "im2col: input must be fp32");
  long total = (long)B * T_out * ldo;
  if (total == 0) return 0;
  int blocks = (int)cdiv64(total, 256);
  if (out_dtype == EEC_F32)
    im2col_k3s2_kernel<float, float><<<blocks, 256, 0, S(stream)>>>((const float*)in, sb, sc, st, (float*)out, ldo, B, C, T_out);
  else
    im2col_k3s2_kernel<float, __nv_bfloat16><<<blocks, 256, 0, S(stream)>>>((const float*)in, sb, sc, st, (__nv_bfloat16*)out, ldo, B, C, T_out);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_col2im_k3s2(const float* dcols, int ldc, float* dx, int B, int C, int T_in, int T_out,
                               eec_stream_t stream) {
  long total = (long)B * T_in * C;
  if (total == 0) return 0;
  col2im_k3s2_kernel<<<(int)cdiv64(total, 256), 256, 0, S(stream)>>>(dcols, ldc, dx, B, C, T_in, T_out);
  EEC_LAUNCH_CHECK();
  return 0;
}
