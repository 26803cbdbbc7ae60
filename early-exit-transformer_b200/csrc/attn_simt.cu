// attn_simt.cu -- CUDA-core self-attention forward/backward in fp32 arithmetic.
// Forward: fp32 parity path (and the checker the tcgen05 kernel is validated against on-GPU).
// Backward: used by both precisions in round 1 (operands may be bf16, math is fp32).
// Reference: nn.MultiheadAttention SDPA branch called from TA:194-200; mask TA:9-15.
#include "common.cuh"

namespace eec {

constexpr int QB = 32;    // queries (or keys, in the dK/dV pass) per block
constexpr int KC = 128;   // keys (or queries) staged per chunk
constexpr int DH = 32;

// grid (ceil(T/QB), H, B), 128 threads: warp w owns rows w*8 .. w*8+7 of the block
template <typename T, bool ACC>
__global__ void __launch_bounds__(128) attn_fwd_simt_kernel(const T* __restrict__ qkv, const int32_t* __restrict__ key_len,
                                                           T* __restrict__ ctx, float* __restrict__ lse, int Tn, int H,
                                                           const DropArgs drop, const ActiveItems act_items) {
  if (act_items.n_dev && (int)blockIdx.z >= active_count(act_items)) return;
  __shared__ float Ks[KC][DH + 1];
  __shared__ float Vs[KC][DH];
  DropKey dkey{};
  if (drop.state) dkey = drop_key(drop);
  const uint64_t tk8 = (uint64_t)((Tn + 7) >> 3) * 8;   // dropout element index of (b, h, t, key) = ((b*H + h)*T + t) * tk8 + key
  __shared__ float Qs[QB][DH];
  __shared__ float Ps[4][KC];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int D3 = 3 * H * DH, D = H * DH;
  const float scale = rsqrtf((float)DH);
  const int klen = min(key_len[b], Tn);
  const T* base = qkv + (long)b * Tn * D3;
  for (int i = tid; i < QB * DH; i += 128) {
    int r = i / DH, d = i % DH;
    Qs[r][d] = (q0 + r < Tn) ? ld_as_float<T>(base + (long)(q0 + r) * D3 + h * DH + d) * scale : 0.f;
  }
  float m[8], l[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; l[i] = 0.f; o[i] = 0.f; }

  for (int k0 = 0; k0 < klen; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < KC * DH; i += 128) {
      int r = i / DH, d = i % DH;
      bool ok = (k0 + r < klen);
      Ks[r][d] = ok ? ld_as_float<T>(base + (long)(k0 + r) * D3 + D + h * DH + d) : 0.f;
      Vs[r][d] = ok ? ld_as_float<T>(base + (long)(k0 + r) * D3 + 2 * D + h * DH + d) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < 8; ++qi) {
      const int qr = w * 8 + qi;
      float s[4];
      float cmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int key = lane + 32 * jj;
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) a = fmaf(Qs[qr][d], Ks[key][d], a);
        s[jj] = (k0 + key < klen) ? a : -INFINITY;
        cmax = fmaxf(cmax, s[jj]);
      }
      cmax = warp_max(cmax);
      const float mn = fmaxf(m[qi], cmax);
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float p = (s[jj] == -INFINITY) ? 0.f : (ACC ? expf(s[jj] - mn) : __expf(s[jj] - mn));
        psum += p;   // (the softmax denominator is taken before dropout)
        if (drop.state) p *= drop_factor1(dkey, drop, ((uint64_t)(b * H + h) * Tn + (q0 + qr)) * tk8 + (k0 + lane + 32 * jj));
        Ps[w][lane + 32 * jj] = p;
      }
      psum = warp_sum(psum);
      const float corr = (m[qi] == -INFINITY) ? 0.f : (ACC ? expf(m[qi] - mn) : __expf(m[qi] - mn));
      l[qi] = l[qi] * corr + psum;
      __syncwarp();
      float acc = 0.f;
#pragma unroll 8
      for (int key = 0; key < KC; ++key) acc = fmaf(Ps[w][key], Vs[key][lane], acc);
      o[qi] = o[qi] * corr + acc;
      m[qi] = mn;
      __syncwarp();
    }
  }
#pragma unroll
  for (int qi = 0; qi < 8; ++qi) {
    const int t = q0 + w * 8 + qi;
    if (t >= Tn) continue;
    const float outv = (l[qi] > 0.f) ? o[qi] / l[qi] : 0.f;
    st_from_float<T>(ctx + ((long)b * Tn + t) * D + h * DH + lane, outv);
    if (lse && lane == 0) lse[((long)b * H + h) * Tn + t] = (l[qi] > 0.f) ? m[qi] + logf(l[qi]) : -INFINITY;
  }
}

// Backward pass 1: D_i = dO_i . O_i ; dQ_i = scale * sum_j dS_ij K_j
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ ctx,
                                                         const T* __restrict__ dctx, const float* __restrict__ lse,
                                                         const int32_t* __restrict__ key_len, T* __restrict__ dqkv,
                                                         float* __restrict__ dvec, int Tn, int H, const DropArgs drop) {
  DropKey dkey{};
  if (drop.state) dkey = drop_key(drop);
  const uint64_t tk8 = (uint64_t)((Tn + 7) >> 3) * 8;
  __shared__ float Ks[KC][DH + 1];
  __shared__ float Vs[KC][DH + 1];
  __shared__ float Qs[QB][DH];
  __shared__ float dOs[QB][DH];
  __shared__ float Ps[4][KC];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int D3 = 3 * H * DH, D = H * DH;
  const float scale = rsqrtf((float)DH);
  const int klen = min(key_len[b], Tn);
  const T* base = qkv + (long)b * Tn * D3;
  for (int i = tid; i < QB * DH; i += 128) {
    int r = i / DH, d = i % DH;
    bool ok = q0 + r < Tn;
    Qs[r][d] = ok ? ld_as_float<T>(base + (long)(q0 + r) * D3 + h * DH + d) * scale : 0.f;
    dOs[r][d] = ok ? ld_as_float<T>(dctx + ((long)b * Tn + q0 + r) * D + h * DH + d) : 0.f;
  }
  __syncthreads();
  float Di[8], Li[8], dq[8];
#pragma unroll
  for (int qi = 0; qi < 8; ++qi) {
    const int t = q0 + w * 8 + qi;
    float ov = (t < Tn) ? ld_as_float<T>(ctx + ((long)b * Tn + t) * D + h * DH + lane) : 0.f;
    Di[qi] = warp_sum(ov * dOs[w * 8 + qi][lane]);
    Li[qi] = (t < Tn) ? lse[((long)b * H + h) * Tn + t] : 0.f;
    dq[qi] = 0.f;
    if (t < Tn && lane == 0) dvec[((long)b * H + h) * Tn + t] = Di[qi];
  }
  for (int k0 = 0; k0 < klen; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < KC * DH; i += 128) {
      int r = i / DH, d = i % DH;
      bool ok = (k0 + r < klen);
      Ks[r][d] = ok ? ld_as_float<T>(base + (long)(k0 + r) * D3 + D + h * DH + d) : 0.f;
      Vs[r][d] = ok ? ld_as_float<T>(base + (long)(k0 + r) * D3 + 2 * D + h * DH + d) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < 8; ++qi) {
      const int qr = w * 8 + qi;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int key = lane + 32 * jj;
        float a = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) {
          a = fmaf(Qs[qr][d], Ks[key][d], a);
          dp = fmaf(dOs[qr][d], Vs[key][d], dp);
        }
        float p = (k0 + key < klen && Li[qi] != -INFINITY) ? expf(a - Li[qi]) : 0.f;
        if (drop.state) dp *= drop_factor1(dkey, drop, ((uint64_t)(b * H + h) * Tn + (q0 + qr)) * tk8 + (k0 + key));
        Ps[w][key] = p * (dp - Di[qi]);
      }
      __syncwarp();
      float acc = 0.f;
#pragma unroll 8
      for (int key = 0; key < KC; ++key) acc = fmaf(Ps[w][key], Ks[key][lane], acc);
      dq[qi] += acc;
      __syncwarp();
    }
  }
#pragma unroll
  for (int qi = 0; qi < 8; ++qi) {
    const int t = q0 + w * 8 + qi;
    if (t < Tn) st_from_float<T>(dqkv + ((long)b * Tn + t) * D3 + h * DH + lane, dq[qi] * scale);
  }
}

// Backward pass 2: block owns QB keys; dV_j = sum_i P_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dctx,
                                                          const float* __restrict__ lse, const float* __restrict__ dvec,
                                                          const int32_t* __restrict__ key_len, T* __restrict__ dqkv,
                                                          int Tn, int H, const DropArgs drop) {
  DropKey dkey{};
  if (drop.state) dkey = drop_key(drop);
  const uint64_t tk8 = (uint64_t)((Tn + 7) >> 3) * 8;
  __shared__ float Qs[KC][DH + 1];
  __shared__ float dOs[KC][DH + 1];
  __shared__ float Ksm[QB][DH];
  __shared__ float Vsm[QB][DH];
  __shared__ float Ps[4][KC];
  __shared__ float dSs[4][KC];
  __shared__ float Ls[KC], Ds[KC];
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * QB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int D3 = 3 * H * DH, D = H * DH;
  const float scale = rsqrtf((float)DH);
  const int klen = min(key_len[b], Tn);
  const T* base = qkv + (long)b * Tn * D3;
  for (int i = tid; i < QB * DH; i += 128) {
    int r = i / DH, d = i % DH;
    bool ok = j0 + r < Tn;
    Ksm[r][d] = ok ? ld_as_float<T>(base + (long)(j0 + r) * D3 + D + h * DH + d) : 0.f;
    Vsm[r][d] = ok ? ld_as_float<T>(base + (long)(j0 + r) * D3 + 2 * D + h * DH + d) : 0.f;
  }
  float dk[8], dv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dk[i] = 0.f; dv[i] = 0.f; }
  const bool any_valid = j0 < klen;
  if (any_valid) {
    for (int i0 = 0; i0 < Tn; i0 += KC) {
      __syncthreads();
      for (int i = tid; i < KC * DH; i += 128) {
        int r = i / DH, d = i % DH;
        bool ok = (i0 + r < Tn);
        Qs[r][d] = ok ? ld_as_float<T>(base + (long)(i0 + r) * D3 + h * DH + d) * scale : 0.f;
        dOs[r][d] = ok ? ld_as_float<T>(dctx + ((long)b * Tn + i0 + r) * D + h * DH + d) : 0.f;
      }
      for (int i = tid; i < KC; i += 128) {
        bool ok = (i0 + i < Tn);
        Ls[i] = ok ? lse[((long)b * H + h) * Tn + i0 + i] : -INFINITY;
        Ds[i] = ok ? dvec[((long)b * H + h) * Tn + i0 + i] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int ki = 0; ki < 8; ++ki) {
        const int kr = w * 8 + ki;
        const bool kvalid = (j0 + kr < klen);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int q = lane + 32 * jj;
          float a = 0.f, dp = 0.f;
#pragma unroll
          for (int d = 0; d < DH; ++d) {
            a = fmaf(Qs[q][d], Ksm[kr][d], a);
            dp = fmaf(dOs[q][d], Vsm[kr][d], dp);
          }
          float p = (kvalid && Ls[q] != -INFINITY) ? expf(a - Ls[q]) : 0.f;
          float f = 1.f;
          if (drop.state) f = drop_factor1(dkey, drop, ((uint64_t)(b * H + h) * Tn + (i0 + q)) * tk8 + (j0 + kr));
          Ps[w][q] = p * f;                      // dV = (P.M)^T dO
          dSs[w][q] = p * (dp * f - Ds[q]);      // dS = P (dP.M - D)
        }
        __syncwarp();
        float av = 0.f, ak = 0.f;
#pragma unroll 8
        for (int q = 0; q < KC; ++q) {
          av = fmaf(Ps[w][q], dOs[q][lane], av);
          ak = fmaf(dSs[w][q], Qs[q][lane], ak);
        }
        dv[ki] += av;
        dk[ki] += ak;  // Qs already carries `scale`
        __syncwarp();
      }
    }
  }
#pragma unroll
  for (int ki = 0; ki < 8; ++ki) {
    const int t = j0 + w * 8 + ki;
    if (t < Tn) {
      st_from_float<T>(dqkv + ((long)b * Tn + t) * D3 + D + h * DH + lane, dk[ki]);
      st_from_float<T>(dqkv + ((long)b * Tn + t) * D3 + 2 * D + h * DH + lane, dv[ki]);
    }
  }
}

}  // namespace eec

using namespace eec;

// EEC_ATTN_TC=0 routes bf16 attention to the CUDA-core kernels (debug cross-check); default = tcgen05
static bool attn_tc_ready() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EEC_ATTN_TC"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
static bool force_simt() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EEC_FORCE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

extern "C" int eec_attn_fwd(const void* qkv, int dtype, const int32_t* key_len, void* ctx, float* lse, int B, int T,
                            int H, int dh, const uint64_t* drop_state, float drop_p, uint32_t drop_site, const void* drop_bits,
                            eec_stream_t stream) {
  EEC_CHECK_ARG(dh == 32, "attn_fwd: head dim must be 32 (got %d)", dh);
  if (B == 0 || T == 0) return 0;
  DropArgs drop = make_drop(drop_state, drop_p, drop_site);
  drop.bits = drop_bits;
  const bool tc_path = dtype == EEC_BF16 && !force_simt() && attn_tc_ready();
  EEC_CHECK_ARG(!(tc_path && drop.state && !drop_bits), "attn_fwd (tensor-core path): dropout needs the keep-mask words of "
                "eec_dropout_bits(R = B*H*T, C = T, Cs = 8*ceil(T/8), W = 32) in drop_bits");
  if (tc_path) return attn_fwd_tc(qkv, key_len, ctx, lse, B, T, H, dh, drop, S(stream));
  dim3 grid(cdiv(T, QB), H, B);
  if (dtype == EEC_F32)
    attn_fwd_simt_kernel<float, true><<<grid, 128, 0, S(stream)>>>((const float*)qkv, key_len, (float*)ctx, lse, T, H, drop, active_items(S(stream)));
  else
    attn_fwd_simt_kernel<__nv_bfloat16, false><<<grid, 128, 0, S(stream)>>>((const __nv_bfloat16*)qkv, key_len, (__nv_bfloat16*)ctx, lse, T, H, drop, active_items(S(stream)));
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_attn_bwd(const void* qkv, const void* ctx, const void* dctx, int dtype, const float* lse,
                            const int32_t* key_len, void* dqkv, float* dvec, float* dq32, int B, int T, int H, int dh,
                            const uint64_t* drop_state, float drop_p, uint32_t drop_site, const void* drop_bits, eec_stream_t stream) {
  EEC_CHECK_ARG(dh == 32, "attn_bwd: head dim must be 32 (got %d)", dh);
  if (B == 0 || T == 0) return 0;
  DropArgs drop = make_drop(drop_state, drop_p, drop_site);
  drop.bits = drop_bits;
  const bool tc_path = dtype == EEC_BF16 && !force_simt() && attn_tc_ready();
  EEC_CHECK_ARG(!(tc_path && drop.state && !drop_bits), "attn_bwd (tensor-core path): dropout needs the forward's keep-mask words in drop_bits");
  if (tc_path)
    return attn_bwd_tc(qkv, ctx, dctx, lse, key_len, dqkv, dvec, dq32, B, T, H, dh, drop, S(stream));
  dim3 grid(cdiv(T, QB), H, B);
  if (dtype == EEC_F32) {
    attn_bwd_dq_kernel<float><<<grid, 128, 0, S(stream)>>>((const float*)qkv, (const float*)ctx, (const float*)dctx, lse, key_len, (float*)dqkv, dvec, T, H, drop);
    EEC_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<float><<<grid, 128, 0, S(stream)>>>((const float*)qkv, (const float*)dctx, lse, dvec, key_len, (float*)dqkv, T, H, drop);
  } else {
    attn_bwd_dq_kernel<__nv_bfloat16><<<grid, 128, 0, S(stream)>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)ctx, (const __nv_bfloat16*)dctx, lse, key_len, (__nv_bfloat16*)dqkv, dvec, T, H, drop);
    EEC_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<__nv_bfloat16><<<grid, 128, 0, S(stream)>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dctx, lse, dvec, key_len, (__nv_bfloat16*)dqkv, T, H, drop);
  }
  EEC_LAUNCH_CHECK();
  return 0;
}
