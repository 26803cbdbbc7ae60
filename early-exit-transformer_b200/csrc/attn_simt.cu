// attn_simt.cu -- CUDA-core self-attention forward/backward in fp32 arithmetic.
// Forward: fp32 parity path (and the checker the tcgen05 kernel is validated against on-GPU).
// Backward: used by both precisions in round 1 (operands may be bf16, math is fp32).
// Reference: nn.MultiheadAttention SDPA branch called from TA:194-200; mask TA:9-15.
#include "common.cuh"

namespace eec {

constexpr int QB = 32;    // queries (or keys, in the dK/dV pass) per block
constexpr int KC = 128;   // keys (or queries) staged per chunk
constexpr int DH = 32;

// Geometry of one attention call (host-built, passed by value).  Self-attention over a packed [rows, 3D] tensor, causal decoder
// self-attention and encoder-decoder cross-attention (Q rows from one tensor, K / V rows from another, Tq != Tk) are the same kernels:
//   Q(b, t, h, d)  = q[(b*Tq + t)*ldq + h*32 + d]      K(b, t', h, d) = k[(b*Tk + t')*ldk + h*32 + d]      (likewise V, and the outputs)
// a key t' is visible to query t iff  t' < key_len[b] (when given)  and  bit t' of key_bits[b] is set (when given)  and  (!causal or t' <= t).
struct AttnGeom {
  const void* q; const void* k; const void* v;
  int ldq, ldk, ldv;
  int Tq, Tk, H;
  const int32_t* key_len;
  const uint32_t* key_bits;    // [B, ceil(Tk/32)] words
  int causal;
};
__device__ __forceinline__ int geom_klen(const AttnGeom& g, int b) { return g.key_len ? min(g.key_len[b], g.Tk) : g.Tk; }
__device__ __forceinline__ bool geom_visible(const AttnGeom& g, int b, int t, int key, int klen) {
  if (key >= klen) return false;
  if (g.causal && key > t) return false;
  if (g.key_bits && !((g.key_bits[(long)b * ((g.Tk + 31) >> 5) + (key >> 5)] >> (key & 31)) & 1u)) return false;
  return true;
}

// grid (ceil(Tq/QB), H, B), 128 threads: warp w owns rows w*8 .. w*8+7 of the block
template <typename T, bool ACC>
__global__ void __launch_bounds__(128) attn_fwd_simt_kernel(const AttnGeom g, T* __restrict__ ctx, int ldo, float* __restrict__ lse,
                                                           const DropArgs drop, const ActiveItems act_items) {
  if (act_items.n_dev && (int)blockIdx.z >= active_count(act_items)) return;
  __shared__ float Ks[KC][DH + 1];
  __shared__ float Vs[KC][DH];
  DropKey dkey{};
  if (drop.state) dkey = drop_key(drop);
  const int Tq = g.Tq, Tk = g.Tk, H = g.H;
  const uint64_t tk8 = (uint64_t)((Tk + 7) >> 3) * 8;   // dropout element index of (b, h, t, key) = ((b*H + h)*Tq + t) * tk8 + key
  __shared__ float Qs[QB][DH];
  __shared__ float Ps[4][KC];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float scale = rsqrtf((float)DH);
  const int klen = geom_klen(g, b);
  const int kend = g.causal ? min(klen, q0 + QB) : klen;      // keys past the block's last query are invisible to all of its rows
  const T* qb = reinterpret_cast<const T*>(g.q) + (long)b * Tq * g.ldq + h * DH;
  const T* kb = reinterpret_cast<const T*>(g.k) + (long)b * Tk * g.ldk + h * DH;
  const T* vb = reinterpret_cast<const T*>(g.v) + (long)b * Tk * g.ldv + h * DH;
  for (int i = tid; i < QB * DH; i += 128) {
    int r = i / DH, d = i % DH;
    Qs[r][d] = (q0 + r < Tq) ? ld_as_float<T>(qb + (long)(q0 + r) * g.ldq + d) * scale : 0.f;
  }
  float m[8], l[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; l[i] = 0.f; o[i] = 0.f; }

  for (int k0 = 0; k0 < kend; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < KC * DH; i += 128) {
      int r = i / DH, d = i % DH;
      bool ok = (k0 + r < klen);
      Ks[r][d] = ok ? ld_as_float<T>(kb + (long)(k0 + r) * g.ldk + d) : 0.f;
      Vs[r][d] = ok ? ld_as_float<T>(vb + (long)(k0 + r) * g.ldv + d) : 0.f;
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
        s[jj] = geom_visible(g, b, q0 + qr, k0 + key, klen) ? a : -INFINITY;
        cmax = fmaxf(cmax, s[jj]);
      }
      cmax = warp_max(cmax);
      const float mn = fmaxf(m[qi], cmax);
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float p = (s[jj] == -INFINITY) ? 0.f : (ACC ? expf(s[jj] - mn) : __expf(s[jj] - mn));
        psum += p;   // (the softmax denominator is taken before dropout)
        if (drop.state) p *= drop_factor1(dkey, drop, ((uint64_t)(b * H + h) * Tq + (q0 + qr)) * tk8 + (k0 + lane + 32 * jj));
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
    if (t >= Tq) continue;
    const float outv = (l[qi] > 0.f) ? o[qi] / l[qi] : 0.f;
    st_from_float<T>(ctx + ((long)b * Tq + t) * ldo + h * DH + lane, outv);
    if (lse && lane == 0) lse[((long)b * H + h) * Tq + t] = (l[qi] > 0.f) ? m[qi] + logf(l[qi]) : -INFINITY;
  }
}

// Backward pass 1: D_i = dO_i . O_i ; dQ_i = scale * sum_j dS_ij K_j
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnGeom g, const T* __restrict__ ctx, const T* __restrict__ dctx, int ldo,
                                                         const float* __restrict__ lse, T* __restrict__ dq_out, int lddq,
                                                         float* __restrict__ dvec, const DropArgs drop) {
  DropKey dkey{};
  if (drop.state) dkey = drop_key(drop);
  const int Tq = g.Tq, Tk = g.Tk, H = g.H;
  const uint64_t tk8 = (uint64_t)((Tk + 7) >> 3) * 8;
  __shared__ float Ks[KC][DH + 1];
  __shared__ float Vs[KC][DH + 1];
  __shared__ float Qs[QB][DH];
  __shared__ float dOs[QB][DH];
  __shared__ float Ps[4][KC];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float scale = rsqrtf((float)DH);
  const int klen = geom_klen(g, b);
  const int kend = g.causal ? min(klen, q0 + QB) : klen;
  const T* qb = reinterpret_cast<const T*>(g.q) + (long)b * Tq * g.ldq + h * DH;
  const T* kb = reinterpret_cast<const T*>(g.k) + (long)b * Tk * g.ldk + h * DH;
  const T* vb = reinterpret_cast<const T*>(g.v) + (long)b * Tk * g.ldv + h * DH;
  for (int i = tid; i < QB * DH; i += 128) {
    int r = i / DH, d = i % DH;
    bool ok = q0 + r < Tq;
    Qs[r][d] = ok ? ld_as_float<T>(qb + (long)(q0 + r) * g.ldq + d) * scale : 0.f;
    dOs[r][d] = ok ? ld_as_float<T>(dctx + ((long)b * Tq + q0 + r) * ldo + h * DH + d) : 0.f;
  }
  __syncthreads();
  float Di[8], Li[8], dq[8];
#pragma unroll
  for (int qi = 0; qi < 8; ++qi) {
    const int t = q0 + w * 8 + qi;
    float ov = (t < Tq) ? ld_as_float<T>(ctx + ((long)b * Tq + t) * ldo + h * DH + lane) : 0.f;
    Di[qi] = warp_sum(ov * dOs[w * 8 + qi][lane]);
    Li[qi] = (t < Tq) ? lse[((long)b * H + h) * Tq + t] : 0.f;
    dq[qi] = 0.f;
    if (t < Tq && lane == 0) dvec[((long)b * H + h) * Tq + t] = Di[qi];
  }
  for (int k0 = 0; k0 < kend; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < KC * DH; i += 128) {
      int r = i / DH, d = i % DH;
      bool ok = (k0 + r < klen);
      Ks[r][d] = ok ? ld_as_float<T>(kb + (long)(k0 + r) * g.ldk + d) : 0.f;
      Vs[r][d] = ok ? ld_as_float<T>(vb + (long)(k0 + r) * g.ldv + d) : 0.f;
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
        float p = (geom_visible(g, b, q0 + qr, k0 + key, klen) && Li[qi] != -INFINITY) ? expf(a - Li[qi]) : 0.f;
        if (drop.state) dp *= drop_factor1(dkey, drop, ((uint64_t)(b * H + h) * Tq + (q0 + qr)) * tk8 + (k0 + key));
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
    if (t < Tq) st_from_float<T>(dq_out + ((long)b * Tq + t) * lddq + h * DH + lane, dq[qi] * scale);
  }
}

// Backward pass 2: block owns QB keys; dV_j = sum_i P_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnGeom g, const T* __restrict__ dctx, int ldo,
                                                          const float* __restrict__ lse, const float* __restrict__ dvec,
                                                          T* __restrict__ dk_out, int lddk, T* __restrict__ dv_out, int lddv,
                                                          const DropArgs drop) {
  DropKey dkey{};
  if (drop.state) dkey = drop_key(drop);
  const int Tq = g.Tq, Tk = g.Tk, H = g.H;
  const uint64_t tk8 = (uint64_t)((Tk + 7) >> 3) * 8;
  __shared__ float Qs[KC][DH + 1];
  __shared__ float dOs[KC][DH + 1];
  __shared__ float Ksm[QB][DH];
  __shared__ float Vsm[QB][DH];
  __shared__ float Ps[4][KC];
  __shared__ float dSs[4][KC];
  __shared__ float Ls[KC], Ds[KC];
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * QB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float scale = rsqrtf((float)DH);
  const int klen = geom_klen(g, b);
  const T* qb = reinterpret_cast<const T*>(g.q) + (long)b * Tq * g.ldq + h * DH;
  const T* kb = reinterpret_cast<const T*>(g.k) + (long)b * Tk * g.ldk + h * DH;
  const T* vb = reinterpret_cast<const T*>(g.v) + (long)b * Tk * g.ldv + h * DH;
  for (int i = tid; i < QB * DH; i += 128) {
    int r = i / DH, d = i % DH;
    bool ok = j0 + r < Tk;
    Ksm[r][d] = ok ? ld_as_float<T>(kb + (long)(j0 + r) * g.ldk + d) : 0.f;
    Vsm[r][d] = ok ? ld_as_float<T>(vb + (long)(j0 + r) * g.ldv + d) : 0.f;
  }
  float dk[8], dv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dk[i] = 0.f; dv[i] = 0.f; }
  const bool any_valid = j0 < klen;
  if (any_valid) {
    const int i_begin = g.causal ? (j0 / KC) * KC : 0;     // queries before the block's first key see none of its keys
    for (int i0 = i_begin; i0 < Tq; i0 += KC) {
      __syncthreads();
      for (int i = tid; i < KC * DH; i += 128) {
        int r = i / DH, d = i % DH;
        bool ok = (i0 + r < Tq);
        Qs[r][d] = ok ? ld_as_float<T>(qb + (long)(i0 + r) * g.ldq + d) * scale : 0.f;
        dOs[r][d] = ok ? ld_as_float<T>(dctx + ((long)b * Tq + i0 + r) * ldo + h * DH + d) : 0.f;
      }
      for (int i = tid; i < KC; i += 128) {
        bool ok = (i0 + i < Tq);
        Ls[i] = ok ? lse[((long)b * H + h) * Tq + i0 + i] : -INFINITY;
        Ds[i] = ok ? dvec[((long)b * H + h) * Tq + i0 + i] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int ki = 0; ki < 8; ++ki) {
        const int kr = w * 8 + ki;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int q = lane + 32 * jj;
          float a = 0.f, dp = 0.f;
#pragma unroll
          for (int d = 0; d < DH; ++d) {
            a = fmaf(Qs[q][d], Ksm[kr][d], a);
            dp = fmaf(dOs[q][d], Vsm[kr][d], dp);
          }
          float p = (geom_visible(g, b, i0 + q, j0 + kr, klen) && Ls[q] != -INFINITY) ? expf(a - Ls[q]) : 0.f;
          float f = 1.f;
          if (drop.state) f = drop_factor1(dkey, drop, ((uint64_t)(b * H + h) * Tq + (i0 + q)) * tk8 + (j0 + kr));
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
    if (t < Tk) {
      st_from_float<T>(dk_out + ((long)b * Tk + t) * lddk + h * DH + lane, dk[ki]);
      st_from_float<T>(dv_out + ((long)b * Tk + t) * lddv + h * DH + lane, dv[ki]);
    }
  }
}

template <typename T>
static int simt_fwd(const AttnGeom& g, int B, void* ctx, int ldo, float* lse, const DropArgs& drop, cudaStream_t st, bool acc) {
  dim3 grid(cdiv(g.Tq, QB), g.H, B);
  if (acc) attn_fwd_simt_kernel<T, true><<<grid, 128, 0, st>>>(g, (T*)ctx, ldo, lse, drop, active_items(st));
  else attn_fwd_simt_kernel<T, false><<<grid, 128, 0, st>>>(g, (T*)ctx, ldo, lse, drop, active_items(st));
  EEC_LAUNCH_CHECK();
  return 0;
}
template <typename T>
static int simt_bwd(const AttnGeom& g, int B, const void* ctx, const void* dctx, int ldo, const float* lse, void* dq, int lddq, void* dk,
                    int lddk, void* dv, int lddv, float* dvec, const DropArgs& drop, cudaStream_t st) {
  attn_bwd_dq_kernel<T><<<dim3(cdiv(g.Tq, QB), g.H, B), 128, 0, st>>>(g, (const T*)ctx, (const T*)dctx, ldo, lse, (T*)dq, lddq, dvec, drop);
  EEC_LAUNCH_CHECK();
  attn_bwd_dkv_kernel<T><<<dim3(cdiv(g.Tk, QB), g.H, B), 128, 0, st>>>(g, (const T*)dctx, ldo, lse, dvec, (T*)dk, lddk, (T*)dv, lddv, drop);
  EEC_LAUNCH_CHECK();
  return 0;
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
  AttnGeom g{};
  const int es = dtype == EEC_F32 ? 4 : 2, D = H * dh;
  g.q = qkv; g.k = (const char*)qkv + (size_t)D * es; g.v = (const char*)qkv + (size_t)2 * D * es;
  g.ldq = g.ldk = g.ldv = 3 * D; g.Tq = g.Tk = T; g.H = H; g.key_len = key_len;
  if (dtype == EEC_F32) return simt_fwd<float>(g, B, ctx, D, lse, drop, S(stream), true);
  return simt_fwd<__nv_bfloat16>(g, B, ctx, D, lse, drop, S(stream), false);
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
  AttnGeom g{};
  const int es = dtype == EEC_F32 ? 4 : 2, D = H * dh;
  g.q = qkv; g.k = (const char*)qkv + (size_t)D * es; g.v = (const char*)qkv + (size_t)2 * D * es;
  g.ldq = g.ldk = g.ldv = 3 * D; g.Tq = g.Tk = T; g.H = H; g.key_len = key_len;
  char* dq = (char*)dqkv;
  if (dtype == EEC_F32)
    return simt_bwd<float>(g, B, ctx, dctx, D, lse, dq, 3 * D, dq + (size_t)D * es, 3 * D, dq + (size_t)2 * D * es, 3 * D, dvec, drop, S(stream));
  return simt_bwd<__nv_bfloat16>(g, B, ctx, dctx, D, lse, dq, 3 * D, dq + (size_t)D * es, 3 * D, dq + (size_t)2 * D * es, 3 * D, dvec, drop, S(stream));
}

/* ---- general attention (decoder self-attention with causal + key masks, encoder-decoder cross-attention): include/eec.h ---- */
static int geom_from_desc(const eec_attn_desc* d, AttnGeom& g) {
  EEC_CHECK_ARG(d != nullptr, "attn: NULL descriptor");
  EEC_CHECK_ARG(d->dh == 32 && d->H >= 1, "attn: head dim must be 32 (got %d)", d->dh);
  EEC_CHECK_ARG(d->q && d->k && d->v, "attn: NULL operand");
  EEC_CHECK_ARG(d->dtype == EEC_F32 || d->dtype == EEC_BF16, "attn: dtype");
  g.q = d->q; g.k = d->k; g.v = d->v; g.ldq = d->ldq; g.ldk = d->ldk; g.ldv = d->ldv;
  g.Tq = d->Tq; g.Tk = d->Tk; g.H = d->H; g.key_len = d->key_len; g.key_bits = d->key_valid_bits; g.causal = d->causal;
  return 0;
}

extern "C" int eec_attn_general_fwd(const eec_attn_desc* d, void* ctx, int ldo, float* lse, eec_stream_t stream) {
  AttnGeom g{};
  if (int r = geom_from_desc(d, g)) return r;
  if (d->B == 0 || d->Tq == 0) return 0;
  DropArgs drop = make_drop(d->drop_state, d->drop_p, d->drop_site);
  drop.bits = d->drop_bits;
  const bool tc_path = d->dtype == EEC_BF16 && !force_simt() && attn_tc_ready() && attn_general_tc_ok(d);
  EEC_CHECK_ARG(!(tc_path && drop.state && !d->drop_bits), "attn_general_fwd (tensor-core path): dropout needs keep-mask words in drop_bits");
  if (tc_path) return attn_general_fwd_tc(d, ctx, ldo, lse, drop, S(stream));
  if (d->dtype == EEC_F32) return simt_fwd<float>(g, d->B, ctx, ldo, lse, drop, S(stream), true);
  return simt_fwd<__nv_bfloat16>(g, d->B, ctx, ldo, lse, drop, S(stream), false);
}

extern "C" int eec_attn_general_bwd(const eec_attn_desc* d, const void* ctx, const void* dctx, int ldo, const float* lse, void* dq,
                                    int lddq, void* dk, int lddk, void* dv, int lddv, float* dvec, float* dq32, eec_stream_t stream) {
  AttnGeom g{};
  if (int r = geom_from_desc(d, g)) return r;
  if (d->B == 0 || d->Tq == 0) return 0;
  DropArgs drop = make_drop(d->drop_state, d->drop_p, d->drop_site);
  drop.bits = d->drop_bits;
  const bool tc_path = d->dtype == EEC_BF16 && !force_simt() && attn_tc_ready() && attn_general_tc_ok(d);
  EEC_CHECK_ARG(!(tc_path && drop.state && !d->drop_bits), "attn_general_bwd (tensor-core path): dropout needs the forward's keep-mask words in drop_bits");
  if (tc_path) return attn_general_bwd_tc(d, ctx, dctx, ldo, lse, dq, lddq, dk, lddk, dv, lddv, dvec, dq32, drop, S(stream));
  if (d->dtype == EEC_F32) return simt_bwd<float>(g, d->B, ctx, dctx, ldo, lse, dq, lddq, dk, lddk, dv, lddv, dvec, drop, S(stream));
  return simt_bwd<__nv_bfloat16>(g, d->B, ctx, dctx, ldo, lse, dq, lddq, dk, lddk, dv, lddv, dvec, drop, S(stream));
}
