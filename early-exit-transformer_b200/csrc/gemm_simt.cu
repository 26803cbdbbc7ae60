// gemm_simt.cu -- fp32 FFMA GEMM with fused epilogue: the fp32 *parity* path (tcgen05 has no
// fp32 MMA; SURVEY §7-H4 requires genuinely fp32-accurate arithmetic for bit-exact greedy CTC).
// Also instantiated for bf16 operands as a debugging cross-check of the tcgen05 kernels
// (EEC_FORCE_SIMT=1); it is never chosen silently for bf16.
#include "common.cuh"

namespace eec {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

struct SimtParams {
  int M, N, K;
  const void* A; long sa_m, sa_k;
  const void* B; long sb_n, sb_k;
  const float* bias;
  int act;
  void* preact; int ldp;
  float alpha;
  const float* residual; int ldr; int res_row_mod;
  void* C; int ldc;
  int accumulate;
  int k_per_split;
  DropArgs drop;   // dropout right after the activation (element index m*N + n); state NULL = off
  ActiveItems act_items;   // rows past the active-item limit are skipped (early-exit inference)
};

// One operand tile [BK k][128 rows] of fp32 values into registers (8 per thread), for either operand major:
//   K-major  (element (row, k) at base[row*ld + k]):  thread -> (row = idx / 4, k = 4 * (idx % 4)), one 16-byte load along k
//   MN-major (element (row, k) at base[k*ld + row]):  thread -> (k = idx / 32, row = 4 * (idx % 32)), one 16-byte load along rows
// idx = tid + 256 * i, i = 0, 1.  Vector loads need 16-byte alignment (VEC, decided on the host) and a fully in-range quad; anything else
// (tile edges, bf16 operands of the debug cross-check) takes the scalar path with zero fill.
template <typename TIn, bool KMAJ, bool VEC>
__device__ __forceinline__ void load_tile(const TIn* __restrict__ base, long ld, int row0, int rows, int k0, int k_end, int tid, float (&r)[8]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + NT * i;
    const int row = KMAJ ? (idx >> 2) : ((idx & 31) << 2);
    const int k = KMAJ ? ((idx & 3) << 2) : (idx >> 5);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KMAJ) {
      if (row0 + row < rows) {
        const TIn* ptr = base + (long)(row0 + row) * ld + k0 + k;
        if (VEC && sizeof(TIn) == 4 && k0 + k + 3 < k_end) {
          v = *reinterpret_cast<const float4*>(ptr);
        } else {
          if (k0 + k < k_end) v.x = ld_as_float<TIn>(ptr);
          if (k0 + k + 1 < k_end) v.y = ld_as_float<TIn>(ptr + 1);
          if (k0 + k + 2 < k_end) v.z = ld_as_float<TIn>(ptr + 2);
          if (k0 + k + 3 < k_end) v.w = ld_as_float<TIn>(ptr + 3);
        }
      }
    } else {
      if (k0 + k < k_end) {
        const TIn* ptr = base + (long)(k0 + k) * ld + row0 + row;
        if (VEC && sizeof(TIn) == 4 && row0 + row + 3 < rows) {
          v = *reinterpret_cast<const float4*>(ptr);
        } else {
          if (row0 + row < rows) v.x = ld_as_float<TIn>(ptr);
          if (row0 + row + 1 < rows) v.y = ld_as_float<TIn>(ptr + 1);
          if (row0 + row + 2 < rows) v.z = ld_as_float<TIn>(ptr + 2);
          if (row0 + row + 3 < rows) v.w = ld_as_float<TIn>(ptr + 3);
        }
      }
    }
    r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
  }
}
template <bool KMAJ>
__device__ __forceinline__ void store_tile(float (*S)[BM + 4], int tid, const float (&r)[8]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + NT * i;
    if (KMAJ) {
      const int row = idx >> 2, k = (idx & 3) << 2;
      S[k][row] = r[4 * i]; S[k + 1][row] = r[4 * i + 1]; S[k + 2][row] = r[4 * i + 2]; S[k + 3][row] = r[4 * i + 3];
    } else {
      const int k = idx >> 5, row = (idx & 31) << 2;
      *reinterpret_cast<float4*>(&S[k][row]) = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    }
  }
}

// 128 x 128 x 16 tiles, 8 x 8 outputs per thread.  The next k-tile's operands are fetched into registers (16-byte loads) before the
// current tile's 1024 FMAs per thread and stored into the OTHER shared-memory buffer afterwards: one barrier per k-tile and the
// global-load latency hidden under the math.
template <typename TIn, typename TOut, typename TPre, bool ACC, bool AK, bool BKM, bool VEC>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(SimtParams p) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const TIn* __restrict__ A = reinterpret_cast<const TIn*>(p.A);
  const TIn* __restrict__ B = reinterpret_cast<const TIn*>(p.B);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= active_rows(p.act_items, p.M)) return;
  const int k_begin = blockIdx.z * p.k_per_split;
  const int k_end = min(p.K, k_begin + p.k_per_split);
  const int ty = tid / 16, tx = tid % 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const long lda = AK ? p.sa_m : p.sa_k, ldb = BKM ? p.sb_n : p.sb_k;

  float ra[8], rb[8];
  load_tile<TIn, AK, VEC>(A, lda, m0, p.M, k_begin, k_end, tid, ra);
  load_tile<TIn, BKM, VEC>(B, ldb, n0, p.N, k_begin, k_end, tid, rb);
  store_tile<AK>(As[0], tid, ra);
  store_tile<BKM>(Bs[0], tid, rb);
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) {
      load_tile<TIn, AK, VEC>(A, lda, m0, p.M, k0 + BK, k_end, tid, ra);
      load_tile<TIn, BKM, VEC>(B, ldb, n0, p.N, k0 + BK, k_end, tid, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tile<AK>(As[buf ^ 1], tid, ra);
      store_tile<BKM>(Bs[buf ^ 1], tid, rb);
      __syncthreads();
      buf ^= 1;
    }
  }

  TOut* __restrict__ C = reinterpret_cast<TOut*>(p.C);
  TPre* __restrict__ P = reinterpret_cast<TPre*>(p.preact);
  const bool first_split = (blockIdx.z == 0);
  DropKey dkey{};
  if (p.drop.state) dkey = drop_key(p.drop);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    int rr = p.res_row_mod ? (m % p.res_row_mod) : m;
    float df[8];
    if (p.drop.state) drop_factors8(dkey, p.drop, ((uint64_t)m * p.N + n0 + tx * 8) >> 3, df);   // (N % 8 == 0 checked on the host)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + tx * 8 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias && first_split) v += p.bias[n];
      if (p.act == EEC_ACT_SILU) {
        if (P) st_from_float<TPre>(P + (long)m * p.ldp + n, v);
        v = v * (ACC ? sigmoid_acc(v) : sigmoidf_(v));
      } else if (p.act == EEC_ACT_DSILU) {
        float h = ld_as_float<TPre>(P + (long)m * p.ldp + n);
        float s = ACC ? sigmoid_acc(h) : sigmoidf_(h);
        v *= s * (1.0f + h * (1.0f - s));
      } else if (p.act == EEC_ACT_RELU) {
        v = fmaxf(v, 0.f);
      } else if (p.act == EEC_ACT_DRELU) {
        v = ld_as_float<TPre>(P + (long)m * p.ldp + n) > 0.f ? v : 0.f;     // P = the forward's activation output relu(h)
      }
      if (p.drop.state) v *= df[j];
      v *= p.alpha;
      if (p.residual && first_split) v += p.residual[(long)rr * p.ldr + n];
      if (p.accumulate) {
        atomicAdd(reinterpret_cast<float*>(p.C) + (long)m * p.ldc + n, v);
      } else {
        st_from_float<TOut>(C + (long)m * p.ldc + n, v);
      }
    }
  }
}

template <typename T, bool ACC>
__global__ void glu_tail_kernel(const T* __restrict__ z, int ldz, T* __restrict__ out, int ldo, int rows, int C,
                                float alpha) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)rows * C;
  if (i >= total) return;
  int r = (int)(i / C), c = (int)(i % C);
  float a = ld_as_float<T>(z + (long)r * ldz + c);
  float g = ld_as_float<T>(z + (long)r * ldz + C + c);
  float s = ACC ? sigmoid_acc(g) : sigmoidf_(g);
  st_from_float<T>(out + (long)r * ldo + c, alpha * a * s);
}

template <typename TIn, typename TOut, typename TPre, bool ACC>
static int launch2(const SimtParams& p, dim3 grid, cudaStream_t st) {
  const bool ak = p.sa_k == 1, bk = p.sb_k == 1;
  // 16-byte operand loads: fp32 operands, 16-byte aligned bases and row pitches, k ranges that start on a multiple of 4
  const long lda = ak ? p.sa_m : p.sa_k, ldb = bk ? p.sb_n : p.sb_k;
  const bool vec = sizeof(TIn) == 4 && (reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.B) & 15) == 0 &&
                   lda % 4 == 0 && ldb % 4 == 0 && p.k_per_split % 4 == 0;
#define EEC_SIMT(AKV, BKV, VECV) gemm_simt_kernel<TIn, TOut, TPre, ACC, AKV, BKV, VECV><<<grid, NT, 0, st>>>(p)
  if constexpr (sizeof(TIn) == 4) {
    if (vec) {
      if (ak && bk) EEC_SIMT(true, true, true); else if (ak) EEC_SIMT(true, false, true); else if (bk) EEC_SIMT(false, true, true); else EEC_SIMT(false, false, true);
      EEC_LAUNCH_CHECK();
      return 0;
    }
  }
  {
    if (ak && bk) EEC_SIMT(true, true, false); else if (ak) EEC_SIMT(true, false, false); else if (bk) EEC_SIMT(false, true, false); else EEC_SIMT(false, false, false);
  }
#undef EEC_SIMT
  EEC_LAUNCH_CHECK();
  return 0;
}
template <typename TIn, typename TOut, typename TPre>
static int launch(const SimtParams& p, dim3 grid, cudaStream_t st, bool /*acc == (TIn is fp32): accurate expf for the parity path*/) {
  return launch2<TIn, TOut, TPre, sizeof(TIn) == 4>(p, grid, st);
}

int gemm_simt(const eec_gemm_desc* d, cudaStream_t st) {
  SimtParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.A = d->A; p.B = d->B;
  p.sa_m = d->a_kmajor ? d->lda : 1; p.sa_k = d->a_kmajor ? 1 : d->lda;
  p.sb_n = d->b_kmajor ? d->ldb : 1; p.sb_k = d->b_kmajor ? 1 : d->ldb;
  p.bias = d->bias; p.act = d->act; p.preact = d->preact; p.ldp = d->ldp;
  p.alpha = d->alpha; p.residual = d->residual; p.ldr = d->ldr; p.res_row_mod = d->res_row_mod;
  p.C = d->C; p.ldc = d->ldc; p.accumulate = d->accumulate;
  const bool glu = (d->act == EEC_ACT_GLU);
  int out_dtype = d->out_dtype;
  if (glu) {
    EEC_CHECK_ARG(d->preact != nullptr, "gemm(simt): GLU needs a preact (z) buffer");
    EEC_CHECK_ARG(d->residual == nullptr && !d->accumulate, "gemm(simt): GLU with residual/accumulate unsupported");
    EEC_CHECK_ARG(d->preact_dtype == d->out_dtype, "gemm(simt): GLU preact dtype must equal out dtype");
    p.act = EEC_ACT_NONE; p.C = d->preact; p.ldc = d->ldp; p.alpha = 1.f; p.preact = nullptr;
    out_dtype = d->preact_dtype;
  }
  if (p.act == EEC_ACT_DSILU || p.act == EEC_ACT_DRELU) EEC_CHECK_ARG(d->preact != nullptr, "gemm: DSILU / DRELU need preact");
  if (d->accumulate) EEC_CHECK_ARG(d->out_dtype == EEC_F32, "gemm: accumulate needs fp32 C");
  p.drop = make_drop(d->drop_state, d->drop_p, d->drop_site);
  p.act_items = d->a_kmajor ? active_items(st) : ActiveItems{nullptr, 0, 0};   // (rows of A are frames only in the K-major forward form)
  if (p.drop.state) EEC_CHECK_ARG(!glu && !d->accumulate && d->N % 8 == 0, "gemm(simt): dropout with GLU / accumulate / N %% 8 != 0 unsupported");
  int tiles = cdiv(p.N, BN) * cdiv(p.M, BM);
  int splits = 1;
  if (d->accumulate && p.K >= 1024 && tiles < 296) {
    splits = min(cdiv(p.K, 512), max(1, 592 / tiles));
  }
  p.k_per_split = cdiv(cdiv(p.K, splits), BK) * BK;
  splits = cdiv(p.K, p.k_per_split);
  dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM), splits);
  const bool acc = (d->in_dtype == EEC_F32);
  const bool pre_bf16 = (d->preact_dtype == EEC_BF16) && !glu;
  if (d->in_dtype == EEC_F32) {
    if (out_dtype == EEC_F32) {
      if (pre_bf16) { if (int r = launch<float, float, __nv_bfloat16>(p, grid, st, acc)) return r; }
      else { if (int r = launch<float, float, float>(p, grid, st, acc)) return r; }
    } else {
      if (pre_bf16) { if (int r = launch<float, __nv_bfloat16, __nv_bfloat16>(p, grid, st, acc)) return r; }
      else { if (int r = launch<float, __nv_bfloat16, float>(p, grid, st, acc)) return r; }
    }
  } else {
    if (out_dtype == EEC_F32) {
      if (pre_bf16) { if (int r = launch<__nv_bfloat16, float, __nv_bfloat16>(p, grid, st, acc)) return r; }
      else { if (int r = launch<__nv_bfloat16, float, float>(p, grid, st, acc)) return r; }
    } else {
      if (pre_bf16) { if (int r = launch<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(p, grid, st, acc)) return r; }
      else { if (int r = launch<__nv_bfloat16, __nv_bfloat16, float>(p, grid, st, acc)) return r; }
    }
  }
  if (glu) {
    int C = d->N / 2;
    long total = (long)d->M * C;
    int blocks = (int)cdiv64(total, 256);
    if (d->out_dtype == EEC_F32) {
      if (acc) glu_tail_kernel<float, true><<<blocks, 256, 0, st>>>((const float*)d->preact, d->ldp, (float*)d->C, d->ldc, d->M, C, d->alpha);
      else glu_tail_kernel<float, false><<<blocks, 256, 0, st>>>((const float*)d->preact, d->ldp, (float*)d->C, d->ldc, d->M, C, d->alpha);
    } else {
      glu_tail_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)d->preact, d->ldp, (__nv_bfloat16*)d->C, d->ldc, d->M, C, d->alpha);
    }
    EEC_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace eec
