// gemm_tc.cu -- bf16 GEMM on the 5th-gen tensor cores with fused epilogues.
//
//   C[M,N] = epi( A(m,k) * B(n,k) )            128 x 256 output tile per CTA, BLOCK_K = 64
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer
// (one lane) + TMEM allocator, warps 2..5 = epilogue (one thread per accumulator row; TMEM lane
// quarter = warp_id % 4).  Operands arrive by TMA into 128B-swizzled shared-memory stages guarded
// by full/empty mbarriers; the accumulator (128 lanes x 256 fp32 columns) lives in TMEM and is
// read back with tcgen05.ld 32x32b.  Two CTAs are resident per SM (2 x 97 KB smem, 2 x 256 TMEM
// columns) so one CTA's epilogue overlaps the other's main loop.
//
// Both operand majors are supported so forward (NT), dgrad (via pre-transposed bf16 weights or
// MN-major B) and wgrad (A and B both MN-major, split-K with fp32 atomics) use the same kernel.
//
// Epilogue modes:
//   EPI_GENERIC : bias, SiLU / dSiLU, alpha, (row-periodic) residual, fp32|bf16 store, optional
//                 pre-activation store, optional fp32 atomic accumulate (split-K wgrad)
//   EPI_GLU     : B tile = rows [n0,n0+128) and [N/2+n0, N/2+n0+128): out = a * sigmoid(g)
//   EPI_LN      : N == 256: x = res + alpha*(acc+bias) -> C (fp32); LayerNorm(x) -> ln_out;
//                 optional second LayerNorm (final_layer_norm followed by the next module's LN)
//   EPI_LOGSOFTMAX : N == 256: out = log_softmax(acc+bias) fp32, optional argmax / entropy
#include <cstring>

#include "tc_common.cuh"

namespace eec {


// ------------------------------------------------------------------ kernel
namespace {
using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;
constexpr int STAGES = 2;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NTHREADS = 192;

enum { EPI_GENERIC = 0, EPI_GLU = 1, EPI_LN = 2, EPI_LOGSOFTMAX = 3 };

struct TcParams {
  int M, N, K;
  int k_blocks_per_split;
  const float* bias;
  int act;
  void* preact; int ldp; int preact_bf16;
  float alpha;
  const float* residual; int ldr; int res_row_mod;
  void* C; int ldc; int out_bf16;
  int accumulate;
  const float* ln_gamma; const float* ln_beta; void* ln_out; int ln_bf16; int ld_ln;
  float* ln_mean; float* ln_rstd;
  const float* ln2_gamma; const float* ln2_beta; float* ln2_mean; float* ln2_rstd;
  float* x_pre;
  int32_t* argmax; float* entropy;
};

__device__ __forceinline__ void store_row32(void* base, long off, bool bf16, const float (&v)[32]) {
  if (bf16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + off;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = v[i * 8 + j];
      st8<__nv_bfloat16>(p + i * 8, t);
    }
  } else {
    float* p = reinterpret_cast<float*>(base) + off;
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(p + i * 4) = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
  }
}
__device__ __forceinline__ void load_row32(const void* base, long off, bool bf16, float (&v)[32]) {
  if (bf16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + off;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t[8];
      ld8<__nv_bfloat16>(p + i * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i * 8 + j] = t[j];
    }
  } else {
    const float* p = reinterpret_cast<const float*>(base) + off;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 f = *reinterpret_cast<const float4*>(p + i * 4);
      v[i * 4] = f.x; v[i * 4 + 1] = f.y; v[i * 4 + 2] = f.z; v[i * 4 + 3] = f.w;
    }
  }
}

template <bool A_KMAJ, bool B_KMAJ, int EPI>
__global__ void __launch_bounds__(NTHREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BLOCK_M;
  const int n_tile = blockIdx.x;
  // EPI_GLU: tile covers output channels [n_tile*128, +128): B rows {n_tile*128..} and {N/2 + n_tile*128..}
  const int n0 = (EPI == EPI_GLU) ? n_tile * 128 : n_tile * BLOCK_N;
  const int total_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = blockIdx.z * p.k_blocks_per_split;
  const int kb_end = min(total_kb, kb_begin + p.k_blocks_per_split);
  const int nkb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, BLOCK_N); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int k = (kb_begin + i) * BLOCK_K;
        if (A_KMAJ) {
          tma_load_2d(sa, &tmA, &full_bar[s], k, m0);                       // box {64 k, 128 m}
        } else {
          tma_load_2d(sa, &tmA, &full_bar[s], m0, k);                       // box {64 m, 64 k} x 2 atoms
          tma_load_2d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k);
        }
        if (B_KMAJ) {
          if (EPI == EPI_GLU) {
            tma_load_2d(sb, &tmB, &full_bar[s], k, n0);                     // box {64 k, 128 n}
            tma_load_2d(sb + 16384, &tmB, &full_bar[s], k, p.N / 2 + n0);
          } else {
            tma_load_2d(sb, &tmB, &full_bar[s], k, n0);                     // box {64 k, 256 n}
          }
        } else {
#pragma unroll
          for (int a = 0; a < BLOCK_N / 64; ++a) tma_load_2d(sb + a * 8192, &tmB, &full_bar[s], n0 + a * 64, k);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, !A_KMAJ, !B_KMAJ);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          const uint64_t ad = A_KMAJ ? make_smem_desc(sa + k * 32, 0, 1024) : make_smem_desc(sa + k * 2048, 8192, 1024);
          const uint64_t bd = B_KMAJ ? make_smem_desc(sb + k * 32, 0, 1024) : make_smem_desc(sb + k * 2048, 8192, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above retire
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread <-> one row
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const int m = m0 + row_in_tile;
    const bool valid = m < p.M;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    float v[32];

    if (EPI == EPI_GENERIC) {
      const long rr = p.res_row_mod ? (m % p.res_row_mod) : m;
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        const int n = n0 + c0;
        if (n >= p.N) break;
        if (nkb > 0) tmem_ld32(trow + c0, v);
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (!valid) continue;
        if (p.bias && blockIdx.z == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + n + j);
        }
        if (p.act == EEC_ACT_SILU) {
          if (p.preact) store_row32(p.preact, (long)m * p.ldp + n, p.preact_bf16, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = siluf_(v[j]);
        } else if (p.act == EEC_ACT_DSILU) {
          float h[32];
          load_row32(p.preact, (long)m * p.ldp + n, p.preact_bf16, h);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= dsiluf_(h[j]);
        }
        if (p.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
        }
        if (p.residual && blockIdx.z == 0) {
          float r[32];
          load_row32(p.residual, rr * p.ldr + n, false, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += r[j];
        }
        if (p.accumulate) {
          float* c = reinterpret_cast<float*>(p.C) + (long)m * p.ldc + n;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(c + j, v[j]);
        } else {
          store_row32(p.C, (long)m * p.ldc + n, p.out_bf16, v);
        }
      }
    } else if (EPI == EPI_GLU) {
      float g[32];
      for (int c0 = 0; c0 < 128; c0 += 32) {
        tmem_ld32(trow + c0, v);
        tmem_ld32(trow + 128 + c0, g);
        if (!valid) continue;
        const int n = n0 + c0;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] += __ldg(p.bias + n + j); g[j] += __ldg(p.bias + p.N / 2 + n + j); }
        }
        if (p.preact) {
          store_row32(p.preact, (long)m * p.ldp + n, p.preact_bf16, v);
          store_row32(p.preact, (long)m * p.ldp + p.N / 2 + n, p.preact_bf16, g);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = p.alpha * v[j] * sigmoidf_(g[j]);
        store_row32(p.C, (long)m * p.ldc + n, p.out_bf16, v);
      }
    } else if (EPI == EPI_LN) {
      // pass 1: x = res + alpha*(acc+bias) -> fp32 store (x_pre if a second LN follows, else C)
      float* xdst = p.ln2_gamma ? p.x_pre : reinterpret_cast<float*>(p.C);
      const int ldx = p.ldc;
      const long rr = p.res_row_mod ? (m % p.res_row_mod) : m;
      float s1 = 0.f, s2 = 0.f;
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        tmem_ld32(trow + c0, v);
        if (!valid) continue;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + c0 + j);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
        if (p.residual) {
          float r[32];
          load_row32(p.residual, rr * p.ldr + c0, false, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
        store_row32(xdst, (long)m * ldx + c0, false, v);
      }
      if (valid) {
        float mu = s1 * (1.f / 256.f);
        float rs = rsqrtf(fmaxf(s2 * (1.f / 256.f) - mu * mu, 0.f) + 1e-5f);
        if (p.ln_mean) { p.ln_mean[m] = mu; p.ln_rstd[m] = rs; }
        float t1 = 0.f, t2 = 0.f;
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          load_row32(xdst, (long)m * ldx + c0, false, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (v[j] - mu) * rs * __ldg(p.ln_gamma + c0 + j) + __ldg(p.ln_beta + c0 + j);
          if (p.ln2_gamma) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { t1 += v[j]; t2 = fmaf(v[j], v[j], t2); }
            store_row32(p.C, (long)m * p.ldc + c0, false, v);
          } else {
            store_row32(p.ln_out, (long)m * p.ld_ln + c0, p.ln_bf16, v);
          }
        }
        if (p.ln2_gamma) {
          mu = t1 * (1.f / 256.f);
          rs = rsqrtf(fmaxf(t2 * (1.f / 256.f) - mu * mu, 0.f) + 1e-5f);
          if (p.ln2_mean) { p.ln2_mean[m] = mu; p.ln2_rstd[m] = rs; }
          for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            load_row32(p.C, (long)m * p.ldc + c0, false, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (v[j] - mu) * rs * __ldg(p.ln2_gamma + c0 + j) + __ldg(p.ln2_beta + c0 + j);
            store_row32(p.ln_out, (long)m * p.ld_ln + c0, p.ln_bf16, v);
          }
        }
      }
    } else {  // EPI_LOGSOFTMAX
      float mx = -INFINITY;
      int mi = 0;
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        tmem_ld32(trow + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = v[j] + __ldg(p.bias + c0 + j);
          if (t > mx) { mx = t; mi = c0 + j; }
        }
      }
      float se = 0.f;
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        tmem_ld32(trow + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) se += __expf(v[j] + __ldg(p.bias + c0 + j) - mx);
      }
      const float lse = mx + __logf(se);
      float h = 0.f;
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        tmem_ld32(trow + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = v[j] + __ldg(p.bias + c0 + j) - lse;
          h -= __expf(v[j]) * v[j];
        }
        if (valid) store_row32(p.C, (long)m * p.ldc + c0, false, v);
      }
      if (valid) {
        if (p.argmax) p.argmax[m] = mi;
        if (p.entropy) p.entropy[m] = h;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

template <bool AK, bool BK, int EPI>
int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, dim3 grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<AK, BK, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  gemm_tc_kernel<AK, BK, EPI><<<grid, NTHREADS, SMEM_BYTES, st>>>(ta, tb, p);
  EEC_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int gemm_tc(const eec_gemm_desc* d, cudaStream_t st, int32_t* argmax, float* entropy, int logsoftmax) {
  EEC_CHECK_ARG(!(d->drop_state && d->drop_p > 0.f), "gemm_tc (v1): dropout unsupported; unset EEC_GEMM_V1");
  EEC_CHECK_ARG(d->in_dtype == EEC_BF16, "gemm_tc: operands must be bf16");
  EEC_CHECK_ARG(d->N % 32 == 0, "gemm_tc: N (%d) must be a multiple of 32", d->N);
  EEC_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "gemm_tc: empty problem %dx%dx%d", d->M, d->N, d->K);
  int epi = EPI_GENERIC;
  if (logsoftmax) epi = EPI_LOGSOFTMAX;
  else if (d->act == EEC_ACT_GLU) epi = EPI_GLU;
  else if (d->ln_out) epi = EPI_LN;
  if (epi == EPI_LN || epi == EPI_LOGSOFTMAX) {
    EEC_CHECK_ARG(d->N == 256, "gemm_tc: row-wise epilogue needs N == 256 (got %d)", d->N);
    EEC_CHECK_ARG(d->act == EEC_ACT_NONE && !d->accumulate, "gemm_tc: row-wise epilogue with act/accumulate unsupported");
    EEC_CHECK_ARG(d->out_dtype == EEC_F32, "gemm_tc: row-wise epilogue writes fp32 C");
    if (epi == EPI_LN && d->ln2_gamma) EEC_CHECK_ARG(d->x_pre != nullptr, "gemm_tc: ln2 needs x_pre");
    if (epi == EPI_LOGSOFTMAX) EEC_CHECK_ARG(d->bias != nullptr, "gemm_tc: logsoftmax epilogue needs bias");
  }
  if (epi == EPI_GLU) {
    EEC_CHECK_ARG(d->N % 256 == 0 && d->b_kmajor, "gemm_tc: GLU needs N %% 256 == 0 and K-major B");
    EEC_CHECK_ARG(!d->residual && !d->accumulate, "gemm_tc: GLU with residual/accumulate unsupported");
  }
  if (d->act == EEC_ACT_DSILU) EEC_CHECK_ARG(d->preact != nullptr, "gemm_tc: DSILU needs preact");
  if (d->accumulate) EEC_CHECK_ARG(d->out_dtype == EEC_F32 && d->act == EEC_ACT_NONE, "gemm_tc: accumulate needs fp32 C, no act");

  CUtensorMap ta, tb;
  if (d->a_kmajor) { if (int r = get_tmap_2d(&ta, d->A, d->K, d->M, (uint64_t)d->lda * 2, 64, 128)) return r; }
  else { if (int r = get_tmap_2d(&ta, d->A, d->M, d->K, (uint64_t)d->lda * 2, 64, 64)) return r; }
  if (d->b_kmajor) {
    if (int r = get_tmap_2d(&tb, d->B, d->K, d->N, (uint64_t)d->ldb * 2, 64, epi == EPI_GLU ? 128 : 256)) return r;
  } else {
    if (int r = get_tmap_2d(&tb, d->B, d->N, d->K, (uint64_t)d->ldb * 2, 64, 64)) return r;
  }
  TcParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.bias = d->bias; p.act = d->act; p.preact = d->preact; p.ldp = d->ldp; p.preact_bf16 = d->preact_dtype == EEC_BF16;
  p.alpha = d->alpha; p.residual = d->residual; p.ldr = d->ldr; p.res_row_mod = d->res_row_mod;
  p.C = d->C; p.ldc = d->ldc; p.out_bf16 = d->out_dtype == EEC_BF16; p.accumulate = d->accumulate;
  p.ln_gamma = d->ln_gamma; p.ln_beta = d->ln_beta; p.ln_out = d->ln_out; p.ln_bf16 = d->ln_dtype == EEC_BF16; p.ld_ln = d->ld_ln;
  p.ln_mean = d->ln_mean; p.ln_rstd = d->ln_rstd;
  p.ln2_gamma = d->ln2_gamma; p.ln2_beta = d->ln2_beta; p.ln2_mean = d->ln2_mean; p.ln2_rstd = d->ln2_rstd; p.x_pre = d->x_pre;
  p.argmax = argmax; p.entropy = entropy;

  const int total_kb = cdiv(d->K, BLOCK_K);
  const int m_tiles = cdiv(d->M, BLOCK_M);
  const int n_tiles = (epi == EPI_GLU) ? d->N / 256 : cdiv(d->N, BLOCK_N);
  int splits = 1;
  if (d->accumulate) {
    const int tiles = m_tiles * n_tiles;
    if (tiles < 296 && total_kb >= 16) splits = min(cdiv(total_kb, 8), max(1, 296 / tiles));
  }
  p.k_blocks_per_split = cdiv(total_kb, splits);
  splits = cdiv(total_kb, p.k_blocks_per_split);
  dim3 grid(n_tiles, m_tiles, splits);

#define EEC_TC_DISPATCH(AK, BK)                                                                       \
  switch (epi) {                                                                                      \
    case EPI_GENERIC: return launch_tc<AK, BK, EPI_GENERIC>(ta, tb, p, grid, st);                      \
    case EPI_LN: return launch_tc<AK, BK, EPI_LN>(ta, tb, p, grid, st);                                \
    default: break;                                                                                   \
  }
  if (epi == EPI_GLU) {
    EEC_CHECK_ARG(d->a_kmajor, "gemm_tc: GLU needs K-major A");
    return launch_tc<true, true, EPI_GLU>(ta, tb, p, grid, st);
  }
  if (epi == EPI_LOGSOFTMAX) {
    EEC_CHECK_ARG(d->a_kmajor && d->b_kmajor, "gemm_tc: logsoftmax epilogue needs K-major operands");
    return launch_tc<true, true, EPI_LOGSOFTMAX>(ta, tb, p, grid, st);
  }
  if (d->a_kmajor && d->b_kmajor) { EEC_TC_DISPATCH(true, true) }
  else if (d->a_kmajor && !d->b_kmajor) { EEC_TC_DISPATCH(true, false) }
  else if (!d->a_kmajor && !d->b_kmajor) { EEC_TC_DISPATCH(false, false) }
  else { EEC_TC_DISPATCH(false, true) }
  set_error("gemm_tc: unsupported configuration");
  return 1;
}

}  // namespace eec
