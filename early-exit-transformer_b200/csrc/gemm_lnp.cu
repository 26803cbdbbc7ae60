// gemm_lnp.cu -- LayerNorm-tail GEMM, CTA-PAIR version (tcgen05 cta_group::2) of gemm_ln3_kernel (gemm_tc3.cu):
//
//   x = residual + alpha * (A W^T + bias) -> fp32 C;  LayerNorm(x) -> ln_out          (N == 256, K-major operands)
//
// the out-projection, pointwise conv 2 and FFN down-projection of a conformer layer (TA:202, :65, :107 followed by :151 / :42 / :103 /
// :211): 48 launches per training step, half of the inference forward.  A pair owns 256 rows; every CTA streams its own 128 rows of
// A and HALF of the weight k-block (128 of the 256 n-rows): 32 KB per k-block instead of 48, four k-blocks in flight instead of three,
// and per tile 1/3 fewer operand bytes through L2 (the K = 2048 form reads the whole 1 MB weight matrix once per tile).  Every CTA
// still owns WHOLE rows (its 128 x 256 accumulator), so the row-wise epilogue of gemm_ln3_kernel is reused unchanged: 8 warps, 128
// x-values per thread kept in registers across the row-statistics exchange, fp32 residual through per-warp TMA staging.
// Pair mechanics as in gemm_ws2.cu / gemm_pair.cu.
#include <stdlib.h>
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64;            // BM = rows per CTA (256 per pair)
constexpr int A_BYTES = BM * BK * 2;                  // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;            // 16 KB: this CTA's half of the weight k-block
constexpr int STAGE = A_BYTES + B_BYTES;              // 32 KB
constexpr int LN_WARPS = 8;
constexpr int LN_NT = 64 + LN_WARPS * 32;             // 320
constexpr int LN_WBUF = 8192;                         // two 4 KB staging buffers per epilogue warp
constexpr int LN_NSTAGE = 4;
constexpr int LN_OFF_STG = LN_NSTAGE * STAGE;                 // 131072
constexpr int LN_OFF_VEC = LN_OFF_STG + LN_WARPS * LN_WBUF;   // float[3][256]: bias, gamma, beta
constexpr int LN_OFF_XCH = LN_OFF_VEC + 3 * 256 * 4;          // float[2 halves][128 rows][2]
constexpr int LN_OFF_BAR = LN_OFF_XCH + 2 * 128 * 2 * 4;
constexpr int LNP_SMEM_BYTES = LN_OFF_BAR + 512;

struct PLN {
  int M, K, m_tiles;
  int rb;          // rows per CTA and tile: 128, 96 or 64 (the MMA is always 256 x 256; rows past rb of a CTA's accumulator are never read)
  const float* bias;
  float alpha;
  int has_res, ln_bf16;
  const float* ln_gamma; const float* ln_beta; float* ln_mean; float* ln_rstd;
  DropArgs drop;   // DROP instantiation only: x = residual + alpha * dropout(A W^T + bias), element index m*256 + n
  ActiveItems act_items;   // super-tiles past the active-item limit are skipped by all three roles (early-exit inference)
};

__device__ __forceinline__ bool try_wait_h(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_h(uint64_t* bar, uint32_t parity) {
  if (try_wait_h(bar, parity)) return;
  const long long t0 = clock64();
  while (!try_wait_h(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("eec: gemm_lnp mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

template <bool DROP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LN_NT, 1)
    gemm_lnp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC,   // fp32 x out (box 32 x 32)
                    const __grid_constant__ CUtensorMap tmR,   // fp32 residual in (box 32 x 32)
                    const __grid_constant__ CUtensorMap tmL,   // LayerNorm out (box 32 x 32)
                    const PLN p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) { printf("eec: gemm_lnp smem base not 1024-aligned\n"); __trap(); }
  float* vecs = reinterpret_cast<float*>(smem + LN_OFF_VEC);
  float* xch = reinterpret_cast<float*>(smem + LN_OFF_XCH);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + LN_OFF_BAR);   // (leader's copy) stage s of BOTH CTAs has landed
  uint64_t* empty_bar = full_bar + LN_NSTAGE;                            // (per CTA)
  uint64_t* tfull_bar = empty_bar + LN_NSTAGE;   // [2] (per CTA)
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] (leader's copy) drained by the epilogue warps of both CTAs
  uint64_t* res_bar = tempty_bar + 2;            // [LN_WARPS][2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 2 * LN_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmR);
    for (int s = 0; s < LN_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * LN_WARPS); }
    for (int w = 0; w < 2 * LN_WARPS; ++w) mbar_init(&res_bar[w], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  for (int i = threadIdx.x; i < 256; i += LN_NT) {
    vecs[i] = p.bias ? p.bias[i] : 0.f;
    vecs[256 + i] = p.ln_gamma[i];
    vecs[512 + i] = p.ln_beta[i];
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int st_tiles = (active_rows(p.act_items, p.M) + 2 * p.rb - 1) / (2 * p.rb);   // (2 rb)-row super-tiles (== all of them unless an active-item limit is set)

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      const int nb = (int)rank * (BN / 2);
      for (int st = pair; st < st_tiles; st += npairs) {
        const int m0 = st * 2 * p.rb + (int)rank * p.rb;
        for (int kb = 0; kb < total_kb; ++kb) {
          wait_h(&empty_bar[s], ph);
          uint8_t* sa = smem + s * STAGE;
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * (p.rb * BK * 2 + B_BYTES));   // the A box holds rb rows
          tma_load_2d_2sm(sa, &tmA, &full_bar[s], kb * BK, m0);
          tma_load_2d_2sm(sa + A_BYTES, &tmB, &full_bar[s], kb * BK, nb);
          if (++s == LN_NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, false, false);
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 0, 1024), bdesc0 = make_smem_desc(smem_u32(smem) + A_BYTES, 0, 1024);
      uint32_t ut = 0, ph = 0;
      int s = 0;
      for (int st = pair; st < st_tiles; st += npairs, ++ut) {
        const uint32_t acc = ut & 1;
        wait_h(&tempty_bar[acc], ((ut >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < total_kb; ++kb) {
          wait_h(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t so = (uint64_t)((s * STAGE) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma2_bf16(d_tmem, adesc0 + so + k * 2, bdesc0 + so + k * 2, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma2_commit_both(&empty_bar[s]);
          if (++s == LN_NSTAGE) { s = 0; ph ^= 1; }
        }
        umma2_commit_both(&tfull_bar[acc]);
      }
    }
  } else {
    // ================================================================== epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const int r = q * 32 + lane;
    const int cb = half * 128;
    uint8_t* buf[2] = {smem + LN_OFF_STG + e * LN_WBUF, smem + LN_OFF_STG + e * LN_WBUF + 4096};
    uint64_t* rbar = res_bar + 2 * e;
    const int sw = lane & 7;
    float x[128];   // this thread's 128 x-values (its row, its column half): TMEM is read once, the values never leave registers
    uint32_t ut = 0;
    for (int st = pair; st < st_tiles; st += npairs, ++ut) {
      const int m0 = st * 2 * p.rb + (int)rank * p.rb;
      const int row0 = m0 + q * 32;
      const int m = m0 + r;
      const bool valid = m < p.M;
      if (q * 32 >= p.rb) {   // (warp-uniform) rb < 128: this lane quarter holds no rows of the tile; only the accumulator hand-shake remains
        mbar_wait(&tfull_bar[ut & 1], (ut >> 1) & 1);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&tempty_bar[ut & 1]);
        continue;
      }
      uint32_t dw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};   // dropout keep-mask words of this thread's 4 x 32 columns
      if (DROP && valid) {
        const uint32_t* db = reinterpret_cast<const uint32_t*>(p.drop.bits);
#pragma unroll
        for (int s = 0; s < 4; ++s) dw[s] = db[(long)((cb >> 5) + s) * p.M + m];
      }
      const uint32_t acc = ut & 1;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cb;
      if (p.has_res && lane == 0) {
        bulk_wait_read<0>();   // the previous tile's stores have finished reading both staging buffers
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          mbar_expect_tx(&rbar[s], 4096);
          tma_load_2d(buf[s], &tmR, &rbar[s], cb + s * 32, row0);
        }
      }
      __syncwarp();
      mbar_wait(&tfull_bar[acc], (ut >> 1) & 1);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        float(&v)[32] = *reinterpret_cast<float(*)[32]>(&x[s * 32]);
        tmem_ld32(trow + s * 32, v);
        if (s == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(&tempty_bar[acc]);   // the whole accumulator slice is in registers
        }
        const float4* bp = reinterpret_cast<const float4*>(vecs + cb + s * 32);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 f = bp[g];
          v[g * 4] = (v[g * 4] + f.x) * p.alpha; v[g * 4 + 1] = (v[g * 4 + 1] + f.y) * p.alpha;
          v[g * 4 + 2] = (v[g * 4 + 2] + f.z) * p.alpha; v[g * 4 + 3] = (v[g * 4 + 3] + f.w) * p.alpha;
        }
        if (DROP) drop_apply_bits<32>(v, dw[s], p.drop.scale);
        uint8_t* b = buf[s & 1];
        uint8_t* row = b + lane * 128;
        if (p.has_res) {
          mbar_wait(&rbar[s & 1], (s >> 1) & 1);   // each buffer's barrier completes twice per tile
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 f = *reinterpret_cast<const float4*>(row + ((g ^ sw) << 4));
            v[g * 4] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
          }
        } else {
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s1 += v[j];
          s2 = fmaf(v[j], v[j], s2);
        }
        // x slab -> same staging buffer (every lane rewrites the row it has just consumed) -> fp32 store
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(row + ((g ^ sw) << 4)) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, b, cb + s * 32, row0);
          bulk_commit();
          if (p.has_res && s < 2) {   // refill this buffer with the residual of slab s + 2 once the store has read it
            bulk_wait_read<0>();
            mbar_expect_tx(&rbar[s & 1], 4096);
            tma_load_2d(b, &tmR, &rbar[s & 1], cb + (s + 2) * 32, row0);
          }
        }
        __syncwarp();
      }
      // row statistics: combine with the warp that owns the other 128 columns of the same rows
      xch[half * 256 + r * 2] = s1;
      xch[half * 256 + r * 2 + 1] = s2;
      bar_sync(1 + q, 64);
      s1 += xch[(half ^ 1) * 256 + r * 2];
      s2 += xch[(half ^ 1) * 256 + r * 2 + 1];
      const float mu = s1 * (1.f / 256.f);
      const float rs = rsqrtf(fmaxf(s2 * (1.f / 256.f) - mu * mu, 0.f) + 1e-5f);
      if (half == 0 && valid && p.ln_mean) { p.ln_mean[m] = mu; p.ln_rstd[m] = rs; }
      bar_sync(1 + q, 64);   // both warps have read the exchange slots: the next tile may overwrite them
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        float(&v)[32] = *reinterpret_cast<float(*)[32]>(&x[s * 32]);
        const float4* gp = reinterpret_cast<const float4*>(vecs + 256 + cb + s * 32);
        const float4* bp2 = reinterpret_cast<const float4*>(vecs + 512 + cb + s * 32);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 gg = gp[g], bb = bp2[g];
          v[g * 4] = (v[g * 4] - mu) * rs * gg.x + bb.x;
          v[g * 4 + 1] = (v[g * 4 + 1] - mu) * rs * gg.y + bb.y;
          v[g * 4 + 2] = (v[g * 4 + 2] - mu) * rs * gg.z + bb.z;
          v[g * 4 + 3] = (v[g * 4 + 3] - mu) * rs * gg.w + bb.w;
        }
        uint8_t* b = buf[s & 1];
        if (lane == 0) bulk_wait_read<1>();   // the store before the newest one used this buffer
        __syncwarp();
        if (p.ln_bf16) {
          uint8_t* row = b + lane * 64;
          const int sw64 = (lane >> 1) & 3;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[g * 8 + 2 * k], v[g * 8 + 2 * k + 1]);
            *reinterpret_cast<uint4*>(row + ((g ^ sw64) << 4)) = u;
          }
        } else {
          uint8_t* row = b + lane * 128;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(row + ((g ^ sw) << 4)) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmL, b, cb + s * 32, row0);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_all();
    tc_fence_before();
  }
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int g_sms_lnp = 0;

}  // namespace

bool gemm_lnp_ok(const eec_gemm_desc* d, cudaStream_t st) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("EEC_GEMM_LNP"); env = (e && e[0] == '0') ? 0 : 1; }
  if (!env) return false;
  if (!g_sms_lnp) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_sms_lnp, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
  }
  if (g_sms_lnp % 2) return false;
  (void)st;
  return true;
}

// LayerNorm-tail epilogue of eec_gemm (N == 256, K-major bf16 operands, fp32 C with ldc 256); validated by gemm_tc2 / gemm_ln3
int gemm_lnp(const eec_gemm_desc* d, cudaStream_t st) {
  CUtensorMap ta, tb, tcm, trm, tlm;
  // Rows per CTA and tile.  M = 23 936 is 93.5 tiles of 256 rows for 74 pairs: the second wave would stream on 27 % of the SMs (measured:
  // 16 of the FFN-down GEMM's 52 us).  With 96 (64) rows per CTA the same 2 (3) waves carry 192 rows per pair instead of 256; the MMA
  // still runs 256 x 256 (these GEMMs are bandwidth-bound), the A box and the epilogue simply stop at rb rows.
  const int pairs = g_sms_lnp / 2;
  static int rb_env = -1;
  if (rb_env < 0) { const char* e = getenv("EEC_LNP_RB"); rb_env = e ? atoi(e) : 0; }
  int rb = 128;
  if (rb_env == 128 || rb_env == 96 || rb_env == 64) rb = rb_env;
  else {
    long best = -1;
    for (int c : {128, 96, 64}) {
      const long waves = cdiv(cdiv(d->M, 2 * c), pairs);
      const long cost = waves * (c + 8);   // + a per-tile constant for the row-wise epilogue
      if (best < 0 || cost < best) { best = cost; rb = c; }
    }
  }
  if (int r = get_tmap_2d(&ta, d->A, d->K, d->M, (uint64_t)d->lda * 2, 64, (uint32_t)rb)) return r;
  if (int r = get_tmap_2d(&tb, d->B, d->K, d->N, (uint64_t)d->ldb * 2, 64, 128)) return r;   // one CTA's half of the weight k-block
  if (int r = get_tmap_box32(&tcm, d->C, false, 256, (uint64_t)d->M, (uint64_t)d->ldc)) return r;
  trm = tcm;
  if (d->residual) { if (int r = get_tmap_box32(&trm, d->residual, false, 256, (uint64_t)d->M, 256)) return r; }
  if (int r = get_tmap_box32(&tlm, d->ln_out, d->ln_dtype == EEC_BF16, 256, (uint64_t)d->M, (uint64_t)d->ld_ln)) return r;
  PLN p{};
  p.M = d->M; p.K = d->K; p.m_tiles = cdiv(d->M, BM); p.rb = rb;
  p.bias = d->bias; p.alpha = d->alpha; p.has_res = d->residual != nullptr; p.ln_bf16 = d->ln_dtype == EEC_BF16;
  p.ln_gamma = d->ln_gamma; p.ln_beta = d->ln_beta; p.ln_mean = d->ln_mean; p.ln_rstd = d->ln_rstd;
  p.act_items = active_items(st);
  p.drop = make_drop(d->drop_state, d->drop_p, d->drop_site);
  p.drop.bits = d->drop_bits;
  if (p.drop.state)
    EEC_CHECK_ARG(d->drop_bits != nullptr, "gemm (LayerNorm tail): dropout needs the keep-mask words of eec_dropout_bits(R = M, C = 256, Cs = 256, W = 32) in drop_bits");
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(gemm_lnp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LNP_SMEM_BYTES));
    EEC_CUDA(cudaFuncSetAttribute(gemm_lnp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LNP_SMEM_BYTES));
    attr_set = true;
  }
  const int grid = 2 * min(cdiv(d->M, 2 * rb), pairs);
  if (p.drop.state) gemm_lnp_kernel<true><<<dim3(grid), dim3(LN_NT), LNP_SMEM_BYTES, st>>>(ta, tb, tcm, trm, tlm, p);
  else gemm_lnp_kernel<false><<<dim3(grid), dim3(LN_NT), LNP_SMEM_BYTES, st>>>(ta, tb, tcm, trm, tlm, p);
  EEC_LAUNCH_CHECK();
  return 0;
}

}  // namespace eec
