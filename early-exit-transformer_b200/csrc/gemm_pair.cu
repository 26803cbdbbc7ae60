// gemm_pair.cu -- streaming bf16 tcgen05 GEMM with an fp32 output, CTA-PAIR version (tcgen05 cta_group::2): the weight gradients
// (dW += dY^T X over the 23 936 frames, split-K, TA:104-108 / in-proj / pointwise convs under autograd) and the long-K data
// gradients (dU = dH W1, K = 2048) of a conformer layer.
//
//   C[M,N] (fp32) (+)= alpha * A(m,k) * B(n,k)      256 x 256 tile per CTA pair, BLOCK_K = 64, 6-stage ring of 32 KB per CTA
//
// Why pairs: these forms stream BOTH operands and are bound by (bytes in flight per SM) / (TMA round trip), not by the tensor pipe
// (DESIGN.md section 4): the single-CTA kernel (gemm_tc3.cu, NST = 4) holds 4 k-blocks x 48 KB per SM.  With cta_group::2 a CTA
// supplies its own 128 rows of A and only HALF of the B tile (128 of the 256 n-rows): 32 KB per k-block, so the same 192 KB hold
// SIX k-blocks, every SM pulls 1/3 fewer operand bytes through L2 per FLOP, and the MMA reads 64 instead of 96 bytes of shared
// memory per clock.  Pair mechanics (2SM TMA completing on the leader's barriers, multicast commit, remote accumulator release) are
// those of gemm_ws2.cu, where they were validated first.
//
//   warp 0      : TMA producer (both CTAs)
//   warp 1      : tcgen05.mma.cta_group::2 issuer (leader CTA only), two 256-column TMEM accumulators
//   warps 2..17 : epilogue: TMEM -> registers -> [32 rows x 16 cols] fp32 staging boxes -> bulk tensor store / reduce-add;
//                 with a_colsum (bias gradient fused into the weight-gradient GEMM) they also read every A stage in shared memory
#include <stdlib.h>
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64;            // BM = rows per CTA (256 per pair)
constexpr int NS = 6;
constexpr int A_BYTES = BM * BK * 2;                  // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;            // 16 KB: this CTA's half of the B tile
constexpr int STAGE = A_BYTES + B_BYTES;              // 32 KB
constexpr int NEW = 16;
constexpr int NTP = 64 + NEW * 32;                    // 576
constexpr int OFF_STG = NS * STAGE;                   // 196608
constexpr int OFF_SCR = OFF_STG + NEW * 2048;         // 229376: float[512] scratch (a_colsum partial sums)
constexpr int OFF_BAR = OFF_SCR + 2048;               // 231424
constexpr int PAIR_SMEM = OFF_BAR + 256;              // 231680 <= 232448

struct PP {
  int M, N, K;
  int mt2_tiles, n_tiles, splits, kb_per_split;       // 256-row super-tiles x 256-column tiles x K splits
  float alpha;
  int accumulate;
  float* a_colsum; float a_colsum_scale;
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
      printf("eec: gemm_pair mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void arrive_peer(uint64_t* bar) {   // the leader arrives on the odd CTA's copy of `bar`
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

// work unit u = split * (mt2_tiles * n_tiles) + mt2 * n_tiles + nt, walked with a stride of (gridDim.x / 2) pairs
struct Unit {
  int split, mt2, nt;
  __device__ __forceinline__ Unit(const PP& p, int u) {
    const int tiles = p.mt2_tiles * p.n_tiles;
    split = u / tiles;
    const int t = u - split * tiles;
    mt2 = t / p.n_tiles;
    nt = t - mt2 * p.n_tiles;
  }
};

template <bool A_KMAJ, bool B_KMAJ, bool OUT_BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTP, 1)
    gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const PP p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) { printf("eec: gemm_pair smem base not 1024-aligned\n"); __trap(); }
  float* scr = reinterpret_cast<float*>(smem + OFF_SCR);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [6] (leader's copy) stage s of BOTH CTAs has landed
  uint64_t* empty_bar = full_bar + NS;                                // [6] (per CTA) its MMAs have completed (+ the a_colsum readers)
  uint64_t* afull_bar = empty_bar + NS;                               // [6] (per CTA, a_colsum only) relayed by the leader's MMA thread
  uint64_t* tfull_bar = afull_bar + NS;                               // [2] (per CTA)
  uint64_t* tempty_bar = tfull_bar + 2;                               // [2] (leader's copy) drained by the epilogue warps of both CTAs
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_units = p.mt2_tiles * p.n_tiles * p.splits;
  const int total_kb = (p.K + BK - 1) / BK;
  const bool colsum = !A_KMAJ && p.a_colsum != nullptr;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], colsum ? 1 + NEW : 1);
      mbar_init(&afull_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * NEW); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int u = pair; u < n_units; u += npairs) {
        const Unit un(p, u);
        const int m0 = un.mt2 * 2 * BM + (int)rank * BM;
        const int nb = un.nt * BN + (int)rank * (BN / 2);
        const int kb0 = un.split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          wait_h(&empty_bar[s], ph);
          uint8_t* sa = smem + s * STAGE;
          uint8_t* sb = sa + A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE);
          const int k = kb * BK;
          if (A_KMAJ) {
            tma_load_2d_2sm(sa, &tmA, &full_bar[s], k, m0);
          } else {
            tma_load_2d_2sm(sa, &tmA, &full_bar[s], m0, k);
            tma_load_2d_2sm(sa + 8192, &tmA, &full_bar[s], m0 + 64, k);
          }
          if (B_KMAJ) {
            tma_load_2d_2sm(sb, &tmB, &full_bar[s], k, nb);
          } else {
            tma_load_2d_2sm(sb, &tmB, &full_bar[s], nb, k);
            tma_load_2d_2sm(sb + 8192, &tmB, &full_bar[s], nb + 64, k);
          }
          if (++s == NS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, !A_KMAJ, !B_KMAJ);
      constexpr uint64_t A_KSTEP = (A_KMAJ ? 32 : 2048) >> 4, B_KSTEP = (B_KMAJ ? 32 : 2048) >> 4;
      const uint32_t sa0 = smem_u32(smem), sb0 = sa0 + A_BYTES;
      const uint64_t adesc0 = A_KMAJ ? make_smem_desc(sa0, 0, 1024) : make_smem_desc(sa0, 8192, 1024);
      const uint64_t bdesc0 = B_KMAJ ? make_smem_desc(sb0, 0, 1024) : make_smem_desc(sb0, 8192, 1024);
      uint32_t ut = 0, ph = 0;
      int s = 0;
      for (int u = pair; u < n_units; u += npairs, ++ut) {
        const Unit un(p, u);
        const int kb0 = un.split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const uint32_t acc = ut & 1;
        wait_h(&tempty_bar[acc], ((ut >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          wait_h(&full_bar[s], ph);
          if (colsum) { mbar_arrive(&afull_bar[s]); arrive_peer(&afull_bar[s]); }   // both CTAs' a_colsum readers may read stage s
          tc_fence_after();
          const uint64_t so = (uint64_t)((s * STAGE) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma2_bf16(d_tmem, adesc0 + so + k * A_KSTEP, bdesc0 + so + k * B_KSTEP, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma2_commit_both(&empty_bar[s]);
          if (++s == NS) { s = 0; ph ^= 1; }
        }
        umma2_commit_both(&tfull_bar[acc]);
      }
    }
  } else {
    // ================================================================== epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;
    const int cg = e >> 2;
    uint8_t* buf = smem + OFF_STG + e * 2048;
    uint8_t* srow = buf + lane * 64;
    const int sw = (lane >> 1) & 3;
    uint32_t ut = 0;
    int cs = 0;
    uint32_t cph = 0;
    for (int u = pair; u < n_units; u += npairs, ++ut) {
      const Unit un(p, u);
      const int m0 = un.mt2 * 2 * BM + (int)rank * BM;
      if (colsum) {
        // bias gradient of the layer whose weight gradient this GEMM computes: column sums of the MN-major A operand
        // ([64 k][128 m] bf16 per stage, 128B-swizzled) taken while the tile sits in shared memory.  Thread (cm, kq) adds
        // 16 k-rows of column cm per k-block; units with n-tile 0 publish (one atomic per column per unit).
        const int tid_e = threadIdx.x - 64, cm = tid_e & 127, kq = tid_e >> 7;
        const int kb0 = un.split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const uint32_t coff = (uint32_t)((cm >> 6) * 8192 + (((cm & 63) & 7) * 2));
        const int cchunk = (cm & 63) >> 3;
        float csum = 0.f;
        for (int kb = kb0; kb < kb1; ++kb) {
          wait_h(&afull_bar[cs], cph);
          if (un.nt == 0) {
            const uint8_t* a = smem + cs * STAGE + coff;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
              const int k = kq * 16 + kk;
              csum += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(a + k * 128 + ((cchunk ^ (k & 7)) << 4)));
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[cs]);
          if (++cs == NS) { cs = 0; cph ^= 1; }
        }
        if (un.nt == 0) {
          scr[kq * 128 + cm] = csum;
          bar_sync(1, NEW * 32);
          if (kq == 0 && m0 + cm < p.M) atomicAdd(p.a_colsum + m0 + cm, p.a_colsum_scale * (scr[cm] + scr[128 + cm] + scr[256 + cm] + scr[384 + cm]));
          bar_sync(1, NEW * 32);
        }
      }
      const uint32_t acc = ut & 1;
      const int row0 = m0 + q * 32;
      const int ncol = un.nt * BN + cg * 64;
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cg * 64;
      wait_h(&tfull_bar[acc], (ut >> 1) & 1);
      tc_fence_after();
      uint32_t ra[16], rb[16], pk[16];
      tmem_ld16_async(tcol, ra);
#pragma unroll
      for (int ss = 0; ss < 4; ++ss) {
        uint32_t(&cur)[16] = (ss & 1) ? rb : ra;
        uint32_t(&nxt)[16] = (ss & 1) ? ra : rb;
        tmem_ld_wait16(cur);
        if (ss < 3) {
          tmem_ld16_async(tcol + (ss + 1) * 16, nxt);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(&tempty_bar[acc]);
        }
        if (OUT_BF16) {
          // bf16 output (data gradients feeding a LayerNorm backward): one [32 rows x 32 cols] box per two sub-slabs
          if (ncol + (ss & ~1) * 16 >= p.N) continue;   // (warp-uniform) column tail
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(cur[j]) * p.alpha, __uint_as_float(cur[j + 1]) * p.alpha);
            pk[(ss & 1) * 8 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          if (ss & 1) {
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(srow + ((c ^ sw) << 4)) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, buf, ncol + (ss >> 1) * 32, row0);
              bulk_commit();
            }
          }
          continue;
        }
        if (ncol + ss * 16 >= p.N) continue;   // (warp-uniform) column tail
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(cur[j]) * p.alpha;
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<float4*>(srow + ((c ^ sw) << 4)) = make_float4(x[c * 4], x[c * 4 + 1], x[c * 4 + 2], x[c * 4 + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p.accumulate) tma_reduce_add_2d(&tmC, buf, ncol + ss * 16, row0);
          else tma_store_2d(&tmC, buf, ncol + ss * 16, row0);
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

template <bool AK, bool BK_, bool OB = false>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const PP& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<AK, BK_, OB>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
    attr_set = true;
  }
  gemm_pair_kernel<AK, BK_, OB><<<dim3(grid), dim3(NTP), PAIR_SMEM, st>>>(ta, tb, tc_, p);   // (static cluster dims 2 x 1 x 1)
  EEC_LAUNCH_CHECK();
  return 0;
}

int g_sms_pair = 0;

}  // namespace

// plain fp32-output GEMMs (store or split-K accumulate; MN-major B): eligible for the pair kernel?
bool gemm_pair_ok(const eec_gemm_desc* d, cudaStream_t st) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("EEC_GEMM_PAIR"); env = (e && e[0] == '0') ? 0 : 1; }
  if (!env) return false;
  if (!g_sms_pair) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_sms_pair, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
  }
  if (g_sms_pair % 2) return false;
  const bool out_bf16 = d->out_dtype == EEC_BF16;
  if (out_bf16 && (!d->a_kmajor || d->accumulate || d->ldc % 8 != 0)) return false;   // bf16 output: plain data gradients only
  if (d->in_dtype != EEC_BF16 || (d->out_dtype != EEC_F32 && !out_bf16) || d->bias || d->act != EEC_ACT_NONE || d->preact || d->residual || d->ln_out) return false;
  if (d->b_kmajor) return false;                       // instantiated: (A K-major | MN-major) x B MN-major = data / weight gradients
  if (d->drop_state && d->drop_p > 0.f) return false;
  if (d->a_colsum && d->a_kmajor) return false;
  if (d->K < 2 * BK) return false;
  if (active_items(st).n_dev) return false;            // (early-exit inference never runs these forms)
  return true;
}

int gemm_pair(const eec_gemm_desc* d, cudaStream_t st) {
  CUtensorMap ta, tb, tcm;
  if (d->a_kmajor) { if (int r = get_tmap_2d(&ta, d->A, d->K, d->M, (uint64_t)d->lda * 2, 64, 128)) return r; }
  else { if (int r = get_tmap_2d(&ta, d->A, d->M, d->K, (uint64_t)d->lda * 2, 64, 64)) return r; }
  if (int r = get_tmap_2d(&tb, d->B, d->N, d->K, (uint64_t)d->ldb * 2, 64, 64)) return r;
  const bool out_bf16 = d->out_dtype == EEC_BF16;
  if (out_bf16) { if (int r = get_tmap_box32(&tcm, d->C, true, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldc)) return r; }
  else if (int r = get_tmap_box(&tcm, d->C, false, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldc, 16, 32, 2 /*SWIZZLE_64B*/)) return r;
  PP p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.mt2_tiles = cdiv(d->M, 2 * BM);
  p.n_tiles = cdiv(d->N, BN);
  const int pairs = g_sms_pair / 2;
  const int total_kb = cdiv(d->K, BK);
  int splits = 1;
  if (d->accumulate) {
    const int tiles = p.mt2_tiles * p.n_tiles;
    if (tiles < pairs && total_kb >= 16) splits = min(cdiv(total_kb, 8), max(1, pairs / tiles));
  }
  p.kb_per_split = cdiv(total_kb, splits);
  p.splits = cdiv(total_kb, p.kb_per_split);
  p.alpha = d->alpha; p.accumulate = d->accumulate;
  p.a_colsum = d->a_colsum; p.a_colsum_scale = d->a_colsum_scale;
  const int grid = 2 * min(p.mt2_tiles * p.n_tiles * p.splits, pairs);
  if (out_bf16) return launch_pair<true, false, true>(ta, tb, tcm, p, grid, st);
  if (d->a_kmajor) return launch_pair<true, false>(ta, tb, tcm, p, grid, st);
  return launch_pair<false, false>(ta, tb, tcm, p, grid, st);
}

}  // namespace eec
