// gemm_tc2.cu -- persistent bf16 tcgen05 GEMM with overlapped, TMA-store-staged epilogues (v2).
//
//   C[M,N] = epi( A(m,k) * B(n,k) )        128 x 256 tile, BLOCK_K = 64, one CTA per SM, static tile loop
//
//   warp 0      : TMA producer (3-stage 128B-swizzled smem ring, runs ahead across tiles)
//   warp 1      : tcgen05.mma issuer; TWO 256-column TMEM accumulators so tile t+1's main loop overlaps
//                 tile t's epilogue
//   warps 2..9  : epilogue; quarter = warp%4 selects the TMEM lane quarter (row = quarter*32 + lane),
//                 half = (warp-2)/4 selects accumulator columns [half*128, half*128+128)
// Epilogue outputs are written 32 columns at a time into swizzled smem staging tiles and leave the SM
// through cp.async.bulk.tensor stores (full-line writes, row clipping by the tensor map); per-channel
// bias lives in smem; SiLU uses tanh.approx (one MUFU op per element).
//
// Epilogue modes
//   EPI_GENERIC    bias, SiLU (+ optional bf16 pre-activation store) / dSiLU (reads bf16 pre-activation),
//                  alpha, fp32 (row-periodic) residual, bf16|fp32 out, or fp32 vector-atomic accumulate (split-K wgrad)
//   EPI_GLU        B rows [n0,n0+128) | [N/2+n0,+128): out = alpha * a * sigmoid(g) (+ optional z store)
//   EPI_LN         N == 256: x = res + alpha*(acc+bias) -> fp32 C; LayerNorm(x) -> ln_out (bf16|fp32), mean/rstd
//   EPI_LOGSOFTMAX N == 256: fp32 log_softmax(acc+bias) (+ argmax, entropy)
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64, NSTAGE = 3;
constexpr int A_BYTES = BM * BK * 2;              // 16 KB
constexpr int B_BYTES = BN * BK * 2;              // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 48 KB
constexpr int STG_BYTES = 16384;                  // one staging tile: 128 rows x (32 fp32 | 32 bf16 [8 KB used])
constexpr int NSTG = 4;                           // 2 per column half
constexpr int OFF_STG = NSTAGE * STAGE_BYTES;
constexpr int OFF_BIAS = OFF_STG + NSTG * STG_BYTES;          // float[2][256]
constexpr int OFF_XCH = OFF_BIAS + 2 * 256 * 4;               // float[2 parity][3 slots][2 halves][128 rows][2]
constexpr int OFF_BAR = OFF_XCH + 2 * 3 * 2 * 128 * 2 * 4;
constexpr int SMEM2_BYTES = OFF_BAR + 256 + 1024;
constexpr int NT2 = 320;

enum { EPI_GENERIC = 0, EPI_GLU = 1, EPI_LN = 2, EPI_LOGSOFTMAX = 3 };

struct P2 {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_per_split;
  const float* bias;
  int act;
  const void* preact_in; int ldp;   // DSILU: bf16 [M,N]
  float alpha;
  const float* residual; int ldr; int res_row_mod;
  int out_bf16, has_c, has_pre, has_ln, ln_bf16;
  float* c_acc; int ldc;            // accumulate target (fp32)
  int accumulate;
  const float* ln_gamma; const float* ln_beta; float* ln_mean; float* ln_rstd;
  int32_t* argmax; float* entropy;
  int debug;   // EEC_GEMM_DEBUG bitmask (perf triage only): 1 = skip bulk store issue, 2 = skip staging entirely, 4 = skip activation math, 8 = no operand TMA, 16 = no MMA issue
};

template <bool A_KMAJ, bool B_KMAJ, int EPI, int CS>
__global__ void __launch_bounds__(NT2, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          const __grid_constant__ CUtensorMap tmC,   // main output
                                                          const __grid_constant__ CUtensorMap tmP,   // pre-activation store
                                                          const __grid_constant__ CUtensorMap tmL,   // LayerNorm output
                                                          const P2 p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tfull_bar = empty_bar + NSTAGE;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work units: (m super-tile of CS m-tiles, n tile, k split); CTA `rank` of a cluster owns m-tile msuper*CS + rank
  const int rank = (CS > 1) ? (int)cluster_ctarank() : 0;
  const int cid = blockIdx.x / CS, n_clusters = gridDim.x / CS;
  const int m_super = (p.m_tiles + CS - 1) / CS;
  const int tiles = m_super * p.n_tiles;
  const int n_units = tiles * p.splits;
  const int total_kb = (p.K + BK - 1) / BK;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CS) - 1u);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CS); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // global k-block counter (stage ring position)
      for (int u = cid; u < n_units; u += n_clusters) {
        const int tile = u % tiles, split = u / tiles;
        const int m0 = ((tile / p.n_tiles) * CS + rank) * BM;
        const int nt = tile % p.n_tiles;
        const int n0 = (EPI == EPI_GLU) ? nt * 128 : nt * BN;
        const int kb0 = split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(&empty_bar[s], ((it / NSTAGE) & 1) ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (p.debug & 8) { mbar_arrive(&full_bar[s]); continue; }
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          const int k = kb * BK;
          if (A_KMAJ) {
            tma_load_2d(sa, &tmA, &full_bar[s], k, m0);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[s], m0, k);
            tma_load_2d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k);
          }
          if (CS == 1) {
            if (B_KMAJ) {
              if (EPI == EPI_GLU) {
                tma_load_2d(sb, &tmB, &full_bar[s], k, n0);
                tma_load_2d(sb + 16384, &tmB, &full_bar[s], k, p.N / 2 + n0);
              } else {
                tma_load_2d(sb, &tmB, &full_bar[s], k, n0);
              }
            } else {
#pragma unroll
              for (int a = 0; a < BN / 64; ++a) tma_load_2d(sb + a * 8192, &tmB, &full_bar[s], n0 + a * 64, k);
            }
          } else {
            // this CTA fetches slice `rank` of the B tile and multicasts it to every CTA of the cluster
            constexpr int SLICE = B_BYTES / CS;
            uint8_t* dst = sb + rank * SLICE;
            if (B_KMAJ) {
              if (EPI == EPI_GLU) {   // CS == 2: rank 0 -> the 128 "a" rows, rank 1 -> the 128 gate rows
                tma_load_2d_mc(dst, &tmB, &full_bar[s], k, (rank == 0) ? n0 : p.N / 2 + n0, MC_MASK);
              } else {
                tma_load_2d_mc(dst, &tmB, &full_bar[s], k, n0 + rank * (BN / CS), MC_MASK);   // box {64 k, 256/CS n}
              }
            } else {
#pragma unroll
              for (int a = 0; a < BN / 64 / CS; ++a)
                tma_load_2d_mc(dst + a * 8192, &tmB, &full_bar[s], n0 + (rank * (BN / 64 / CS) + a) * 64, k, MC_MASK);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, !A_KMAJ, !B_KMAJ);
      uint32_t it = 0, ut = 0;
      for (int u = cid; u < n_units; u += n_clusters, ++ut) {
        const int split = u / tiles;
        const int kb0 = split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const uint32_t acc = ut & 1;
        mbar_wait(&tempty_bar[acc], ((ut >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(&full_bar[s], (it / NSTAGE) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          if (!(p.debug & 16))
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = A_KMAJ ? make_smem_desc(sa + k * 32, 0, 1024) : make_smem_desc(sa + k * 2048, 8192, 1024);
            const uint64_t bd = B_KMAJ ? make_smem_desc(sb + k * 32, 0, 1024) : make_smem_desc(sb + k * 2048, 8192, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CS == 1) umma_commit(&empty_bar[s]);
          else umma_commit_mc(&empty_bar[s], MC_MASK);   // frees this stage in every CTA that multicasts into it
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ================================================================== epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;   // 0..255
    Stager st;
    st.buf[0] = smem + OFF_STG + (half * 2) * STG_BYTES;
    st.buf[1] = smem + OFF_STG + (half * 2 + 1) * STG_BYTES;
    st.next = 0;
    st.bar_id = 1 + half;
    st.leader = (et == half * 128);
    st.r = r;
    st.debug = p.debug;
    float v[32];
    uint32_t ut = 0;
    for (int u = cid; u < n_units; u += n_clusters, ++ut) {
      const int tile = u % tiles, split = u / tiles;
      const int m0 = ((tile / p.n_tiles) * CS + rank) * BM;
      const int nt = tile % p.n_tiles;
      const int n0 = (EPI == EPI_GLU) ? nt * 128 : nt * BN;
      const uint32_t acc = ut & 1;
      const int m = m0 + r;
      const bool valid = m < p.M;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      float* bs = bias_s + acc * 256;
      // stage the bias slice of this tile (fp32) in smem
      if (p.bias) {
        if (EPI == EPI_GLU) bs[et] = (et < 128) ? p.bias[n0 + et] : p.bias[p.N / 2 + n0 + (et - 128)];
        else bs[et] = (n0 + et < p.N) ? p.bias[n0 + et] : 0.f;
      } else {
        bs[et] = 0.f;
      }
      bar_sync(3, 256);
      mbar_wait(&tfull_bar[acc], (ut >> 1) & 1);
      tc_fence_after();
      const bool first_split = (split == 0);

      if (EPI == EPI_GENERIC) {
        const long rr = p.res_row_mod ? (m % p.res_row_mod) : m;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = half * 128 + cc * 32;
          const int n = n0 + c0;
          if (n >= p.N) break;  // uniform across the half
          tmem_ld32(trow + c0, v);
          if (first_split) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bs[c0 + j];
          }
          if (p.act == EEC_ACT_SILU) {
            if (p.has_pre) st.store(&tmP, n, m0, v, true);
            if (!(p.debug & 4)) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= sigmoid_fast(v[j]);
            }
          } else if (p.act == EEC_ACT_DSILU) {
            if (valid) {
              const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(p.preact_in) + (long)m * p.ldp + n;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float h8[8];
                ld8<__nv_bfloat16>(hp + g * 8, h8);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float s = sigmoid_fast(h8[j]);
                  v[g * 8 + j] *= s * fmaf(h8[j], 1.0f - s, 1.0f);
                }
              }
            }
          }
          if (p.alpha != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
          }
          if (p.residual && first_split && valid) {
            const float* rp = p.residual + rr * p.ldr + n;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 f = *reinterpret_cast<const float4*>(rp + g * 4);
              v[g * 4] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
            }
          }
          if (p.accumulate) {
            if (valid) {
              float* cp = p.c_acc + (long)m * p.ldc + n;
#pragma unroll
              for (int g = 0; g < 8; ++g) red_add_v4(cp + g * 4, v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
            }
          } else {
            st.store(&tmC, n, m0, v, p.out_bf16);
          }
        }
      } else if (EPI == EPI_GLU) {
        float g[32];
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = half * 64 + cc * 32;   // output channel offset within the 128-channel tile
          tmem_ld32(trow + c0, v);
          tmem_ld32(trow + 128 + c0, g);
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] += bs[c0 + j]; g[j] += bs[128 + c0 + j]; }
          if (p.has_pre) {
            st.store(&tmP, n0 + c0, m0, v, true);
            st.store(&tmP, p.N / 2 + n0 + c0, m0, g, true);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = p.alpha * v[j] * sigmoid_fast(g[j]);
          st.store(&tmC, n0 + c0, m0, v, p.out_bf16);
        }
      } else if (EPI == EPI_LN) {
        const long rr = p.res_row_mod ? (m % p.res_row_mod) : m;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = half * 128 + cc * 32;
          tmem_ld32(trow + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (v[j] + bs[c0 + j]) * p.alpha;
          if (p.residual && valid) {
            const float* rp = p.residual + rr * p.ldr + c0;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 f = *reinterpret_cast<const float4*>(rp + g * 4);
              v[g * 4] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
          tmem_st32(trow + c0, v);                 // keep x in TMEM for the normalisation pass
          st.store(&tmC, c0, m0, v, false);        // fp32 residual stream
        }
        float* slot = xch + (((ut & 1) * 3 + 0) * 2) * 256;
        slot[half * 256 + r * 2] = s1;
        slot[half * 256 + r * 2 + 1] = s2;
        bar_sync(3, 256);
        s1 += slot[(half ^ 1) * 256 + r * 2];
        s2 += slot[(half ^ 1) * 256 + r * 2 + 1];
        const float mu = s1 * (1.f / 256.f);
        const float rs = rsqrtf(fmaxf(s2 * (1.f / 256.f) - mu * mu, 0.f) + 1e-5f);
        if (half == 0 && valid && p.ln_mean) { p.ln_mean[m] = mu; p.ln_rstd[m] = rs; }
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = half * 128 + cc * 32;
          tmem_ld32(trow + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (v[j] - mu) * rs * __ldg(p.ln_gamma + c0 + j) + __ldg(p.ln_beta + c0 + j);
          st.store(&tmL, c0, m0, v, p.ln_bf16);
        }
      } else {  // EPI_LOGSOFTMAX
        float mx = -INFINITY;
        int mi = 0;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = half * 128 + cc * 32;
          tmem_ld32(trow + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float t = v[j] + bs[c0 + j];
            if (t > mx) { mx = t; mi = c0 + j; }
          }
        }
        float* slot0 = xch + (((ut & 1) * 3 + 0) * 2) * 256;
        slot0[half * 256 + r * 2] = mx;
        slot0[half * 256 + r * 2 + 1] = __int_as_float(mi);
        bar_sync(3, 256);
        {
          const float omx = slot0[(half ^ 1) * 256 + r * 2];
          const int omi = __float_as_int(slot0[(half ^ 1) * 256 + r * 2 + 1]);
          if (omx > mx || (omx == mx && omi < mi)) { mx = omx; mi = omi; }
        }
        float se = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = half * 128 + cc * 32;
          tmem_ld32(trow + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) se += __expf(v[j] + bs[c0 + j] - mx);
        }
        float* slot1 = xch + (((ut & 1) * 3 + 1) * 2) * 256;
        slot1[half * 256 + r * 2] = se;
        bar_sync(3, 256);
        se += slot1[(half ^ 1) * 256 + r * 2];
        const float lse = mx + __logf(se);
        float h = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = half * 128 + cc * 32;
          tmem_ld32(trow + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = v[j] + bs[c0 + j] - lse;
            h -= __expf(v[j]) * v[j];
          }
          st.store(&tmC, c0, m0, v, false);
        }
        if (p.entropy) {
          float* slot2 = xch + (((ut & 1) * 3 + 2) * 2) * 256;
          slot2[half * 256 + r * 2] = h;
          bar_sync(3, 256);
          h += slot2[(half ^ 1) * 256 + r * 2];
          if (half == 0 && valid) p.entropy[m] = h;
        }
        if (half == 0 && valid && p.argmax) p.argmax[m] = mi;
      }
      // accumulator drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (st.leader) bulk_wait_all();
    tc_fence_before();
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // no CTA exits while a peer can still multicast into its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool AK, bool BK_, int EPI, int CS>
int launch2cs(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const CUtensorMap& tp, const CUtensorMap& tl,
              const P2& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<AK, BK_, EPI, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NT2);
  cfg.dynamicSmemBytes = SMEM2_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  count_launch();
  EEC_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<AK, BK_, EPI, CS>, ta, tb, tc_, tp, tl, p));
  return 0;
}

template <bool AK, bool BK_, int EPI>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const CUtensorMap& tp, const CUtensorMap& tl,
            const P2& p, int grid, cudaStream_t st, int cs) {
  if (cs == 2) return launch2cs<AK, BK_, EPI, 2>(ta, tb, tc_, tp, tl, p, grid, st);
  return launch2cs<AK, BK_, EPI, 1>(ta, tb, tc_, tp, tl, p, grid, st);
}

int g_num_sms = 0;

}  // namespace

int gemm_tc2(const eec_gemm_desc* d, cudaStream_t st, int32_t* argmax, float* entropy, int logsoftmax) {
  EEC_CHECK_ARG(d->in_dtype == EEC_BF16, "gemm_tc2: operands must be bf16");
  EEC_CHECK_ARG(d->N % 32 == 0, "gemm_tc2: N (%d) must be a multiple of 32", d->N);
  EEC_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "gemm_tc2: empty problem %dx%dx%d", d->M, d->N, d->K);
  int epi = EPI_GENERIC;
  if (logsoftmax) epi = EPI_LOGSOFTMAX;
  else if (d->act == EEC_ACT_GLU) epi = EPI_GLU;
  else if (d->ln_out) epi = EPI_LN;
  EEC_CHECK_ARG(d->ln2_gamma == nullptr, "gemm_tc2: chained second LayerNorm is only available on the fp32 path");
  if (epi == EPI_LN || epi == EPI_LOGSOFTMAX) {
    EEC_CHECK_ARG(d->N == 256, "gemm_tc2: row-wise epilogue needs N == 256 (got %d)", d->N);
    EEC_CHECK_ARG(d->act == EEC_ACT_NONE && !d->accumulate, "gemm_tc2: row-wise epilogue with act/accumulate unsupported");
    EEC_CHECK_ARG(d->out_dtype == EEC_F32 && d->ldc == 256, "gemm_tc2: row-wise epilogue writes fp32 C with ldc 256");
    if (epi == EPI_LOGSOFTMAX) EEC_CHECK_ARG(d->bias != nullptr, "gemm_tc2: logsoftmax epilogue needs bias");
    EEC_CHECK_ARG(d->a_kmajor && d->b_kmajor, "gemm_tc2: row-wise epilogue needs K-major operands");
  }
  if (epi == EPI_GLU) {
    EEC_CHECK_ARG(d->N % 256 == 0 && d->b_kmajor && d->a_kmajor, "gemm_tc2: GLU needs N %% 256 == 0 and K-major operands");
    EEC_CHECK_ARG(!d->residual && !d->accumulate, "gemm_tc2: GLU with residual/accumulate unsupported");
  }
  if (d->act == EEC_ACT_DSILU || d->act == EEC_ACT_DRELU) EEC_CHECK_ARG(d->preact != nullptr && d->preact_dtype == EEC_BF16, "gemm_tc2: DSILU / DRELU need a bf16 preact");
  if (d->act == EEC_ACT_RELU || d->act == EEC_ACT_DRELU) EEC_CHECK_ARG(d->out_dtype == EEC_BF16 && !d->accumulate, "gemm (tcgen05): RELU / DRELU epilogues write bf16, no accumulate");
  if (d->act == EEC_ACT_SILU && d->preact) EEC_CHECK_ARG(d->preact_dtype == EEC_BF16, "gemm_tc2: preact store must be bf16");
  if (d->accumulate) EEC_CHECK_ARG(d->out_dtype == EEC_F32 && d->act == EEC_ACT_NONE, "gemm_tc2: accumulate needs fp32 C, no act");

  {
    // GENERIC / GLU epilogues run on the v3 kernel (16 epilogue warps, warp-private staging); EEC_GEMM_V2=1 keeps v2 for A/B runs
    static int v2_env = -1;
    if (v2_env < 0) { const char* e = getenv("EEC_GEMM_V2"); v2_env = (e && e[0] == '1') ? 1 : 0; }
    // (v3 stages dSiLU / SiLU+pre-activation outputs in bf16 boxes only; the fp32-output variants of those stay here)
    const bool v3_ok = d->out_dtype == EEC_BF16 || !(d->act == EEC_ACT_DSILU || d->act == EEC_ACT_DRELU || (d->act == EEC_ACT_SILU && d->preact));
    if (!v2_env && v3_ok && (epi == EPI_GENERIC || epi == EPI_GLU)) return gemm_tc3(d, st);
    if (!v2_env && epi == EPI_LN && d->res_row_mod == 0 && (!d->residual || d->ldr == 256)) return gemm_ln3(d, st);
  }
  EEC_CHECK_ARG(!d->a_colsum, "gemm_tc2: a_colsum is implemented by the v3 kernel only (unset EEC_GEMM_V2)");
  EEC_CHECK_ARG(!(d->drop_state && d->drop_p > 0.f), "gemm_tc2: dropout is implemented by the v3 kernels only (this descriptor routes to v2)");
  if (!g_num_sms) {
    int dev = 0;
    EEC_CUDA(cudaGetDevice(&dev));
    EEC_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // cluster of 2 CTAs along M (B-tile multicast) whenever there are at least two m-tiles
  static int cs_env = -1;
  if (cs_env < 0) { const char* e = getenv("EEC_GEMM_CLUSTER"); cs_env = e ? atoi(e) : 2; }
  const int cs = (cs_env >= 2 && cdiv(d->M, BM) >= 2) ? 2 : 1;
  CUtensorMap ta, tb, tcm, tpm, tlm;
  if (d->a_kmajor) { if (int r = get_tmap_2d(&ta, d->A, d->K, d->M, (uint64_t)d->lda * 2, 64, 128)) return r; }
  else { if (int r = get_tmap_2d(&ta, d->A, d->M, d->K, (uint64_t)d->lda * 2, 64, 64)) return r; }
  if (d->b_kmajor) {
    if (int r = get_tmap_2d(&tb, d->B, d->K, d->N, (uint64_t)d->ldb * 2, 64, (epi == EPI_GLU || cs == 2) ? 128 : 256)) return r;
  } else {
    if (int r = get_tmap_2d(&tb, d->B, d->N, d->K, (uint64_t)d->ldb * 2, 64, 64)) return r;
  }
  const int n_out = (epi == EPI_GLU) ? d->N / 2 : d->N;
  tcm = ta; tpm = ta; tlm = ta;  // placeholders when unused
  const bool out_bf16 = d->out_dtype == EEC_BF16;
  if (!d->accumulate) {
    if (int r = get_tmap_store(&tcm, d->C, out_bf16, (uint64_t)n_out, (uint64_t)d->M, (uint64_t)d->ldc)) return r;
  }
  const bool store_pre = d->preact && (d->act == EEC_ACT_SILU || epi == EPI_GLU);
  if (store_pre) { if (int r = get_tmap_store(&tpm, d->preact, true, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldp)) return r; }
  if (epi == EPI_LN) {
    if (int r = get_tmap_store(&tlm, d->ln_out, d->ln_dtype == EEC_BF16, 256, (uint64_t)d->M, (uint64_t)d->ld_ln)) return r;
  }
  P2 p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.m_tiles = cdiv(d->M, BM);
  p.n_tiles = (epi == EPI_GLU) ? d->N / 256 : cdiv(d->N, BN);
  const int total_kb = cdiv(d->K, BK);
  int splits = 1;
  if (d->accumulate) {
    const int tiles = p.m_tiles * p.n_tiles;
    if (tiles < g_num_sms && total_kb >= 16) splits = min(cdiv(total_kb, 8), max(1, g_num_sms / tiles));
  }
  p.kb_per_split = cdiv(total_kb, splits);
  p.splits = cdiv(total_kb, p.kb_per_split);
  p.bias = d->bias; p.act = d->act; p.preact_in = d->preact; p.ldp = d->ldp; p.alpha = d->alpha;
  p.residual = d->residual; p.ldr = d->ldr; p.res_row_mod = d->res_row_mod;
  p.out_bf16 = out_bf16; p.has_pre = store_pre; p.ln_bf16 = d->ln_dtype == EEC_BF16;
  p.c_acc = reinterpret_cast<float*>(d->C); p.ldc = d->ldc; p.accumulate = d->accumulate;
  p.ln_gamma = d->ln_gamma; p.ln_beta = d->ln_beta; p.ln_mean = d->ln_mean; p.ln_rstd = d->ln_rstd;
  p.argmax = argmax; p.entropy = entropy;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("EEC_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.debug = dbg;
  const int n_units = cdiv(p.m_tiles, cs) * p.n_tiles * p.splits;
  const int grid = min(n_units, g_num_sms / cs) * cs;

#define EEC_TC2_DISPATCH(AK, BK_)                                                                       \
  switch (epi) {                                                                                        \
    case EPI_GENERIC: return launch2<AK, BK_, EPI_GENERIC>(ta, tb, tcm, tpm, tlm, p, grid, st, cs);          \
    default: break;                                                                                     \
  }
  if (epi == EPI_GLU) return launch2<true, true, EPI_GLU>(ta, tb, tcm, tpm, tlm, p, grid, st, cs);
  if (epi == EPI_LN) return launch2<true, true, EPI_LN>(ta, tb, tcm, tpm, tlm, p, grid, st, cs);
  if (epi == EPI_LOGSOFTMAX) return launch2<true, true, EPI_LOGSOFTMAX>(ta, tb, tcm, tpm, tlm, p, grid, st, cs);
  if (d->a_kmajor && d->b_kmajor) { EEC_TC2_DISPATCH(true, true) }
  else if (d->a_kmajor && !d->b_kmajor) { EEC_TC2_DISPATCH(true, false) }
  else if (!d->a_kmajor && !d->b_kmajor) { EEC_TC2_DISPATCH(false, false) }
  else { EEC_TC2_DISPATCH(false, true) }
  set_error("gemm_tc2: unsupported configuration");
  return 1;
}

}  // namespace eec
