// attn_bwd_tc.cu -- self-attention backward on the tcgen05 tensor cores (bf16 operands, fp32 math).
//
// One CTA per (128-key block j, head, utterance); it keeps K_j, V_j resident and walks the query blocks i:
//   for each 64-key half hk of the block (S / dP are single-buffered 64-column tiles so that the CTA needs only 256 TMEM
//   columns and TWO CTAs share an SM: one CTA's softmax math overlaps the other's MMAs and barrier latencies):
//     S  = Q_i K_j[hk]^T      (M=128 q, N=64 k, K=32)         -> TMEM [0,64)
//     dP = dO_i V_j[hk]^T     (M=128 q, N=64 k, K=32)         -> TMEM [64,128)
//     8 softmax warps: P = exp2(S*c - lse_i), dS = P (dP - D_i) -> bf16 P / dS tiles in 128B-swizzled smem
//   dV_j += P^T  dO_i       (M=128 k, N=32, K=128 q; A = P  read MN-major, B = dO_i MN-major) -> TMEM [128,160)
//   dK_j += dS^T Q_i        (M=128 k, N=32, K=128 q; A = dS read MN-major, B = Q_i  MN-major) -> TMEM [160,192)
//   dQ_i  = dS  K_j         (M=128 q, N=32, K=128 k; A = dS K-major,       B = K_j  MN-major) -> TMEM [192,224)
// dQ partials of the key blocks are summed with fp32 vector atomics into a workspace and converted
// (with the 1/sqrt(dh) scale) by a small tail kernel; dK/dV are written directly by the owning CTA.
// The same smem P/dS tile is consumed both K-major (dQ) and MN-major (dV, dK): no transposes.
// D_i = dO_i . O_i comes from a tiny pre-pass.  Reference: autograd of nn.MultiheadAttention (TA:194-200).
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int BT = 128;       // block of queries / keys
constexpr int DHD = 32;
constexpr int TILE_QD = BT * DHD * 2;      // 8 KB  ([128][32] bf16)
constexpr int TILE_P = BT * BT * 2;        // 32 KB
constexpr int ST = 2;                      // Q/dO stages
constexpr int BW_SMEM = 2 * TILE_QD /*K,V*/ + ST * 2 * TILE_QD /*Q,dO*/ + 2 * TILE_P /*P,dS*/ + 256;   // 2 CTAs/SM: 2 x (112.25 KB + 1 KB reserved) <= 228 KB
constexpr int BW_THREADS = 320;           // TMA warp, MMA warp, 8 softmax warps (2 per TMEM lane quarter: 64 of the 128 keys each)
constexpr int N_MATH = 256;
constexpr uint32_t C_S = 0, C_DP = 64, C_DV = 128, C_DK = 160, C_DQ = 192;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t SW64 = 4, SW128 = 2;

// -DEEC_ATTN_BWD_TIMELINE (make attn_bwd_timeline): CTA (0,0,0) stamps clock64 at the hand-overs of its first query blocks and prints them
#ifdef EEC_ATTN_BWD_TIMELINE
__device__ long long g_btl[96];
#define BTL(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_btl[slot] = clock64(); } while (0)
#else
#define BTL(slot) do { } while (0)
#endif

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// D[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d]      (one warp per frame row: lane owns 8 channels = 1/4 head); the same pass zeroes the row
// of the fp32 dQ workspace the main kernel accumulates into (it used to be a separate memset node per layer)
__global__ void __launch_bounds__(256) attn_dvec_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int ldo,
                                                        float* __restrict__ dvec, float* __restrict__ dq32, int B, int T, int H) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B * T) return;
  {
    float4* z = reinterpret_cast<float4*>(dq32 + (long)row * 256 + lane * 8);
    z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float a[8], g[8];
  ld8<__nv_bfloat16>(o + (long)row * ldo + lane * 8, a);
  ld8<__nv_bfloat16>(d_o + (long)row * ldo + lane * 8, g);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(a[i], g[i], s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if ((lane & 3) == 0) {
    const int b = row / T, t = row % T, h = lane >> 2;
    dvec[((long)b * H + h) * T + t] = s;
  }
}

// dq[:, 0:256] (row pitch lddq) = bf16(scale * dq32)
__global__ void dq_convert_kernel(const float* __restrict__ dq32, __nv_bfloat16* __restrict__ dqkv, int lddq, long rows, float scale) {
  pdl_trigger();
  pdl_wait();
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= rows * 256) return;
  long r = i / 256;
  int c = (int)(i % 256);
  float v[8];
  ld8<float>(dq32 + i, v);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] *= scale;
  st8<__nv_bfloat16>(dqkv + r * lddq + c, v);
}

// geometry of a call: see attn_tc.cu (self-attention over the packed projection, causal decoder self-attention, cross-attention)
struct BwGeom {
  int Tq, Tk, H;
  int q_col, k_col, v_col;
  const int32_t* key_len;
  const uint32_t* key_bits;
  int causal;
  __nv_bfloat16* dk; int lddk;      // head 0 of dK / dV at these pointers
  __nv_bfloat16* dv; int lddv;
};

template <bool DROP, bool GENERAL>
__global__ void __launch_bounds__(BW_THREADS, 2) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                 const __grid_constant__ CUtensorMap tm_kv,
                                                                 const __grid_constant__ CUtensorMap tm_do, const BwGeom g,
                                                                 const float* __restrict__ lse,
                                                                 const float* __restrict__ dvec, float* __restrict__ dq32,
                                                                 const DropArgs drop) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) BTL(0);
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) { printf("eec: attn_bwd smem base not 1024-aligned\n"); __trap(); }
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE_QD;
  uint8_t* sQ = sV + TILE_QD;                 // [ST]
  uint8_t* sdO = sQ + ST * TILE_QD;           // [ST]
  uint8_t* sP = sdO + ST * TILE_QD;
  uint8_t* sdS = sP + TILE_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + TILE_P);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;     // [2]
  uint64_t* qdo_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* sdp_free = bars + 6;     // N_MATH arrivals
  uint64_t* pds_full = bars + 7;     // P / dS of key half 0 written (one phase per query block)
  uint64_t* pds_full1 = bars + 10;   // ... of key half 1 (two barriers: a parity wait on ONE barrier with two phases per block can alias)
  uint64_t* dq_full = bars + 8;
  uint64_t* acc_full = bars + 9;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BT;
  const int H = g.H, Tq = g.Tq, Tk = g.Tk;
  const int klen = g.key_len ? min(g.key_len[b], Tk) : Tk;
  const int D = H * DHD;
  const int rowq = b * Tq, rowk = b * Tk;
  const int nq = (Tq + BT - 1) / BT;
  const int i_begin = (GENERAL && g.causal) ? k0 / BT : 0;     // queries before the block's first key see none of its keys
  const bool active = k0 < klen && i_begin < nq;   // block has at least one key some query may see

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_do);
    mbar_init(kv_full, 1);
    for (int s = 0; s < ST; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, N_MATH / 32);   // one arrival per softmax warp (after __syncwarp): 256 per-thread arrivals on one mbarrier serialise
    mbar_init(pds_full, N_MATH / 32);
    mbar_init(pds_full1, N_MATH / 32);
    mbar_init(dq_full, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) BTL(1);

  if (warp == 0) {
    if (lane == 0 && active) {
      mbar_expect_tx(kv_full, 2 * TILE_QD);
      tma_load_2d(sK, &tm_kv, kv_full, g.k_col + h * DHD, rowk + k0);
      tma_load_2d(sV, &tm_kv, kv_full, g.v_col + h * DHD, rowk + k0);
      for (int i = 0; i < nq - i_begin; ++i) {        // i counts the query blocks this CTA visits (barrier phases); block index i_begin + i
        const int s = i % ST;
        mbar_wait(&qdo_empty[s], ((i / ST) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[s], 2 * TILE_QD);
        tma_load_2d(sQ + s * TILE_QD, &tm_q, &qdo_full[s], g.q_col + h * DHD, rowq + (i_begin + i) * BT);
        tma_load_2d(sdO + s * TILE_QD, &tm_do, &qdo_full[s], h * DHD, rowq + (i_begin + i) * BT);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && active) {
      constexpr uint32_t id_s = make_idesc_bf16(BT, 64, false, false);    // S, dP (64-key half)
      constexpr uint32_t id_t = make_idesc_bf16(BT, DHD, true, true);     // dV, dK : A^T (MN-major), B MN-major
      constexpr uint32_t id_q = make_idesc_bf16(BT, DHD, false, true);    // dQ     : A K-major,      B MN-major
      mbar_wait(kv_full, 0);
      BTL(2);
      const uint32_t ak = smem_u32(sK), av = smem_u32(sV), ap = smem_u32(sP), ads = smem_u32(sdS);
      const int cnt = nq - i_begin;
      // UMMA descriptors built ONCE (the single issuing thread spent ~70 clk per MMA re-deriving them: 1.7 k clk for the 24 MMAs of a block);
      // per MMA only a 16-byte-unit offset is added to the address field
      const uint64_t d_k64 = make_smem_desc(ak, 0, 512, SW64), d_v64 = make_smem_desc(av, 0, 512, SW64);
      const uint64_t d_pT = make_smem_desc(ap, 16384, 1024, SW128), d_dsT = make_smem_desc(ads, 16384, 1024, SW128);   // MN-major reads (dV, dK)
      const uint64_t d_dsK = make_smem_desc(ads, 0, 1024, SW128);                                                       // K-major read (dQ)
      uint64_t d_q64[ST], d_do64[ST];
#pragma unroll
      for (int s = 0; s < ST; ++s) {
        d_q64[s] = make_smem_desc(smem_u32(sQ + s * TILE_QD), 0, 512, SW64);
        d_do64[s] = make_smem_desc(smem_u32(sdO + s * TILE_QD), 0, 512, SW64);
      }
      // S / dP of one 64-key half of query block i: TMEM [0,64) / [64,128), single-buffered (the softmax warps hand them back through sdp_free)
      auto issue_sdp = [&](int i, int hk) {
        const uint64_t dq = (i % ST) ? d_q64[1] : d_q64[0], ddo = (i % ST) ? d_do64[1] : d_do64[0];
        const int n = 2 * i + hk;                    // half-step counter: phases of sdp_full / sdp_free / pds_full
        if (n > 0) mbar_wait(sdp_free, (n - 1) & 1);  // the softmax warps have read the previous S / dP out of TMEM
        tc_fence_after();
        const uint64_t koff = (uint64_t)((hk * 64 * 64) >> 4);          // 64 key rows of 64 B
#pragma unroll
        for (int k = 0; k < DHD / 16; ++k) umma_bf16(tmem_base + C_S, dq + k * 2, d_k64 + koff + k * 2, id_s, k);
#pragma unroll
        for (int k = 0; k < DHD / 16; ++k) umma_bf16(tmem_base + C_DP, ddo + k * 2, d_v64 + koff + k * 2, id_s, k);
        umma_commit(sdp_full);
      };
      if (cnt > 0) {
        mbar_wait(&qdo_full[0], 0);
        BTL(8);
        issue_sdp(0, 0);
        BTL(9);
      }
      for (int i = 0; i < cnt; ++i) {
        const int s = i % ST;
        const uint64_t dq = s ? d_q64[1] : d_q64[0], ddo = s ? d_do64[1] : d_do64[0];
        issue_sdp(i, 1);
        BTL(8 + i * 8 + 2);
        // the first half of the NEXT query block before this block's dV / dK / dQ: its softmax math then overlaps those 24 MMAs
        // (measured before: S ready 2.8 k clk after the hand-over, the whole chain of a block strictly serial: 8 k clk per block)
        if (i + 1 < cnt) {
          mbar_wait(&qdo_full[(i + 1) % ST], ((i + 1) / ST) & 1);
          BTL(8 + (i + 1) * 8 + 0);
          issue_sdp(i + 1, 0);
          BTL(8 + (i + 1) * 8 + 1);
        }
        mbar_wait(pds_full, i & 1);    // both halves' P / dS of block i are in shared memory
        mbar_wait(pds_full1, i & 1);
        BTL(8 + i * 8 + 3);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BT / 16; ++k) {  // contraction over the 128 queries of block i
          umma_bf16(tmem_base + C_DV, d_pT + k * 128, ddo + k * 64, id_t, (i > 0 || k > 0) ? 1u : 0u);
          umma_bf16(tmem_base + C_DK, d_dsT + k * 128, dq + k * 64, id_t, (i > 0 || k > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)    // contraction over the 128 keys of block j
          umma_bf16(tmem_base + C_DQ, d_dsK + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2), d_k64 + k * 64, id_q, k > 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(&qdo_empty[s]);
        BTL(8 + i * 8 + 4);
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;          // 32 of the 64 key columns of a half-step
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    const float sc2 = rsqrtf((float)DHD) * LOG2E;
    float s[32], dp[32];
    // dropout on the probabilities (DROP): with M = mask * scale regenerated from the forward's counters,
    //   dV uses P.M (the tile written to sP), dS = P (dP.M - D) and D = dO . O is unchanged (O = (P.M) V).
    const uint32_t* dbits = reinterpret_cast<const uint32_t*>(drop.bits);   // keep-mask words of the forward (eec_dropout_bits, W = 32)
    const long dR = (long)gridDim.z * H * Tq;
    const uint32_t* kbits = GENERAL && g.key_bits ? g.key_bits + (long)b * ((Tk + 31) >> 5) : nullptr;
    if (active) {
      // this row's log-sum-exp and dO.O of the NEXT query block are loaded one iteration ahead (they sit on the critical path otherwise:
      // ~10 % of the kernel's stall samples in round 1 were the L2 round trip of these two loads)
      const float* lse_bh = lse + ((long)b * H + h) * Tq;
      const float* dv_bh = dvec + ((long)b * H + h) * Tq;
      const int t_first = i_begin * BT + r;
      float l2_next = t_first < Tq ? lse_bh[t_first] : 0.f;
      float di_next = t_first < Tq ? dv_bh[t_first] : 0.f;
      // dQ partial of (query block ii, this key block): 128 rows x 32 fp32 in TMEM; warp (q, half) adds columns [16 half, +16) of its 32 rows
      auto drain_dq = [&](int ii) {
        const int tp = (i_begin + ii) * BT + r;
        uint32_t dqv[16];
        tc_fence_after();
        tmem_ld16_async(trow + C_DQ + half * 16, dqv);
        tmem_ld_wait16(dqv);
        if (tp < Tq) {
          float* dst = dq32 + ((long)(rowq + tp)) * D + h * DHD + half * 16;
#pragma unroll
          for (int gg = 0; gg < 4; ++gg)
            red_add_v4(dst + gg * 4, __uint_as_float(dqv[gg * 4]), __uint_as_float(dqv[gg * 4 + 1]), __uint_as_float(dqv[gg * 4 + 2]), __uint_as_float(dqv[gg * 4 + 3]));
        }
        tc_fence_before();
      };
      for (int i = 0; i < nq - i_begin; ++i) {
        const int t = (i_begin + i) * BT + r;
        const bool rvalid = t < Tq;
        const long drow = (long)(b * H + h) * Tq + t;
        const float l2 = l2_next * LOG2E;
        const float di = di_next;
        if (t + BT < Tq) { l2_next = lse_bh[t + BT]; di_next = dv_bh[t + BT]; } else { l2_next = 0.f; di_next = 0.f; }
#pragma unroll 1
        for (int hk = 0; hk < 2; ++hk) {
          const int n = 2 * i + hk;
          const int c0 = hk * 64 + half * 32;    // first key column (within the block) of this warp's slab
          uint32_t dword = 0u;                    // keep-mask word of (row t, these 32 keys): in flight while S / dP are being computed
          if (DROP && rvalid && k0 + c0 < klen) dword = dbits[(long)((k0 + c0) >> 5) * dR + drow];
          mbar_wait(sdp_full, n & 1);
          if (threadIdx.x == 64) BTL(40 + i * 16 + hk * 6 + 0);
          tc_fence_after();
          tmem_ld32x2(trow + C_S + half * 32, trow + C_DP + half * 32, s, dp);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sdp_free);   // S / dP are in registers: the MMA warp may start the next half
          if (threadIdx.x == 64) BTL(40 + i * 16 + hk * 6 + 1);
          if (GENERAL) {
            // visibility word of (row t, these 32 keys): length, causal and per-key validity masks
            const int base = k0 + c0;
            uint32_t mk = !rvalid ? 0u : (base + 32 <= klen) ? ~0u : (base >= klen ? 0u : ((1u << (klen - base)) - 1u));
            if (g.causal) {
              const int lim = t - base;
              mk &= lim >= 31 ? ~0u : (lim < 0 ? 0u : ((2u << lim) - 1u));
            }
            if (kbits && base < Tk) mk &= kbits[base >> 5];
            const bool lfin = l2 > -INFINITY;         // a row without any visible key has lse = -inf: its probabilities are 0
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const bool ok = ((mk >> e) & 1u) && lfin;
              const float f = DROP ? ((dword & (1u << e)) ? drop.scale : 0.f) : 1.f;
              float p;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(s[e], sc2, -l2)));
              p = ok ? p : 0.f;
              s[e] = p * f;
              dp[e] = ok ? p * (dp[e] * f - di) : 0.f;
            }
          } else if (DROP) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const bool ok = rvalid && (k0 + c0 + e < klen);
              const float f = (dword & (1u << e)) ? drop.scale : 0.f;
              float p;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(s[e], sc2, -l2)));
              p = ok ? p : 0.f;
              s[e] = p * f;
              dp[e] = ok ? p * (dp[e] * f - di) : 0.f;
            }
          } else if (rvalid && k0 + c0 + 32 <= klen) {   // common case: no masked key in this slab
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              float p;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(s[e], sc2, -l2)));
              s[e] = p;
              dp[e] = p * (dp[e] - di);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const bool ok = rvalid && (k0 + c0 + e < klen);
              float p;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(s[e], sc2, -l2)));
              p = ok ? p : 0.f;
              s[e] = p;
              dp[e] = ok ? p * (dp[e] - di) : 0.f;
            }
          }
          if (hk == 0 && i > 0) {
            // P / dS of the previous query block are read by its dV / dK / dQ MMAs until dq_full flips: only now may this block's tiles
            // overwrite them -- and only now is the previous block's dQ partial complete.  Draining it HERE (after this half's math, all
            // eight warps, 16 columns each) keeps the exponentials of block i under the MMAs of block i - 1.
            mbar_wait(dq_full, (i - 1) & 1);
            drain_dq(i - 1);
          }
          if (threadIdx.x == 64) BTL(40 + i * 16 + hk * 6 + 2);
          uint8_t* prow = sP + hk * 16384 + r * 128;
          uint8_t* drow = sdS + hk * 16384 + r * 128;
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const int off = ((half * 4 + gg) ^ (r & 7)) << 4;
            uint4 u, w;
            __nv_bfloat162* hu = reinterpret_cast<__nv_bfloat162*>(&u);
            __nv_bfloat162* hw = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hu[e] = __floats2bfloat162_rn(s[gg * 8 + 2 * e], s[gg * 8 + 2 * e + 1]);
              hw[e] = __floats2bfloat162_rn(dp[gg * 8 + 2 * e], dp[gg * 8 + 2 * e + 1]);
            }
            *reinterpret_cast<uint4*>(prow + off) = u;
            *reinterpret_cast<uint4*>(drow + off) = w;
          }
          if (threadIdx.x == 64) BTL(40 + i * 16 + hk * 6 + 3);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(hk ? pds_full1 : pds_full);
          if (threadIdx.x == 64) BTL(40 + i * 16 + hk * 6 + 4);
        }
      }
      if (nq - i_begin > 0) {   // the last block's dQ partial
        mbar_wait(dq_full, (nq - i_begin - 1) & 1);
        drain_dq(nq - i_begin - 1);
      }
      mbar_wait(acc_full, 0);
      if (threadIdx.x == 64) BTL(3);
      tc_fence_after();
    }
    // dK / dV rows of this key block (zeros for fully masked blocks): the half-0 warps write dV, the half-1 warps dK
    const int tk = k0 + r;
    if (active) {
      tmem_ld32(trow + (half == 0 ? C_DV : C_DK), s);
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) s[e] = 0.f;
    }
    if (tk < Tk) {
      const float scale = (half == 0) ? 1.0f : rsqrtf((float)DHD);
      __nv_bfloat16* dst = (half == 0 ? g.dv + ((long)(rowk + tk)) * g.lddv : g.dk + ((long)(rowk + tk)) * g.lddk) + h * DHD;
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        float a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = s[gg * 8 + e] * scale;
        st8<__nv_bfloat16>(dst + gg * 8, a);
      }
    }
    tc_fence_before();
  }
  if (threadIdx.x == 64) BTL(4);
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
#ifdef EEC_ATTN_BWD_TIMELINE
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 32) {
    const long long t0 = g_btl[0];
    printf("attn_bwd CTA(0,0,0) clk since start: setup %lld, K/V landed %lld, acc_full seen %lld, outputs written %lld, end %lld\n", g_btl[1] - t0, g_btl[2] - t0,
           g_btl[3] - t0, g_btl[4] - t0, clock64() - t0);
    for (int i = 0; i < 3; ++i)
      printf("  i=%d MMA: Q/dO landed %lld, S/dP(h0) issued %lld, S/dP(h1) issued %lld, P/dS complete %lld, dV/dK/dQ issued %lld | softmax h0: S ready %lld, in regs %lld, math done %lld, stored %lld, arrived %lld; h1: %lld %lld %lld %lld %lld; dQ ready %lld, reduced %lld\n",
             i, g_btl[8 + i * 8] - t0, g_btl[9 + i * 8] - t0, g_btl[10 + i * 8] - t0, g_btl[11 + i * 8] - t0, g_btl[12 + i * 8] - t0, g_btl[40 + i * 16] - t0,
             g_btl[41 + i * 16] - t0, g_btl[42 + i * 16] - t0, g_btl[43 + i * 16] - t0, g_btl[44 + i * 16] - t0, g_btl[46 + i * 16] - t0, g_btl[47 + i * 16] - t0,
             g_btl[48 + i * 16] - t0, g_btl[49 + i * 16] - t0, g_btl[50 + i * 16] - t0, g_btl[52 + i * 16] - t0, g_btl[53 + i * 16] - t0);
  }
#endif
}

static int bwd_launch(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& td, const BwGeom& g, int B, const void* ctx,
                      const void* dctx, int ldo, const float* lse, float* dvec, float* dq32, void* dq, int lddq, const DropArgs& drop,
                      bool general, cudaStream_t st) {
  const long rows = (long)B * g.Tq;
  launch_pdl(attn_dvec_kernel, dim3((int)cdiv64(rows * 32, 256)), dim3(256), 0, st, (const __nv_bfloat16*)ctx, (const __nv_bfloat16*)dctx, ldo, dvec, dq32, B, g.Tq, g.H);
  EEC_LAUNCH_CHECK();
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM));
    EEC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM));
    EEC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM));
    EEC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(g.Tk, BT), g.H, B);
  if (general && drop.state) launch_pdl(attn_bwd_tc_kernel<true, true>, dim3(grid), dim3(BW_THREADS), BW_SMEM, st, tq, tkv, td, g, lse, dvec, dq32, drop);
  else if (general) launch_pdl(attn_bwd_tc_kernel<false, true>, dim3(grid), dim3(BW_THREADS), BW_SMEM, st, tq, tkv, td, g, lse, dvec, dq32, drop);
  else if (drop.state) launch_pdl(attn_bwd_tc_kernel<true, false>, dim3(grid), dim3(BW_THREADS), BW_SMEM, st, tq, tkv, td, g, lse, dvec, dq32, drop);
  else launch_pdl(attn_bwd_tc_kernel<false, false>, dim3(grid), dim3(BW_THREADS), BW_SMEM, st, tq, tkv, td, g, lse, dvec, dq32, drop);
  EEC_LAUNCH_CHECK();
  launch_pdl(dq_convert_kernel, dim3((int)cdiv64(rows * 32, 256)), dim3(256), 0, st, dq32, (__nv_bfloat16*)dq, lddq, rows, rsqrtf((float)DHD));
  EEC_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int attn_bwd_tc(const void* qkv, const void* ctx, const void* dctx, const float* lse, const int32_t* key_len, void* dqkv,
                float* dvec, float* dq32, int B, int T, int H, int dh, const DropArgs& drop, cudaStream_t st) {
  EEC_CHECK_ARG(dh == DHD && H * dh == 256, "attn_bwd_tc: needs 8 heads of 32");
  EEC_CHECK_ARG(dq32 != nullptr, "attn_bwd_tc: dq32 workspace is NULL");
  const long rows = (long)B * T;
  CUtensorMap tq, td;
  if (int r = get_tmap_2d(&tq, qkv, 768, (uint64_t)rows, 768 * 2, DHD, 128, /*SWIZZLE_64B*/ 2)) return r;
  if (int r = get_tmap_2d(&td, dctx, 256, (uint64_t)rows, 256 * 2, DHD, 128, 2)) return r;
  BwGeom g{};
  g.Tq = g.Tk = T; g.H = H; g.q_col = 0; g.k_col = 256; g.v_col = 512; g.key_len = key_len;
  g.dk = (__nv_bfloat16*)dqkv + 256; g.dv = (__nv_bfloat16*)dqkv + 512; g.lddk = g.lddv = 768;
  return bwd_launch(tq, tq, td, g, B, ctx, dctx, 256, lse, dvec, dq32, dqkv, 768, drop, false, st);
}

int attn_general_bwd_tc(const eec_attn_desc* d, const void* ctx, const void* dctx, int ldo, const float* lse, void* dq, int lddq, void* dk,
                        int lddk, void* dv, int lddv, float* dvec, float* dq32, const DropArgs& drop, cudaStream_t st) {
  EEC_CHECK_ARG(d->dh == DHD && d->H * d->dh == 256, "attn_general_bwd_tc: needs 8 heads of 32");
  EEC_CHECK_ARG(dq32 != nullptr, "attn_general_bwd_tc: dq32 workspace is NULL");
  const uintptr_t k = (uintptr_t)d->k, v = (uintptr_t)d->v;
  const void* kv_base = k < v ? d->k : d->v;
  BwGeom g{};
  g.Tq = d->Tq; g.Tk = d->Tk; g.H = d->H;
  g.q_col = 0; g.k_col = (int)((k - (uintptr_t)kv_base) / 2); g.v_col = (int)((v - (uintptr_t)kv_base) / 2);
  g.key_len = d->key_len; g.key_bits = d->key_valid_bits; g.causal = d->causal;
  g.dk = (__nv_bfloat16*)dk; g.lddk = lddk; g.dv = (__nv_bfloat16*)dv; g.lddv = lddv;
  CUtensorMap tq, tkv, td;
  if (int r = get_tmap_2d(&tq, d->q, (uint64_t)d->H * d->dh, (uint64_t)d->B * d->Tq, (uint64_t)d->ldq * 2, DHD, 128, 2)) return r;
  const uint64_t kv_cols = (uint64_t)(g.k_col > g.v_col ? g.k_col : g.v_col) + (uint64_t)d->H * d->dh;
  if (int r = get_tmap_2d(&tkv, kv_base, kv_cols, (uint64_t)d->B * d->Tk, (uint64_t)d->ldk * 2, DHD, 128, 2)) return r;
  if (int r = get_tmap_2d(&td, dctx, (uint64_t)d->H * d->dh, (uint64_t)d->B * d->Tq, (uint64_t)ldo * 2, DHD, 128, 2)) return r;
  return bwd_launch(tq, tkv, td, g, d->B, ctx, dctx, ldo, lse, dvec, dq32, dq, lddq, drop, true, st);
}

}  // namespace eec
