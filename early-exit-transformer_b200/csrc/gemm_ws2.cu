// gemm_ws2.cu -- weight-stationary bf16 tcgen05 GEMM for the K = 256 projections, CTA-PAIR version (tcgen05 cta_group::2).
//
//   C[M,N] (bf16) = epi( A[M,256] * B(n,k) )      256 x 256 tile per CTA pair (2 SMs of one TPC), 128 rows per CTA
//
// Why pairs (measured with the single-CTA kernel gemm_ws.cu, profiles/r02_gemm_ws_*.txt): a 128 x 256 x 16 tcgen05.mma reads 4 KB of A
// and 8 KB of B from shared memory every 128 clk = 96 B/clk of the SM's 128 B/clk, next to the TMA writes of the A ring, the staging
// stores and the bulk-store reads -- the tensor pipe was active 2.6 k clk per tile instead of 2.0 k and the tile period sat at 3.5 k.
// With cta_group::2 every CTA supplies its own 128 rows of A and only HALF of the weight tile (128 of the 256 n-rows): 64 B/clk per
// SM, the resident weights shrink to 64 KB per CTA and the A ring grows to 8 x 16 KB = two full tiles of look-ahead.
//   * cluster (2,1,1); pair p owns weight tile p % n_tiles and walks the 256-row super-tiles r, r + cnt, ...
//   * both CTAs load their A rows / B half with cp.async.bulk.tensor...cta_group::2, completing on the LEADER's mbarriers;
//   * the leader's warp 1 issues tcgen05.mma.cta_group::2 (M = 256) and commits with a multicast arrive to both CTAs' barriers
//     (smem slot free, accumulator full); the epilogue warps of both CTAs release the accumulator on the leader's barrier;
//   * the epilogue is the one of gemm_ws.cu (template modes, bias folded into the tanh argument, staggered store fences).
// Measured alternatives for the output path (B200, 23936 x 2048 x 256 + SiLU, this kernel at 31.4 us): one [32 x 64] box per warp and
// tile with 4 KB of staging and a 6-stage A ring: 32.0 us; re-reading the staging buffer and storing through the LSU (8 rows x 64 B per
// instruction, no proxy fence): 39.3 us; 256-bit stores straight from the row-per-thread registers: 4.1 k clk per tile on their own
// (tools/tma_store_bw.cu).  The bulk tensor store of [32 x 32] boxes stays.
#include <stdlib.h>
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64, NKB = 4;   // K == 256; BM = rows per CTA (256 per pair)
constexpr int NSA = 8;                                // A ring: two tiles
constexpr int A_STAGE = BM * BK * 2;                  // 16 KB
constexpr int B_KBLK = (BN / 2) * BK * 2;             // 16 KB: this CTA's half of the weight tile's k-block
constexpr int NEW = 16;
constexpr int NTW = 64 + NEW * 32;                    // 576 threads
constexpr int OFF_B = 0;
constexpr int OFF_A = OFF_B + NKB * B_KBLK;           // 65536
constexpr int OFF_STG = OFF_A + NSA * A_STAGE;        // 196608
constexpr int OFF_BIAS = OFF_STG + NEW * 2048;        // 229376
constexpr int OFF_BAR = OFF_BIAS + 256 * 4;           // 230400
constexpr int WS_SMEM = OFF_BAR + 256;                // 230656 <= 232448

enum { WS_BIAS = 0, WS_SILU = 1, WS_SILU_PRE = 2, WS_DSILU = 3, WS_GLU = 4, WS_GLU_PRE = 5 };

struct PW {
  int M, N, m_tiles, n_tiles;
  const float* bias;
  float alpha;
  __nv_bfloat16* pre; int ldp;     // SILU_PRE / GLU_PRE: written; DSILU: read
  ActiveItems act_items;
  DropArgs drop;   // DROP instantiations: dropout right after the activation (element index m*N + n), keep-mask words of eec_dropout_bits (W = 16)
  int knobs;       // EEC_WS_KNOBS (perf triage): 1 = producer / MMA threads poll their barriers, 2 = no staggered stores, 4 = no L2 prefetch of A
  long long* tl;   // EEC_GEMM_TL=1 on a -DEEC_GEMM_TIMELINE build (perf triage): clock64 / globaltimer accumulators of CTA 0
};

// wait with the hardware suspend hint: the thread sleeps until the phase completes (or the hint expires) instead of polling
__device__ __forceinline__ bool mbar_try_wait_h(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait_h(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_h(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_h(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("eec: gemm_ws mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
               "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float2 v) {
  const __nv_bfloat162 h = __float22bfloat162_rn(v);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {   // exact: bf16 -> fp32 is a 16-bit shift
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// 16 accumulator columns of this lane's row -> 8 packed bf16 pairs.  `b` = this sub-slab's 16 bias values in shared memory
// (warp-uniform address: broadcast; already scaled by 0.5 for the SiLU modes).
template <int MODE, bool DROP = false>
__device__ __forceinline__ void epi16(const uint32_t (&c)[16], const float* b, uint32_t (&o)[8], uint32_t (&pre_o)[8],
                                      const uint32_t (&pre_i)[8], float alpha, uint32_t dword = 0xffffu, float dscale = 1.f) {
  const float4* bp = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE != WS_DSILU) f = bp[g];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = g * 4 + e * 2;
      const float2 x = make_float2(__uint_as_float(c[j]), __uint_as_float(c[j + 1]));
      const float2 bb = e ? make_float2(f.z, f.w) : make_float2(f.x, f.y);
      float2 r;
      if (MODE == WS_BIAS) {
        r = __fadd2_rn(x, bb);
      } else if (MODE == WS_SILU || MODE == WS_SILU_PRE) {
        const float2 h = __ffma2_rn(x, make_float2(0.5f, 0.5f), bb);            // (x + bias) / 2
        if (MODE == WS_SILU_PRE) pre_o[j >> 1] = pack_bf16(__fadd2_rn(h, h));    // x + bias (exact doubling)
        r = __ffma2_rn(h, make_float2(tanh_fast(h.x), tanh_fast(h.y)), h);       // v sigmoid(v) = h + h tanh(h)
      } else {   // WS_DSILU: alpha * x * dSiLU(z),  dSiLU(z) = (1 + t)/2 + (z/2)(1 - t^2)/2, t = tanh(z/2)
        const float cc = 0.5f * alpha;
        const float2 hh = __fmul2_rn(unpack_bf16(pre_i[j >> 1]), make_float2(0.5f, 0.5f));
        const float2 t = make_float2(tanh_fast(hh.x), tanh_fast(hh.y));
        const float2 q = __ffma2_rn(__fmul2_rn(t, t), make_float2(-cc, -cc), make_float2(cc, cc));
        const float2 d = __ffma2_rn(hh, q, __ffma2_rn(t, make_float2(cc, cc), make_float2(cc, cc)));
        r = __fmul2_rn(x, d);
      }
      if (DROP) {   // keep-mask bit j of this thread's 16-column word; kept values are scaled by 1 / (1 - p)
        r.x = (dword & (1u << j)) ? r.x * dscale : 0.f;
        r.y = (dword & (2u << j)) ? r.y * dscale : 0.f;
      }
      o[j >> 1] = pack_bf16(r);
    }
  }
}

// (2SM TMA loads, cta_group::2 MMA / commit and the remote arrive live in tc_common.cuh)  TMEM of a pair: one warp of EACH CTA allocates
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int MODE, bool B_KMAJ, bool DROP = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTW, 1)
    gemm_ws2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const PW p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) { printf("eec: gemm_ws2 smem base not 1024-aligned\n"); __trap(); }
  float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [8]  (leader's copy is used) A stage s of BOTH CTAs has landed
  uint64_t* empty_bar = full_bar + NSA;                               // [8]  (per CTA) its MMAs have completed
  uint64_t* tfull_bar = empty_bar + NSA;                              // [2]  (per CTA) accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;                               // [2]  (leader's copy) drained by the epilogue warps of both CTAs
  uint64_t* bfull_bar = tempty_bar + 2;                               // [4]  (leader's copy) k-block kb of both weight halves has landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bfull_bar + NKB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef EEC_GEMM_TIMELINE
  auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  const bool tlc = p.tl && blockIdx.x == 0;
  if (tlc && threadIdx.x == 0) p.tl[8] = gtime();
#define WS_TL(cond, stmt) do { if (tlc && (cond)) { stmt; } } while (0)
#else
#define WS_TL(cond, stmt) do { } while (0)
#endif
  const uint32_t rank = cluster_ctarank();      // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nt = pair % p.n_tiles;              // this pair's weight tile
  const int r0 = pair / p.n_tiles;              // its rank among the pairs that share it
  const int cnt = (npairs - nt + p.n_tiles - 1) / p.n_tiles;
  const int n0 = nt * BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < NSA; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * NEW); }
    for (int s = 0; s < NKB; ++s) mbar_init(&bfull_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr_smem, 512);
  pdl_wait();
  if (threadIdx.x < 256) {
    float b;
    if (MODE == WS_GLU || MODE == WS_GLU_PRE) {
      const int i = threadIdx.x;
      b = p.bias ? 0.5f * p.bias[i < 128 ? nt * 128 + i : p.N / 2 + nt * 128 + (i - 128)] : 0.f;   // halved: value/2 and gate/2 are what the epilogue needs
    } else {
      b = p.bias ? p.bias[n0 + threadIdx.x] : 0.f;
      if (MODE == WS_SILU || MODE == WS_SILU_PRE) b *= 0.5f;
    }
    bias_s[threadIdx.x] = b;
  }
  tc_fence_before();
  cluster_sync_all();     // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int st_eff = (active_rows(p.act_items, p.M) + 2 * BM - 1) / (2 * BM);   // 256-row super-tiles
  WS_TL(threadIdx.x == 0, p.tl[9] = gtime());
  long long w0_ = 0, w1_ = 0, w2_ = 0, w3_ = 0, w4_ = 0, t_ = 0, tb_ = 0;
  (void)w0_; (void)w1_; (void)w2_; (void)w3_; (void)w4_; (void)t_; (void)tb_;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      bool first = true;
      // this CTA's half of the weight tile: n-rows [nb, nb + 128).  GLU: the leader holds the 128 value rows of output channels
      // [nt*128, +128), its peer the 128 gate rows (N/2 further down): accumulator columns [0,128) = value, [128,256) = gate in both CTAs
      const int nb = (MODE == WS_GLU || MODE == WS_GLU_PRE) ? (rank == 0 ? nt * 128 : p.N / 2 + nt * 128) : n0 + (int)rank * (BN / 2);
      for (int st = r0; st < st_eff; st += cnt) {
        const int m0 = st * 2 * BM + (int)rank * BM;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          if (first) {
            uint8_t* sb = smem + OFF_B + kb * B_KBLK;
            if (rank == 0) mbar_expect_tx(&bfull_bar[kb], 2 * B_KBLK);
            if (B_KMAJ) {
              tma_load_2d_2sm(sb, &tmB, &bfull_bar[kb], kb * BK, nb);
            } else {
#pragma unroll
              for (int a = 0; a < BN / 128; ++a) tma_load_2d_2sm(sb + a * 8192, &tmB, &bfull_bar[kb], nb + a * 64, kb * BK);
            }
          }
          mbar_wait_h(&empty_bar[s], ph);
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * A_STAGE);
          tma_load_2d_2sm(smem + OFF_A + s * A_STAGE, &tmA, &full_bar[s], kb * BK, m0);
          if (++s == NSA) { s = 0; ph ^= 1; }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, false, !B_KMAJ);
      constexpr uint64_t B_KSTEP = (B_KMAJ ? 32 : 2048) >> 4;
      uint64_t bdesc[NKB];
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const uint32_t sb = smem_u32(smem + OFF_B + kb * B_KBLK);
        bdesc[kb] = B_KMAJ ? make_smem_desc(sb, 0, 1024) : make_smem_desc(sb, 8192, 1024);
      }
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem + OFF_A), 0, 1024);
      WS_TL(true, p.tl[12] = gtime(); tb_ = clock64());
      uint32_t ut = 0, ph = 0;
      int s = 0;
      for (int st = r0; st < st_eff; st += cnt, ++ut) {
        const uint32_t acc = ut & 1;
        WS_TL(true, t_ = clock64());
        mbar_wait_h(&tempty_bar[acc], ((ut >> 1) & 1) ^ 1);
        WS_TL(true, w0_ += clock64() - t_);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          WS_TL(true, t_ = clock64());
          if (ut == 0) mbar_wait_h(&bfull_bar[kb], 0);
          mbar_wait_h(&full_bar[s], ph);
          WS_TL(true, w1_ += clock64() - t_);
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)((s * A_STAGE) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            if (!(p.knobs & 8)) umma2_bf16(d_tmem, ad + k * 2, bdesc[kb] + k * B_KSTEP, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma2_commit_both(&empty_bar[s]);
          if (++s == NSA) { s = 0; ph ^= 1; }
        }
        umma2_commit_both(&tfull_bar[acc]);
      }
      WS_TL(true, p.tl[10] = gtime(); p.tl[0] = clock64() - tb_; p.tl[1] = w0_; p.tl[2] = w1_; p.tl[3] = ut);
    }
  } else {
    // ================================================================== epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;
    const int cg = e >> 2;
    uint8_t* buf = smem + OFF_STG + e * 2048;
    uint8_t* srow = buf + lane * 64;
    const int sw = (lane >> 1) & 3;
    const float* bsl = bias_s + cg * 64;
    const int ncol = n0 + cg * 64;   // first output column of this warp
    // one [32 rows x 32 cols] bf16 box: 8 + 8 packed pairs of this lane's row -> swizzled staging buffer -> bulk tensor store.
    // The generic->async proxy fence + store issue cost ~360 clk per box (measured), during which the warp issues no MUFU work.
    auto store_box = [&](const uint32_t(&a)[8], const uint32_t(&b)[8], int col, int row) {
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      if (lane == 0) bulk_wait_read<0>();
      WS_TL(e == 0 && lane == 0, w3_ += clock64() - t_);
      __syncwarp();
      *reinterpret_cast<uint4*>(srow + ((0 ^ sw) << 4)) = make_uint4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<uint4*>(srow + ((1 ^ sw) << 4)) = make_uint4(a[4], a[5], a[6], a[7]);
      *reinterpret_cast<uint4*>(srow + ((2 ^ sw) << 4)) = make_uint4(b[0], b[1], b[2], b[3]);
      *reinterpret_cast<uint4*>(srow + ((3 ^ sw) << 4)) = make_uint4(b[4], b[5], b[6], b[7]);
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmC, buf, col, row);
        bulk_commit();
      }
      WS_TL(e == 0 && lane == 0, w4_ += clock64() - t_);
    };
    // The four warps of an SM sub-partition (same lane quarter, column groups 0..3) start every tile together and would stay in
    // lock-step: all in their MUFU phase (the pipe is 4x oversubscribed), then all in their store fences (the pipe idles).  The odd
    // column groups therefore store each box one sub-slab LATER than the even ones (the last box of a tile after the first
    // sub-slab of the next tile), so that one pair's fences overlap the other pair's tanh evaluations.
    constexpr bool CAN_DEFER = (MODE == WS_SILU || MODE == WS_BIAS);
    const bool defer = CAN_DEFER && (cg & 1) && !(p.knobs & 2);
    const bool spin = false; (void)spin;
    uint32_t pa0[8], pa1[8], pb0[8], pb1[8];   // packed outputs of sub-slabs 0..3
    int prev_row0 = -1;
    uint32_t ut = 0;
    for (int st = r0; st < st_eff; st += cnt, ++ut) {
      const int row0 = st * 2 * BM + (int)rank * BM + q * 32;
      const int m = row0 + lane;
      const bool valid = m < p.M;
      uint32_t pin[4][8];
      if (MODE == WS_DSILU) {
        const __nv_bfloat16* pp = p.pre + (long)m * p.ldp + ncol;
        // this lane's 128 bytes of the NEXT tile's pre-activation: into L2 now, so that the loads below are L2 hits a tile later
        if (st + cnt < st_eff && m + cnt * 2 * BM < p.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + (long)cnt * 2 * BM * p.ldp));
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
          if (valid) ldg256(pp + ss * 16, pin[ss]);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) pin[ss][j] = 0u;
          }
        }
      }
      __nv_bfloat16* pout = (MODE == WS_SILU_PRE) ? p.pre + (long)m * p.ldp + ncol : nullptr;
      uint32_t dwv[4] = {0xffffu, 0xffffu, 0xffffu, 0xffffu};   // in flight while the main loop of this tile runs
      if (DROP && valid) {
        const uint16_t* db = reinterpret_cast<const uint16_t*>(p.drop.bits);
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) dwv[ss] = db[(long)((ncol >> 4) + ss) * p.M + m];
      }
      if (MODE == WS_GLU || MODE == WS_GLU_PRE) {
        // out[:, nt*128 + cg*32 .. +32) = (a + ba) * sigmoid(g + bg): value columns cg*32.., gate columns 128 + cg*32.. of the accumulator;
        // with ah = (a + ba)/2, gh = (g + bg)/2:  out = ah + ah * tanh(gh); the stored pre-activation z = [2 ah | 2 gh] (exact doubling)
        const uint32_t acc_ = ut & 1;
        const uint32_t tc0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc_ * BN + cg * 32;
        mbar_wait_h(&tfull_bar[acc_], (ut >> 1) & 1);
        tc_fence_after();
        uint32_t va[16], vg[16], wa[16], wg[16];
        tmem_ld16_async(tc0, va);
        tmem_ld16_async(tc0 + 128, vg);
        tmem_ld16_async(tc0 + 16, wa);
        tmem_ld16_async(tc0 + 128 + 16, wg);
        tmem_ld_wait16(wg);   // (one wait retires all four loads; only wg is threaded through the wait statement -- checked in the SASS:
                              //  the first consumers of va / vg / wa sit behind it; tie all four arrays to the wait when this block changes)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&tempty_bar[acc_]);
        const int ocol = nt * 128 + cg * 32;
        __nv_bfloat16* zrow = (MODE == WS_GLU_PRE) ? p.pre + (long)m * p.ldp : nullptr;
#pragma unroll
        for (int ss = 0; ss < 2; ++ss) {
          const uint32_t(&ca)[16] = ss ? wa : va;
          const uint32_t(&cgt)[16] = ss ? wg : vg;
          const float4* ba = reinterpret_cast<const float4*>(bias_s + cg * 32 + ss * 16);
          const float4* bg = reinterpret_cast<const float4*>(bias_s + 128 + cg * 32 + ss * 16);
          uint32_t za[8], zg[8];
          uint32_t(&o)[8] = ss ? pa1 : pa0;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 fa = ba[g4], fg = bg[g4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int j = g4 * 4 + h2 * 2;
              const float2 ah = __ffma2_rn(make_float2(__uint_as_float(ca[j]), __uint_as_float(ca[j + 1])), make_float2(0.5f, 0.5f),
                                           h2 ? make_float2(fa.z, fa.w) : make_float2(fa.x, fa.y));
              const float2 gh = __ffma2_rn(make_float2(__uint_as_float(cgt[j]), __uint_as_float(cgt[j + 1])), make_float2(0.5f, 0.5f),
                                           h2 ? make_float2(fg.z, fg.w) : make_float2(fg.x, fg.y));
              if (MODE == WS_GLU_PRE) { za[j >> 1] = pack_bf16(__fadd2_rn(ah, ah)); zg[j >> 1] = pack_bf16(__fadd2_rn(gh, gh)); }
              o[j >> 1] = pack_bf16(__ffma2_rn(ah, make_float2(tanh_fast(gh.x), tanh_fast(gh.y)), ah));
            }
          }
          if (MODE == WS_GLU_PRE && valid) { stg256(zrow + ocol + ss * 16, za); stg256(zrow + p.N / 2 + ocol + ss * 16, zg); }
        }
        store_box(pa0, pa1, ocol, row0);
        continue;
      }
      const uint32_t acc = ut & 1;
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cg * 64;
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      mbar_wait_h(&tfull_bar[acc], (ut >> 1) & 1);
      WS_TL(e == 0 && lane == 0, tb_ = clock64(); w0_ += tb_ - t_);
      tc_fence_after();
      uint32_t ra[16], rb[16], po0[8], po1[8];
      tmem_ld16_async(tcol, ra);
      // ---- sub-slab 0
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      tmem_ld_wait16(ra);
      WS_TL(e == 0 && lane == 0, w2_ += clock64() - t_);
      tmem_ld16_async(tcol + 16, rb);
      epi16<MODE, DROP>(ra, bsl, pa0, po0, pin[0], p.alpha, dwv[0], p.drop.scale);
      if (MODE == WS_SILU_PRE && valid) stg256(pout, po0);
      if (defer && prev_row0 >= 0) store_box(pb0, pb1, ncol + 32, prev_row0);
      // ---- sub-slab 1
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      tmem_ld_wait16(rb);
      WS_TL(e == 0 && lane == 0, w2_ += clock64() - t_);
      tmem_ld16_async(tcol + 32, ra);
      epi16<MODE, DROP>(rb, bsl + 16, pa1, po1, pin[(MODE == WS_DSILU) ? 1 : 0], p.alpha, dwv[1], p.drop.scale);
      if (MODE == WS_SILU_PRE && valid) stg256(pout + 16, po1);
      if (!defer) store_box(pa0, pa1, ncol, row0);
      // ---- sub-slab 2
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      tmem_ld_wait16(ra);
      WS_TL(e == 0 && lane == 0, w2_ += clock64() - t_);
      tmem_ld16_async(tcol + 48, rb);
      epi16<MODE, DROP>(ra, bsl + 32, pb0, po0, pin[(MODE == WS_DSILU) ? 2 : 0], p.alpha, dwv[2], p.drop.scale);
      if (MODE == WS_SILU_PRE && valid) stg256(pout + 32, po0);
      if (defer) store_box(pa0, pa1, ncol, row0);
      // ---- sub-slab 3
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      tmem_ld_wait16(rb);
      WS_TL(e == 0 && lane == 0, w2_ += clock64() - t_);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(&tempty_bar[acc]);   // the accumulator slice is in registers: the leader's MMA warp may reuse it
      epi16<MODE, DROP>(rb, bsl + 48, pb1, po1, pin[(MODE == WS_DSILU) ? 3 : 0], p.alpha, dwv[3], p.drop.scale);
      if (MODE == WS_SILU_PRE && valid) stg256(pout + 48, po1);
      if (!defer) store_box(pb0, pb1, ncol + 32, row0);
      prev_row0 = row0;
      WS_TL(e == 0 && lane == 0, w1_ += clock64() - tb_);
    }
    if (defer && prev_row0 >= 0) store_box(pb0, pb1, ncol + 32, prev_row0);
    WS_TL(e == 0 && lane == 0, p.tl[4] = w0_; p.tl[5] = w1_; p.tl[6] = w2_; p.tl[7] = w3_; p.tl[13] = w4_);
    if (lane == 0) bulk_wait_all();
    tc_fence_before();
  }
  cluster_sync_all();   // the leader's MMAs write the peer's TMEM and read its shared memory: neither CTA may leave early
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
  WS_TL(threadIdx.x == 32, p.tl[11] = gtime());
}

template <int MODE, bool B_KMAJ, bool DROP = false>
int launch_ws2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const PW& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(gemm_ws2_kernel<MODE, B_KMAJ, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
    attr_set = true;
  }
  gemm_ws2_kernel<MODE, B_KMAJ, DROP><<<dim3(grid), dim3(NTW), WS_SMEM, st>>>(ta, tb, tc_, p);   // (static cluster dims 2 x 1 x 1)
  EEC_LAUNCH_CHECK();
  return 0;
}

int g_sms_ws2 = 0;

}  // namespace

// CTA-pair version of gemm_ws (same eligibility: gemm_ws_ok); needs an even number of SMs
int gemm_ws2(const eec_gemm_desc* d, cudaStream_t st) {
  if (!g_sms_ws2) {
    int dev = 0;
    EEC_CUDA(cudaGetDevice(&dev));
    EEC_CUDA(cudaDeviceGetAttribute(&g_sms_ws2, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap ta, tb, tcm;
  if (int r = get_tmap_2d(&ta, d->A, d->K, d->M, (uint64_t)d->lda * 2, 64, 128)) return r;
  if (d->b_kmajor) { if (int r = get_tmap_2d(&tb, d->B, d->K, d->N, (uint64_t)d->ldb * 2, 64, 128)) return r; }   // one CTA's half of a weight tile
  else { if (int r = get_tmap_2d(&tb, d->B, d->N, d->K, (uint64_t)d->ldb * 2, 64, 64)) return r; }
  const bool glu = d->act == EEC_ACT_GLU;   // output [M, N/2]
  if (int r = get_tmap_box32(&tcm, d->C, true, (uint64_t)(glu ? d->N / 2 : d->N), (uint64_t)d->M, (uint64_t)d->ldc)) return r;
  PW p{};
  p.M = d->M; p.N = d->N; p.m_tiles = cdiv(d->M, BM); p.n_tiles = d->N / BN;
  p.bias = d->bias; p.alpha = d->alpha;
  p.pre = reinterpret_cast<__nv_bfloat16*>(d->preact); p.ldp = d->ldp;
  p.act_items = (d->act != EEC_ACT_DSILU) ? active_items(st) : ActiveItems{nullptr, 0, 0};   // forward forms: rows of A are frames
  const bool store_pre = d->act == EEC_ACT_SILU && d->preact;
  if (glu && d->preact) EEC_CHECK_ARG(d->ldp % 16 == 0 && (reinterpret_cast<uintptr_t>(d->preact) & 31) == 0, "gemm_ws2: GLU pre-activation needs 32-byte aligned rows");
  if (store_pre) EEC_CHECK_ARG(d->ldp % 16 == 0 && (reinterpret_cast<uintptr_t>(d->preact) & 31) == 0, "gemm_ws: pre-activation needs 32-byte aligned rows");
  if (d->act == EEC_ACT_DSILU) EEC_CHECK_ARG((reinterpret_cast<uintptr_t>(d->preact) & 31) == 0, "gemm_ws: pre-activation needs 32-byte aligned rows");
  static int knobs = -1;
  if (knobs < 0) { const char* e = getenv("EEC_WS_KNOBS"); knobs = e ? atoi(e) : 0; }
  p.knobs = knobs;
  static int tl_env = -1;
  static long long* tl_buf = nullptr;
  if (tl_env < 0) { const char* e = getenv("EEC_GEMM_TL"); tl_env = (e && e[0] == '1') ? 1 : 0; }
  struct Report {
    long long* b; cudaStream_t s; int M, N;
    ~Report() {
      long long h[16];
      if (!b || cudaStreamSynchronize(s) != cudaSuccess || cudaMemcpy(h, b, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return;
      const double t = h[3] ? (double)h[3] : 1.0;
      fprintf(stderr, "gemm_ws2 %dx%dx256 CTA0: %lld tiles; per tile (clk): MMA thread %.0f = wait accumulator %.0f + wait A %.0f + issue; epilogue warp: wait tfull %.0f, busy %.0f (tmem ld waits %.0f, staging-free waits %.0f, fence + store issue %.0f)\n",
              M, N, h[3], h[0] / t, h[1] / t, h[2] / t, h[4] / t, h[5] / t, h[6] / t, h[7] / t, h[13] / t);
      fprintf(stderr, "        A k-block 0: TMA issue -> seen by the MMA thread %.0f clk (%lld waits)\n", h[15] ? (double)h[14] / h[15] : 0.0, h[15]);
      fprintf(stderr, "        CTA0 globaltimer (ns): setup %lld, weight tile landed +%lld, main loop %lld, drain+teardown %lld, total %lld\n", h[9] - h[8], h[12] - h[9],
              h[10] - h[12], h[11] - h[10], h[11] - h[8]);
    }
  } report{nullptr, st, d->M, d->N};
  if (tl_env) {
    if (!tl_buf) EEC_CUDA(cudaMalloc(&tl_buf, 128));
    EEC_CUDA(cudaMemsetAsync(tl_buf, 0, 128, st));
    p.tl = tl_buf;
    report.b = tl_buf;
  }
  const int grid = 2 * min(cdiv(p.m_tiles, 2) * p.n_tiles, g_sms_ws2 / 2);   // CTA pairs
  EEC_CHECK_ARG(grid / 2 >= p.n_tiles, "gemm_ws2: fewer CTA pairs (%d) than weight tiles (%d)", grid / 2, p.n_tiles);
  p.drop = make_drop(d->drop_state, d->drop_p, d->drop_site);
  p.drop.bits = d->drop_bits;
  if (p.drop.state) {
    EEC_CHECK_ARG(d->drop_bits != nullptr, "gemm (tensor-core path): dropout needs the keep-mask words of eec_dropout_bits(R = M, C = N, Cs = N, W = 16) in drop_bits");
    if (d->act == EEC_ACT_SILU && store_pre) return launch_ws2<WS_SILU_PRE, true, true>(ta, tb, tcm, p, grid, st);
    if (d->act == EEC_ACT_DSILU) return launch_ws2<WS_DSILU, false, true>(ta, tb, tcm, p, grid, st);
    set_error("gemm_ws2: dropout is fused into the SiLU + pre-activation and dSiLU forms only");
    return 1;
  }
  if (d->act == EEC_ACT_GLU) return d->preact ? launch_ws2<WS_GLU_PRE, true>(ta, tb, tcm, p, grid, st) : launch_ws2<WS_GLU, true>(ta, tb, tcm, p, grid, st);
  if (d->act == EEC_ACT_NONE) return d->b_kmajor ? launch_ws2<WS_BIAS, true>(ta, tb, tcm, p, grid, st) : launch_ws2<WS_BIAS, false>(ta, tb, tcm, p, grid, st);
  if (d->act == EEC_ACT_SILU) return store_pre ? launch_ws2<WS_SILU_PRE, true>(ta, tb, tcm, p, grid, st) : launch_ws2<WS_SILU, true>(ta, tb, tcm, p, grid, st);
  return launch_ws2<WS_DSILU, false>(ta, tb, tcm, p, grid, st);
}

}  // namespace eec
