// gemm_ws.cu -- weight-stationary bf16 tcgen05 GEMM for the K = 256 projections of a conformer layer
// (FFN up-projection TA:104-105 and its dSiLU data gradient, the attention in-projection TA:194, pointwise conv 1 + GLU TA:44-51).
//
//   C[M,N] (bf16) = epi( A[M,256] * B(n,k) )       128 x 256 tile, one CTA per SM, every CTA keeps ONE weight tile
//
// Why a second kernel next to gemm_tc3: measured on B200 (profiles/r02_gemm_triage.txt) the v3 kernel needs 28.7 us for the
// 23936 x 2048 x 256 product with its epilogue switched OFF -- per 128 x 256 tile it streams 64 KB of A and 128 KB of B through
// shared memory (1.9 k clk at the measured 103 B/clk/SM) for 2.0 k clk of tcgen05.mma -- and its run-time-dispatched epilogue
// executes 755 instructions per warp and tile (27 % of them useful).  Here
//   * the [256 n x 256 k] weight tile (128 KB) is loaded ONCE per CTA and stays in shared memory: CTA c owns n-tile c % n_tiles and
//     walks the m-tiles r, r + cnt, ...; only the 64 KB A tile streams (4 x 16 KB ring = exactly one tile of look-ahead);
//   * the epilogue is a template: no run-time mode tests, bias folded into the tanh argument (h = 0.5 x + 0.5 b: one FFMA2),
//     packed fp32x2 math, one 2 KB staging box per warp;
//   * SiLU + pre-activation: the bf16 pre-activation leaves straight from registers (one 256-bit store per thread and 16 columns:
//     whole 32-byte sectors), the activation through the staged TMA store; dSiLU: the pre-activation arrives through 256-bit
//     loads issued before the accumulator wait (both forms are HBM-bound: 208 MB per launch);
//   * mbarrier waits use the suspend-time hint (no hot polling next to the epilogue warps).
//
//   warp 0      : TMA producer (weight tile once, then the A ring)
//   warp 1      : tcgen05.mma issuer, two 256-column TMEM accumulators
//   warps 2..17 : epilogue; quarter = warp % 4 (TMEM lanes), column group = (warp - 2) / 4 -> accumulator columns [cg*64, +64)
#include <stdlib.h>
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64, NKB = 4;   // K == 256
constexpr int A_STAGE = BM * BK * 2;                  // 16 KB
constexpr int B_KBLK = BN * BK * 2;                   // 32 KB
constexpr int NEW = 16;
constexpr int NTW = 64 + NEW * 32;                    // 576 threads
constexpr int OFF_B = 0;
constexpr int OFF_A = OFF_B + NKB * B_KBLK;           // 131072
constexpr int OFF_STG = OFF_A + NKB * A_STAGE;        // 196608
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
__device__ __forceinline__ void mbar_wait_s(uint64_t* bar, uint32_t parity, bool spin) {
  if (spin) mbar_wait(bar, parity);
  else {
    if (mbar_try_wait_h(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_h(bar, parity)) {
      if (clock64() - t0 > 4000000000LL) { printf("eec: gemm_ws mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
  }
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
template <int MODE>
__device__ __forceinline__ void epi16(const uint32_t (&c)[16], const float* b, uint32_t (&o)[8], uint32_t (&pre_o)[8],
                                      const uint32_t (&pre_i)[8], float alpha) {
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
      o[j >> 1] = pack_bf16(r);
    }
  }
}

template <int MODE, bool B_KMAJ>
__global__ void __launch_bounds__(NTW, 1) gemm_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                         const __grid_constant__ CUtensorMap tmC, const PW p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) { printf("eec: gemm_ws smem base not 1024-aligned\n"); __trap(); }
  float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [4]  A k-block kb of the current tile has landed
  uint64_t* empty_bar = full_bar + NKB;                               // [4]  its MMAs have completed
  uint64_t* tfull_bar = empty_bar + NKB;                              // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                               // [2]
  uint64_t* bfull_bar = tempty_bar + 2;                               // [4]  k-block kb of the weight tile has landed
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
  const int nt = blockIdx.x % p.n_tiles;        // this CTA's weight tile
  const int r0 = blockIdx.x / p.n_tiles;        // its rank among the CTAs that share it
  const int cnt = ((int)gridDim.x - nt + p.n_tiles - 1) / p.n_tiles;
  const int n0 = nt * BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < NKB; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], NEW); }
    for (int s = 0; s < NKB; ++s) mbar_init(&bfull_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  pdl_wait();
  if (threadIdx.x < 256) {
    float b = p.bias ? p.bias[n0 + threadIdx.x] : 0.f;
    if (MODE == WS_SILU || MODE == WS_SILU_PRE) b *= 0.5f;
    bias_s[threadIdx.x] = b;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int mt_eff = (active_rows(p.act_items, p.M) + BM - 1) / BM;
  const bool spin = p.knobs & 1;
  volatile long long* tl_issue = reinterpret_cast<volatile long long*>(smem + OFF_BAR + 192);
  (void)tl_issue;
  WS_TL(threadIdx.x == 0, p.tl[9] = gtime());
  long long w0_ = 0, w1_ = 0, w2_ = 0, w3_ = 0, w4_ = 0, t_ = 0, tb_ = 0;
  (void)w0_; (void)w1_; (void)w2_; (void)w3_; (void)w4_; (void)t_; (void)tb_;

  if (warp == 0) {
    if (lane == 0) {
      // weight tile: one barrier per k-block, interleaved with the first A tile, so that the first MMAs start after 48 KB instead of 192 KB
      uint32_t ph = 1;
      bool first = true;
      for (int mt = r0; mt < mt_eff; mt += cnt, ph ^= 1) {
        // the ring holds ONE tile of look-ahead (64 KB in flight per SM): against a 1.9 us HBM round trip that alone streams 34 GB/s
        // per SM (measured: 1.7 k clk of operand wait per tile), so the tiles after the next are pulled into L2 ahead of time
        if (first && !(p.knobs & 4)) {
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            if (mt + cnt < mt_eff) tma_prefetch_l2_2d(&tmA, kb * BK, (mt + cnt) * BM);
            if (mt + 2 * cnt < mt_eff) tma_prefetch_l2_2d(&tmA, kb * BK, (mt + 2 * cnt) * BM);
          }
        }
        if (mt + 3 * cnt < mt_eff && !(p.knobs & 4)) {
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) tma_prefetch_l2_2d(&tmA, kb * BK, (mt + 3 * cnt) * BM);
        }
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          if (first) {
            uint8_t* sb = smem + OFF_B + kb * B_KBLK;
            mbar_expect_tx(&bfull_bar[kb], B_KBLK);
            if (B_KMAJ) {
              tma_load_2d(sb, &tmB, &bfull_bar[kb], kb * BK, n0);
            } else {
#pragma unroll
              for (int a = 0; a < BN / 64; ++a) tma_load_2d(sb + a * 8192, &tmB, &bfull_bar[kb], n0 + a * 64, kb * BK);
            }
          }
          mbar_wait_s(&empty_bar[kb], ph, spin);
          WS_TL(kb == 0, tl_issue[0] = clock64());
          mbar_expect_tx(&full_bar[kb], A_STAGE);
          tma_load_2d(smem + OFF_A + kb * A_STAGE, &tmA, &full_bar[kb], kb * BK, mt * BM);
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, false, !B_KMAJ);
      constexpr uint64_t B_KSTEP = (B_KMAJ ? 32 : 2048) >> 4;
      uint64_t adesc[NKB], bdesc[NKB];
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        adesc[kb] = make_smem_desc(smem_u32(smem + OFF_A + kb * A_STAGE), 0, 1024);
        const uint32_t sb = smem_u32(smem + OFF_B + kb * B_KBLK);
        bdesc[kb] = B_KMAJ ? make_smem_desc(sb, 0, 1024) : make_smem_desc(sb, 8192, 1024);
      }
      WS_TL(true, p.tl[12] = gtime(); tb_ = clock64());
      uint32_t ut = 0;
      for (int mt = r0; mt < mt_eff; mt += cnt, ++ut) {
        const uint32_t acc = ut & 1;
        WS_TL(true, t_ = clock64());
        mbar_wait_s(&tempty_bar[acc], ((ut >> 1) & 1) ^ 1, spin);
        WS_TL(true, w0_ += clock64() - t_);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          WS_TL(true, t_ = clock64());
          if (ut == 0) mbar_wait_h(&bfull_bar[kb], 0);
          const bool had_to_wait = !mbar_try_wait(&full_bar[kb], ut & 1);
          mbar_wait_s(&full_bar[kb], ut & 1, spin);
          WS_TL(true, w1_ += clock64() - t_);
          WS_TL(kb == 0 && had_to_wait && ut > 0, w2_ += clock64() - tl_issue[0]; ++w3_);
          (void)had_to_wait;
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, adesc[kb] + k * 2, bdesc[kb] + k * B_KSTEP, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[kb]);
        }
        umma_commit(&tfull_bar[acc]);
      }
      WS_TL(true, p.tl[10] = gtime(); p.tl[0] = clock64() - tb_; p.tl[1] = w0_; p.tl[2] = w1_; p.tl[3] = ut; p.tl[14] = w2_; p.tl[15] = w3_);
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
    uint32_t pa0[8], pa1[8], pb0[8], pb1[8];   // packed outputs of sub-slabs 0..3
    int prev_row0 = -1;
    uint32_t ut = 0;
    for (int mt = r0; mt < mt_eff; mt += cnt, ++ut) {
      const int row0 = mt * BM + q * 32;
      const int m = row0 + lane;
      const bool valid = m < p.M;
      uint32_t pin[4][8];
      if (MODE == WS_DSILU) {
        const __nv_bfloat16* pp = p.pre + (long)m * p.ldp + ncol;
        // this lane's 128 bytes of the NEXT tile's pre-activation: into L2 now, so that the loads below are L2 hits a tile later
        if (mt + cnt < mt_eff && m + cnt * BM < p.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + (long)cnt * BM * p.ldp));
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
      epi16<MODE>(ra, bsl, pa0, po0, pin[0], p.alpha);
      if (MODE == WS_SILU_PRE && valid) stg256(pout, po0);
      if (defer && prev_row0 >= 0) store_box(pb0, pb1, ncol + 32, prev_row0);
      // ---- sub-slab 1
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      tmem_ld_wait16(rb);
      WS_TL(e == 0 && lane == 0, w2_ += clock64() - t_);
      tmem_ld16_async(tcol + 32, ra);
      epi16<MODE>(rb, bsl + 16, pa1, po1, pin[(MODE == WS_DSILU) ? 1 : 0], p.alpha);
      if (MODE == WS_SILU_PRE && valid) stg256(pout + 16, po1);
      if (!defer) store_box(pa0, pa1, ncol, row0);
      // ---- sub-slab 2
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      tmem_ld_wait16(ra);
      WS_TL(e == 0 && lane == 0, w2_ += clock64() - t_);
      tmem_ld16_async(tcol + 48, rb);
      epi16<MODE>(ra, bsl + 32, pb0, po0, pin[(MODE == WS_DSILU) ? 2 : 0], p.alpha);
      if (MODE == WS_SILU_PRE && valid) stg256(pout + 32, po0);
      if (defer) store_box(pa0, pa1, ncol, row0);
      // ---- sub-slab 3
      WS_TL(e == 0 && lane == 0, t_ = clock64());
      tmem_ld_wait16(rb);
      WS_TL(e == 0 && lane == 0, w2_ += clock64() - t_);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // the accumulator slice is in registers: the MMA warp may reuse it
      epi16<MODE>(rb, bsl + 48, pb1, po1, pin[(MODE == WS_DSILU) ? 3 : 0], p.alpha);
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
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  WS_TL(threadIdx.x == 32, p.tl[11] = gtime());
}

template <int MODE, bool B_KMAJ>
int launch_ws(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const PW& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(gemm_ws_kernel<MODE, B_KMAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
    attr_set = true;
  }
  launch_pdl(gemm_ws_kernel<MODE, B_KMAJ>, dim3(grid), dim3(NTW), WS_SMEM, st, ta, tb, tc_, p);
  EEC_LAUNCH_CHECK();
  return 0;
}

int g_sms_ws = 0;

}  // namespace

// can this descriptor run on the weight-stationary kernel?  (checked by gemm_tc3 before its own path)
bool gemm_ws_ok(const eec_gemm_desc* d) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("EEC_GEMM_WS"); env = (e && e[0] == '0') ? 0 : 1; }
  if (!env) return false;
  if (d->K != 256 || !d->a_kmajor || d->N % 256 != 0 || d->out_dtype != EEC_BF16 || d->in_dtype != EEC_BF16) return false;
  if (d->N / 256 > 64) return false;                                  // one weight tile per CTA (pair): more n-tiles than CTAs stay on the v3 kernel
  if (d->preact && ((reinterpret_cast<uintptr_t>(d->preact) & 31) != 0 || d->ldp % 16 != 0)) return false;   // 256-bit row accesses
  if ((reinterpret_cast<uintptr_t>(d->C) & 15) != 0 || d->ldc % 8 != 0) return false;
  if (d->accumulate || d->residual || d->a_colsum || d->ln_out) return false;
  if (d->drop_state && d->drop_p > 0.f) {   // dropout: the CTA-pair kernel has it for the two training forms (SiLU + stored pre-activation, dSiLU)
    const char* e = getenv("EEC_GEMM_WS");
    if (e && e[0] == '1') return false;
    int sms = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms % 2) return false;
    if (!d->drop_bits) return false;
    if (!((d->act == EEC_ACT_SILU && d->preact) || d->act == EEC_ACT_DSILU)) return false;
  }
  if (d->act == EEC_ACT_NONE) return d->alpha == 1.0f && !d->preact;   // (MN-major B: the N = K = 256 data gradients with a bf16 output)
  if (d->act == EEC_ACT_SILU) return d->b_kmajor && d->alpha == 1.0f && (!d->preact || d->preact_dtype == EEC_BF16);
  if (d->act == EEC_ACT_GLU) {   // value * sigmoid(gate) of pointwise conv 1: CTA-pair kernel only (the leader holds the value rows, its peer the gate rows)
    const char* e = getenv("EEC_GEMM_WS");
    if (e && e[0] == '1') return false;
    int sms = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms % 2) return false;
    return d->b_kmajor && d->alpha == 1.0f && (!d->preact || (d->preact_dtype == EEC_BF16 && d->ldp % 16 == 0)) && !(d->drop_state && d->drop_p > 0.f);
  }
  if (d->act == EEC_ACT_DSILU) return !d->b_kmajor && !d->bias && d->preact && d->preact_dtype == EEC_BF16 && d->ldp % 16 == 0;
  return false;
}

int gemm_ws(const eec_gemm_desc* d, cudaStream_t st) {
  {
    static int pair = -1;   // default: the CTA-pair (cta_group::2) version, gemm_ws2.cu; EEC_GEMM_WS=1 keeps this single-CTA kernel
    if (pair < 0) {
      const char* e = getenv("EEC_GEMM_WS");
      int sms = 0, dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      pair = (e && e[0] == '1') ? 0 : (sms % 2 == 0 ? 1 : 0);
    }
    if (pair) return gemm_ws2(d, st);
  }
  if (!g_sms_ws) {
    int dev = 0;
    EEC_CUDA(cudaGetDevice(&dev));
    EEC_CUDA(cudaDeviceGetAttribute(&g_sms_ws, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap ta, tb, tcm;
  if (int r = get_tmap_2d(&ta, d->A, d->K, d->M, (uint64_t)d->lda * 2, 64, 128)) return r;
  if (d->b_kmajor) { if (int r = get_tmap_2d(&tb, d->B, d->K, d->N, (uint64_t)d->ldb * 2, 64, 256)) return r; }
  else { if (int r = get_tmap_2d(&tb, d->B, d->N, d->K, (uint64_t)d->ldb * 2, 64, 64)) return r; }
  if (int r = get_tmap_box32(&tcm, d->C, true, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldc)) return r;
  PW p{};
  p.M = d->M; p.N = d->N; p.m_tiles = cdiv(d->M, BM); p.n_tiles = d->N / BN;
  p.bias = d->bias; p.alpha = d->alpha;
  p.pre = reinterpret_cast<__nv_bfloat16*>(d->preact); p.ldp = d->ldp;
  p.act_items = (d->act != EEC_ACT_DSILU) ? active_items(st) : ActiveItems{nullptr, 0, 0};   // forward forms: rows of A are frames
  const bool store_pre = d->act == EEC_ACT_SILU && d->preact;
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
      fprintf(stderr, "gemm_ws %dx%dx256 CTA0: %lld tiles; per tile (clk): MMA thread %.0f = wait accumulator %.0f + wait A %.0f + issue; epilogue warp: wait tfull %.0f, busy %.0f (tmem ld waits %.0f, staging-free waits %.0f, fence + store issue %.0f)\n",
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
  const int grid = min(p.m_tiles * p.n_tiles, g_sms_ws);
  EEC_CHECK_ARG(grid >= p.n_tiles, "gemm_ws: fewer CTAs (%d) than weight tiles (%d)", grid, p.n_tiles);
  if (d->act == EEC_ACT_NONE) return d->b_kmajor ? launch_ws<WS_BIAS, true>(ta, tb, tcm, p, grid, st) : launch_ws<WS_BIAS, false>(ta, tb, tcm, p, grid, st);
  if (d->act == EEC_ACT_SILU) return store_pre ? launch_ws<WS_SILU_PRE, true>(ta, tb, tcm, p, grid, st) : launch_ws<WS_SILU, true>(ta, tb, tcm, p, grid, st);
  return launch_ws<WS_DSILU, false>(ta, tb, tcm, p, grid, st);
}

}  // namespace eec
