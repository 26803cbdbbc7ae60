// gemm_tc3.cu -- persistent bf16 tcgen05 GEMM, v3 epilogue: 16 epilogue warps with warp-private TMA staging.
//
//   C[M,N] = epi( A(m,k) * B(n,k) )        128 x 256 tile, BLOCK_K = 64, one CTA per SM, static tile loop
//
//   warp 0       : TMA producer (3-stage 128B-swizzled smem ring, runs ahead across tiles)
//   warp 1       : tcgen05.mma issuer; TWO 256-column TMEM accumulators (tile t+1's main loop overlaps tile t's epilogue)
//   warps 2..17  : epilogue; quarter = warp%4 selects the TMEM lane quarter (row = quarter*32 + lane),
//                  column group = (warp-2)/4 selects accumulator columns [cg*64, cg*64+64)
//
// What changed against v2 (measured there: the epilogue skeleton alone cost ~2.5 us per tile, more than the tile's MMAs):
//   * every epilogue warp owns a 4 KB staging buffer and issues its own [32 rows x 32 cols] bulk tensor stores:
//     no named barriers between warps, only __syncwarp + one proxy fence per box;
//   * both 32-column slabs of a warp are read from TMEM by back-to-back tcgen05.ld with ONE wait, and the
//     accumulator is handed back to the MMA warp right after that read (before any math or store);
//   * the bias vector lives in shared memory for the whole kernel (no per-tile global load + block barrier);
//   * dSiLU reads its bf16 pre-activation through TMA loads issued before the accumulator wait (full-line reads,
//     latency hidden behind the main loop) instead of row-strided per-thread global loads;
//   * split-K accumulation leaves through cp.reduce.async.bulk.tensor (fp32 add in L2) instead of per-thread red.v4.
//
// Epilogue modes
//   EPI_GENERIC    bias, SiLU (+ optional bf16 pre-activation store) / dSiLU (reads bf16 pre-activation),
//                  alpha, fp32 (row-periodic) residual, bf16|fp32 out, or fp32 reduce-add (split-K wgrad)
//   EPI_GLU        B rows [n0,n0+128) | [N/2+n0,+128): out = alpha * a * sigmoid(g) (+ optional z store)
// (the row-wise LayerNorm / log-softmax epilogues stay in gemm_tc2.cu)
#include <stdlib.h>
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2;              // 16 KB
constexpr int B_BYTES = BN * BK * 2;              // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 48 KB
constexpr int NEW = 16;                           // epilogue warps
constexpr int MAX_BIAS = 2048;
// Shared-memory layout for an NST-deep operand ring.  NST = 3: 4 KB of staging per epilogue warp (one fp32 box or two bf16
// boxes of 32 columns) + the bias vector.  NST = 4 (plain fp32-output GEMMs: weight gradients, the long-K dgrad): these
// kernels are bound by TMA latency x bytes in flight per SM (a CTA with 144 KB in flight streams 65 GB/s, 1.9 us per stage
// round trip), so the ring gets a fourth stage and the staging shrinks to one 2 KB box of 16 fp32 columns per warp.
template <int NST> struct L3 {
  static constexpr int WB = (NST == 4) ? 2048 : 4096;
  static constexpr int OFF_STG = NST * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_STG + NEW * WB;                       // float[MAX_BIAS] | 2 KB of scratch
  static constexpr int OFF_BAR = OFF_BIAS + ((NST == 4) ? 2048 : MAX_BIAS * 4);
  static constexpr int BYTES = OFF_BAR + 512;                               // base must be 1024-byte aligned (checked)
};
constexpr int NT3 = 64 + NEW * 32;                // 576

enum { EPI_GENERIC = 0, EPI_GLU = 1 };

struct P3 {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_per_split;
  int step_split, step_mt, step_nt;   // one grid stride (gridDim.x units) expressed in (split, m-tile, n-tile) steps: no divisions in the tile loops
  const float* bias;
  int act;
  float alpha;
  const float* residual; int ldr; int res_row_mod;
  int out_bf16, has_pre, accumulate;
  float* a_colsum; float a_colsum_scale;   // MN-major A only: a_colsum[m] += scale * sum_k A(m,k) from the smem A tiles
  int debug;       // EEC_GEMM_DEBUG bitmask (perf triage only): 1 = no bulk store issue, 2 = no staging, 4 = no activation math
  long long* tl;   // EEC_GEMM_TL=1 (perf triage only): clock64 accumulators of CTA 0, see gemm_tc3()
  DropArgs drop;   // DROP instantiations only: dropout right after the activation, element index m*N + n
  ActiveItems act_items;   // m-tiles past the active-item limit are skipped by all three roles (early-exit inference)
  int l2_pf;               // operand L2 prefetch distance in k-blocks (0 = off): the producer asks L2 for k-block kb + l2_pf when it loads kb
};

// per-warp staging: values -> swizzled smem box -> bulk tensor store / reduce of a [32 rows x 32 cols] box
struct WStager {
  uint8_t* buf;
  int lane;
  int sub;        // next bf16 sub-buffer (2 KB each)
  bool pend_f32;  // the newest outstanding bulk group reads the whole 4 KB buffer
  int debug;      // EEC_GEMM_DEBUG (perf triage only): 1 = no bulk store issue, 2 = no staging at all

  __device__ __forceinline__ void store_bf16(const CUtensorMap* tm, int x, int y, const float (&v)[32]) {
    if (debug & 2) return;
    if (lane == 0) {
      if (pend_f32) bulk_wait_read<0>();
      else bulk_wait_read<1>();   // the group before the newest one used this sub-buffer
    }
    __syncwarp();
    uint8_t* b = buf + sub * 2048;
    uint8_t* row = b + lane * 64;
    const int sw = (lane >> 1) & 3;
    if (!(debug & 16))
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
      *reinterpret_cast<uint4*>(row + ((g ^ sw) << 4)) = u;
    }
    if (!(debug & 8)) fence_proxy_async();
    __syncwarp();
    if (lane == 0 && !(debug & 1)) {
      tma_store_2d(tm, b, x, y);
      bulk_commit();
    }
    sub ^= 1;
    pend_f32 = false;
  }
  template <bool REDUCE>
  __device__ __forceinline__ void store_f32(const CUtensorMap* tm, int x, int y, const float (&v)[32]) {
    if (debug & 2) return;
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
    uint8_t* row = buf + lane * 128;
    const int sw = lane & 7;
#pragma unroll
    for (int g = 0; g < 8; ++g)
      *reinterpret_cast<float4*>(row + ((g ^ sw) << 4)) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0 && !(debug & 1)) {
      if (REDUCE) tma_reduce_add_2d(tm, buf, x, y);
      else tma_store_2d(tm, buf, x, y);
      bulk_commit();
    }
    pend_f32 = true;
  }
};

__device__ __forceinline__ void add_bias32(float (&v)[32], const float* b) {
  const float4* bp = reinterpret_cast<const float4*>(b);   // warp-uniform address: smem broadcast
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 f = bp[g];
    v[g * 4] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
  }
}
// 16 values of this lane's row -> column half `half` of a [32 rows x 32 cols] staging box (bf16: 64B swizzle, fp32: 128B swizzle)
__device__ __forceinline__ void put16_bf16(uint8_t* box, int half, int lane, const float (&x)[16]) {
  uint8_t* row = box + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(x[c * 8 + 2 * e], x[c * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(row + (((half * 2 + c) ^ sw) << 4)) = u;
  }
}
__device__ __forceinline__ void put16_f32(uint8_t* box, int half, int lane, const float (&x)[16]) {
  uint8_t* row = box + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<float4*>(row + (((half * 4 + c) ^ sw) << 4)) = make_float4(x[c * 4], x[c * 4 + 1], x[c * 4 + 2], x[c * 4 + 3]);
}
// x[j] *= alpha * dSiLU(h[j]) with h = 16 bf16 of this lane's row (column half `half`) of a 64B-swizzled [32 x 32] bf16 box.
// With t = tanh(h/2):  alpha * dSiLU(h) = c (1 + t) + (h/2) c (1 - t^2),  c = alpha / 2: one MUFU + six packed fp32x2 ops per pair
__device__ __forceinline__ void mul_dsilu16(float (&x)[16], const uint8_t* box, int half, int lane, float alpha) {
  const uint8_t* row = box + lane * 64;
  const int sw = (lane >> 1) & 3;
  const float c = 0.5f * alpha;
  const float2 c2 = make_float2(c, c), nc2 = make_float2(-c, -c), half2 = make_float2(0.5f, 0.5f);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + (((half * 2 + cc) ^ sw) << 4));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 hh = __fmul2_rn(__bfloat1622float2(h[e]), half2);
      const float2 t = make_float2(tanh_fast(hh.x), tanh_fast(hh.y));
      const float2 r = __ffma2_rn(t, c2, c2);                       // c (1 + t)
      const float2 q = __ffma2_rn(__fmul2_rn(t, nc2), t, c2);       // c (1 - t^2)
      const float2 d = __ffma2_rn(hh, q, r);
      const float2 o = __fmul2_rn(make_float2(x[cc * 8 + 2 * e], x[cc * 8 + 2 * e + 1]), d);
      x[cc * 8 + 2 * e] = o.x; x[cc * 8 + 2 * e + 1] = o.y;
    }
  }
}

// backward of ReLU through the activation OUTPUT a = relu(h) (a > 0 <=> h > 0): x = alpha * x * [a > 0]
__device__ __forceinline__ void mul_drelu16(float (&x)[16], const uint8_t* box, int half, int lane, float alpha) {
  const uint8_t* row = box + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + (((half * 2 + cc) ^ sw) << 4));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = __bfloat1622float2(h[e]);
      x[cc * 8 + 2 * e] = a.x > 0.f ? x[cc * 8 + 2 * e] * alpha : 0.f;
      x[cc * 8 + 2 * e + 1] = a.y > 0.f ? x[cc * 8 + 2 * e + 1] * alpha : 0.f;
    }
  }
}

// work unit u = split * (m_tiles * n_tiles) + mt * n_tiles + nt, walked with a stride of gridDim.x units
struct UnitIter {
  int u, mt, nt, split;
  __device__ __forceinline__ UnitIter(const P3& p) {
    u = blockIdx.x;
    const int tiles = p.m_tiles * p.n_tiles;
    split = u / tiles;
    const int t = u - split * tiles;
    mt = t / p.n_tiles;
    nt = t - mt * p.n_tiles;
  }
  __device__ __forceinline__ void next(const P3& p) {
    u += gridDim.x;
    split += p.step_split;
    mt += p.step_mt;
    nt += p.step_nt;
    if (nt >= p.n_tiles) { nt -= p.n_tiles; ++mt; }
    if (mt >= p.m_tiles) { mt -= p.m_tiles; ++split; }
  }
};

template <bool A_KMAJ, bool B_KMAJ, int EPI, int NST, bool DROP = false>
__global__ void __launch_bounds__(NT3, 1) gemm_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          const __grid_constant__ CUtensorMap tmC,   // main output (box 32 x 32)
                                                          const __grid_constant__ CUtensorMap tmP,   // bf16 pre-activation: store (SiLU/GLU) or load (dSiLU)
                                                          const P3 p) {
  pdl_trigger();
  constexpr int NSTAGE = NST;
  using LY = L3<NST>;
  extern __shared__ __align__(1024) uint8_t smem_al[];
  uint8_t* smem = smem_al;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) { printf("eec: gemm_tc3 smem base not 1024-aligned\n"); __trap(); }
  float* bias_s = reinterpret_cast<float*>(smem + LY::OFF_BIAS);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + LY::OFF_BAR);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tfull_bar = empty_bar + NSTAGE;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* load_bar = tempty_bar + 2;        // [NEW] per-warp pre-activation loads
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(load_bar + NEW);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
#ifndef EEC_GEMM_TIMELINE
  (void)gtime;
#endif
  if (p.tl && blockIdx.x == 0 && threadIdx.x == 0) p.tl[8] = gtime();
  const int tiles = p.m_tiles * p.n_tiles;
  const int n_units = tiles * p.splits;
  const int total_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    // with a_colsum the 16 epilogue warps also read every A stage: the slot is free after the MMAs AND those readers
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], (!A_KMAJ && p.a_colsum) ? 1 + NEW : 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], NEW); }
    for (int w = 0; w < NEW; ++w) mbar_init(&load_bar[w], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  pdl_wait();   // everything above overlapped the previous kernel's tail; from here on its outputs are read
  if (p.bias) {
    for (int i = threadIdx.x; i < p.N; i += NT3) bias_s[i] = p.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (p.tl && blockIdx.x == 0 && threadIdx.x == 0) p.tl[9] = gtime();
  const int mt_eff = (active_rows(p.act_items, p.M) + BM - 1) / BM;   // == p.m_tiles unless an active-item limit is set

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;   // stage ring position + parity of the "empty" wait
      for (UnitIter ui(p); ui.u < n_units; ui.next(p)) {
        if (ui.mt >= mt_eff) continue;
        const int m0 = ui.mt * BM;
        const int n0 = (EPI == EPI_GLU) ? ui.nt * 128 : ui.nt * BN;
        const int kb0 = ui.split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, s = (s + 1 == NSTAGE) ? 0 : s + 1, ph ^= (s == 0)) {
          mbar_wait(&empty_bar[s], ph);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (p.debug & 32) { mbar_arrive(&full_bar[s]); continue; }   // triage: MMAs run on whatever is in smem
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          const int k = kb * BK;
          if (p.l2_pf && kb + p.l2_pf < kb1) {
            // the ring holds NST x 48 KB per SM and an HBM round trip is ~1.9 us, so a long-K GEMM streams at (bytes in flight) /
            // latency; asking L2 for the operands a few k-blocks ahead turns the ring's loads into L2 hits (shorter round trip)
            const int kp = k + p.l2_pf * BK;
            if (A_KMAJ) tma_prefetch_l2_2d(&tmA, kp, m0);
            else { tma_prefetch_l2_2d(&tmA, m0, kp); tma_prefetch_l2_2d(&tmA, m0 + 64, kp); }
            if (B_KMAJ) {
              tma_prefetch_l2_2d(&tmB, kp, n0);
              if (EPI == EPI_GLU) tma_prefetch_l2_2d(&tmB, kp, p.N / 2 + n0);
            } else {
#pragma unroll
              for (int a = 0; a < BN / 64; ++a) tma_prefetch_l2_2d(&tmB, n0 + a * 64, kp);
            }
          }
          if (A_KMAJ) {
            tma_load_2d(sa, &tmA, &full_bar[s], k, m0);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[s], m0, k);
            tma_load_2d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k);
          }
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
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, !A_KMAJ, !B_KMAJ);
      uint32_t ut = 0;
      int s = 0;
      uint32_t ph = 0;
      // UMMA descriptors of every stage, built once (per MMA only the 16-element k offset is added)
      uint64_t adesc[NSTAGE], bdesc[NSTAGE];
#pragma unroll
      for (int i = 0; i < NSTAGE; ++i) {
        const uint32_t sa = smem_u32(smem + i * STAGE_BYTES), sb = sa + A_BYTES;
        adesc[i] = A_KMAJ ? make_smem_desc(sa, 0, 1024) : make_smem_desc(sa, 8192, 1024);
        bdesc[i] = B_KMAJ ? make_smem_desc(sb, 0, 1024) : make_smem_desc(sb, 8192, 1024);
      }
      constexpr uint64_t A_KSTEP = (A_KMAJ ? 32 : 2048) >> 4, B_KSTEP = (B_KMAJ ? 32 : 2048) >> 4;   // descriptor address field is in 16-byte units
#ifdef EEC_GEMM_TIMELINE
      const bool prof = p.tl && blockIdx.x == 0;
#else
      constexpr bool prof = false;
#endif
      long long w_tempty = 0, w_full = 0, t_ = 0, t_begin = prof ? clock64() : 0;
      for (UnitIter ui(p); ui.u < n_units; ui.next(p), ++ut) {
        if (ui.mt >= mt_eff) { --ut; continue; }   // (skipped units do not advance the accumulator / phase counters)
        const int kb0 = ui.split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const uint32_t acc = ut & 1;
        if (prof) t_ = clock64();
        mbar_wait(&tempty_bar[acc], ((ut >> 1) & 1) ^ 1);   // epilogue has read this accumulator out of TMEM
        if (prof) w_tempty += clock64() - t_;
        if (!(p.debug & 64)) tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (prof) t_ = clock64();
          mbar_wait(&full_bar[s], ph);
          if (prof) w_full += clock64() - t_;
          if (!(p.debug & 128)) tc_fence_after();
          uint64_t ad = adesc[0], bd = bdesc[0];   // (explicit selects keep the descriptor arrays in registers)
#pragma unroll
          for (int i = 1; i < NSTAGE; ++i)
            if (s == i) { ad = adesc[i]; bd = bdesc[i]; }
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, ad + k * A_KSTEP, bd + k * B_KSTEP, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
      }
      if (prof) p.tl[10] = gtime();
      if (prof) { p.tl[0] = clock64() - t_begin; p.tl[1] = w_tempty; p.tl[2] = w_full; p.tl[3] = ut; }
    }
  } else {
    // ================================================================== epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;
    const int cg = e >> 2;
    WStager st;
    st.buf = smem + LY::OFF_STG + e * LY::WB;
    st.lane = lane;
    st.sub = 0;
    st.pend_f32 = false;
    st.debug = p.debug;
    uint64_t* lbar = &load_bar[e];
    uint32_t lphase = 0;
    float v[32], w[32];
    uint32_t ut = 0;
#ifdef EEC_GEMM_TIMELINE
    const bool prof = p.tl && blockIdx.x == 0 && e == 0 && lane == 0;
    long long e_wait = 0, e_ld = 0, e_rest = 0, t0_ = 0, t1_ = 0, t2_ = 0;
#define TL_STAMP(t) do { if (prof) t = clock64(); } while (0)
#define TL_ACC(a, d) do { if (prof) a += (d); } while (0)
#else
#define TL_STAMP(t) do { } while (0)
#define TL_ACC(a, d) do { } while (0)
#endif
    int cs = 0;             // a_colsum: this warp's position in the operand ring (walks every k-block of every unit)
    uint32_t cph = 0;
    for (UnitIter ui(p); ui.u < n_units; ui.next(p), ++ut) {
      if (ui.mt >= mt_eff) { --ut; continue; }
      TL_STAMP(t0_); TL_ACC(e_rest, ut ? t0_ - t2_ : 0);
      const int split = ui.split;
      const int m0 = ui.mt * BM;
      const int nt = ui.nt;
      if (!A_KMAJ && p.a_colsum) {
        // bias gradient of the layer whose weight gradient this GEMM computes: column sums of the MN-major A operand
        // ([64 k][128 m] bf16 per stage, 128B-swizzled) taken while the tile sits in shared memory.  Thread (cm, kq) adds
        // 16 k-rows of column cm per k-block; units with n-tile 0 publish (one atomic per column per unit).
        const int tid_e = threadIdx.x - 64, cm = tid_e & 127, kq = tid_e >> 7;
        const int kb0 = split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const uint32_t coff = (uint32_t)((cm >> 6) * 8192 + (((cm & 63) & 7) * 2));
        const int cchunk = (cm & 63) >> 3;
        float csum = 0.f;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[cs], cph);
          if (nt == 0) {
            const uint8_t* a = smem + cs * STAGE_BYTES + coff;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
              const int k = kq * 16 + kk;
              csum += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(a + k * 128 + ((cchunk ^ (k & 7)) << 4)));
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[cs]);
          if (++cs == NSTAGE) { cs = 0; cph ^= 1; }
        }
        if (nt == 0) {
          bias_s[kq * 128 + cm] = csum;     // (a weight-gradient GEMM has no bias vector: the area is free)
          bar_sync(1, NEW * 32);
          if (kq == 0 && m0 + cm < p.M)
            atomicAdd(p.a_colsum + m0 + cm, p.a_colsum_scale * (bias_s[cm] + bias_s[128 + cm] + bias_s[256 + cm] + bias_s[384 + cm]));
          bar_sync(1, NEW * 32);
        }
      }
      const int n0 = (EPI == EPI_GLU) ? nt * 128 : nt * BN;
      const uint32_t acc = ut & 1;
      const int row0 = m0 + q * 32;
      const int m = row0 + lane;
      const bool valid = m < p.M;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const bool first_split = (split == 0);

      if (EPI == EPI_GENERIC) {
        // 64 accumulator columns per warp = 2 output boxes of 32 columns, read from TMEM 16 columns at a time with the next
        // read in flight while the current 16 values are processed (TMEM drains at 64 B/clk/SM: it must overlap the math)
        const int cb = cg * 64;
        const int nb = n0 + cb;
        const int nslab = (nb + 32 < p.N) ? 2 : (nb < p.N) ? 1 : 0;   // warp-uniform
        const bool drelu = p.act == EEC_ACT_DRELU, relu = p.act == EEC_ACT_RELU;
        const bool dsilu = p.act == EEC_ACT_DSILU || drelu, silu = p.act == EEC_ACT_SILU;   // (dsilu: "multiply by a TMA-loaded box" data flow)
        if (dsilu && nslab) {
          // pre-activation boxes of both slabs: in flight while the main loop of this tile runs
          if (lane == 0) {
            bulk_wait_read<0>();   // the previous tile's stores have finished reading the staging buffer
            mbar_expect_tx(lbar, nslab * 2048);
            tma_load_2d(st.buf, &tmP, lbar, nb, row0);
            if (nslab == 2) tma_load_2d(st.buf + 2048, &tmP, lbar, nb + 32, row0);
            // the staging buffer is busy until this tile's stores have been read, so the NEXT tile's boxes cannot be loaded yet:
            // pull them into L2 now, so that the load issued at the next tile start is an L2 hit instead of an HBM round trip
            UnitIter nx = ui;
            nx.next(p);
            if (nx.u < n_units) {
              const int nrow = nx.mt * BM + q * 32, nn = nx.nt * BN + cb;
              if (nn < p.N) tma_prefetch_l2_2d(&tmP, nn, nrow);
              if (nn + 32 < p.N) tma_prefetch_l2_2d(&tmP, nn + 32, nrow);
            }
          }
          __syncwarp();
        }
        // dropout keep-mask of this thread's 64 outputs: four 16-bit words written by eec_dropout_bits ([n/16][m] layout: the 32 rows
        // of a warp read 64 contiguous bytes); loaded before the accumulator wait so that the latency hides behind the main loop
        uint32_t dw[2] = {0xffffffffu, 0xffffffffu};
        if (DROP && valid) {
          const uint16_t* db = reinterpret_cast<const uint16_t*>(p.drop.bits);
#pragma unroll
          for (int ss = 0; ss < 4; ++ss) {
            const int n = nb + ss * 16;
            const uint32_t wv = (n < p.N) ? (uint32_t)db[(long)(n >> 4) * p.M + m] : 0xffffu;
            dw[ss >> 1] = (ss & 1) ? ((dw[ss >> 1] & 0xffffu) | (wv << 16)) : ((dw[ss >> 1] & 0xffff0000u) | wv);
          }
        }
        mbar_wait(&tfull_bar[acc], (ut >> 1) & 1);
        TL_STAMP(t1_); TL_ACC(e_wait, t1_ - t0_);
        tc_fence_after();
        uint32_t ra[16], rb[16];
        const uint32_t tcol = trow + cb;
        if (p.debug & 256) {   // triage: no TMEM reads at all
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          continue;
        }
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
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // accumulator is in registers: the MMA warp may reuse it
            TL_STAMP(t2_); TL_ACC(e_ld, t2_ - t1_);
          }
          const int sl = ss >> 1, half = ss & 1;
          if (sl >= nslab) continue;
          const int n = nb + ss * 16;
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(cur[j]);
          if (NST == 4) {
            // plain fp32 output (store or split-K reduce-add): one 2 KB box of 16 columns per sub-slab (64-byte rows, 64B swizzle)
            if (p.alpha != 1.0f) {
#pragma unroll
              for (int j = 0; j < 16; ++j) x[j] *= p.alpha;
            }
            if (p.residual && first_split && valid) {
              const long rr = p.res_row_mod ? (m % p.res_row_mod) : m;
              const float* rp = p.residual + rr * p.ldr + n;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 f = *reinterpret_cast<const float4*>(rp + g * 4);
                x[g * 4] += f.x; x[g * 4 + 1] += f.y; x[g * 4 + 2] += f.z; x[g * 4 + 3] += f.w;
              }
            }
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            uint8_t* row = st.buf + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<float4*>(row + ((c ^ sw) << 4)) = make_float4(x[c * 4], x[c * 4 + 1], x[c * 4 + 2], x[c * 4 + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (p.accumulate) tma_reduce_add_2d(&tmC, st.buf, n, row0);
              else tma_store_2d(&tmC, st.buf, n, row0);
              bulk_commit();
            }
            continue;
          }
          if (first_split && p.bias) {
            const float4* bp = reinterpret_cast<const float4*>(bias_s + n);   // warp-uniform address: smem broadcast
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 f = bp[g];
              const float2 lo = __fadd2_rn(make_float2(x[g * 4], x[g * 4 + 1]), make_float2(f.x, f.y));
              const float2 hi = __fadd2_rn(make_float2(x[g * 4 + 2], x[g * 4 + 3]), make_float2(f.z, f.w));
              x[g * 4] = lo.x; x[g * 4 + 1] = lo.y; x[g * 4 + 2] = hi.x; x[g * 4 + 3] = hi.y;
            }
          }
          // staging sub-buffers (2 KB bf16 boxes): SiLU+pre-activation store: pre -> 0, out -> 1; dSiLU: slab sl reuses the
          // sub-buffer its pre-activation arrived in; otherwise the two slabs alternate.  fp32 boxes take the whole 4 KB.
          const int osub = p.has_pre ? 1 : dsilu ? sl : st.sub;
          if (half == 0 && !dsilu) {
            if (lane == 0) {
              if (!p.out_bf16 || p.has_pre) bulk_wait_read<0>();
              else bulk_wait_read<1>();
            }
            __syncwarp();
          }
          if (silu) {
            if (p.has_pre) put16_bf16(st.buf, half, lane, x);
            if (!(p.debug & 4)) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {   // x * sigmoid(x) = h + h * tanh(h), h = x / 2; packed fp32x2 FMUL2 / FFMA2 (sm_100)
                const float2 h = __fmul2_rn(make_float2(x[j], x[j + 1]), make_float2(0.5f, 0.5f));
                const float2 r = __ffma2_rn(h, make_float2(tanh_fast(h.x), tanh_fast(h.y)), h);
                x[j] = r.x; x[j + 1] = r.y;
              }
            }
          } else if (dsilu) {
            if (ss == 0) {
              mbar_wait(lbar, lphase);
              lphase ^= 1;
            }
            if (drelu) mul_drelu16(x, st.buf + sl * 2048, half, lane, p.alpha);
            else mul_dsilu16(x, st.buf + sl * 2048, half, lane, p.alpha);   // (alpha folded in)
          } else if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (DROP) drop_apply_bits<16>(x, (dw[ss >> 1] >> ((ss & 1) * 16)) & 0xffffu, p.drop.scale);
          if (p.alpha != 1.0f && !dsilu) {
            const float2 al = make_float2(p.alpha, p.alpha);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float2 r = __fmul2_rn(make_float2(x[j], x[j + 1]), al);
              x[j] = r.x; x[j + 1] = r.y;
            }
          }
          if (p.residual && first_split && valid) {
            const long rr = p.res_row_mod ? (m % p.res_row_mod) : m;
            const float* rp = p.residual + rr * p.ldr + n;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 f = *reinterpret_cast<const float4*>(rp + g * 4);
              x[g * 4] += f.x; x[g * 4 + 1] += f.y; x[g * 4 + 2] += f.z; x[g * 4 + 3] += f.w;
            }
          }
          uint8_t* ob = p.out_bf16 ? st.buf + osub * 2048 : st.buf;
          if (p.debug & 2) continue;
          if (p.out_bf16) put16_bf16(ob, half, lane, x);
          else put16_f32(ob, half, lane, x);
          if (half == 1) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(p.debug & 1)) {
              const int nx = nb + sl * 32;
              if (p.has_pre) tma_store_2d(&tmP, st.buf, nx, row0);   // same bulk group as the output box
              if (p.accumulate) tma_reduce_add_2d(&tmC, ob, nx, row0);
              else tma_store_2d(&tmC, ob, nx, row0);
              bulk_commit();
            }
            st.sub ^= 1;
          }
        }
      } else {  // EPI_GLU: v = "a" half, w = gate half of the same 32 output channels
        const int c = cg * 32;
        mbar_wait(&tfull_bar[acc], (ut >> 1) & 1);
        TL_STAMP(t1_); TL_ACC(e_wait, t1_ - t0_);
        tc_fence_after();
        tmem_ld32x2(trow + c, trow + 128 + c, v, w);
        TL_STAMP(t2_); TL_ACC(e_ld, t2_ - t1_);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (p.bias) {
          add_bias32(v, bias_s + n0 + c);
          add_bias32(w, bias_s + p.N / 2 + n0 + c);
        }
        if (p.has_pre) {
          st.store_bf16(&tmP, n0 + c, row0, v);
          st.store_bf16(&tmP, p.N / 2 + n0 + c, row0, w);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = p.alpha * v[j] * sigmoid_fast(w[j]);
        if (p.out_bf16) st.store_bf16(&tmC, n0 + c, row0, v);
        else st.store_f32<false>(&tmC, n0 + c, row0, v);
      }
    }
#ifdef EEC_GEMM_TIMELINE
    if (prof) { e_rest += clock64() - t2_; p.tl[4] = e_wait; p.tl[5] = e_ld; p.tl[6] = e_rest; }
#endif
    if (lane == 0) bulk_wait_all();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (p.tl && blockIdx.x == 0 && threadIdx.x == 32) p.tl[11] = gtime();
}


// ================================================================================================================
// LayerNorm-tail GEMM (N == 256, K-major operands):  x = residual + alpha * (A W^T + bias) -> fp32 C;  LayerNorm(x) -> ln_out
// (out-projection, pointwise conv 2 and FFN down-projection of a conformer layer: TA:202, :65, :107 followed by :151/:42/:103/:211).
// 8 epilogue warps: warp (quarter q, half h) owns rows [32q, 32q+32) x columns [128h, 128h+128) of the tile and keeps its
// 128 x-values per thread in REGISTERS across the row-statistics exchange, so the accumulator is read from TMEM exactly once
// (TMEM drains at 64 B/clk: the v2 kernel's TMEM round trip cost two more passes).  The fp32 residual arrives through
// per-warp TMA loads into two 4 KB staging buffers (first two slabs prefetched before the accumulator is ready); the same
// buffers stage the x store (each lane overwrites the residual row it just consumed) and the LayerNorm output store.
constexpr int LN_WARPS = 8;
constexpr int LN_NT = 64 + LN_WARPS * 32;                 // 320
constexpr int LN_WBUF = 8192;                             // two 4 KB staging buffers per epilogue warp
constexpr int LN_NSTAGE = 3;
constexpr int LN_OFF_STG = LN_NSTAGE * STAGE_BYTES;
constexpr int LN_OFF_VEC = LN_OFF_STG + LN_WARPS * LN_WBUF;   // float[3][256]: bias, gamma, beta
constexpr int LN_OFF_XCH = LN_OFF_VEC + 3 * 256 * 4;          // float[2 halves][128 rows][2]
constexpr int LN_OFF_BAR = LN_OFF_XCH + 2 * 128 * 2 * 4;
constexpr int LN_SMEM_BYTES = LN_OFF_BAR + 512 + 1024;

struct PLN {
  int M, K, m_tiles;
  const float* bias;
  float alpha;
  int has_res, ln_bf16;
  const float* ln_gamma; const float* ln_beta; float* ln_mean; float* ln_rstd;
  int l2_pf;       // A-operand L2 prefetch distance in k-blocks (0 = off), see gemm_tc3_kernel
  DropArgs drop;   // DROP instantiation only: x = residual + alpha * dropout(A W^T + bias), element index m*256 + n
  ActiveItems act_items;
};

template <bool DROP>
__global__ void __launch_bounds__(LN_NT, 1) gemm_ln3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                            const __grid_constant__ CUtensorMap tmC,   // fp32 x out (box 32 x 32)
                                                            const __grid_constant__ CUtensorMap tmR,   // fp32 residual in (box 32 x 32)
                                                            const __grid_constant__ CUtensorMap tmL,   // LayerNorm out (box 32 x 32)
                                                            const PLN p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* vecs = reinterpret_cast<float*>(smem + LN_OFF_VEC);
  float* xch = reinterpret_cast<float*>(smem + LN_OFF_XCH);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + LN_OFF_BAR);
  uint64_t* empty_bar = full_bar + LN_NSTAGE;
  uint64_t* tfull_bar = empty_bar + LN_NSTAGE;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* res_bar = tempty_bar + 2;         // [LN_WARPS][2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 2 * LN_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmR);
    for (int s = 0; s < LN_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], LN_WARPS); }
    for (int w = 0; w < 2 * LN_WARPS; ++w) mbar_init(&res_bar[w], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  pdl_wait();   // everything above overlapped the previous kernel's tail
  for (int i = threadIdx.x; i < 256; i += LN_NT) {
    vecs[i] = p.bias ? p.bias[i] : 0.f;
    vecs[256 + i] = p.ln_gamma[i];
    vecs[512 + i] = p.ln_beta[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int m_tiles = (active_rows(p.act_items, p.M) + BM - 1) / BM;   // == p.m_tiles unless an active-item limit is set

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        const int m0 = mt * BM;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph);
          uint8_t* sa = smem + s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          if (p.l2_pf && kb + p.l2_pf < total_kb) tma_prefetch_l2_2d(&tmA, (kb + p.l2_pf) * BK, m0);   // (the weight tile is L2 resident anyway)
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m0);
          tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], kb * BK, 0);
          if (++s == LN_NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, false, false);
      uint32_t ut = 0, ph = 0;
      int s = 0;
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++ut) {
        const uint32_t acc = ut & 1;
        mbar_wait(&tempty_bar[acc], ((ut >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 0, 1024), make_smem_desc(sb + k * 32, 0, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (++s == LN_NSTAGE) { s = 0; ph ^= 1; }
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
    const int cb = half * 128;
    uint8_t* buf[2] = {smem + LN_OFF_STG + e * LN_WBUF, smem + LN_OFF_STG + e * LN_WBUF + 4096};
    uint64_t* rbar = res_bar + 2 * e;
    const int sw = lane & 7;
    float x[128];   // this thread's 128 x-values (its row, its column half): TMEM is read once, the values never leave registers
    uint32_t ut = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++ut) {
      const int m0 = mt * BM;
      const int row0 = m0 + q * 32;
      const int m = m0 + r;
      const bool valid = m < p.M;
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
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // the whole accumulator slice is in registers
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
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool AK, bool BK_, int EPI, int NST, bool DROP = false>
int launch3n(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const CUtensorMap& tp, const P3& p, int grid,
             cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(gemm_tc3_kernel<AK, BK_, EPI, NST, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L3<NST>::BYTES));
    attr_set = true;
  }
  launch_pdl(gemm_tc3_kernel<AK, BK_, EPI, NST, DROP>, dim3(grid), dim3(NT3), L3<NST>::BYTES, st, ta, tb, tc_, tp, p);
  EEC_LAUNCH_CHECK();
  return 0;
}
template <bool AK, bool BK_, int EPI>
int launch3(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const CUtensorMap& tp, const P3& p, int grid,
            cudaStream_t st, bool four_stages = false) {
  if (EPI == EPI_GENERIC && four_stages) return launch3n<AK, BK_, EPI_GENERIC, 4>(ta, tb, tc_, tp, p, grid, st);
  return launch3n<AK, BK_, EPI, 3>(ta, tb, tc_, tp, p, grid, st);
}

int g_sms3 = 0;

}  // namespace

// GENERIC / GLU epilogues of eec_gemm (bf16 operands); the caller (gemm_tc2) has validated the descriptor
int gemm_tc3(const eec_gemm_desc* d, cudaStream_t st) {
  if (gemm_ws_ok(d)) return gemm_ws(d, st);   // K = 256 projections with a bf16 output: weight-stationary kernel (gemm_ws.cu)
  if (gemm_pair_ok(d, st)) return gemm_pair(d, st);   // fp32-output data / weight gradients: CTA-pair streaming kernel (gemm_pair.cu)
  const int epi = (d->act == EEC_ACT_GLU) ? EPI_GLU : EPI_GENERIC;
  EEC_CHECK_ARG(!d->bias || d->N <= MAX_BIAS, "gemm_tc3: N (%d) > %d with a bias vector is unsupported", d->N, MAX_BIAS);
  if (!g_sms3) {
    int dev = 0;
    EEC_CUDA(cudaGetDevice(&dev));
    EEC_CUDA(cudaDeviceGetAttribute(&g_sms3, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap ta, tb, tcm, tpm;
  if (d->a_kmajor) { if (int r = get_tmap_2d(&ta, d->A, d->K, d->M, (uint64_t)d->lda * 2, 64, 128)) return r; }
  else { if (int r = get_tmap_2d(&ta, d->A, d->M, d->K, (uint64_t)d->lda * 2, 64, 64)) return r; }
  if (d->b_kmajor) {
    if (int r = get_tmap_2d(&tb, d->B, d->K, d->N, (uint64_t)d->ldb * 2, 64, epi == EPI_GLU ? 128 : 256)) return r;
  } else {
    if (int r = get_tmap_2d(&tb, d->B, d->N, d->K, (uint64_t)d->ldb * 2, 64, 64)) return r;
  }
  const int n_out = (epi == EPI_GLU) ? d->N / 2 : d->N;
  const bool out_bf16 = d->out_dtype == EEC_BF16;
  const bool store_pre_ = d->preact && (d->act == EEC_ACT_SILU || epi == EPI_GLU);
  // plain fp32-output GEMMs (weight gradients, the long-K dgrad) run the 4-stage ring with 16-column output boxes
  static int nst_env = -1;
  if (nst_env < 0) { const char* e = getenv("EEC_GEMM_STAGES"); nst_env = e ? atoi(e) : 0; }
  const bool four = nst_env != 3 && epi == EPI_GENERIC && !out_bf16 && !d->bias && d->act == EEC_ACT_NONE && !store_pre_;
  if (four) { if (int r = get_tmap_box(&tcm, d->C, false, (uint64_t)n_out, (uint64_t)d->M, (uint64_t)d->ldc, 16, 32, 2 /*SWIZZLE_64B*/)) return r; }
  else if (int r = get_tmap_box32(&tcm, d->C, out_bf16, (uint64_t)n_out, (uint64_t)d->M, (uint64_t)d->ldc)) return r;
  const bool store_pre = d->preact && (d->act == EEC_ACT_SILU || epi == EPI_GLU);
  tpm = tcm;
  if (store_pre || d->act == EEC_ACT_DSILU || d->act == EEC_ACT_DRELU) {
    if (int r = get_tmap_box32(&tpm, d->preact, true, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldp)) return r;
  }
  P3 p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.m_tiles = cdiv(d->M, BM);
  p.n_tiles = (epi == EPI_GLU) ? d->N / 256 : cdiv(d->N, BN);
  const int total_kb = cdiv(d->K, BK);
  int splits = 1;
  if (d->accumulate) {
    const int tiles = p.m_tiles * p.n_tiles;
    if (tiles < g_sms3 && total_kb >= 16) splits = min(cdiv(total_kb, 8), max(1, g_sms3 / tiles));
  }
  p.kb_per_split = cdiv(total_kb, splits);
  p.splits = cdiv(total_kb, p.kb_per_split);
  p.bias = d->bias; p.act = d->act; p.alpha = d->alpha;
  p.residual = d->residual; p.ldr = d->ldr; p.res_row_mod = d->res_row_mod;
  p.out_bf16 = out_bf16; p.has_pre = store_pre; p.accumulate = d->accumulate;
  EEC_CHECK_ARG(!d->a_colsum || (!d->a_kmajor && !d->b_kmajor && !d->bias && epi == EPI_GENERIC),
                "gemm_tc3: a_colsum needs the weight-gradient form (MN-major A and B, no bias)");
  p.a_colsum = d->a_colsum; p.a_colsum_scale = d->a_colsum_scale;
  p.act_items = (d->a_kmajor && !d->accumulate) ? active_items(st) : ActiveItems{nullptr, 0, 0};   // forward form only: rows of A are frames
  p.drop = make_drop(d->drop_state, d->drop_p, d->drop_site);
  p.drop.bits = d->drop_bits;
  if (p.drop.state) {
    EEC_CHECK_ARG(epi == EPI_GENERIC && !four && d->a_kmajor && !d->accumulate && d->N % 16 == 0,
                  "gemm_tc3: dropout is fused into the SiLU / dSiLU / plain bf16-output epilogues of K-major-A GEMMs only");
    EEC_CHECK_ARG(d->drop_bits != nullptr, "gemm (tensor-core path): dropout needs the keep-mask words of eec_dropout_bits(R = M, C = N, Cs = N, W = 16) in drop_bits");
  }
  const int n_units = p.m_tiles * p.n_tiles * p.splits;
  const int grid = min(n_units, g_sms3);
  {
    const int tiles = p.m_tiles * p.n_tiles;
    p.step_split = grid / tiles;
    const int rem = grid - p.step_split * tiles;
    p.step_mt = rem / p.n_tiles;
    p.step_nt = rem - p.step_mt * p.n_tiles;
  }
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("EEC_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.debug = dbg;
  static int pf_env = -1;
  if (pf_env < 0) { const char* e = getenv("EEC_GEMM_L2PF"); pf_env = e ? atoi(e) : 0; }
  p.l2_pf = (p.kb_per_split >= 16) ? pf_env : 0;   // long-K forms only (weight gradients, K = 2048 dgrad)
  static int tl_env = -1;
  static long long* tl_buf = nullptr;
  if (tl_env < 0) { const char* e = getenv("EEC_GEMM_TL"); tl_env = (e && e[0] == '1') ? 1 : 0; }
  if (tl_env) {
    if (!tl_buf) EEC_CUDA(cudaMalloc(&tl_buf, 128));
    p.tl = tl_buf;
    struct Report {
      long long* b; cudaStream_t s; int M, N, K;
      ~Report() {
        long long h[12];
        if (cudaStreamSynchronize(s) != cudaSuccess || cudaMemcpy(h, b, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return;
        const double t = h[3] ? (double)h[3] : 1.0;
        fprintf(stderr, "gemm_tc3 %dx%dx%d CTA0: %lld units; per unit (clk): MMA thread total %.0f = wait accumulator %.0f + wait operands %.0f + issue;"
                " epilogue warp: wait tfull %.0f, tmem ld %.0f, math+stores %.0f\n", M, N, K, h[3], h[0] / t, h[1] / t, h[2] / t, h[4] / t, h[5] / t, h[6] / t);
        fprintf(stderr, "          CTA0 globaltimer (ns): setup %lld, main loop %lld, drain+teardown %lld, total %lld\n", h[9] - h[8], h[10] - h[9], h[11] - h[10], h[11] - h[8]);
      }
    } report{tl_buf, st, d->M, d->N, d->K};
    if (epi == EPI_GLU) return launch3<true, true, EPI_GLU>(ta, tb, tcm, tpm, p, grid, st);
    if (d->a_kmajor && d->b_kmajor) return launch3<true, true, EPI_GENERIC>(ta, tb, tcm, tpm, p, grid, st, four);
    if (d->a_kmajor && !d->b_kmajor) return launch3<true, false, EPI_GENERIC>(ta, tb, tcm, tpm, p, grid, st, four);
    if (!d->a_kmajor && !d->b_kmajor) return launch3<false, false, EPI_GENERIC>(ta, tb, tcm, tpm, p, grid, st, four);
    return launch3<false, true, EPI_GENERIC>(ta, tb, tcm, tpm, p, grid, st, four);
  }
  if (p.drop.state) {
    if (d->b_kmajor) return launch3n<true, true, EPI_GENERIC, 3, true>(ta, tb, tcm, tpm, p, grid, st);
    return launch3n<true, false, EPI_GENERIC, 3, true>(ta, tb, tcm, tpm, p, grid, st);
  }
  if (epi == EPI_GLU) return launch3<true, true, EPI_GLU>(ta, tb, tcm, tpm, p, grid, st);
  if (d->a_kmajor && d->b_kmajor) return launch3<true, true, EPI_GENERIC>(ta, tb, tcm, tpm, p, grid, st, four);
  if (d->a_kmajor && !d->b_kmajor) return launch3<true, false, EPI_GENERIC>(ta, tb, tcm, tpm, p, grid, st, four);
  if (!d->a_kmajor && !d->b_kmajor) return launch3<false, false, EPI_GENERIC>(ta, tb, tcm, tpm, p, grid, st, four);
  return launch3<false, true, EPI_GENERIC>(ta, tb, tcm, tpm, p, grid, st, four);
}

// LayerNorm-tail epilogue of eec_gemm (N == 256, K-major bf16 operands, fp32 C with ldc 256); validated by gemm_tc2
int gemm_ln3(const eec_gemm_desc* d, cudaStream_t st) {
  EEC_CHECK_ARG(d->res_row_mod == 0, "gemm_ln3: row-periodic residual unsupported");
  EEC_CHECK_ARG(!d->residual || d->ldr == 256, "gemm_ln3: residual must have ld 256");
  if (gemm_lnp_ok(d, st)) return gemm_lnp(d, st);   // CTA-pair version (gemm_lnp.cu); this kernel stays for early-exit compaction / odd SM counts
  if (!g_sms3) {
    int dev = 0;
    EEC_CUDA(cudaGetDevice(&dev));
    EEC_CUDA(cudaDeviceGetAttribute(&g_sms3, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap ta, tb, tcm, trm, tlm;
  if (int r = get_tmap_2d(&ta, d->A, d->K, d->M, (uint64_t)d->lda * 2, 64, 128)) return r;
  if (int r = get_tmap_2d(&tb, d->B, d->K, d->N, (uint64_t)d->ldb * 2, 64, 256)) return r;
  if (int r = get_tmap_box32(&tcm, d->C, false, 256, (uint64_t)d->M, (uint64_t)d->ldc)) return r;
  trm = tcm;
  if (d->residual) { if (int r = get_tmap_box32(&trm, d->residual, false, 256, (uint64_t)d->M, 256)) return r; }
  if (int r = get_tmap_box32(&tlm, d->ln_out, d->ln_dtype == EEC_BF16, 256, (uint64_t)d->M, (uint64_t)d->ld_ln)) return r;
  PLN p{};
  p.M = d->M; p.K = d->K; p.m_tiles = cdiv(d->M, BM);
  p.bias = d->bias; p.alpha = d->alpha; p.has_res = d->residual != nullptr; p.ln_bf16 = d->ln_dtype == EEC_BF16;
  p.ln_gamma = d->ln_gamma; p.ln_beta = d->ln_beta; p.ln_mean = d->ln_mean; p.ln_rstd = d->ln_rstd;
  {
    static int pf_env = -1;
    if (pf_env < 0) { const char* e = getenv("EEC_GEMM_L2PF"); pf_env = e ? atoi(e) : 0; }
    p.l2_pf = (cdiv(d->K, BK) >= 16) ? pf_env : 0;
  }
  p.drop = make_drop(d->drop_state, d->drop_p, d->drop_site);
  p.drop.bits = d->drop_bits;
  if (p.drop.state)
    EEC_CHECK_ARG(d->drop_bits != nullptr, "gemm (LayerNorm tail): dropout needs the keep-mask words of eec_dropout_bits(R = M, C = 256, Cs = 256, W = 32) in drop_bits");
  p.act_items = active_items(st);
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(gemm_ln3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM_BYTES));
    EEC_CUDA(cudaFuncSetAttribute(gemm_ln3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM_BYTES));
    attr_set = true;
  }
  const int grid = min(p.m_tiles, g_sms3);
  if (p.drop.state) launch_pdl(gemm_ln3_kernel<true>, dim3(grid), dim3(LN_NT), LN_SMEM_BYTES, st, ta, tb, tcm, trm, tlm, p);
  else launch_pdl(gemm_ln3_kernel<false>, dim3(grid), dim3(LN_NT), LN_SMEM_BYTES, st, ta, tb, tcm, trm, tlm, p);
  EEC_LAUNCH_CHECK();
  return 0;
}

}  // namespace eec
