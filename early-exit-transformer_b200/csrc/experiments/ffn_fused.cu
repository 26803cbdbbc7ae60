// ffn_fused.cu -- the conformer feed-forward module as ONE persistent tcgen05 kernel (bf16 operands).
//
//   x_out = x_in + alpha * ( SiLU(u W1^T + b1) W2^T + b2 )          (TA:102-109, 185-187, 207-209)
//   ln_out = LayerNorm(x_out; gamma, beta)                           (the next module's / the layer's final LayerNorm)
//
// The [N, 2048] hidden activation never leaves the SM: a CTA owns a 128-row tile and walks the hidden
// dimension in 16 chunks of 128.  Per chunk c
//   G1: H_c[128x128]   = U[128x256] W1_c^T          tcgen05.mma, accumulator in TMEM (double-buffered)
//   epilogue warps    : TMEM -> +b1 -> (optional bf16 pre-activation store for backward) -> SiLU -> bf16 A_c in
//                       128B-swizzled shared memory (double-buffered), i.e. directly the next MMA's A operand
//   G2: Y[128x256]    += A_c[128x128] W2_c^T         accumulator in TMEM columns [0,256) for the whole tile
// and at the end of the tile the epilogue warps apply bias/alpha/residual and the LayerNorm tail and
// leave through TMA stores.  The MMA warp issues G1(c+1) before G2(c), so the tensor pipe works on the next
// chunk while the epilogue warps activate the current one.
//
//   warp 0      : TMA producer (U tile once per tile; W1/W2 sub-tiles [128 rows x 64 k] through a ring of 16 KB slots)
//   warp 1      : MMA issuer + TMEM owner (512 columns: Y 256 | H0 128 | H1 128)
//   warps 2..9  : epilogue; quarter = warp%4 -> TMEM lanes, half = (warp-2)/4 -> 64 of the chunk's 128 columns
//
// HBM traffic per 128-row tile: U 64 KB + residual 128 KB in, x_out 128 KB + ln_out 64|128 KB out (+ 512 KB
// bf16 pre-activation when training); W1/W2 (2 MB bf16) stream from L2 once per tile.
#include <stdlib.h>
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int FM = 128;          // rows per tile
constexpr int FD = 256;          // d_model
constexpr int FC = 128;          // hidden chunk width
constexpr int SUB = 16384;       // one [128 x 64] bf16 sub-tile, 128B-swizzled
constexpr int FNT = 320;
constexpr uint32_t Y_COL = 0, H_COL = 256;

constexpr int MAX_F = 2048;      // b1 is staged in shared memory
template <bool STORE_H>
struct Lay {
  // 3 slots x 16 KB in flight already saturate the 64 B/clk an SM can take from L2 (tools/tma_bw.cu)
  static constexpr int RING = STORE_H ? 3 : 5;
  static constexpr int OFF_U = 0;                              // 4 sub-tiles: U[128 x 256]
  static constexpr int OFF_A = OFF_U + 4 * SUB;                // 2 buffers x 2 sub-tiles (LayerNorm-tail staging at tile end)
  static constexpr int OFF_H = OFF_A + 4 * SUB;                // STORE_H: one sub-tile per column half
  static constexpr int OFF_RING = OFF_H + (STORE_H ? 2 * SUB : 0);
  static constexpr int OFF_B1 = OFF_RING + RING * SUB;         // float[MAX_F]
  static constexpr int OFF_VEC = OFF_B1 + MAX_F * 4;           // float[3][256]: b2, ln gamma, ln beta
  static constexpr int OFF_XCH = OFF_VEC + 3 * 256 * 4;        // float[2 halves][128 rows][2]
  static constexpr int OFF_BAR = OFF_XCH + 2 * 128 * 2 * 4;
  static constexpr int BYTES = OFF_BAR + 256;
};

struct FP {
  int N, n_tiles, F;
  const float* b1;
  const float* b2;
  const float* residual;
  float alpha;
  const float* ln_g;
  const float* ln_b;
  float* ln_mean;
  float* ln_rstd;
  int ln_bf16;
  long long* tl;   // debug timeline (EEC_FFN_DEBUG & 16): clock64 stamps of CTA 0's epilogue leader
  int debug;   // EEC_FFN_DEBUG bitmask (perf triage only): 1 = no weight TMA, 2 = no activation/A-store, 4 = no G2 MMAs, 8 = no G1 MMAs
};

// one row (64 bf16 = 128 B) of a 128B-swizzled [128 x 64] sub-tile
__device__ __forceinline__ void put_row64(uint8_t* sub, int r, const float (&a)[32], const float (&b)[32]) {
  uint8_t* row = sub + r * 128;
  const int sw = r & 7;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(a[g * 8 + 2 * e], a[g * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(row + ((g ^ sw) << 4)) = u;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(b[g * 8 + 2 * e], b[g * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(row + (((g + 4) ^ sw) << 4)) = u;
  }
}

template <bool STORE_H>
__global__ void __launch_bounds__(FNT, 1) ffn_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmU,
                                                               const __grid_constant__ CUtensorMap tmW1,
                                                               const __grid_constant__ CUtensorMap tmW2,
                                                               const __grid_constant__ CUtensorMap tmH,    // bf16 pre-activation store
                                                               const __grid_constant__ CUtensorMap tmX,    // fp32 x_out store
                                                               const __grid_constant__ CUtensorMap tmL,    // LayerNorm output store
                                                               const __grid_constant__ CUtensorMap tmR,    // fp32 residual load (same boxes as tmX)
                                                               const FP p) {
  using L = Lay<STORE_H>;
  constexpr int RING = L::RING;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sU = smem + L::OFF_U;
  uint8_t* sA = smem + L::OFF_A;
  uint8_t* sH = smem + L::OFF_H;
  uint8_t* sR = smem + L::OFF_RING;
  float* b1s = reinterpret_cast<float*>(smem + L::OFF_B1);
  float* vecs = reinterpret_cast<float*>(smem + L::OFF_VEC);
  float* xch = reinterpret_cast<float*>(smem + L::OFF_XCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full = bars;                 // [RING]
  uint64_t* empty = bars + 6;            // [RING]
  uint64_t* u_full = bars + 12;
  uint64_t* u_empty = bars + 13;
  uint64_t* h_full = bars + 14;          // [2]
  uint64_t* h_empty = bars + 16;         // [2]
  uint64_t* a_full = bars + 18;          // [2]
  uint64_t* a_empty = bars + 20;         // [2]
  uint64_t* y_full = bars + 22;
  uint64_t* y_empty = bars + 23;
  uint64_t* r_full = bars + 24;          // [2 halves][2 buffers]: residual boxes landed in the tail staging tiles
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = p.F / FC;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("eec: ffn_fused smem base not 1024-aligned\n"); __trap(); }
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < RING; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(u_full, 1); mbar_init(u_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], 1); mbar_init(&h_empty[i], 8);
      mbar_init(&a_full[i], 8); mbar_init(&a_empty[i], 1);
    }
    mbar_init(y_full, 1); mbar_init(y_empty, 8);
    for (int i = 0; i < 4; ++i) mbar_init(&r_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < p.F; i += FNT) b1s[i] = p.b1[i];
  for (int i = threadIdx.x; i < 256; i += FNT) { vecs[i] = p.b2[i]; vecs[256 + i] = p.ln_g[i]; vecs[512 + i] = p.ln_b[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
        const int m0 = tile * FM;
        mbar_wait(u_empty, (lt & 1) ^ 1);
        mbar_expect_tx(u_full, 4 * SUB);
#pragma unroll
        for (int s = 0; s < 4; ++s) tma_load_2d(sU + s * SUB, &tmU, u_full, s * 64, m0);
        for (int step = 0; step <= n_chunks; ++step) {
          if (step < n_chunks) {
#pragma unroll
            for (int s = 0; s < 4; ++s, ++it) {
              const int slot = it % RING;
              mbar_wait(&empty[slot], ((it / RING) & 1) ^ 1);
              if (p.debug & 1) { mbar_arrive(&full[slot]); continue; }
              mbar_expect_tx(&full[slot], SUB);
              tma_load_2d(sR + slot * SUB, &tmW1, &full[slot], s * 64, step * FC);
            }
          }
          if (step >= 1) {
            const int c = step - 1;
#pragma unroll
            for (int s = 0; s < 4; ++s, ++it) {
              const int slot = it % RING;
              mbar_wait(&empty[slot], ((it / RING) & 1) ^ 1);
              if (p.debug & 1) { mbar_arrive(&full[slot]); continue; }
              mbar_expect_tx(&full[slot], SUB);
              tma_load_2d(sR + slot * SUB, &tmW2, &full[slot], c * FC + (s >> 1) * 64, (s & 1) * 128);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(FM, 128, false, false);
      uint32_t it = 0, lt = 0;
      long long w_full = 0, w_afull = 0, w_hempty = 0, w_y = 0, t_;
      const bool prof = p.tl && blockIdx.x == 0;
#define TWAIT(acc, stmt) do { if (prof) t_ = clock64(); stmt; if (prof) acc += clock64() - t_; } while (0)
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
        mbar_wait(u_full, lt & 1);
        tc_fence_after();
        for (int step = 0; step <= n_chunks; ++step) {
          if (step < n_chunks) {
            const uint32_t gc = lt * n_chunks + step, hb = gc & 1;
            TWAIT(w_hempty, mbar_wait(&h_empty[hb], ((gc >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + H_COL + hb * FC;
#pragma unroll
            for (int s = 0; s < 4; ++s, ++it) {
              const int slot = it % RING;
              TWAIT(w_full, mbar_wait(&full[slot], (it / RING) & 1));
              tc_fence_after();
              const uint32_t sa = smem_u32(sU + s * SUB), sb = smem_u32(sR + slot * SUB);
              if (!(p.debug & 8)) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 0, 1024), make_smem_desc(sb + k * 32, 0, 1024), idesc,
                          (s > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit(&empty[slot]);
            }
            umma_commit(&h_full[hb]);
            if (step == n_chunks - 1) umma_commit(u_empty);   // every read of this tile's U has been issued
          }
          if (step >= 1) {
            const int c = step - 1;
            const uint32_t gc = lt * n_chunks + c, ab = gc & 1;
            if (c == 0) TWAIT(w_y, mbar_wait(y_empty, (lt & 1) ^ 1));      // the previous tile's Y has been drained
            TWAIT(w_afull, mbar_wait(&a_full[ab], (gc >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 4; ++s, ++it) {
              const int slot = it % RING;
              TWAIT(w_full, mbar_wait(&full[slot], (it / RING) & 1));
              tc_fence_after();
              const uint32_t sa = smem_u32(sA + (ab * 2 + (s >> 1)) * SUB), sb = smem_u32(sR + slot * SUB);
              const uint32_t d_tmem = tmem_base + Y_COL + (s & 1) * 128;
              if (!(p.debug & 4)) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 0, 1024), make_smem_desc(sb + k * 32, 0, 1024), idesc,
                          (c > 0 || s >= 2 || k > 0) ? 1u : 0u);
              }
              umma_commit(&empty[slot]);
            }
            umma_commit(&a_empty[ab]);
            if (c == n_chunks - 1) umma_commit(y_full);
          }
        }
      }
      if (prof) { p.tl[32] = w_full; p.tl[33] = w_afull; p.tl[34] = w_hempty; p.tl[35] = w_y; }
    }
  } else {
    // ================================================================== epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const bool leader = (et == half * 128);
    Stager st;
    st.buf[0] = sA + (half * 2) * SUB;
    st.buf[1] = sA + (half * 2 + 1) * SUB;
    st.next = 0;
    st.bar_id = 1 + half;
    st.leader = leader;
    st.r = r;
    st.debug = 0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float v[32], w[32];
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
      const int m0 = tile * FM;
      const int m = m0 + r;
      const bool valid = m < p.N;
      if (p.tl && blockIdx.x == 0 && et == 0) p.tl[lt * 8 + 0] = clock64();
      // the tile tail reads the fp32 residual rows: pull them into L2 now, a whole chunk loop ahead of their use
      if (leader) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tma_prefetch_l2_2d(&tmR, half * 128 + cc * 32, m0);
      }
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t gc = lt * n_chunks + c, hb = gc & 1;
        mbar_wait(&h_full[hb], (gc >> 1) & 1);
        tc_fence_after();
        tmem_ld64(lane_base + H_COL + hb * FC + half * 64, v, w);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_empty[hb]);   // H_c is in registers: the MMA warp may overwrite this buffer
        const float4* bp = reinterpret_cast<const float4*>(b1s + c * FC + half * 64);   // warp-uniform: smem broadcast
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b0 = bp[g], b1v = bp[8 + g];
          v[g * 4] += b0.x; v[g * 4 + 1] += b0.y; v[g * 4 + 2] += b0.z; v[g * 4 + 3] += b0.w;
          w[g * 4] += b1v.x; w[g * 4 + 1] += b1v.y; w[g * 4 + 2] += b1v.z; w[g * 4 + 3] += b1v.w;
        }
        if (STORE_H) {
          uint8_t* hs = sH + half * SUB;
          if (leader) bulk_wait_read<0>();     // the previous chunk's store has finished reading the staging tile
          bar_sync(1 + half, 128);
          put_row64(hs, r, v, w);
          fence_proxy_async();
          bar_sync(1 + half, 128);
          if (leader) {
            tma_store_2d(&tmH, hs, c * FC + half * 64, m0);
            bulk_commit();
          }
        }
        if (!(p.debug & 2)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] *= sigmoid_fast(v[j]); w[j] *= sigmoid_fast(w[j]); }
        }
        mbar_wait(&a_empty[hb], ((gc >> 1) & 1) ^ 1);   // G2 of chunk c-2 has finished reading this A buffer
        if (!(p.debug & 2)) put_row64(sA + (hb * 2 + half) * SUB, r, v, w);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[hb]);
      }
      if (p.tl && blockIdx.x == 0 && et == 0) p.tl[lt * 8 + 1] = clock64();
      // ------------------------------------------------ tile tail: x = residual + alpha*(Y + b2); LayerNorm(x)
      // pass 1: residual boxes [128 rows x 32 cols] fp32 arrive by TMA in the (now idle) A buffers (full-line HBM reads,
      //         conflict-free swizzled smem reads by the row-owning threads); x goes back to TMEM, row statistics in registers
      // pass 2: x_out and LayerNorm(x) leave through the same staging tiles as TMA stores
      mbar_wait(y_full, lt & 1);
      tc_fence_after();
      if (p.tl && blockIdx.x == 0 && et == 0) p.tl[lt * 8 + 2] = clock64();
      uint8_t* tb[2] = {sA + (half * 2) * SUB, sA + (half * 2 + 1) * SUB};
      uint64_t* rf = r_full + half * 2;
      if (leader) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          mbar_expect_tx(&rf[cc], SUB);
          tma_load_2d(tb[cc], &tmR, &rf[cc], half * 128 + cc * 32, m0);
        }
      }
      const uint32_t trow = lane_base + Y_COL;
      float s1 = 0.f, s2 = 0.f;
      const int sw = r & 7;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int c0 = half * 128 + cc * 32;
        const uint8_t* row = tb[cc & 1] + r * 128;
        tmem_ld32(trow + c0, v);
        const float4* b2p = reinterpret_cast<const float4*>(vecs + c0);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 bb = b2p[g];
          v[g * 4] = (v[g * 4] + bb.x) * p.alpha; v[g * 4 + 1] = (v[g * 4 + 1] + bb.y) * p.alpha;
          v[g * 4 + 2] = (v[g * 4 + 2] + bb.z) * p.alpha; v[g * 4 + 3] = (v[g * 4 + 3] + bb.w) * p.alpha;
        }
        mbar_wait(&rf[cc & 1], (cc >> 1) & 1);    // each buffer is filled twice per tile: parities 0, 1
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 f = *reinterpret_cast<const float4*>(row + ((g ^ sw) << 4));
          v[g * 4] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
        tmem_st32(trow + c0, v);
        if (cc + 2 < 4) {
          bar_sync(1 + half, 128);                // every row of this box has been read: refill the buffer
          if (leader) {
            mbar_expect_tx(&rf[cc & 1], SUB);
            tma_load_2d(tb[cc & 1], &tmR, &rf[cc & 1], c0 + 64, m0);
          }
        }
      }
      if (p.tl && blockIdx.x == 0 && et == 0) p.tl[lt * 8 + 3] = clock64();
      xch[half * 256 + r * 2] = s1;
      xch[half * 256 + r * 2 + 1] = s2;
      bar_sync(3, 256);
      s1 += xch[(half ^ 1) * 256 + r * 2];
      s2 += xch[(half ^ 1) * 256 + r * 2 + 1];
      const float mu = s1 * (1.f / 256.f);
      const float rs = rsqrtf(fmaxf(s2 * (1.f / 256.f) - mu * mu, 0.f) + 1e-5f);
      if (half == 0 && valid && p.ln_mean) { p.ln_mean[m] = mu; p.ln_rstd[m] = rs; }
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int c0 = half * 128 + cc * 32;
        tmem_ld32(trow + c0, v);
        if (cc == 3) {   // Y fully read: hand the accumulator back before the last staging round trips
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(y_empty);
        }
        st.store(&tmX, c0, m0, v, false);
        const float4* gp = reinterpret_cast<const float4*>(vecs + 256 + c0);
        const float4* bp2 = reinterpret_cast<const float4*>(vecs + 512 + c0);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 gg = gp[g], bb = bp2[g];
          v[g * 4] = (v[g * 4] - mu) * rs * gg.x + bb.x; v[g * 4 + 1] = (v[g * 4 + 1] - mu) * rs * gg.y + bb.y;
          v[g * 4 + 2] = (v[g * 4 + 2] - mu) * rs * gg.z + bb.z; v[g * 4 + 3] = (v[g * 4 + 3] - mu) * rs * gg.w + bb.w;
        }
        st.store(&tmL, c0, m0, v, p.ln_bf16);
      }
      if (p.tl && blockIdx.x == 0 && et == 0) p.tl[lt * 8 + 4] = clock64();
      // the staging tiles alias the A buffers: every bulk store must have read them before the next tile's chunk 0
      if (leader) bulk_wait_read<0>();
      bar_sync(3, 256);
      if (p.tl && blockIdx.x == 0 && et == 0) p.tl[lt * 8 + 5] = clock64();
    }
    if (leader) bulk_wait_all();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_sms = 0;

template <bool STORE_H>
int launch_ffn(const CUtensorMap& tu, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& th, const CUtensorMap& tx,
               const CUtensorMap& tl, const CUtensorMap& tr, const FP& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(ffn_fused_fwd_kernel<STORE_H>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<STORE_H>::BYTES));
    attr_set = true;
  }
  ffn_fused_fwd_kernel<STORE_H><<<grid, FNT, Lay<STORE_H>::BYTES, st>>>(tu, tw1, tw2, th, tx, tl, tr, p);
  EEC_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace eec

using namespace eec;

extern "C" int eec_ffn_fwd(const void* u, const void* w1, const float* b1, const void* w2, const float* b2,
                           const float* residual, float alpha, const float* ln_gamma, const float* ln_beta, float* x_out,
                           void* ln_out, int ln_dtype, float* ln_mean, float* ln_rstd, void* hpre, int rows, int d, int f,
                           eec_stream_t stream) {
  EEC_CHECK_ARG(d == FD, "ffn_fwd: d_model must be 256 (got %d)", d);
  EEC_CHECK_ARG(f > 0 && f % FC == 0 && f <= MAX_F, "ffn_fwd: d_feed_forward must be a multiple of 128 in [128, %d] (got %d)", MAX_F, f);
  EEC_CHECK_ARG(u && w1 && b1 && w2 && b2 && residual && ln_gamma && ln_beta && x_out && ln_out, "ffn_fwd: NULL argument");
  if (rows == 0) return 0;
  if (!g_sms) {
    int dev = 0;
    EEC_CUDA(cudaGetDevice(&dev));
    EEC_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap tu, tw1, tw2, th, tx, tl, tr;
  if (int r = get_tmap_2d(&tu, u, FD, (uint64_t)rows, FD * 2, 64, 128)) return r;
  if (int r = get_tmap_2d(&tw1, w1, FD, (uint64_t)f, FD * 2, 64, 128)) return r;
  if (int r = get_tmap_2d(&tw2, w2, (uint64_t)f, FD, (uint64_t)f * 2, 64, 128)) return r;
  th = tu;
  if (hpre) { if (int r = get_tmap_2d(&th, hpre, (uint64_t)f, (uint64_t)rows, (uint64_t)f * 2, 64, 128)) return r; }
  if (int r = get_tmap_store(&tx, x_out, false, FD, (uint64_t)rows, FD)) return r;
  if (int r = get_tmap_store(&tl, ln_out, ln_dtype == EEC_BF16, FD, (uint64_t)rows, FD)) return r;
  if (int r = get_tmap_store(&tr, residual, false, FD, (uint64_t)rows, FD)) return r;
  FP p{};
  p.N = rows; p.n_tiles = cdiv(rows, FM); p.F = f;
  p.b1 = b1; p.b2 = b2; p.residual = residual; p.alpha = alpha;
  p.ln_g = ln_gamma; p.ln_b = ln_beta; p.ln_mean = ln_mean; p.ln_rstd = ln_rstd; p.ln_bf16 = ln_dtype == EEC_BF16;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("EEC_FFN_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.debug = dbg;
  static long long* tl_buf = nullptr;
  if ((dbg & 16) && !tl_buf) { EEC_CUDA(cudaMalloc(&tl_buf, 64 * 8)); EEC_CUDA(cudaMemset(tl_buf, 0, 64 * 8)); }
  p.tl = (dbg & 16) ? tl_buf : nullptr;
  const int grid = min(p.n_tiles, g_sms);
  const int rc = hpre ? launch_ffn<true>(tu, tw1, tw2, th, tx, tl, tr, p, grid, S(stream))
                      : launch_ffn<false>(tu, tw1, tw2, th, tx, tl, tr, p, grid, S(stream));
  if (rc == 0 && p.tl) {
    long long h[40];
    EEC_CUDA(cudaStreamSynchronize(S(stream)));
    EEC_CUDA(cudaMemcpy(h, p.tl, sizeof(h), cudaMemcpyDeviceToHost));
    for (int t = 0; t < 2; ++t)
      fprintf(stderr, "ffn timeline tile %d: chunks %lld  wait_y %lld  tail1 %lld  tail2 %lld  drain %lld  (clk)\n", t, h[t * 8 + 1] - h[t * 8],
              h[t * 8 + 2] - h[t * 8 + 1], h[t * 8 + 3] - h[t * 8 + 2], h[t * 8 + 4] - h[t * 8 + 3], h[t * 8 + 5] - h[t * 8 + 4]);
    fprintf(stderr, "ffn MMA-thread waits (both tiles): full %lld  a_full %lld  h_empty %lld  y_empty %lld (clk)\n", h[32], h[33], h[34], h[35]);
  }
  return rc;
}
