// attn_tc.cu -- self-attention forward on the tcgen05 tensor cores (bf16 in, fp32 softmax).
//
// One CTA per (128-query tile, head, utterance).  For every 128-key block:
//   S = Q K^T            tcgen05.mma  M=128 N=128 K=32   (Q,K: K-major, 64B-swizzled TMA tiles)  -> TMEM cols [0,128)
//   softmax warps (one thread per query row) read S with tcgen05.ld, apply the key-length mask,
//   keep a running max / sum (exp2 domain), write P (bf16) into 128B-swizzled shared memory
//   O_blk = P V          tcgen05.mma  M=128 N=32  K=128  (P: K-major smem; V: MN-major TMA tile)  -> TMEM cols [128,160)
//   the softmax threads fold O_blk into their fp32 register accumulator with the online-softmax rescale.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2..5 = softmax/epilogue.
// Fully masked utterances (key_len == 0) produce zeros, like torch's CPU SDPA.
// Padded QUERY rows are computed like any other row; only keys are masked.
#include <stdlib.h>
#include "tc_common.cuh"

namespace eec {
namespace {
using namespace tc;

constexpr int QT = 128;            // queries per CTA
constexpr int KB = 128;            // keys per block
constexpr int DHEAD = 32;
constexpr int KV_STAGES = 2;
constexpr int Q_BYTES = QT * DHEAD * 2;       // 8 KB
constexpr int K_BYTES = KB * DHEAD * 2;       // 8 KB
constexpr int V_BYTES = KB * DHEAD * 2;       // 8 KB
constexpr int P_BYTES = QT * KB * 2;          // 32 KB (two 64-key swizzle atoms of 16 KB)
constexpr int AT_SMEM = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + P_BYTES + 1024 + 256;
constexpr int AT_THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t S_COL = 0, O_COL = 128;
constexpr uint32_t SW64 = 4, SW128 = 2;

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Geometry of a call (host-built).  Self-attention over the packed [rows, 3D] projection, the decoder's causal self-attention and its
// cross-attention over the encoder states (Tq != Tk, Q and K/V in different tensors) run the same kernel: two tensor maps (Q tensor,
// K/V tensor) and the column at which each block starts.
struct TcGeom {
  int Tq, Tk, H, B;
  int q_col, k_col, v_col;          // first column of head 0 of Q (in the Q tensor) and of K / V (in the K/V tensor)
  const int32_t* key_len;           // [B] or NULL
  const uint32_t* key_bits;         // [B, ceil(Tk/32)] or NULL (GENERAL only)
  int causal;                       // (GENERAL only)
  int ldo;                          // row pitch of ctx in elements
};
__device__ __forceinline__ int tc_klen(const TcGeom& g, int b) { return g.key_len ? min(g.key_len[b], g.Tk) : g.Tk; }
// key blocks an item has to visit: those holding a key < klen and, under the causal mask, a key <= the tile's last query
__device__ __forceinline__ int tc_nblk(const TcGeom& g, int klen, int q0) {
  int n = (klen + KB - 1) / KB;
  if (g.causal) n = min(n, (min(q0 + QT, g.Tq) - 1) / KB + 1);
  return n;
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent kernel: a CTA that lives for one (128 queries, head, utterance) item of ~3 key blocks spends ~45 % of its warp-stall
// samples on TMEM allocation / barrier set-up / tear-down (ncu source view, round 1).  Here a grid of 2 CTAs per SM walks the work
// items: one TMEM allocation and one barrier initialisation per CTA, barrier phases running on across items, and a q_empty barrier so
// that the producer refills the Q tile only after the item's last S MMA has completed.
// GENERAL adds the decoder's masks (causal, per-key validity bits) to the softmax; the encoder instantiation does not carry them.
template <bool DROP, bool GENERAL>
__global__ void __launch_bounds__(AT_THREADS, 2) attn_fwd_tcp_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                 const __grid_constant__ CUtensorMap tm_kv, const TcGeom g,
                                                                 __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse,
                                                                 const DropArgs drop, const ActiveItems act_items) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;                       // [stage]
  uint8_t* sV = sK + KV_STAGES * K_BYTES;           // [stage]
  uint8_t* sP = sV + KV_STAGES * V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                     // [2]
  uint64_t* kv_empty = bars + 3;                    // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* q_empty = bars + 9;                     // every S MMA of the current work item has been issued and has completed: sQ may be refilled
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = g.H, Tq = g.Tq, Tk = g.Tk;
  const int nqt = (Tq + QT - 1) / QT;
  const int Beff = act_items.n_dev ? min(g.B, active_count(act_items)) : g.B;    // utterances past the active-item limit are not work items
  const int n_items = nqt * H * Beff;
  // work item w = (b * H + h) * nqt + qt, walked with a stride of gridDim.x.  Barrier phases run on across items: `it` counts the
  // items of this CTA that loaded a Q tile (nblk > 0), `jbase` the key blocks it has processed so far.

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, jbase = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int qt = w % nqt, bh = w / nqt, h = bh % H, b = bh / H;
        const int q0 = qt * QT;
        const int nblk = tc_nblk(g, tc_klen(g, b), q0);
        if (nblk == 0) continue;
        if (it > 0) mbar_wait(q_empty, (it - 1) & 1);        // the previous item's S MMAs are done with sQ
        mbar_expect_tx(q_full, Q_BYTES);
        tma_load_2d(sQ, &tm_q, q_full, g.q_col + h * DHEAD, b * Tq + q0);
        for (int j = 0; j < nblk; ++j) {
          const uint32_t jb = jbase + j;
          const int s = jb % KV_STAGES;
          mbar_wait(&kv_empty[s], ((jb / KV_STAGES) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], K_BYTES + V_BYTES);
          tma_load_2d(sK + s * K_BYTES, &tm_kv, &kv_full[s], g.k_col + h * DHEAD, b * Tk + j * KB);
          tma_load_2d(sV + s * V_BYTES, &tm_kv, &kv_full[s], g.v_col + h * DHEAD, b * Tk + j * KB);
        }
        ++it;
        jbase += nblk;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(QT, KB, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(QT, DHEAD, false, true);
      uint32_t it = 0, jbase = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int b = (w / nqt) / H;
      const int nblk = tc_nblk(g, tc_klen(g, b), (w % nqt) * QT);
      if (nblk == 0) continue;
      mbar_wait(q_full, it & 1);
      for (int j = 0; j < nblk; ++j) {
        const uint32_t jb = jbase + j;
        const int s = jb % KV_STAGES;
        mbar_wait(&kv_full[s], (jb / KV_STAGES) & 1);
        if (jb > 0) mbar_wait(s_free, (jb - 1) & 1);   // softmax threads have drained S of the previous block (of this or the previous item)
        tc_fence_after();
        const uint32_t aq = smem_u32(sQ), bk = smem_u32(sK + s * K_BYTES), bv = smem_u32(sV + s * V_BYTES);
#pragma unroll
        for (int k = 0; k < DHEAD / 16; ++k)
          umma_bf16(tmem_base + S_COL, make_smem_desc(aq + k * 32, 0, 512, SW64), make_smem_desc(bk + k * 32, 0, 512, SW64),
                    idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        if (j == nblk - 1) umma_commit(q_empty);     // last S MMA of the item: once it completes sQ is free
        mbar_wait(p_full, jb & 1);                   // P_j is in smem (and O_{j-1} has been consumed)
        tc_fence_after();
        const uint32_t ap = smem_u32(sP);
#pragma unroll
        for (int k = 0; k < KB / 16; ++k)
          umma_bf16(tmem_base + O_COL, make_smem_desc(ap + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024, SW128),
                    make_smem_desc(bv + k * 1024, 0, 512, SW64), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
      }
      ++it;
      jbase += nblk;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue: one thread per query row
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sc = rsqrtf((float)DHEAD) * 1.4426950408889634f;  // 1/sqrt(dh) * log2(e)
    uint32_t jbase = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
    const int qt = w % nqt, bh = w / nqt, h = bh % H, b = bh / H;
    const int q0 = qt * QT;
    const int klen = tc_klen(g, b);
    const int nblk = tc_nblk(g, klen, q0);
    float m_run = -INFINITY, l_run = 0.f;
    float o[DHEAD];
#pragma unroll
    for (int i = 0; i < DHEAD; ++i) o[i] = 0.f;
    float v[32];
    // dropout on the probabilities (DROP): the row sum keeps the undropped values, the P tile fed to the P V product is masked
    // keep-mask words (eec_dropout_bits, W = 32): word (k/32, row) at bits[(k/32)*R + row], row = (b*H + h)*T + t, R = B*H*T
    const uint32_t* dbits = reinterpret_cast<const uint32_t*>(drop.bits);
    const long drow = (long)(b * H + h) * Tq + (q0 + r), dR = (long)g.B * H * Tq;
    const bool dvalid = (q0 + r) < Tq;
    const uint32_t* kbits = GENERAL && g.key_bits ? g.key_bits + (long)b * ((Tk + 31) >> 5) : nullptr;
    for (int j = 0; j < nblk; ++j) {
      const int nvalid = min(KB, klen - j * KB);   // only the last key block can be partial
      uint32_t dword[4] = {0u, 0u, 0u, 0u};          // this row's keep-mask words of the block: in flight while S is being computed
      if (DROP && dvalid) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c * 32 < nvalid) dword[c] = dbits[(long)((j * KB) / 32 + c) * dR + drow];
      }
      const uint32_t jb = jbase + j;
      // GENERAL: visibility word of this row for each 32-key span of the block (length, causal and per-key validity masks)
      uint32_t vis[4] = {~0u, ~0u, ~0u, ~0u};
      if (GENERAL) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int base = j * KB + c * 32;
          uint32_t mk = (base + 32 <= klen) ? ~0u : (base >= klen ? 0u : ((1u << (klen - base)) - 1u));
          if (g.causal) {
            const int lim = q0 + r - base;        // key base + i is visible iff i <= lim
            mk &= lim >= 31 ? ~0u : (lim < 0 ? 0u : ((2u << lim) - 1u));
          }
          if (kbits && base < Tk) mk &= kbits[base >> 5];
          vis[c] = mk;
        }
      }
      mbar_wait(s_full, jb & 1);
      tc_fence_after();
      // pass 1: row max (raw scores; the positive scale is applied once) over the valid keys of this block
      float mraw = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < KB; c0 += 32) {
        if (c0 >= nvalid) break;               // uniform
        tmem_ld32(trow + S_COL + c0, v);
        if (GENERAL) {
          const uint32_t mk = c0 == 0 ? vis[0] : c0 == 32 ? vis[1] : c0 == 64 ? vis[2] : vis[3];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if ((mk >> i) & 1u) mraw = fmaxf(mraw, v[i]);
        } else if (c0 + 32 <= nvalid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < nvalid) mraw = fmaxf(mraw, v[i]);
        }
      }
      const float mx = fmaxf(m_run, mraw * sc);
      // pass 2: p = exp2(s*sc - mx), row sum, bf16 P tile in swizzled smem
      float psum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < KB; c0 += 32) {
        if (c0 < nvalid) {
          tmem_ld32(trow + S_COL + c0, v);
          if (GENERAL) {
            const uint32_t mk = c0 == 0 ? vis[0] : c0 == 32 ? vis[1] : c0 == 64 ? vis[2] : vis[3];
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = ((mk >> i) & 1u) ? ex2_fast(fmaf(v[i], sc, -mx)) : 0.f; psum += v[i]; }
          } else if (c0 + 32 <= nvalid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = ex2_fast(fmaf(v[i], sc, -mx)); psum += v[i]; }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = (c0 + i < nvalid) ? ex2_fast(fmaf(v[i], sc, -mx)) : 0.f; psum += v[i]; }
          }
          if (DROP) drop_apply_bits<32>(v, c0 == 0 ? dword[0] : c0 == 32 ? dword[1] : c0 == 64 ? dword[2] : dword[3], drop.scale);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        uint8_t* prow = sP + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {  // 4 chunks of 8 keys (16 B) in this 32-key span
          const int chunk = ((c0 & 63) >> 3) + gg;
          uint4 u;
          __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) hh[e] = __floats2bfloat162_rn(v[gg * 8 + 2 * e], v[gg * 8 + 2 * e + 1]);
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = u;
        }
      }
      tc_fence_before();
      mbar_arrive(s_free);
      fence_proxy_async();
      mbar_arrive(p_full);
      const float corr = (m_run == -INFINITY) ? 0.f : ex2_fast(m_run - mx);
      l_run = l_run * corr + psum;
      m_run = mx;
      mbar_wait(o_full, jb & 1);
      tc_fence_after();
      tmem_ld32(trow + O_COL, v);
#pragma unroll
      for (int i = 0; i < DHEAD; ++i) o[i] = fmaf(o[i], corr, v[i]);
    }
    const int t = q0 + r;
    if (t < Tq) {
      const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;
      __nv_bfloat16* dst = ctx + ((long)b * Tq + t) * g.ldo + h * DHEAD;
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        float tt[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) tt[e] = o[gg * 8 + e] * inv;
        st8<__nv_bfloat16>(dst + gg * 8, tt);
      }
      if (lse) lse[((long)b * H + h) * Tq + t] = (l_run > 0.f) ? (m_run + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
    }
    jbase += nblk;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// v2 forward (default).  At head dim 32 the tensor pipe needs only ~256 clk per 128 x 128 key block, while draining the fp32 S tile from
// TMEM costs 64 KB / (64 B/clk) = 1024 clk per pass and one exp per element costs 1024 clk of MUFU per warp: the kernel above reads S
// twice (max pass, exp pass), reads O after every block, evaluates ex2 in fp32 and walks serial max / sum chains, so a block's
// MMA -> softmax -> MMA chain is ~6.5 k clk of mostly latency.  Here:
//   * ONE TMEM pass: a thread loads its whole 128-key row of S into registers (two 64-column loads, one wait each) and releases the S
//     tile at once, so the QK^T of block j + 1 runs under the softmax of block j;
//   * O and the row sums never leave TMEM between blocks: P V accumulates into O across the key blocks, and a second tiny MMA against a
//     tile of ones accumulates the row sums L = sum_k P (of the bf16 values the tensor core actually multiplies) next to it;
//   * lazy rescaling: the running maximum is only moved -- and O, L rescaled in TMEM -- when a block's maximum exceeds it by more than
//     2^4; otherwise the stale maximum keeps being used (P <= 16, exact after the final division by L);
//   * row maximum over four independent chains (three-input FMNMX).  (ex2.approx.ftz.bf16x2 was tried for the exponentials: sm_100a
//     issues TWO MUFU.EX2.BF16 per packed pair, so it saves nothing and costs 0.3 % of accuracy; the exponentials stay fp32.)
// In-kernel clock64 timeline of a key block with two co-resident CTAs (make attn_timeline, tools/attn_timeline.py, B200): S load 190 clk,
// max 270 (710 with a partial block's masking), wait for the previous P V ~200-600, exponentials + P tile 1.9 k clk -- the MUFU pipe
// (128 ex2 per thread, 16 per clk and SM, shared by the softmax warps of both CTAs) is what a block's period of ~3.5 k clk is made of.
// DROP keeps the row sum in registers (it is taken BEFORE dropout, so the ones-MMA over the dropped P cannot provide it).
constexpr int ONES_BYTES = KB * DHEAD * 2;      // 8 KB of bf16 ones, addressed with the V tile's descriptor shape (N = 16)
constexpr int AT2_SMEM = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + P_BYTES + ONES_BYTES + 1024 + 256;
constexpr uint32_t L_COL = 160;
constexpr float LAZY_LOG2 = 4.0f;
constexpr int AT2_MAX_ITEMS = 128;      // work items per CTA (the host enlarges the grid beyond 2 CTAs per SM if needed)

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld16_async(taddr, r);
  tmem_ld_wait16(r);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

#ifdef EEC_ATTN_TIMELINE
// perf triage (make attn_timeline): CTA 0 stamps clock64 at every hand-over of a key block; row = event id, col = block counter
__device__ long long g_attn_tl[8][64];
#define ATL(ev, idx) do { if (blockIdx.x == 0 && (idx) < 64) g_attn_tl[ev][idx] = clock64(); } while (0)
#else
#define ATL(ev, idx) do { } while (0)
#endif

template <bool DROP, bool GENERAL>
__global__ void __launch_bounds__(AT_THREADS, 2) attn_fwd_v2_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                const __grid_constant__ CUtensorMap tm_kv, const TcGeom g,
                                                                __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse,
                                                                const DropArgs drop, const ActiveItems act_items) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;                       // [stage]
  uint8_t* sV = sK + KV_STAGES * K_BYTES;           // [stage]
  uint8_t* sP = sV + KV_STAGES * V_BYTES;
  uint8_t* sOnes = sP + P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + ONES_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                     // [2]
  uint64_t* kv_empty = bars + 3;                    // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;                      // the softmax threads hold S of the block in registers
  uint64_t* p_full = bars + 7;                      // P of the block is in smem (and O, L carry any rescale)
  uint64_t* o_full = bars + 8;                      // P V (and the row-sum MMA) of the block have completed
  uint64_t* q_empty = bars + 9;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);
  // this CTA's work items, decoded once (utterance, head, first query, visible key length, key blocks): the three warp roles walk the
  // same list, and none of them stalls on the key_len load / the index divisions at an item boundary
  __shared__ int4 items_s[AT2_MAX_ITEMS];
  __shared__ int items_klen[AT2_MAX_ITEMS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = g.H, Tq = g.Tq, Tk = g.Tk;
  const int nqt = (Tq + QT - 1) / QT;
  const int Beff = act_items.n_dev ? min(g.B, active_count(act_items)) : g.B;
  const int n_items = nqt * H * Beff;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);     // one arrival per softmax warp (after __syncwarp): 128 per-thread arrivals on one mbarrier serialise
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < ONES_BYTES / 4; i += AT_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;   // bf16 1.0 pairs
  const int n_my = (n_items > (int)blockIdx.x) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;    // <= AT2_MAX_ITEMS (host)
  for (int i = threadIdx.x; i < n_my; i += AT_THREADS) {
    const int w = blockIdx.x + i * gridDim.x;
    const int qt = w % nqt, bh = w / nqt, h = bh % H, b = bh / H;
    const int klen = tc_klen(g, b);
    items_s[i] = make_int4(b, h, qt * QT, tc_nblk(g, klen, qt * QT));
    items_klen[i] = klen;
  }
  fence_proxy_async();
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, jbase = 0;
      for (int wi = 0; wi < n_my; ++wi) {
        const int4 itm = items_s[wi];
        const int b = itm.x, h = itm.y, q0 = itm.z, nblk = itm.w;
        if (nblk == 0) continue;
        if (it > 0) mbar_wait(q_empty, (it - 1) & 1);
        mbar_expect_tx(q_full, Q_BYTES);
        tma_load_2d(sQ, &tm_q, q_full, g.q_col + h * DHEAD, b * Tq + q0);
        for (int j = 0; j < nblk; ++j) {
          const uint32_t jb = jbase + j;
          const int s = jb % KV_STAGES;
          mbar_wait(&kv_empty[s], ((jb / KV_STAGES) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], K_BYTES + V_BYTES);
          tma_load_2d(sK + s * K_BYTES, &tm_kv, &kv_full[s], g.k_col + h * DHEAD, b * Tk + j * KB);
          tma_load_2d(sV + s * V_BYTES, &tm_kv, &kv_full[s], g.v_col + h * DHEAD, b * Tk + j * KB);
        }
        ++it;
        jbase += nblk;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(QT, KB, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(QT, DHEAD, false, true);
      constexpr uint32_t idesc_l = make_idesc_bf16(QT, 16, false, true);
      const uint32_t aq = smem_u32(sQ), ap = smem_u32(sP), a1 = smem_u32(sOnes);
      // UMMA descriptors built once (the issuing thread spent ~80 clk per MMA re-deriving them: 1.3 k clk for the 16 MMAs of a block, on
      // the path between "P written" and "P V done"); per MMA only an offset in 16-byte units is added
      const uint64_t d_q = make_smem_desc(aq, 0, 512, SW64), d_p = make_smem_desc(ap, 0, 1024, SW128), d_1 = make_smem_desc(a1, 0, 512, SW64);
      uint64_t d_k[KV_STAGES], d_v[KV_STAGES];
#pragma unroll
      for (int s = 0; s < KV_STAGES; ++s) {
        d_k[s] = make_smem_desc(smem_u32(sK + s * K_BYTES), 0, 512, SW64);
        d_v[s] = make_smem_desc(smem_u32(sV + s * V_BYTES), 0, 512, SW64);
      }
      auto pick = [](const uint64_t(&d)[KV_STAGES], int s) {   // (explicit selects keep the descriptor arrays in registers)
        uint64_t r = d[0];
#pragma unroll
        for (int i = 1; i < KV_STAGES; ++i)
          if (s == i) r = d[i];
        return r;
      };
      uint32_t it = 0, jbase = 0;
      // S MMA of block jb (its K tile is stage jb % 2); the S tile is free once the softmax threads hold block jb - 1 in registers
      auto issue_s = [&](uint32_t jb, bool last_of_item) {
        const int s = jb % KV_STAGES;
        mbar_wait(&kv_full[s], (jb / KV_STAGES) & 1);
        if (jb > 0) mbar_wait(s_free, (jb - 1) & 1);
        tc_fence_after();
        const uint64_t dk = pick(d_k, s);
#pragma unroll
        for (int k = 0; k < DHEAD / 16; ++k) umma_bf16(tmem_base + S_COL, d_q + k * 2, dk + k * 2, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        if (last_of_item) umma_commit(q_empty);      // once it completes sQ may be refilled
      };
      for (int wi = 0; wi < n_my; ++wi) {
        const int nblk = items_s[wi].w;
        if (nblk == 0) continue;
        mbar_wait(q_full, it & 1);
        issue_s(jbase, nblk == 1);
        for (int j = 0; j < nblk; ++j) {
          const uint32_t jb = jbase + j;
          if (j + 1 < nblk) issue_s(jb + 1, j + 2 == nblk);     // runs under the softmax of block j
          const int s = jb % KV_STAGES;
          mbar_wait(p_full, jb & 1);
          ATL(0, jb);                                           // MMA thread: P of block jb seen
          tc_fence_after();
          const uint64_t dv = pick(d_v, s);
          const uint32_t accf = j > 0 ? 1u : 0u;                // the first block of an item overwrites O and L
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            const uint64_t pd = d_p + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2);
            umma_bf16(tmem_base + O_COL, pd, dv + k * 64, idesc_o, (k > 0) ? 1u : accf);
            umma_bf16(tmem_base + L_COL, pd, d_1 + k * 64, idesc_l, (k > 0) ? 1u : accf);
          }
          umma_commit(o_full);
          umma_commit(&kv_empty[s]);
          ATL(1, jb);                                           // MMA thread: P V of block jb issued + committed
        }
        ++it;
        jbase += nblk;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue: one thread per query row
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sc = rsqrtf((float)DHEAD) * 1.4426950408889634f;  // 1/sqrt(dh) * log2(e)
    uint32_t jbase = 0;
    const uint32_t* dbits = reinterpret_cast<const uint32_t*>(drop.bits);
    for (int wi = 0; wi < n_my; ++wi) {
      const int4 itm = items_s[wi];
      const int b = itm.x, h = itm.y, q0 = itm.z, nblk = itm.w;
      const int klen = items_klen[wi];
      float m_run = -INFINITY;      // the (possibly stale) maximum every P of this row has been taken against
      float l_run = 0.f;            // DROP only: row sum of the undropped probabilities
      const long drow = (long)(b * H + h) * Tq + (q0 + r), dR = (long)g.B * H * Tq;
      const bool dvalid = (q0 + r) < Tq;
      const uint32_t* kbits = GENERAL && g.key_bits ? g.key_bits + (long)b * ((Tk + 31) >> 5) : nullptr;
      for (int j = 0; j < nblk; ++j) {
        const uint32_t jb = jbase + j;
        const int nvalid = min(KB, klen - j * KB);
        uint32_t dword[4] = {0u, 0u, 0u, 0u};
        if (DROP && dvalid) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c * 32 < nvalid) dword[c] = dbits[(long)((j * KB) / 32 + c) * dR + drow];
        }
        uint32_t vis[4] = {~0u, ~0u, ~0u, ~0u};
        bool all_vis = nvalid == KB;
        if (GENERAL || !all_vis) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int base = j * KB + c * 32;
            uint32_t mk = (base + 32 <= klen) ? ~0u : (base >= klen ? 0u : ((1u << (klen - base)) - 1u));
            if (GENERAL) {
              if (g.causal) {
                const int lim = q0 + r - base;
                mk &= lim >= 31 ? ~0u : (lim < 0 ? 0u : ((2u << lim) - 1u));
              }
              if (kbits && base < Tk) mk &= kbits[base >> 5];
            }
            vis[c] = mk;
          }
          all_vis = (vis[0] & vis[1] & vis[2] & vis[3]) == ~0u;
        }
        if (threadIdx.x == 64) ATL(2, jb);           // softmax: starts waiting for S of block jb
        mbar_wait(s_full, jb & 1);
        if (threadIdx.x == 64) ATL(3, jb);           // softmax: S ready
        tc_fence_after();
        float x[128];
        tmem_ld32x2(trow + S_COL, trow + S_COL + 32, *reinterpret_cast<float(*)[32]>(&x[0]), *reinterpret_cast<float(*)[32]>(&x[32]));
        tmem_ld32x2(trow + S_COL + 64, trow + S_COL + 96, *reinterpret_cast<float(*)[32]>(&x[64]), *reinterpret_cast<float(*)[32]>(&x[96]));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);          // S lives in registers: the next block's Q K^T may overwrite the tile
        if (threadIdx.x == 64) ATL(4, jb);           // softmax: S in registers
        if (!all_vis) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (!((vis[i >> 5] >> (i & 31)) & 1u)) x[i] = -INFINITY;
        }
        float m0 = x[0], m1 = x[1], m2 = x[2], m3 = x[3];
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          m0 = fmaxf(m0, x[i]); m1 = fmaxf(m1, x[i + 1]); m2 = fmaxf(m2, x[i + 2]); m3 = fmaxf(m3, x[i + 3]);
        }
        const float mblk = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sc;
        // lazy rescale: move the maximum only when the block exceeds it by more than 2^LAZY_LOG2
        const bool move = (mblk > m_run + LAZY_LOG2) || (m_run == -INFINITY && mblk > -INFINITY);
        const float m_new = move ? mblk : m_run;
        const float corr = (move && m_run > -INFINITY) ? ex2_fast(m_run - m_new) : 1.0f;
        if (j > 0) {
          if (threadIdx.x == 64) ATL(5, jb);         // softmax: max done, waits for P V of the previous block
          mbar_wait(o_full, (jb - 1) & 1);           // P V of the previous block is done: sP may be rewritten, O / L may be rescaled
          if (threadIdx.x == 64) ATL(6, jb);
          tc_fence_after();
          if (__any_sync(0xffffffffu, corr != 1.0f)) {
            float t[16];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              tmem_ld16(trow + O_COL + hh * 16, t);
#pragma unroll
              for (int i = 0; i < 16; ++i) t[i] *= corr;
              tmem_st16(trow + O_COL + hh * 16, t);
            }
            if (!DROP) {
              tmem_ld16(trow + L_COL, t);
#pragma unroll
              for (int i = 0; i < 16; ++i) t[i] *= corr;
              tmem_st16(trow + L_COL, t);
            }
            tc_fence_before();
          }
        }
        if (DROP) l_run *= corr;
        m_run = m_new;
        const float msub = (m_run == -INFINITY) ? 0.f : m_run;   // (a row without a visible key so far: every x is -inf -> P = 0)
        float psum = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < KB; c0 += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = pack_bf16x2(ex2_fast(fmaf(x[c0 + 2 * i], sc, -msub)), ex2_fast(fmaf(x[c0 + 2 * i + 1], sc, -msub)));
          if (DROP) {
            const uint32_t dwv = c0 == 0 ? dword[0] : c0 == 32 ? dword[1] : c0 == 64 ? dword[2] : dword[3];
            const __nv_bfloat162 sc2 = __float2bfloat162_rn(drop.scale);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&pk[i]));
              psum += f.x + f.y;
              const uint32_t keep = ((dwv >> (2 * i)) & 1u ? 0x0000ffffu : 0u) | ((dwv >> (2 * i + 1)) & 1u ? 0xffff0000u : 0u);
              __nv_bfloat162 pv = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&pk[i]), sc2);
              pk[i] = *reinterpret_cast<uint32_t*>(&pv) & keep;
            }
          }
          uint8_t* prow = sP + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const int chunk = ((c0 & 63) >> 3) + gg;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = make_uint4(pk[gg * 4], pk[gg * 4 + 1], pk[gg * 4 + 2], pk[gg * 4 + 3]);
          }
        }
        if (DROP) l_run += psum;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (threadIdx.x == 64) ATL(7, jb);           // softmax: P written
      }
      // ---- epilogue of the item: O / L out of TMEM once
      if (nblk > 0) {
        mbar_wait(o_full, (jbase + nblk - 1) & 1);
        tc_fence_after();
        float o[32], lt[16];
        tmem_ld32(trow + O_COL, o);
        float lsum = l_run;
        if (!DROP) {
          tmem_ld16(trow + L_COL, lt);
          lsum = lt[0];
        }
        tc_fence_before();
        const int t = q0 + r;
        if (t < Tq) {
          const float inv = (lsum > 0.f) ? 1.0f / lsum : 0.f;
          __nv_bfloat16* dst = ctx + ((long)b * Tq + t) * g.ldo + h * DHEAD;
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            float tt[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) tt[e] = o[gg * 8 + e] * inv;
            st8<__nv_bfloat16>(dst + gg * 8, tt);
          }
          if (lse) lse[((long)b * H + h) * Tq + t] = (lsum > 0.f) ? (m_run + log2f(lsum)) * 0.6931471805599453f : -INFINITY;
        }
      } else {
        const int t = q0 + r;
        if (t < Tq) {      // no visible key block at all (key_len == 0): zeros, like torch's CPU SDPA
          __nv_bfloat16* dst = ctx + ((long)b * Tq + t) * g.ldo + h * DHEAD;
          float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) st8<__nv_bfloat16>(dst + gg * 8, z);
          if (lse) lse[((long)b * H + h) * Tq + t] = -INFINITY;
        }
      }
      jbase += nblk;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static int fwd_launch(const CUtensorMap& tq, const CUtensorMap& tkv, const TcGeom& g, void* ctx, float* lse, const DropArgs& drop,
                      bool general, cudaStream_t st) {
  static bool attr_set = false;
  static int sms = 0;
  if (!attr_set) {
    EEC_CUDA(cudaFuncSetAttribute(attn_fwd_tcp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    EEC_CUDA(cudaFuncSetAttribute(attn_fwd_tcp_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    EEC_CUDA(cudaFuncSetAttribute(attn_fwd_tcp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    attr_set = true;
  }
  const int items = cdiv(g.Tq, QT) * g.H * g.B;
  const dim3 pgrid(min(items, 2 * sms));
  static int v2 = -1;
  if (v2 < 0) { const char* e = getenv("EEC_ATTN_V2"); v2 = (e && e[0] == '0') ? 0 : 1; }   // 0 = the two-pass kernel (A/B runs)
  if (v2) {
    const dim3 pgrid(max(min(items, 2 * sms), cdiv(items, AT2_MAX_ITEMS)));
    static bool attr2 = false;
    if (!attr2) {
      EEC_CUDA(cudaFuncSetAttribute(attn_fwd_v2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
      EEC_CUDA(cudaFuncSetAttribute(attn_fwd_v2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
      EEC_CUDA(cudaFuncSetAttribute(attn_fwd_v2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
      EEC_CUDA(cudaFuncSetAttribute(attn_fwd_v2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
      attr2 = true;
    }
    if (general && drop.state) launch_pdl(attn_fwd_v2_kernel<true, true>, pgrid, dim3(AT_THREADS), AT2_SMEM, st, tq, tkv, g, (__nv_bfloat16*)ctx, lse, drop, active_items(st));
    else if (general) launch_pdl(attn_fwd_v2_kernel<false, true>, pgrid, dim3(AT_THREADS), AT2_SMEM, st, tq, tkv, g, (__nv_bfloat16*)ctx, lse, drop, active_items(st));
    else if (drop.state) launch_pdl(attn_fwd_v2_kernel<true, false>, pgrid, dim3(AT_THREADS), AT2_SMEM, st, tq, tkv, g, (__nv_bfloat16*)ctx, lse, drop, active_items(st));
    else launch_pdl(attn_fwd_v2_kernel<false, false>, pgrid, dim3(AT_THREADS), AT2_SMEM, st, tq, tkv, g, (__nv_bfloat16*)ctx, lse, drop, active_items(st));
    EEC_LAUNCH_CHECK();
    return 0;
  }
  EEC_CHECK_ARG(!(general && drop.state), "attention (two-pass kernel, EEC_ATTN_V2=0): dropout with the decoder masks is implemented by the v2 kernel only");
  if (general) launch_pdl(attn_fwd_tcp_kernel<false, true>, pgrid, dim3(AT_THREADS), AT_SMEM, st, tq, tkv, g, (__nv_bfloat16*)ctx, lse, drop, active_items(st));
  else if (drop.state) launch_pdl(attn_fwd_tcp_kernel<true, false>, pgrid, dim3(AT_THREADS), AT_SMEM, st, tq, tkv, g, (__nv_bfloat16*)ctx, lse, drop, active_items(st));
  else launch_pdl(attn_fwd_tcp_kernel<false, false>, pgrid, dim3(AT_THREADS), AT_SMEM, st, tq, tkv, g, (__nv_bfloat16*)ctx, lse, drop, active_items(st));
  EEC_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int attn_fwd_tc(const void* qkv, const int32_t* key_len, void* ctx, float* lse, int B, int T, int H, int dh, const DropArgs& drop,
                cudaStream_t st) {
  EEC_CHECK_ARG(dh == DHEAD, "attn_fwd_tc: head dim must be 32");
  CUtensorMap tm;
  const int D = H * dh, D3 = 3 * D;
  if (int r = get_tmap_2d(&tm, qkv, (uint64_t)D3, (uint64_t)B * T, (uint64_t)D3 * 2, DHEAD, 128, /*SWIZZLE_64B*/ 2)) return r;
  TcGeom g{};
  g.Tq = g.Tk = T; g.H = H; g.B = B; g.q_col = 0; g.k_col = D; g.v_col = 2 * D; g.key_len = key_len; g.ldo = D;
  return fwd_launch(tm, tm, g, ctx, lse, drop, false, st);
}

// The decoder's attention on the tensor cores needs K and V in ONE row-major tensor (a packed projection output) and 16-byte aligned rows
bool attn_general_tc_ok(const eec_attn_desc* d) {
  const uintptr_t q = (uintptr_t)d->q, k = (uintptr_t)d->k, v = (uintptr_t)d->v;
  const uintptr_t lo = k < v ? k : v, hi = k < v ? v : k;
  return d->dh == DHEAD && d->ldk == d->ldv && (hi - lo) < (uintptr_t)d->ldk * 2 && (hi - lo) % 2 == 0 && (q % 16) == 0 && (lo % 16) == 0 &&
         ((hi - lo) % 16) == 0 && (d->ldq % 8) == 0 && (d->ldk % 8) == 0 && d->H * d->dh <= d->ldq;
}

static int general_geom(const eec_attn_desc* d, TcGeom& g, CUtensorMap& tq, CUtensorMap& tkv) {
  const uintptr_t k = (uintptr_t)d->k, v = (uintptr_t)d->v;
  const void* kv_base = k < v ? d->k : d->v;
  g.Tq = d->Tq; g.Tk = d->Tk; g.H = d->H; g.B = d->B;
  g.q_col = 0; g.k_col = (int)((k - (uintptr_t)kv_base) / 2); g.v_col = (int)((v - (uintptr_t)kv_base) / 2);
  g.key_len = d->key_len; g.key_bits = d->key_valid_bits; g.causal = d->causal;
  // the maps start at the given pointers (a column offset inside a wider packed tensor): [H*dh (+ the K/V column distance) columns, rows]
  // with the wide tensor's row pitch
  if (int r = get_tmap_2d(&tq, d->q, (uint64_t)d->H * d->dh, (uint64_t)d->B * d->Tq, (uint64_t)d->ldq * 2, DHEAD, 128, 2)) return r;
  const uint64_t kv_cols = (uint64_t)(g.k_col > g.v_col ? g.k_col : g.v_col) + (uint64_t)d->H * d->dh;
  if (int r = get_tmap_2d(&tkv, kv_base, kv_cols, (uint64_t)d->B * d->Tk, (uint64_t)d->ldk * 2, DHEAD, 128, 2)) return r;
  return 0;
}

int attn_general_fwd_tc(const eec_attn_desc* d, void* ctx, int ldo, float* lse, const DropArgs& drop, cudaStream_t st) {
  TcGeom g{};
  CUtensorMap tq, tkv;
  if (int r = general_geom(d, g, tq, tkv)) return r;
  g.ldo = ldo;
  return fwd_launch(tq, tkv, g, ctx, lse, drop, true, st);
}

#ifdef EEC_ATTN_TIMELINE
extern "C" int eec_debug_attn_timeline(long long* host_out) {   // 8 x 64 stamps of CTA 0's last launch
  return cudaMemcpyFromSymbol(host_out, g_attn_tl, sizeof(long long) * 8 * 64) == cudaSuccess ? 0 : 1;
}
#endif

}  // namespace eec
