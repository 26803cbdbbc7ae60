// ctc.cu -- multi-exit CTC forward+backward in one kernel (train.py:57-65 ->
// nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True)), greedy decode
// (util/beam_infer.py:21-23) and the on-device early-exit selection (north star; no reference).
//
// CTC: one CTA per (exit, utterance); thread s owns extended-target state s (2U+1 states).
// alpha rows are kept in a shared-memory ping-pong and spilled to a workspace; the beta sweep
// fuses the occupancy reduction and writes d(loss)/d(logits) = scale*(softmax - occupancy)
// directly, so log-softmax backward never runs as a separate kernel.
#include <stdlib.h>
#include "common.cuh"
#ifndef EEC_CTC_PF
#define EEC_CTC_PF 4
#endif

namespace eec {

__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// grid (B, E).  lp/grad: [E][B][T][V]; nll: [E][B]; loss_out: [E]
template <int NT>
__global__ void __launch_bounds__(NT) ctc_kernel(const float* __restrict__ lp_all, const int64_t* __restrict__ targets,
                                                 const int64_t* __restrict__ target_len, int B, int T, int V, int Lmax,
                                                 int blank, float gscale, float* __restrict__ nll_all,
                                                 float* __restrict__ loss_out, float* __restrict__ grad_all,
                                                 float* __restrict__ alpha_ws, int Smax) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* buf0 = sm;              // [Smax + 2] with 2 leading -inf pads
  float* buf1 = sm + (Smax + 2);
  float* occ = sm + 2 * (Smax + 2);            // [V]
  int* ext = reinterpret_cast<int*>(occ + V);  // [Smax]
  __shared__ float s_ll;
  const int b = blockIdx.x, e = blockIdx.y, tid = threadIdx.x;
  const long eb = (long)e * B + b;
  const float* lp = lp_all + eb * T * V;
  float* grad = grad_all ? grad_all + eb * T * V : nullptr;
  float* aw = alpha_ws + eb * (long)T * Smax;
  int U = (int)target_len[b];
  if (U > Lmax) U = Lmax;
  const int S = 2 * U + 1;
  for (int s = tid; s < Smax; s += NT) ext[s] = (s < S && (s & 1)) ? (int)targets[(long)b * Lmax + (s >> 1)] : blank;
  if (tid < 2) { buf0[tid] = -INFINITY; buf1[tid] = -INFINITY; }
  __syncthreads();
  const int s = tid;
  const bool active = s < S;
  const int my = active ? ext[s] : blank;
  const bool skip = active && s >= 2 && my != blank && my != ext[s - 2];
  float* prev = buf0 + 2;
  float* cur = buf1 + 2;
  // ---- alpha
  if (s < Smax) {
    float a = -INFINITY;
    if (active && s < 2) a = lp[my];
    prev[s] = a;
    aw[s] = a;
  }
  __syncthreads();
  for (int t = 1; t < T; ++t) {
    if (s < Smax) {
      float a = -INFINITY;
      if (active) a = lse3(prev[s], prev[s - 1], skip ? prev[s - 2] : -INFINITY) + lp[(long)t * V + my];
      cur[s] = a;
      aw[(long)t * Smax + s] = a;
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  if (tid == 0) s_ll = (S >= 2) ? lse2(prev[S - 1], prev[S - 2]) : prev[S - 1];
  __syncthreads();
  const float ll = s_ll;
  const bool feasible = (ll != -INFINITY);
  const float denom = (float)B * (float)max(U, 1);
  if (tid == 0) {
    nll_all[eb] = feasible ? -ll : 0.f;
    if (feasible && loss_out) atomicAdd(loss_out + e, -ll / denom);
  }
  if (!grad) return;
  if (!feasible) {  // zero_infinity: zero gradient rows
    for (long i = tid; i < (long)T * V; i += NT) grad[i] = 0.f;
    return;
  }
  // ---- beta sweep + occupancy + gradient.  Reuse the ping-pong with 2 TRAILING pads.
  __syncthreads();
  float* bprev = buf0;  // [Smax] + 2 trailing
  float* bcur = buf1;
  if (tid < 2) { buf0[Smax + tid] = -INFINITY; buf1[Smax + tid] = -INFINITY; }
  const bool skip_fwd = active && (s + 2 < S) && ext[s + 2] != blank && ext[s + 2] != my;
  const float sc = gscale / denom;
  for (int t = T - 1; t >= 0; --t) {
    float bv = -INFINITY;
    if (active) {
      const float l = lp[(long)t * V + my];
      if (t == T - 1) bv = (s >= S - 2) ? l : -INFINITY;
      else bv = lse3(bprev[s], bprev[s + 1], skip_fwd ? bprev[s + 2] : -INFINITY) + l;
    }
    if (s < Smax) bcur[s] = bv;
    for (int c = tid; c < V; c += NT) occ[c] = 0.f;
    __syncthreads();
    if (active) {
      const float a = aw[(long)t * Smax + s];
      if (a != -INFINITY && bv != -INFINITY) atomicAdd(&occ[my], expf(a + bv - lp[(long)t * V + my] - ll));
    }
    __syncthreads();
    for (int c = tid; c < V; c += NT) grad[(long)t * V + c] = sc * (expf(lp[(long)t * V + c]) - occ[c]);
    float* tmp = bprev; bprev = bcur; bcur = tmp;
    // next iteration's writes to bcur/occ are ordered by the two barriers above
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-per-utterance CTC (the fast path, 2L+1 <= 256): lane l owns SPL consecutive extended-target states
// in registers (even = blank, odd = label), neighbours come from ONE warp shuffle per alpha step (two per
// beta step), log-sum-exp uses MUFU ex2/lg2, emissions and alpha rows are prefetched PF steps ahead.
// No block barriers and no shared memory: 2*T dependent steps of ~150 cycles per utterance, all
// E*B utterances in flight at once.  The dense part of the gradient (scale*softmax) is written by a
// separate streaming kernel; this kernel subtracts the sparse occupancies with fp32 RED atomics.
// The recursion runs in the log2 domain with a FINITE "minus infinity" sentinel so that every
// log-sum-exp is branch-free (no inf-inf NaNs) and the compiler can interleave the lane's states.
constexpr float CTC_NEG = -1.0e30f;
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;
__device__ __forceinline__ float fex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float flg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float flse2(float a, float b) {
  const float m = fmaxf(a, b);
  return m + flg2(fex2(a - m) + fex2(b - m));
}
__device__ __forceinline__ float flse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  return m + flg2(fex2(a - m) + fex2(b - m) + fex2(c - m));
}

__global__ void ctc_grad_init_kernel(const float4* __restrict__ lp, const int64_t* __restrict__ target_len, float4* __restrict__ grad,
                                     int B, long tv4, float gscale, long total4) {
  pdl_trigger();
  pdl_wait();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)((i / tv4) % B);
    const float sc = gscale / ((float)B * (float)max((int)target_len[b], 1));
    const float4 v = lp[i];
    grad[i] = make_float4(sc * __expf(v.x), sc * __expf(v.y), sc * __expf(v.z), sc * __expf(v.w));
  }
}

// RING (V == 256): the emission rows travel global -> shared memory as 1 KB bulk copies (cp.async.bulk + mbarrier) through a
// 16-row ring per warp.  Measured before (B200, T' = 374): the alpha chain alone took 263 us = 1.35 k clk per time step although a step
// is 146 instructions -- register prefetches of the scattered emissions (4 / 8 / 12 steps ahead: no difference) cannot hide the L2 /
// HBM latency, because every wait on a load scoreboard also waits for the NEWER loads that share it.  mbarrier phases complete in
// order, so the ring really keeps 15 rows in flight; it also reads 1 KB per step instead of ~97 scattered 32-byte sectors.
constexpr int CTC_RING = 16;
__device__ __forceinline__ void ctc_row_load(float* dst, const float* src, uint64_t* bar) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(1024) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(1024), "r"(b)
               : "memory");
}
__device__ __forceinline__ void ctc_row_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
}

template <int SPL, bool RING>
__global__ void __launch_bounds__(64) ctc_warp_kernel(const float* __restrict__ lp_all, const int64_t* __restrict__ targets,
                                                      const int64_t* __restrict__ target_len, int EB, int B, int T, int V, int Lmax,
                                                      int blank, float gscale, float* __restrict__ nll_all,
                                                      float* __restrict__ loss_out, float* __restrict__ grad_all,
                                                      float* __restrict__ alpha_ws, float* __restrict__ beta_ws) {
  pdl_trigger();
  pdl_wait();
  // One CTA of TWO warps per (exit, utterance): warp 0 runs the alpha recursion forward in time while warp 1 runs the beta
  // recursion backward in time (the two T-step dependency chains are the whole cost of this kernel and are independent);
  // both leave their lattices in the workspace, then the two warps share the embarrassingly parallel occupancy pass.
  constexpr int NL = SPL / 2;   // label states per lane
  constexpr int PF = EEC_CTC_PF;         // prefetch distance (time steps)
  constexpr int SW = 32 * SPL;  // workspace row width
  __shared__ float s_ll;
  __shared__ __align__(128) float ring[RING ? 2 : 1][RING ? CTC_RING : 1][RING ? 256 : 1];
  __shared__ uint64_t rbar[2][CTC_RING];
  const int wg = blockIdx.x;
  const int role = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (RING) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 2 * CTC_RING; ++i) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&rbar[0][0] + i);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1));
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  const int e = wg / B, b = wg % B;
  const float* lp = lp_all + (long)wg * T * V;
  float* grad = grad_all ? grad_all + (long)wg * T * V : nullptr;
  float* aw = alpha_ws + (long)wg * T * SW + lane * SPL;
  float* bw = beta_ws + (long)wg * T * SW + lane * SPL;
  int U = (int)target_len[b];
  if (U > Lmax) U = Lmax;
  const int S = 2 * U + 1;
  const int s0 = lane * SPL;
  int lab[NL];
  bool skip[NL], skipf[NL];
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    const int li = lane * NL + j;   // label index of state s0 + 2j + 1
    lab[j] = (li < U) ? (int)targets[(long)b * Lmax + li] : blank;
    const int lprev = (li >= 1 && li - 1 < U) ? (int)targets[(long)b * Lmax + li - 1] : -1;
    const int lnext = (li + 1 < U) ? (int)targets[(long)b * Lmax + li + 1] : -1;
    skip[j] = (li < U) && (li >= 1) && lab[j] != lprev;
    skipf[j] = (li + 1 < U) && lab[j] != lnext;
  }
  const float denom = (float)B * (float)max(U, 1);
  float emb[PF], eml[PF][NL];
  // emission rows of this warp's recursion, in ITS time order: row(n) = base + dir * n
  const int rdir = (role == 0) ? 1 : -1, rbase = (role == 0) ? 0 : T - 1;
  float(*rg)[RING ? 256 : 1] = ring[RING ? role : 0];
  uint64_t* rb = rbar[role];
  auto ring_fetch = [&](int n, float& lb, float(&ll)[NL]) {      // wait for row n of the sequence, read this lane's emissions (log2 domain)
    const int sl = n & (CTC_RING - 1);
    ctc_row_wait(&rb[sl], (n / CTC_RING) & 1);
    lb = rg[sl][blank] * LOG2E_F;
#pragma unroll
    for (int j = 0; j < NL; ++j) ll[j] = rg[sl][lab[j]] * LOG2E_F;
  };
  auto ring_refill = [&](int n_done) {                          // every lane has read row n_done: its slot takes row n_done + CTC_RING
    __syncwarp();
    if (lane == 0 && n_done + CTC_RING < T)
      ctc_row_load(rg[n_done & (CTC_RING - 1)], lp + (long)(rbase + rdir * (n_done + CTC_RING)) * V, &rb[n_done & (CTC_RING - 1)]);
  };
  if (RING && (role == 0 || grad)) {
    if (lane == 0)
      for (int n = 0; n < CTC_RING && n < T; ++n) ctc_row_load(rg[n], lp + (long)(rbase + rdir * n) * V, &rb[n]);
    __syncwarp();
  }
  if (role == 0) {
    // ---------------- alpha (forward in time)
    float a[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) a[i] = CTC_NEG;
    auto alpha_step = [&](int t, float lb_, const float(&ll_)[NL]) {
      float prev_last = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);
      if (lane == 0) prev_last = CTC_NEG;
      float na[SPL];
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const float am1 = (i == 0) ? prev_last : a[i - 1];
        if ((i & 1) == 0) {
          na[i] = flse2(a[i], am1) + lb_;
        } else {
          const float am2 = (i == 1) ? prev_last : a[i - 2];
          na[i] = flse3(a[i], am1, skip[i >> 1] ? am2 : CTC_NEG) + ll_[i >> 1];
        }
        na[i] = (s0 + i >= S) ? CTC_NEG : fmaxf(na[i], CTC_NEG);
      }
#pragma unroll
      for (int i = 0; i < SPL; ++i) { a[i] = na[i]; aw[(long)t * SW + i] = na[i]; }
    };
    if (RING) {
      float lb_, ll_[NL], nlb = CTC_NEG, nll[NL];
      ring_fetch(0, lb_, ll_);
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s = s0 + i;
        a[i] = CTC_NEG;
        if (s == 0) a[i] = lb_;
        if (s == 1 && S > 1) a[i] = ll_[0];
      }
#pragma unroll
      for (int i = 0; i < SPL; ++i) aw[i] = a[i];
#pragma unroll
      for (int j = 0; j < NL; ++j) nll[j] = CTC_NEG;
      if (T > 1) ring_fetch(1, nlb, nll);
      for (int t = 1; t < T; ++t) {
        lb_ = nlb;
#pragma unroll
        for (int j = 0; j < NL; ++j) ll_[j] = nll[j];
        ring_refill(t - 1);
        if (t + 1 < T) ring_fetch(t + 1, nlb, nll);   // one step ahead: the shared-memory reads overlap this step's recursion
        alpha_step(t, lb_, ll_);
      }
    } else {
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s = s0 + i;
        a[i] = CTC_NEG;
        if (s == 0) a[i] = lp[blank] * LOG2E_F;
        if (s == 1 && S > 1) a[i] = lp[lab[0]] * LOG2E_F;
      }
#pragma unroll
      for (int i = 0; i < SPL; ++i) aw[i] = a[i];
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        const int t = 1 + k;
        if (t < T) {
          emb[k] = lp[(long)t * V + blank] * LOG2E_F;
#pragma unroll
          for (int j = 0; j < NL; ++j) eml[k][j] = lp[(long)t * V + lab[j]] * LOG2E_F;
        }
      }
      for (int t0 = 1; t0 < T; t0 += PF) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
          const int t = t0 + k;
          if (t < T) {
            const float lb_ = emb[k];
            float ll_[NL];
#pragma unroll
            for (int j = 0; j < NL; ++j) ll_[j] = eml[k][j];
            if (t + PF < T) {
              emb[k] = lp[(long)(t + PF) * V + blank] * LOG2E_F;
#pragma unroll
              for (int j = 0; j < NL; ++j) eml[k][j] = lp[(long)(t + PF) * V + lab[j]] * LOG2E_F;
            }
            alpha_step(t, lb_, ll_);
          }
        }
      }
    }
    // log-likelihood = lse(alpha_T-1[S-1], alpha_T-1[S-2])
    float mine = CTC_NEG;   // log2 domain
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      const int s = s0 + i;
      if (s == S - 1 || (s == S - 2 && S >= 2)) mine = flse2(mine, a[i]);
    }
    float mx = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = fex2(mine - mx);
    se = warp_sum(se);
    const float llv = mx + flg2(se);                 // log2 P(y|x)
    if (lane == 0) {
      const bool feas = llv > 0.5f * CTC_NEG;
      s_ll = llv;
      nll_all[wg] = feas ? -llv * LN2_F : 0.f;
      if (feas && loss_out) atomicAdd(loss_out + e, -llv * LN2_F / denom);
    }
  } else if (grad) {
    // ---------------- beta (backward in time); the workspace receives beta_t(s) - emission_t(s)
    float bt[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) bt[i] = CTC_NEG;
    auto beta_step = [&](int t, float lb_, const float(&ll_)[NL]) {
      float nb[SPL];
      if (t == T - 1) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const int s = s0 + i;
          nb[i] = (s < S && s >= S - 2) ? (((i & 1) == 0) ? lb_ : ll_[i >> 1]) : CTC_NEG;
        }
      } else {
        float n0 = __shfl_down_sync(0xffffffffu, bt[0], 1);
        float n1 = __shfl_down_sync(0xffffffffu, bt[1], 1);
        if (lane == 31) { n0 = CTC_NEG; n1 = CTC_NEG; }
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const float bp1 = (i == SPL - 1) ? n0 : bt[i + 1];
          if ((i & 1) == 0) {
            nb[i] = flse2(bt[i], bp1) + lb_;
          } else {
            const float bp2 = (i == SPL - 1) ? n1 : bt[(i + 2 < SPL) ? i + 2 : i];
            nb[i] = flse3(bt[i], bp1, skipf[i >> 1] ? bp2 : CTC_NEG) + ll_[i >> 1];
          }
          nb[i] = (s0 + i >= S) ? CTC_NEG : fmaxf(nb[i], CTC_NEG);
        }
      }
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        bt[i] = nb[i];
        bw[(long)t * SW + i] = nb[i] - (((i & 1) == 0) ? lb_ : ll_[i >> 1]);
      }
    };
    if (RING) {
      float lb_, ll_[NL], nlb = CTC_NEG, nll[NL];
#pragma unroll
      for (int j = 0; j < NL; ++j) nll[j] = CTC_NEG;
      ring_fetch(0, nlb, nll);
      for (int n = 0; n < T; ++n) {
        lb_ = nlb;
#pragma unroll
        for (int j = 0; j < NL; ++j) ll_[j] = nll[j];
        if (n >= 1) ring_refill(n - 1);
        if (n + 1 < T) ring_fetch(n + 1, nlb, nll);
        beta_step(T - 1 - n, lb_, ll_);
      }
    } else {
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        const int t = T - 1 - k;
        if (t >= 0) {
          emb[k] = lp[(long)t * V + blank] * LOG2E_F;
#pragma unroll
          for (int j = 0; j < NL; ++j) eml[k][j] = lp[(long)t * V + lab[j]] * LOG2E_F;
        }
      }
      for (int t0 = T - 1; t0 >= 0; t0 -= PF) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
          const int t = t0 - k;
          if (t >= 0) {
            const float lb_ = emb[k];
            float ll_[NL];
#pragma unroll
            for (int j = 0; j < NL; ++j) ll_[j] = eml[k][j];
            if (t - PF >= 0) {
              emb[k] = lp[(long)(t - PF) * V + blank] * LOG2E_F;
#pragma unroll
              for (int j = 0; j < NL; ++j) eml[k][j] = lp[(long)(t - PF) * V + lab[j]] * LOG2E_F;
            }
            beta_step(t, lb_, ll_);
          }
        }
      }
    }
  }
  __syncthreads();
  if (!grad) return;
  const float ll = s_ll;
  const bool feasible = ll > 0.5f * CTC_NEG;
  if (!feasible) {  // zero_infinity: the dense init wrote scale*softmax; zero the whole slab
    float4* g4 = reinterpret_cast<float4*>(grad);
    for (long i = threadIdx.x; i < (long)T * V / 4; i += 64) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  // ---------------- occupancy: posterior of state s at time t = 2^(alpha + (beta - emission) - ll); no dependency between time steps
  const float sc = gscale / denom;
  constexpr int OU = 8;   // time steps in flight per warp (the lattice rows come back from L2 / HBM)
  for (int t0 = role * OU; t0 < T; t0 += 2 * OU) {
    float av[OU][SPL], bv[OU][SPL];
#pragma unroll
    for (int k = 0; k < OU; ++k)
      if (t0 + k < T) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) { av[k][i] = aw[(long)(t0 + k) * SW + i]; bv[k][i] = bw[(long)(t0 + k) * SW + i]; }
      }
#pragma unroll
    for (int k = 0; k < OU; ++k) {
      const int t = t0 + k;
      if (t < T) {
        float occ_blank = 0.f;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const float o = fex2(fmaxf(av[k][i] + bv[k][i] - ll, -126.f));
          if ((i & 1) == 0) occ_blank += o;
          else if (o > 1e-30f) atomicAdd(grad + (long)t * V + lab[i >> 1], -sc * o);
        }
        occ_blank = warp_sum(occ_blank);
        if (lane == 0 && occ_blank > 1e-30f) atomicAdd(grad + (long)t * V + blank, -sc * occ_blank);
      }
    }
  }
}

// greedy collapse: one warp per utterance
__global__ void greedy_collapse_kernel(const int32_t* __restrict__ argmax, int32_t* __restrict__ tokens,
                                       int32_t* __restrict__ n_tokens, int B, int T, int blank) {
  pdl_trigger();
  pdl_wait();
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int32_t* a = argmax + (long)b * T;
  int32_t* out = tokens + (long)b * T;
  int count = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    int v = (t < T) ? a[t] : blank;
    int pv = (t > 0 && t < T) ? a[t - 1] : -1;
    bool keep = (t < T) && (v != pv) && (v != blank);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) out[count + __popc(m & ((1u << lane) - 1u))] = v;
    count += __popc(m);
  }
  for (int t = count + lane; t < T; t += 32) out[t] = -1;
  if (lane == 0) n_tokens[b] = count;
}

// early-exit selection + stable compaction bookkeeping; ONE block of 1024 threads.
__global__ void __launch_bounds__(1024) exit_select_kernel(const float* __restrict__ entropy, const int32_t* __restrict__ argmax,
                                                           const int32_t* __restrict__ key_len_alive,
                                                           const int32_t* __restrict__ row_map, int32_t* __restrict__ n_alive,
                                                           int exit_idx, int is_last, float threshold,
                                                           int32_t* __restrict__ exit_index, int32_t* __restrict__ tokens,
                                                           int32_t* __restrict__ n_tokens, int32_t* __restrict__ new_row_map,
                                                           int32_t* __restrict__ new_key_len, int32_t* __restrict__ gather_idx,
                                                           float* __restrict__ mean_entropy_out, int B, int T, int blank) {
  pdl_trigger();
  pdl_wait();
  __shared__ int done[1024];
  __shared__ int pos[1024];
  __shared__ int s_new;
  const int n = *n_alive;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = w; i < n; i += 32) {
    const int kl = min(key_len_alive[i], T);
    float h = 0.f;
    for (int t = lane; t < kl; t += 32) h += entropy[(long)i * T + t];
    h = warp_sum(h) / (float)max(kl, 1);
    const int fin = (is_last || h < threshold) ? 1 : 0;
    if (lane == 0) {
      done[i] = fin;
      if (mean_entropy_out) mean_entropy_out[(long)exit_idx * B + row_map[i]] = h;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0;
    for (int i = 0; i < n; ++i) {
      pos[i] = k;
      if (!done[i]) ++k;
    }
    s_new = k;
  }
  __syncthreads();
  // survivors: compacted bookkeeping (read old maps before anyone overwrites: separate out arrays)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (!done[i]) {
      const int j = pos[i];
      new_row_map[j] = row_map[i];
      new_key_len[j] = key_len_alive[i];
      gather_idx[j] = i;
    }
  }
  // finalised rows: greedy tokens of this exit into the ORIGINAL row
  for (int i = w; i < n; i += 32) {
    if (!done[i]) continue;
    const int orig = row_map[i];
    const int32_t* a = argmax + (long)i * T;
    int32_t* out = tokens + (long)orig * T;
    int count = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      int v = (t < T) ? a[t] : blank;
      int pv = (t > 0 && t < T) ? a[t - 1] : -1;
      bool keep = (t < T) && (v != pv) && (v != blank);
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) out[count + __popc(m & ((1u << lane) - 1u))] = v;
      count += __popc(m);
    }
    for (int t = count + lane; t < T; t += 32) out[t] = -1;
    if (lane == 0) { n_tokens[orig] = count; exit_index[orig] = exit_idx; }
  }
  __syncthreads();
  if (threadIdx.x == 0) *n_alive = s_new;
}

__global__ void gather_rows_kernel(const float4* __restrict__ x, float4* __restrict__ y, const int32_t* __restrict__ gather_idx,
                                   const int32_t* __restrict__ n_alive, long row_vec) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.y;
  if (j >= *n_alive) return;
  const long src = (long)gather_idx[j] * row_vec, dst = (long)j * row_vec;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < row_vec; i += (long)gridDim.x * blockDim.x) y[dst + i] = x[src + i];
}

// ---- front-end im2col / col2im for Conv1d(k=3, s=2)
template <typename TI, typename TO>
__global__ void im2col_k3s2_kernel(const TI* __restrict__ in, long sb, long sc, long st, TO* __restrict__ out, int ldo, int B,
                                   int C, int T_out) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)B * T_out * ldo;
  if (i >= total) return;
  int k = (int)(i % ldo);
  long r = i / ldo;
  int t = (int)(r % T_out), b = (int)(r / T_out);
  float v = 0.f;
  if (k < 3 * C) {
    int c = k / 3, j = k % 3;
    v = ld_as_float<TI>(in + (long)b * sb + (long)c * sc + (long)(2 * t + j) * st);
  }
  st_from_float<TO>(out + i, v);
}

// Fast path A: channel-major input with contiguous time (the fbank, in[b][c][f], st == 1).  Block = (32 output frames, utterance):
// the 65-frame span of every channel is read along time (coalesced), transposed through shared memory, and every output row
// leaves as 16-byte stores.  (The element-per-thread kernel above touched a different cache line per thread: 66 us per launch.)
constexpr int IM_TT = 32;                 // output frames per block
constexpr int IM_SPAN = 2 * IM_TT + 1;    // input frames they read
template <typename TO>
__global__ void __launch_bounds__(256) im2col_k3s2_tmajor_kernel(const float* __restrict__ in, long sb, long sc, TO* __restrict__ out, int ldo,
                                                              int C, int T_out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float im_tile[];      // [C][IM_SPAN + 1]
  const int b = blockIdx.y, t0 = blockIdx.x * IM_TT;
  const int f0 = 2 * t0, f_last = 2 * (T_out - 1) + 2;   // last input frame any output of this utterance reads
  const float* src = in + (long)b * sb;
  for (int idx = threadIdx.x; idx < C * IM_SPAN; idx += 256) {
    const int c = idx / IM_SPAN, x = idx - c * IM_SPAN;
    im_tile[c * (IM_SPAN + 1) + x] = (f0 + x <= f_last) ? src[(long)c * sc + f0 + x] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < IM_TT; r += 8) {
    const int t = t0 + r;
    if (t >= T_out) break;
    TO* orow = out + ((long)b * T_out + t) * ldo;
    for (int k0 = lane * 8; k0 < ldo; k0 += 256) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e, c = k / 3, j = k - 3 * c;
        v[e] = (k < 3 * C) ? im_tile[c * (IM_SPAN + 1) + 2 * r + j] : 0.f;
      }
      st8<TO>(orow + k0, v);
    }
  }
}
// Fast path B: frame-major input (in[b][f][c], sc == 1): one warp per output row reads its three input rows with 128-bit loads and
// writes the (c*3 + j)-interleaved row with 16-byte stores.
template <typename TO>
__global__ void __launch_bounds__(256) im2col_k3s2_fmajor_kernel(const float* __restrict__ in, long sb, long st, TO* __restrict__ out, int ldo,
                                                              int B, int C, int T_out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float im_rows[];      // [8 warps][3][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + warp;
  if (row >= (long)B * T_out) return;
  const int b = (int)(row / T_out), t = (int)(row % T_out);
  float* mine = im_rows + warp * 3 * C;
  const float* src = in + (long)b * sb + (long)(2 * t) * st;
  for (int j = 0; j < 3; ++j)
    for (int c = lane * 4; c < C; c += 128)
      *reinterpret_cast<float4*>(mine + j * C + c) = *reinterpret_cast<const float4*>(src + (long)j * st + c);
  __syncwarp();
  TO* orow = out + row * ldo;
  for (int k0 = lane * 8; k0 < ldo; k0 += 256) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k0 + e, c = k / 3, j = k - 3 * c;
      v[e] = (k < 3 * C) ? mine[j * C + c] : 0.f;
    }
    st8<TO>(orow + k0, v);
  }
}

__global__ void col2im_k3s2_kernel(const float* __restrict__ dcols, int ldc, float* __restrict__ dx, int B, int C, int T_in,
                                   int T_out) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)B * T_in * C;
  if (i >= total) return;
  int c = (int)(i % C);
  long r = i / C;
  int f = (int)(r % T_in), b = (int)(r / T_in);
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    int ft = f - j;
    if (ft >= 0 && (ft & 1) == 0 && (ft >> 1) < T_out) v += dcols[((long)b * T_out + (ft >> 1)) * ldc + c * 3 + j];
  }
  dx[i] = v;
}

}  // namespace eec

using namespace eec;

static int ctc_smax(int Lmax) { return ((2 * Lmax + 1 + 31) / 32) * 32; }
// states per lane of the warp kernel (0 = too long, use the block kernel)
static int ctc_spl(int Lmax) {
  const int S = 2 * Lmax + 1;
  if (S <= 64) return 2;
  if (S <= 128) return 4;
  if (S <= 192) return 6;
  if (S <= 256) return 8;
  return 0;
}

extern "C" int64_t eec_ctc_workspace_bytes(int E, int B, int T, int Lmax) {
  const int spl = ctc_spl(Lmax);
  const int width = spl ? 32 * spl : ctc_smax(Lmax);
  // warp kernel: an alpha lattice AND a beta lattice (the two recursions run concurrently on two warps)
  return (int64_t)(spl ? 2 : 1) * E * B * T * width * (int64_t)sizeof(float);
}

extern "C" int eec_ctc_fwd_bwd(const float* lp, const int64_t* targets, const int64_t* target_len, int E, int B, int T,
                               int V, int Lmax, int blank, float gscale, float* nll, float* loss_out, float* grad,
                               void* workspace, eec_stream_t stream) {
  if (E == 0 || B == 0) return 0;
  EEC_CHECK_ARG(T >= 1, "ctc: T must be >= 1");
  EEC_CHECK_ARG(workspace != nullptr, "ctc: workspace is NULL");
  const int spl = ctc_spl(Lmax);
  if (spl && V % 4 == 0) {
    float* wsf = reinterpret_cast<float*>(workspace);
    if (grad) {
      const long total4 = (long)E * B * T * V / 4;
      const int blocks = (int)min((long)148 * 16, cdiv64(total4, 256));
      launch_pdl(ctc_grad_init_kernel, dim3(blocks), dim3(256), 0, S(stream), (const float4*)lp, target_len, (float4*)grad, B, (long)T * V / 4, gscale, total4);
      EEC_LAUNCH_CHECK();
    }
    const int EB = E * B;
    float* wsb = wsf + (long)EB * T * 32 * spl;
    static int ring_env = -1;
    if (ring_env < 0) { const char* e = getenv("EEC_CTC_RING"); ring_env = (e && e[0] == '0') ? 0 : 1; }
    const bool ring = ring_env && V == 256 && (reinterpret_cast<uintptr_t>(lp) & 15) == 0;   // 1 KB emission rows as bulk copies
#define EEC_CTC_WARP(SPLV)                                                                                              \
  do {                                                                                                                  \
    if (ring) launch_pdl(ctc_warp_kernel<SPLV, true>, dim3(EB), dim3(64), 0, S(stream), lp, targets, target_len, EB, B, T, V, Lmax, blank, gscale, nll, \
                         loss_out, grad, wsf, wsb);                                                                     \
    else launch_pdl(ctc_warp_kernel<SPLV, false>, dim3(EB), dim3(64), 0, S(stream), lp, targets, target_len, EB, B, T, V, Lmax, blank, gscale, nll, \
                    loss_out, grad, wsf, wsb);                                                                          \
  } while (0)
    if (spl == 2) EEC_CTC_WARP(2);
    else if (spl == 4) EEC_CTC_WARP(4);
    else if (spl == 6) EEC_CTC_WARP(6);
    else EEC_CTC_WARP(8);
    EEC_LAUNCH_CHECK();
    return 0;
  }
  const int Smax = ctc_smax(Lmax);
  EEC_CHECK_ARG(Smax <= 1024, "ctc: target length %d too long (2L+1 must be <= 1024)", Lmax);
  size_t smem = (size_t)(2 * (Smax + 2) + V) * sizeof(float) + (size_t)Smax * sizeof(int);
  dim3 grid(B, E);
  float* ws = reinterpret_cast<float*>(workspace);
  if (Smax <= 256)
    launch_pdl(ctc_kernel<256>, dim3(grid), dim3(256), smem, S(stream), lp, targets, target_len, B, T, V, Lmax, blank, gscale, nll, loss_out, grad, ws, Smax);
  else if (Smax <= 512)
    launch_pdl(ctc_kernel<512>, dim3(grid), dim3(512), smem, S(stream), lp, targets, target_len, B, T, V, Lmax, blank, gscale, nll, loss_out, grad, ws, Smax);
  else
    launch_pdl(ctc_kernel<1024>, dim3(grid), dim3(1024), smem, S(stream), lp, targets, target_len, B, T, V, Lmax, blank, gscale, nll, loss_out, grad, ws, Smax);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_greedy_collapse(const int32_t* argmax, int32_t* tokens, int32_t* n_tokens, int B, int T, int blank,
                                   eec_stream_t stream) {
  if (B == 0) return 0;
  launch_pdl(greedy_collapse_kernel, dim3(cdiv(B * 32, 128)), dim3(128), 0, S(stream), argmax, tokens, n_tokens, B, T, blank);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_exit_select(const float* entropy, const int32_t* argmax, const int32_t* key_len_alive,
                               const int32_t* row_map, int32_t* n_alive, int exit_idx, int is_last, float threshold,
                               int32_t* exit_index, int32_t* tokens, int32_t* n_tokens, int32_t* new_row_map,
                               int32_t* new_key_len, int32_t* gather_idx, float* mean_entropy_out, int B, int T, int blank,
                               eec_stream_t stream) {
  EEC_CHECK_ARG(B <= 1024, "exit_select: batch must be <= 1024 (got %d)", B);
  if (B == 0) return 0;
  launch_pdl(exit_select_kernel, dim3(1), dim3(1024), 0, S(stream), entropy, argmax, key_len_alive, row_map, n_alive, exit_idx, is_last, threshold,
                                               exit_index, tokens, n_tokens, new_row_map, new_key_len, gather_idx,
                                               mean_entropy_out, B, T, blank);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_gather_rows(const float* x, float* y, const int32_t* gather_idx, const int32_t* n_alive, int B,
                               int64_t row_elems, eec_stream_t stream) {
  EEC_CHECK_ARG(row_elems % 4 == 0, "gather_rows: row_elems %% 4");
  if (B == 0) return 0;
  long rv = row_elems / 4;
  dim3 grid((unsigned)min((long)32, cdiv64(rv, 256)), B);
  launch_pdl(gather_rows_kernel, dim3(grid), dim3(256), 0, S(stream), (const float4*)x, (float4*)y, gather_idx, n_alive, rv);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_im2col_k3s2(const void* in, int in_dtype, int64_t sb, int64_t sc, int64_t st, void* out, int out_dtype,
                               int ldo, int B, int C, int T_out, eec_stream_t stream) {
  EEC_CHECK_ARG(ldo >= 3 * C, "im2col: ldo < 3*C");
  EEC_CHECK_ARG(in_dtype == EEC_F32, "im2col: input must be fp32");
  long total = (long)B * T_out * ldo;
  if (total == 0) return 0;
  const bool aligned = ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (aligned && st == 1 && (size_t)C * (IM_SPAN + 1) * 4 <= 48 * 1024) {   // channel-major, time contiguous (the fbank)
    const dim3 grid(cdiv(T_out, IM_TT), B);
    const size_t smem = (size_t)C * (IM_SPAN + 1) * 4;
    if (out_dtype == EEC_F32) launch_pdl(im2col_k3s2_tmajor_kernel<float>, grid, dim3(256), smem, S(stream), (const float*)in, (long)sb, (long)sc, (float*)out, ldo, C, T_out);
    else launch_pdl(im2col_k3s2_tmajor_kernel<__nv_bfloat16>, grid, dim3(256), smem, S(stream), (const float*)in, (long)sb, (long)sc, (__nv_bfloat16*)out, ldo, C, T_out);
    EEC_LAUNCH_CHECK();
    return 0;
  }
  if (aligned && sc == 1 && C % 4 == 0 && st % 4 == 0 && sb % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (size_t)C * 96 <= 48 * 1024) {   // frame-major
    const dim3 grid((unsigned)cdiv64((long)B * T_out, 8));
    const size_t smem = (size_t)8 * 3 * C * 4;
    if (out_dtype == EEC_F32) launch_pdl(im2col_k3s2_fmajor_kernel<float>, grid, dim3(256), smem, S(stream), (const float*)in, (long)sb, (long)st, (float*)out, ldo, B, C, T_out);
    else launch_pdl(im2col_k3s2_fmajor_kernel<__nv_bfloat16>, grid, dim3(256), smem, S(stream), (const float*)in, (long)sb, (long)st, (__nv_bfloat16*)out, ldo, B, C, T_out);
    EEC_LAUNCH_CHECK();
    return 0;
  }
  int blocks = (int)cdiv64(total, 256);
  if (out_dtype == EEC_F32)
    launch_pdl(im2col_k3s2_kernel<float, float>, dim3(blocks), dim3(256), 0, S(stream), (const float*)in, sb, sc, st, (float*)out, ldo, B, C, T_out);
  else
    launch_pdl(im2col_k3s2_kernel<float, __nv_bfloat16>, dim3(blocks), dim3(256), 0, S(stream), (const float*)in, sb, sc, st, (__nv_bfloat16*)out, ldo, B, C, T_out);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_col2im_k3s2(const float* dcols, int ldc, float* dx, int B, int C, int T_in, int T_out,
                               eec_stream_t stream) {
  long total = (long)B * T_in * C;
  if (total == 0) return 0;
  launch_pdl(col2im_k3s2_kernel, dim3((int)cdiv64(total, 256)), dim3(256), 0, S(stream), dcols, ldc, dx, B, C, T_in, T_out);
  EEC_LAUNCH_CHECK();
  return 0;
}
