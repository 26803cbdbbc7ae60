// convmod.cu -- interior of the conformer convolution module (TA:52-65): depthwise conv k=31
// with SAME zero padding per utterance, BatchNorm1d (eval: folded; train: batch statistics
// over all B*T frames incl. padding), SiLU, GLU backward, and the matching backward kernels.
//
// Depthwise kernel shape: one thread per channel (a warp reads 32 contiguous channels of one
// frame = one coalesced line), a block walks a tile of TT output frames; every input frame is
// loaded ONCE into a register and scattered into the <=31 accumulators it feeds (loops are
// fully unrolled, so tap indices are compile-time).  Algorithmic bytes: read g + write out.
#include <stdlib.h>
#include "common.cuh"

namespace eec {

constexpr int KW = 31;
constexpr int HALF = 15;
constexpr int TT = 32;     // output frames per block (62 input frames incl. halo)
constexpr float BN_EPS = 1e-5f;

enum { DW_EVAL = 0, DW_STATS = 1, DW_BWD_DATA = 2 };

// copy rows [t_first, t_first + TT+30) x 256 channels of one utterance into smem (row pitch 256 elements)
template <typename TI>
__device__ __forceinline__ void stage_tile(TI* tile, const TI* __restrict__ src, int t_first, int T, int C) {
  constexpr int ROWS = TT + KW - 1;
  constexpr int CPR = 256 * sizeof(TI) / 16;   // 16-byte chunks per row
  for (int i = threadIdx.x; i < ROWS * CPR; i += 256) {
    const int r = i / CPR, c = i % CPR;
    const int t = t_first + r;
    uint8_t* dst = reinterpret_cast<uint8_t*>(tile) + (size_t)i * 16;
    if (t >= 0 && t < T) {
      const uint8_t* s = reinterpret_cast<const uint8_t*>(src + (long)t * C) + c * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(s) : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}
template <typename TI> constexpr int dw_smem_bytes() { return (TT + KW - 1) * 256 * (int)sizeof(TI); }

template <typename TI, typename TO, int MODE>
__global__ void __launch_bounds__(256, 3) dwconv_kernel(const TI* __restrict__ g, const float* __restrict__ w,
                                                     const float* __restrict__ bias, const float* __restrict__ bn_scale_src_w,
                                                     const float* __restrict__ bn_b, const float* __restrict__ run_mean,
                                                     const float* __restrict__ run_var, TO* __restrict__ out,
                                                     double* __restrict__ sums, int T, int C, const ActiveItems act_items) {
  pdl_trigger();
  pdl_wait();
  if (MODE == DW_EVAL && act_items.n_dev && (int)blockIdx.y >= active_count(act_items)) return;   // utterance past the active-item limit
  const int ch = blockIdx.z * 256 + threadIdx.x;
  const int b = blockIdx.y, t0 = blockIdx.x * TT;
  float wr[KW];
#pragma unroll
  for (int j = 0; j < KW; ++j) wr[j] = (MODE == DW_BWD_DATA) ? w[ch * KW + (KW - 1 - j)] : w[ch * KW + j];
  const float bv = (MODE == DW_BWD_DATA) ? 0.f : bias[ch];
  float acc[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t) acc[t] = bv;
  // stage the (TT+30) x 256 input tile in shared memory: 16-byte cp.async per thread, zero rows outside [0,T)
  extern __shared__ __align__(16) uint8_t dw_smem[];
  TI* tile = reinterpret_cast<TI*>(dw_smem);
  stage_tile<TI>(tile, g + (long)b * T * C + blockIdx.z * 256, t0 - HALF, T, C);
#pragma unroll
  for (int r = 0; r < TT + KW - 1; ++r) {
    const float x = ld_as_float<TI>(tile + r * 256 + threadIdx.x);
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      const int j = r - t;
      if (j >= 0 && j < KW) acc[t] = fmaf(wr[j], x, acc[t]);
    }
  }
  TO* ob = out + (long)b * T * C + ch;
  if (MODE == DW_EVAL) {
    const float sc = bn_scale_src_w[ch] * rsqrtf(run_var[ch] + BN_EPS);
    const float sh = bn_b[ch] - run_mean[ch] * sc;
#pragma unroll
    for (int t = 0; t < TT; ++t)
      if (t0 + t < T) {
        float n = fmaf(acc[t], sc, sh);
        st_from_float<TO>(ob + (long)(t0 + t) * C, n * sigmoid_acc(n));
      }
  } else if (MODE == DW_STATS) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int t = 0; t < TT; ++t)
      if (t0 + t < T) {
        st_from_float<TO>(ob + (long)(t0 + t) * C, acc[t]);
        s1 += acc[t];
        s2 = fmaf(acc[t], acc[t], s2);
      }
    atomicAdd(sums + ch, (double)s1);
    atomicAdd(sums + C + ch, (double)s2);
  } else {
#pragma unroll
    for (int t = 0; t < TT; ++t)
      if (t0 + t < T) st_from_float<TO>(ob + (long)(t0 + t) * C, acc[t]);
  }
}

// Persistent, double-buffered variant for bf16 inputs (forward: eval and train pass A).  The kernel above stages a tile, waits, then
// computes: with three blocks per SM (register-limited) the FMA pipe idles while tiles are in flight (ncu: 55-63 % issue active).
// Here every block walks tiles with a stride of gridDim.x and the cp.async copy of tile i+1 runs under the 992 FMAs per thread of
// tile i; the depthwise weights stay in registers across tiles and the BatchNorm statistics of pass A are accumulated per thread
// over all of a block's tiles (one pair of double atomics per thread instead of one per tile).
template <typename TO, int MODE>
__global__ void __launch_bounds__(256, 3) dwconv_pipe_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ bn_w,
                                                          const float* __restrict__ bn_b, const float* __restrict__ run_mean,
                                                          const float* __restrict__ run_var, TO* __restrict__ out,
                                                          double* __restrict__ sums, int B, int T, const ActiveItems act_items) {
  pdl_trigger();
  pdl_wait();
  constexpr int C = 256, ROWS = TT + KW - 1, CPR = C * 2 / 16;   // 32 16-byte chunks per bf16 row
  extern __shared__ __align__(16) uint8_t dw_smem[];
  const int ch = threadIdx.x;
  const int nt = (T + TT - 1) / TT;
  const int Beff = (MODE == DW_EVAL && act_items.n_dev) ? min(B, active_count(act_items)) : B;
  const int tiles = nt * Beff;
  float wr[KW];
#pragma unroll
  for (int j = 0; j < KW; ++j) wr[j] = w[ch * KW + j];
  const float bv = bias[ch];
  float sc = 0.f, sh = 0.f;
  if (MODE == DW_EVAL) {
    sc = bn_w[ch] * rsqrtf(run_var[ch] + BN_EPS);
    sh = bn_b[ch] - run_mean[ch] * sc;
  }
  auto issue = [&](int tile, int buf) {
    const int b = tile / nt, t_first = (tile - b * nt) * TT - HALF;
    const __nv_bfloat16* src = g + (long)b * T * C;
    uint8_t* base = dw_smem + buf * (ROWS * C * 2);
    for (int i = threadIdx.x; i < ROWS * CPR; i += 256) {
      const int r = i / CPR, c = i % CPR;
      const int t = t_first + r;
      uint8_t* dst = base + (size_t)i * 16;
      if (t >= 0 && t < T) {
        const uint8_t* sp = reinterpret_cast<const uint8_t*>(src + (long)t * C) + c * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(sp) : "memory");
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float s1 = 0.f, s2 = 0.f;
  int tile = blockIdx.x, buf = 0;
  if (tile < tiles) issue(tile, 0);
  for (; tile < tiles; tile += gridDim.x, buf ^= 1) {
    const int nxt = tile + gridDim.x;
    if (nxt < tiles) {
      issue(nxt, buf ^ 1);                                   // (its previous readers passed the barrier at the end of the last iteration)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const __nv_bfloat16* tl = reinterpret_cast<const __nv_bfloat16*>(dw_smem + buf * (ROWS * C * 2));
    float acc[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) acc[t] = bv;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float x = __bfloat162float(tl[r * C + ch]);
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        const int j = r - t;
        if (j >= 0 && j < KW) acc[t] = fmaf(wr[j], x, acc[t]);
      }
    }
    const int b = tile / nt, t0 = (tile - b * nt) * TT;
    TO* ob = out + (long)b * T * C + ch;
    if (MODE == DW_EVAL) {
#pragma unroll
      for (int t = 0; t < TT; ++t)
        if (t0 + t < T) {
          const float n = fmaf(acc[t], sc, sh);
          st_from_float<TO>(ob + (long)(t0 + t) * C, n * sigmoid_acc(n));
        }
    } else {
#pragma unroll
      for (int t = 0; t < TT; ++t)
        if (t0 + t < T) {
          st_from_float<TO>(ob + (long)(t0 + t) * C, acc[t]);
          s1 += acc[t];
          s2 = fmaf(acc[t], acc[t], s2);
        }
    }
    __syncthreads();                                          // every thread has read `buf` before the next iteration refills it
  }
  if (MODE == DW_STATS && blockIdx.x < tiles) {
    atomicAdd(sums + ch, (double)s1);
    atomicAdd(sums + C + ch, (double)s2);
  }
}

// dW[ch][j] += sum_{b,t} dc[b,t,ch] * g[b,t+j-15,ch];  dbias[ch] += sum dc
// grid (time tiles, B, C/256): one block per (tile, utterance) writes its 32 x 256 partial sums ([31 taps + bias][ch], coalesced)
// into the workspace; dwconv_wgrad_reduce_kernel sums the partials.  (Per-block fp32 atomics onto the 7 936 addresses cost more
// than the convolution itself: 85 us at 3 M atomics, 171 us at 6 M.)
template <typename TI>
__global__ void __launch_bounds__(256, 3) dwconv_wgrad_kernel(const float* __restrict__ dc, const TI* __restrict__ g,
                                                           float* __restrict__ partial, int T, int C) {
  pdl_trigger();
  pdl_wait();
  const int ch = blockIdx.z * 256 + threadIdx.x;
  const int b = blockIdx.y, t0 = blockIdx.x * TT;
  float acc[KW];
  float sb = 0.f;
#pragma unroll
  for (int j = 0; j < KW; ++j) acc[j] = 0.f;
  extern __shared__ __align__(16) uint8_t dw_smem[];
  TI* tile = reinterpret_cast<TI*>(dw_smem);
  const float* db = dc + (long)b * T * C + ch;
  float d[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    d[t] = (t0 + t < T) ? db[(long)(t0 + t) * C] : 0.f;
    sb += d[t];
  }
  stage_tile<TI>(tile, g + (long)b * T * C + blockIdx.z * 256, t0 - HALF, T, C);
#pragma unroll
  for (int r = 0; r < TT + KW - 1; ++r) {
    const float x = ld_as_float<TI>(tile + r * 256 + threadIdx.x);
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      const int j = r - t;
      if (j >= 0 && j < KW) acc[j] = fmaf(d[t], x, acc[j]);
    }
  }
  // partial[((z * nblk + blk) * 32 + j) * 256 + threadIdx.x]
  const long blk = (long)blockIdx.y * gridDim.x + blockIdx.x;
  float* pp = partial + ((long)blockIdx.z * gridDim.x * gridDim.y + blk) * 32 * 256 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < KW; ++j) pp[j * 256] = acc[j];
  pp[KW * 256] = sb;
}

// grid (32, C/256, RCH), 256 threads: block (j, z, c) sums partial[z][i][j][:] over its chunk of the nblk blocks and adds the
// result into dw / dbias (RCH fp32 atomics per output element in total)
constexpr int RCH = 8;
__global__ void __launch_bounds__(256) dwconv_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                  float* __restrict__ dbias, int nblk) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.x, z = blockIdx.y;
  const int per = (nblk + RCH - 1) / RCH;
  const int i0 = blockIdx.z * per, i1 = min(nblk, i0 + per);
  const float* pp = partial + ((long)z * nblk * 32 + j) * 256 + threadIdx.x;
  float s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.f;
  int i = i0;
  for (; i + 8 <= i1; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += pp[(long)(i + k) * 32 * 256];
  }
  for (; i < i1; ++i) s[0] += pp[(long)i * 32 * 256];
  const float tot = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  const int ch = z * 256 + threadIdx.x;
  if (j < KW) atomicAdd(dw + ch * KW + j, tot);
  else atomicAdd(dbias + ch, tot);
}

// ---- streaming kernels of the conv module (train-mode BatchNorm + SiLU forward / backward, GLU backward).
// A warp owns half frames: lane l holds channels [4l, 4l+4) of a 128-channel half of the 256-channel slice, so a row half
// is ONE fully coalesced 512 B (fp32) / 256 B (bf16) access per tensor; a warp keeps RB rows in flight and strides over
// the rows of a persistent grid (3 CTAs per SM), so loads, math and stores of different warps overlap instead of moving
// in lock-step.
// Per-channel parameters live in registers.  bf16 activations use the one-MUFU sigmoid (tanh.approx); the fp32 parity
// path keeps the accurate one.
constexpr int RB = 4;            // rows in flight per warp
constexpr int SW_WARPS = 8;      // warps per CTA
__device__ __forceinline__ float sigmoid_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(0.5f * x));
  return fmaf(y, 0.5f, 0.5f);
}
template <bool FAST> __device__ __forceinline__ float sigm(float x) { return FAST ? sigmoid_tanh(x) : sigmoid_acc(x); }

constexpr int CPL = 4;           // channels per lane
template <typename T> struct Row8;   // CPL consecutive channels of one row
template <> struct Row8<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 f = *reinterpret_cast<const float4*>(p); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Row8<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 u;
    *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// batch statistics -> per-channel mean / rstd (no double division or square root on the hot path)
template <bool FAST>
__device__ __forceinline__ void bn_mean_rstd(const double* sums, int C, int ch, double inv_n, float& mean, float& rstd, double& var_d) {
  const double mean_d = sums[ch] * inv_n;
  var_d = fma(-mean_d, mean_d, sums[C + ch] * inv_n);
  if (var_d < 0.0) var_d = 0.0;
  mean = (float)mean_d;
  rstd = FAST ? rsqrtf((float)var_d + BN_EPS) : 1.0f / sqrtf((float)var_d + BN_EPS);
}

template <typename TO>
__global__ void __launch_bounds__(SW_WARPS * 32, 3) bn_silu_train_kernel(const float* __restrict__ c, const double* __restrict__ sums,
                                                            const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                            float* __restrict__ run_mean, float* __restrict__ run_var,
                                                            int64_t* __restrict__ nbt, float momentum,
                                                            float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                            TO* __restrict__ out, int rows, int C, long stat_rows) {
  pdl_trigger();
  pdl_wait();
  constexpr bool FAST = sizeof(TO) == 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch0 = blockIdx.y * 256 + (warp & 1) * 128 + lane * CPL;
  const double inv_n = 1.0 / (double)stat_rows;   // (data parallel with synchronised statistics: rows of ALL ranks)
  float sc[CPL], sh[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int ch = ch0 + k;
    float mean, rstd;
    double var_d;
    bn_mean_rstd<FAST>(sums, C, ch, inv_n, mean, rstd, var_d);
    sc[k] = rstd * bn_w[ch];
    sh[k] = fmaf(-mean, sc[k], bn_b[ch]);
    if (blockIdx.x == 0 && warp < 2) {   // warps 0 and 1 cover the two channel halves
      save_mean[ch] = mean;
      save_rstd[ch] = rstd;
      if (run_mean) {
        run_mean[ch] = (1.f - momentum) * run_mean[ch] + momentum * mean;
        const double unb = (stat_rows > 1) ? var_d * ((double)stat_rows / ((double)stat_rows - 1.0)) : var_d;
        run_var[ch] = (1.f - momentum) * run_var[ch] + momentum * (float)unb;
        if (ch == 0 && nbt) *nbt += 1;
      }
    }
  }
  const int stride = gridDim.x * (SW_WARPS / 2);
  for (int r = blockIdx.x * (SW_WARPS / 2) + (warp >> 1); r < rows; r += stride * RB) {
    float v[RB][CPL];
#pragma unroll
    for (int i = 0; i < RB; ++i)
      if (r + i * stride < rows) Row8<float>::ld(c + (long)(r + i * stride) * C + ch0, v[i]);
#pragma unroll
    for (int i = 0; i < RB; ++i)
      if (r + i * stride < rows) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const float y = fmaf(v[i][k], sc[k], sh[k]);
          v[i][k] = y * sigm<FAST>(y);
        }
        Row8<TO>::st(out + (long)(r + i * stride) * C + ch0, v[i]);
      }
  }
}

// APPLY = false: sums2[ch] += sum dn, sums2[C + ch] += sum dn * nhat   (dn = d/dn of SiLU(BN(c)))
// APPLY = true : dc = gamma * rstd * (dn - mean(dn) - nhat * mean(dn * nhat)); dgamma / dbeta from sums2
template <typename TI, bool APPLY>
__global__ void __launch_bounds__(SW_WARPS * 32, 3) bn_silu_bwd_kernel(const TI* __restrict__ ds, const float* __restrict__ c,
                                                          const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                          const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                          double* __restrict__ sums2, float* __restrict__ dc,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int C, long stat_rows,
                                                          const double* __restrict__ sums2_local) {
  pdl_trigger();
  pdl_wait();
  constexpr bool FAST = sizeof(TI) == 2;
  __shared__ float red[2][SW_WARPS / 2][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch0 = blockIdx.y * 256 + (warp & 1) * 128 + lane * CPL;
  float a1[CPL], a0[CPL], gam[CPL], bet[CPL], m1[CPL], m2[CPL], s1[CPL], s2[CPL];   // nhat = c * a1 + a0
  const float inv_n = 1.0f / (float)stat_rows;
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int ch = ch0 + k;
    a1[k] = save_rstd[ch];
    a0[k] = -save_mean[ch] * a1[k];
    gam[k] = bn_w[ch]; bet[k] = bn_b[ch];
    s1[k] = 0.f; s2[k] = 0.f;
    if (APPLY) {
      const float t1 = (float)sums2[ch], t2 = (float)sums2[C + ch];
      m1[k] = t1 * inv_n;
      m2[k] = t2 * inv_n;
      if (blockIdx.x == 0 && warp < 2) {   // warps 0 and 1 cover the two channel halves
        // synchronised statistics: the parameter gradients are THIS rank's sums (the gradient all-reduce averages them)
        atomicAdd(dbeta + ch, sums2_local ? (float)sums2_local[ch] : t1);
        atomicAdd(dgamma + ch, sums2_local ? (float)sums2_local[C + ch] : t2);
      }
    }
  }
  const int stride = gridDim.x * (SW_WARPS / 2);
  for (int r = blockIdx.x * (SW_WARPS / 2) + (warp >> 1); r < rows; r += stride * RB) {
    float v[RB][CPL], g[RB][CPL];
#pragma unroll
    for (int i = 0; i < RB; ++i)
      if (r + i * stride < rows) {
        Row8<float>::ld(c + (long)(r + i * stride) * C + ch0, v[i]);
        Row8<TI>::ld(ds + (long)(r + i * stride) * C + ch0, g[i]);
      }
#pragma unroll
    for (int i = 0; i < RB; ++i)
      if (r + i * stride < rows) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const float nh = fmaf(v[i][k], a1[k], a0[k]);
          const float y = fmaf(nh, gam[k], bet[k]);
          const float sg = sigm<FAST>(y);
          const float dn = g[i][k] * sg * fmaf(y, 1.f - sg, 1.f);
          if (APPLY) v[i][k] = gam[k] * a1[k] * (dn - m1[k] - nh * m2[k]);
          else { s1[k] += dn; s2[k] = fmaf(dn, nh, s2[k]); }
        }
        if (APPLY) Row8<float>::st(dc + (long)(r + i * stride) * C + ch0, v[i]);
      }
  }
  if (!APPLY) {
#pragma unroll
    for (int k = 0; k < CPL; ++k) { red[0][warp >> 1][(warp & 1) * 128 + lane * CPL + k] = s1[k]; red[1][warp >> 1][(warp & 1) * 128 + lane * CPL + k] = s2[k]; }
    __syncthreads();
    const int ch = threadIdx.x;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < SW_WARPS / 2; ++w) { a += red[0][w][ch]; b += red[1][w][ch]; }
    atomicAdd(sums2 + blockIdx.y * 256 + ch, (double)a);
    atomicAdd(sums2 + C + blockIdx.y * 256 + ch, (double)b);
  }
}

// z = [a | gate] (rows x 2C), dg (rows x C) -> dz = [dg * sigmoid(gate) | dg * a * sigmoid(gate) * (1 - sigmoid(gate))]
template <typename T>
__global__ void __launch_bounds__(SW_WARPS * 32, 3) glu_bwd_kernel(const T* __restrict__ z, const T* __restrict__ dg, T* __restrict__ dz, int rows, int C) {
  pdl_trigger();
  pdl_wait();
  constexpr bool FAST = sizeof(T) == 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch0 = blockIdx.y * 256 + (warp & 1) * 128 + lane * CPL;
  const int stride = gridDim.x * (SW_WARPS / 2);
  for (int r = blockIdx.x * (SW_WARPS / 2) + (warp >> 1); r < rows; r += stride * RB) {
    float a[RB][CPL], gt[RB][CPL], d[RB][CPL];
#pragma unroll
    for (int i = 0; i < RB; ++i)
      if (r + i * stride < rows) {
        const long rr = r + i * stride;
        Row8<T>::ld(z + rr * 2 * C + ch0, a[i]);
        Row8<T>::ld(z + rr * 2 * C + C + ch0, gt[i]);
        Row8<T>::ld(dg + rr * C + ch0, d[i]);
      }
#pragma unroll
    for (int i = 0; i < RB; ++i)
      if (r + i * stride < rows) {
        const long rr = r + i * stride;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const float sg = sigm<FAST>(gt[i][k]);
          const float da = d[i][k] * sg;
          gt[i][k] = da * a[i][k] * (1.f - sg);
          a[i][k] = da;
        }
        Row8<T>::st(dz + rr * 2 * C + ch0, a[i]);
        Row8<T>::st(dz + rr * 2 * C + C + ch0, gt[i]);
      }
  }
}

}  // namespace eec

using namespace eec;

// persistent grid of the streaming kernels: 3 CTAs per SM (24 warps), never more CTAs than row batches
static int stream_grid(int rows) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return max(1, min(sms * 3, cdiv(rows, SW_WARPS / 2)));
}

// dynamic-smem launch of a dwconv kernel whose staged input type is TI (fp32 tiles exceed the 48 KB default)
#define EEC_DW_LAUNCH(kernel, TI, ...)                                                                                   \
  do {                                                                                                                   \
    static bool attr_ = false;                                                                                           \
    if (!attr_) {                                                                                                        \
      EEC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes<TI>()));          \
      attr_ = true;                                                                                                      \
    }                                                                                                                    \
    launch_pdl(kernel, dim3(grid), dim3(256), dw_smem_bytes<TI>(), S(stream), __VA_ARGS__);                                                   \
  } while (0)

static int dw_pipe_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EEC_DW_PIPE"); v = e ? atoi(e) : 1; }   // 0: one tile per block; 1 (default): persistent, double-buffered tiles
  return v;
}
static bool dw_pipe_enabled() { return dw_pipe_mode() != 0; }
template <typename TO, int MODE>
static int dw_pipe_launch(const __nv_bfloat16* g, const float* w, const float* bias, const float* bn_w, const float* bn_b,
                          const float* run_mean, const float* run_var, TO* out, double* sums, int B, int T, cudaStream_t st) {
  constexpr int SMEM = 2 * (TT + KW - 1) * 256 * 2;
  static bool attr = false;
  if (!attr) {
    EEC_CUDA(cudaFuncSetAttribute(dwconv_pipe_kernel<TO, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  const int tiles = cdiv(T, TT) * B;
  if (tiles == 0) return 0;
  launch_pdl(dwconv_pipe_kernel<TO, MODE>, dim3(min(tiles, sms * 3)), dim3(256), SMEM, st, g, w, bias, bn_w, bn_b, run_mean, run_var, out, sums, B, T,
             active_items(st));
  EEC_LAUNCH_CHECK();
  return 0;
}

#define DW_ARGS_OK()                                                                 \
  EEC_CHECK_ARG(K == KW, "dwconv: depthwise_kernel_size must be 31 (got %d)", K);    \
  EEC_CHECK_ARG(C % 256 == 0, "dwconv: channels must be a multiple of 256 (got %d)", C); \
  if (B == 0 || T == 0) return 0;                                                    \
  dim3 grid(cdiv(T, TT), B, C / 256)

extern "C" int eec_dwconv_bn_silu_eval(const void* g, int dtype, const float* w, const float* bias, const float* bn_w,
                                       const float* bn_b, const float* run_mean, const float* run_var, void* out, int B,
                                       int T, int C, int K, eec_stream_t stream) {
  DW_ARGS_OK();
  if (dtype == EEC_BF16 && C == 256 && dw_pipe_enabled()) {
    if (int r = dw_pipe_launch<__nv_bfloat16, DW_EVAL>((const __nv_bfloat16*)g, w, bias, bn_w, bn_b, run_mean, run_var, (__nv_bfloat16*)out, nullptr, B, T, S(stream))) return r;
    return 0;
  }
  if (dtype == EEC_F32)
    EEC_DW_LAUNCH((dwconv_kernel<float, float, DW_EVAL>), float, (const float*)g, w, bias, bn_w, bn_b, run_mean, run_var, (float*)out, nullptr, T, C, active_items(S(stream)));
  else
    EEC_DW_LAUNCH((dwconv_kernel<__nv_bfloat16, __nv_bfloat16, DW_EVAL>), __nv_bfloat16, (const __nv_bfloat16*)g, w, bias, bn_w, bn_b, run_mean, run_var, (__nv_bfloat16*)out, nullptr, T, C, active_items(S(stream)));
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_dwconv_stats(const void* g, int dtype, const float* w, const float* bias, float* c, double* sums,
                                int B, int T, int C, int K, eec_stream_t stream) {
  DW_ARGS_OK();
  if (dtype == EEC_BF16 && C == 256 && dw_pipe_enabled()) {
    if (int r = dw_pipe_launch<float, DW_STATS>((const __nv_bfloat16*)g, w, bias, nullptr, nullptr, nullptr, nullptr, c, sums, B, T, S(stream))) return r;
    return 0;
  }
  if (dtype == EEC_F32)
    EEC_DW_LAUNCH((dwconv_kernel<float, float, DW_STATS>), float, (const float*)g, w, bias, nullptr, nullptr, nullptr, nullptr, c, sums, T, C, ActiveItems{nullptr, 0, 0});
  else
    EEC_DW_LAUNCH((dwconv_kernel<__nv_bfloat16, float, DW_STATS>), __nv_bfloat16, (const __nv_bfloat16*)g, w, bias, nullptr, nullptr, nullptr, nullptr, c, sums, T, C, ActiveItems{nullptr, 0, 0});
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_bn_silu_train(const float* c, const double* sums, const float* bn_w, const float* bn_b,
                                 float* run_mean, float* run_var, int64_t* num_batches_tracked, float momentum,
                                 float* save_mean, float* save_rstd, void* out, int dtype, int rows, int C,
                                 int64_t stat_rows, eec_stream_t stream) {
  EEC_CHECK_ARG(C % 256 == 0, "bn_silu_train: C %% 256");
  if (rows == 0) return 0;
  const long sr = stat_rows > 0 ? (long)stat_rows : (long)rows;
  dim3 grid(stream_grid(rows), C / 256);
  if (dtype == EEC_F32)
    launch_pdl(bn_silu_train_kernel<float>, dim3(grid), dim3(SW_WARPS * 32), 0, S(stream), c, sums, bn_w, bn_b, run_mean, run_var, num_batches_tracked, momentum, save_mean, save_rstd, (float*)out, rows, C, sr);
  else
    launch_pdl(bn_silu_train_kernel<__nv_bfloat16>, dim3(grid), dim3(SW_WARPS * 32), 0, S(stream), c, sums, bn_w, bn_b, run_mean, run_var, num_batches_tracked, momentum, save_mean, save_rstd, (__nv_bfloat16*)out, rows, C, sr);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_bn_silu_bwd_stats(const void* ds, int dtype, const float* c, const float* save_mean,
                                     const float* save_rstd, const float* bn_w, const float* bn_b, double* sums2,
                                     int rows, int C, eec_stream_t stream) {
  EEC_CHECK_ARG(C % 256 == 0, "bn_silu_bwd_stats: C %% 256");
  if (rows == 0) return 0;
  dim3 grid(stream_grid(rows), C / 256);
  if (dtype == EEC_F32)
    launch_pdl(bn_silu_bwd_kernel<float, false>, dim3(grid), dim3(SW_WARPS * 32), 0, S(stream), (const float*)ds, c, save_mean, save_rstd, bn_w, bn_b, sums2, nullptr, nullptr, nullptr, rows, C, (long)rows, nullptr);
  else
    launch_pdl(bn_silu_bwd_kernel<__nv_bfloat16, false>, dim3(grid), dim3(SW_WARPS * 32), 0, S(stream), (const __nv_bfloat16*)ds, c, save_mean, save_rstd, bn_w, bn_b, sums2, nullptr, nullptr, nullptr, rows, C, (long)rows, nullptr);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_bn_silu_bwd_apply(const void* ds, int dtype, const float* c, const float* save_mean,
                                     const float* save_rstd, const float* bn_w, const float* bn_b, const double* sums2,
                                     float* dc, float* dgamma, float* dbeta, int rows, int C, int64_t stat_rows,
                                     const double* sums2_local, eec_stream_t stream) {
  EEC_CHECK_ARG(C % 256 == 0, "bn_silu_bwd_apply: C %% 256");
  if (rows == 0) return 0;
  const long sr = stat_rows > 0 ? (long)stat_rows : (long)rows;
  dim3 grid(stream_grid(rows), C / 256);
  if (dtype == EEC_F32)
    launch_pdl(bn_silu_bwd_kernel<float, true>, dim3(grid), dim3(SW_WARPS * 32), 0, S(stream), (const float*)ds, c, save_mean, save_rstd, bn_w, bn_b, const_cast<double*>(sums2), dc, dgamma, dbeta, rows, C, sr, sums2_local);
  else
    launch_pdl(bn_silu_bwd_kernel<__nv_bfloat16, true>, dim3(grid), dim3(SW_WARPS * 32), 0, S(stream), (const __nv_bfloat16*)ds, c, save_mean, save_rstd, bn_w, bn_b, const_cast<double*>(sums2), dc, dgamma, dbeta, rows, C, sr, sums2_local);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t eec_dwconv_bwd_workspace_bytes(int B, int T, int C) {
  return (int64_t)cdiv(T, TT) * B * (C / 256) * 32 * 256 * (int64_t)sizeof(float);
}

extern "C" int eec_dwconv_bwd(const float* dc, const void* g, int dtype, const float* w, void* dg, float* dw,
                              float* dbias, int B, int T, int C, int K, void* workspace, eec_stream_t stream) {
  DW_ARGS_OK();
  EEC_CHECK_ARG(workspace != nullptr, "dwconv_bwd: workspace is NULL (eec_dwconv_bwd_workspace_bytes)");
  float* ws = reinterpret_cast<float*>(workspace);
  if (dtype == EEC_F32) {
    EEC_DW_LAUNCH((dwconv_kernel<float, float, DW_BWD_DATA>), float, dc, w, nullptr, nullptr, nullptr, nullptr, nullptr, (float*)dg, nullptr, T, C, ActiveItems{nullptr, 0, 0});
    EEC_LAUNCH_CHECK();
    EEC_DW_LAUNCH((dwconv_wgrad_kernel<float>), float, dc, (const float*)g, ws, T, C);
  } else {
    EEC_DW_LAUNCH((dwconv_kernel<float, __nv_bfloat16, DW_BWD_DATA>), float, dc, w, nullptr, nullptr, nullptr, nullptr, nullptr, (__nv_bfloat16*)dg, nullptr, T, C, ActiveItems{nullptr, 0, 0});
    EEC_LAUNCH_CHECK();
    EEC_DW_LAUNCH((dwconv_wgrad_kernel<__nv_bfloat16>), __nv_bfloat16, dc, (const __nv_bfloat16*)g, ws, T, C);
  }
  EEC_LAUNCH_CHECK();
  launch_pdl(dwconv_wgrad_reduce_kernel, dim3(32, C / 256, RCH), dim3(256), 0, S(stream), (const float*)ws, dw, dbias, (int)(grid.x * grid.y));
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_glu_bwd(const void* z, const void* dg, void* dz, int dtype, int rows, int C, eec_stream_t stream) {
  EEC_CHECK_ARG(C % 256 == 0, "glu_bwd: C %% 256");
  if (rows == 0) return 0;
  dim3 grid(stream_grid(rows), C / 256);
  if (dtype == EEC_F32) launch_pdl(glu_bwd_kernel<float>, dim3(grid), dim3(SW_WARPS * 32), 0, S(stream), (const float*)z, (const float*)dg, (float*)dz, rows, C);
  else launch_pdl(glu_bwd_kernel<__nv_bfloat16>, dim3(grid), dim3(SW_WARPS * 32), 0, S(stream), (const __nv_bfloat16*)z, (const __nv_bfloat16*)dg, (__nv_bfloat16*)dz, rows, C);
  EEC_LAUNCH_CHECK();
  return 0;
}
