// rowwise.cu -- bandwidth-bound row kernels: LayerNorm fwd/bwd, log-softmax (+argmax/entropy),
// casts, column sums, small glue.  One warp per 256-wide row, 8 contiguous channels per lane
// (128-bit loads), statistics in fp32 with a two-pass variance held in registers.
#include <stdlib.h>
#include "common.cuh"

namespace eec {

constexpr float LN_EPS = 1e-5f;

// ------------------------------------------------------------------ LayerNorm forward
template <typename TOut>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, TOut* __restrict__ out,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                                            const ActiveItems act_items) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= active_rows(act_items, rows)) return;
  const int c0 = lane * 8;
  float v[8], g[8], b[8];
  ld8<float>(x + (long)warp * 256 + c0, v);
  ld8<float>(gamma + c0, g);
  ld8<float>(beta + c0, b);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mu = warp_sum(s) * (1.0f / 256.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float d = v[i] - mu; q += d * d; }
  const float rs = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + LN_EPS);
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = (v[i] - mu) * rs * g[i] + b[i];
  st8<TOut>(out + (long)warp * 256 + c0, o);
  if (lane == 0) {
    if (mean) mean[warp] = mu;
    if (rstd) rstd[warp] = rs;
  }
}

// ------------------------------------------------------------------ LayerNorm backward
template <typename TC, bool DROP, typename TDY = float>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const TDY* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            int dx_accumulate, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, TC* __restrict__ dx_bf16,
                                                            float* __restrict__ dx_colsum, float colsum_scale, int rows,
                                                            const DropArgs drop) {
  pdl_trigger();
  pdl_wait();
  DropKey dkey{};
  if (DROP) dkey = drop_key(drop);
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c0 = lane * 8;
  float g[8], dg[8], db[8], cs[8];
  ld8<float>(gamma + c0, g);
#pragma unroll
  for (int i = 0; i < 8; ++i) { dg[i] = 0.f; db[i] = 0.f; cs[i] = 0.f; }
  // Two rows per warp and iteration, every load of both rows issued before the first use: the kernel is a pure stream (18 B per element)
  // and one row per warp left too few bytes in flight (3.5 TB/s cold).
  constexpr int RPI = 2;
  const int rstride = gridDim.x * 8;
  for (int r0 = blockIdx.x * 8 + wib; r0 < rows; r0 += RPI * rstride) {
    float d[RPI][8], v[RPI][8], old[RPI][8], mu[RPI], rs[RPI];
#pragma unroll
    for (int k = 0; k < RPI; ++k) {
      const int r = r0 + k * rstride;
      if (r < rows) {
        ld8<TDY>(dy + (long)r * 256 + c0, d[k]);
        ld8<float>(x + (long)r * 256 + c0, v[k]);
        if (dx_accumulate) ld8<float>(dx + (long)r * 256 + c0, old[k]);
        mu[k] = mean[r]; rs[k] = rstd[r];
      }
    }
#pragma unroll
    for (int k = 0; k < RPI; ++k) {
      const int r = r0 + k * rstride;
      if (r >= rows) break;
      float s1 = 0.f, s2 = 0.f, xh[8], dg_[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[i] = (v[k][i] - mu[k]) * rs[k];
        dg_[i] = d[k][i] * g[i];
        s1 += dg_[i];
        s2 += dg_[i] * xh[i];
        dg[i] += d[k][i] * xh[i];
        db[i] += d[k][i];
      }
      s1 = warp_sum(s1) * (1.0f / 256.0f);
      s2 = warp_sum(s2) * (1.0f / 256.0f);
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rs[k] * (dg_[i] - s1 - xh[i] * s2);
      if (dx_accumulate) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += old[k][i];
      }
      st8<float>(dx + (long)r * 256 + c0, o);
      if (DROP) {   // the copy / column sums feed the backward of a projection whose output was dropped out (dx itself is the residual gradient)
        float f[8];
        drop_factors8(dkey, drop, (uint64_t)r * 32 + lane, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= f[i];
      }
      if (dx_bf16) st8<TC>(dx_bf16 + (long)r * 256 + c0, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) cs[i] += o[i];
    }
  }
  if (dx_colsum) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[wib][c0 + i] = cs[i];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(dx_colsum + threadIdx.x, s * colsum_scale);
    __syncthreads();
  }
  if (dgamma) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[wib][c0 + i] = dg[i];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(dgamma + threadIdx.x, s);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[wib][c0 + i] = db[i];
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(dbeta + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ log-softmax (+argmax, entropy), V = 256
__global__ void __launch_bounds__(256) logsoftmax_kernel(const float* __restrict__ logits, float* __restrict__ out,
                                                         int32_t* __restrict__ argmax, float* __restrict__ entropy,
                                                         int rows) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int c0 = lane * 8;
  float v[8];
  ld8<float>(logits + (long)warp * 256 + c0, v);
  float m = v[0];
  int mi = c0;
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (v[i] > m) { m = v[i]; mi = c0 + i; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, m, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += expf(v[i] - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  float o[8], h = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i] = v[i] - lse;
    h -= expf(o[i]) * o[i];
  }
  st8<float>(out + (long)warp * 256 + c0, o);
  if (entropy) {
    h = warp_sum(h);
    if (lane == 0) entropy[warp] = h;
  }
  if (argmax && lane == 0) argmax[warp] = mi;
}

// generic log-softmax backward: dlogits = g - exp(lp) * sum(g)   (V = 256)
__global__ void __launch_bounds__(256) logsoftmax_bwd_kernel(const float* __restrict__ g, const float* __restrict__ lp,
                                                             float* __restrict__ dl, int rows) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int c0 = lane * 8;
  float gv[8], l[8];
  ld8<float>(g + (long)warp * 256 + c0, gv);
  ld8<float>(lp + (long)warp * 256 + c0, l);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += gv[i];
  s = warp_sum(s);
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = gv[i] - expf(l[i]) * s;
  st8<float>(dl + (long)warp * 256 + c0, o);
}

// ------------------------------------------------------------------ casts
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long n) {
  pdl_trigger();
  pdl_wait();
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float v[8];
    ld8<TI>(in + i, v);
    st8<TO>(out + i, v);
  } else {
    for (; i < n; ++i) st_from_float<TO>(out + i, ld_as_float<TI>(in + i));
  }
}

// ------------------------------------------------------------------ column sums: out[c] += sum_r in[r,c]
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ in, int ld, float* __restrict__ out, int rows,
                                                     int cols, int rows_per_block, float scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < cols)
    for (int r = r0 + ty; r < r1; r += 8) s += ld_as_float<T>(in + (long)r * ld + c);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    if (c < cols) atomicAdd(out + c, t * scale);
  }
}

// cols % 256 == 0: block = 32 column-octets x 8 row lanes, 128-bit loads, 4 rows in flight per thread
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ in, int ld, float* __restrict__ out, int rows,
                                                         int rows_per_block, float scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][256];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int rb = r0 + rl; rb < r1; rb += 32) {
    float v[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rb + 8 * i;
      if (r < r1) ld8<T>(in + (long)r * ld + c0, v[i]);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[i][k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cg * 8 + k] = acc[k];
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
  atomicAdd(out + blockIdx.x * 256 + threadIdx.x, t * scale);
}

// ------------------------------------------------------------------ dropout (standalone: positional-encoding site, tests)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) dropout_kernel(const TI* __restrict__ in, TO* __restrict__ out, long n, const DropArgs drop) {
  pdl_trigger();
  pdl_wait();
  const DropKey key = drop_key(drop);
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g * 8 < n; g += (long)gridDim.x * blockDim.x) {
    const long i = g * 8;
    float f[8];
    drop_factors8(key, drop, (uint64_t)g, f);
    if (i + 8 <= n) {
      float v[8];
      ld8<TI>(in + i, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= f[k];
      st8<TO>(out + i, v);
    } else {
      for (int k = 0; i + k < n; ++k) st_from_float<TO>(out + i + k, ld_as_float<TI>(in + i + k) * f[k]);
    }
  }
}
// Keep-mask words for the tensor-core kernels: the GEMM / attention epilogues are instruction-issue bound, so they do not run Philox
// themselves (measured: +75 % on the FFN epilogues); this kernel evaluates the SAME counter-based draws once per forward into
// 1 bit per element, laid out so that the consumer's thread (one row, W consecutive columns) reads one coalesced word:
//   logical tensor [R, C], element index r*Cs + c;  word (c/W, r) -> bits[(c/W)*R + r];  bit j <-> element (c/W)*W + j is kept.
template <int W, typename TW>
__global__ void __launch_bounds__(256) dropout_bits_kernel(TW* __restrict__ bits, long R, int C, long Cs, const DropArgs drop) {
  pdl_trigger();
  pdl_wait();
  const DropKey key = drop_key(drop);
  // the kernel is bound by the integer ALU pipe, so everything that is not Philox itself is kept off it: the ten round keys are
  // formed once per thread (not per call), a draw is compared in place (hi half: w >= thr << 16; lo half: w << 16 >= thr << 16) and
  // a kept element ORs a constant into the word under a predicate (two ALU instructions per element)
  uint32_t rk0[10], rk1[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) { rk0[i] = key.k0 + (uint32_t)i * 0x9E3779B9u; rk1[i] = key.k1 + (uint32_t)i * 0xBB67AE85u; }
  const uint32_t thr_hi = drop.thr << 16;
  const long cw = blockIdx.y;                                  // word column: grid.y = ceil(C / W) (no divisions in the loop)
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (long)gridDim.x * blockDim.x) {
    const uint64_t g0 = ((uint64_t)r * Cs + (uint64_t)cw * W) >> 3;
    uint32_t word = 0;
#pragma unroll
    for (int q = 0; q < W / 8; ++q) {
      const uint64_t g = g0 + q;
      uint32_t c0 = (uint32_t)g, c1 = (uint32_t)(g >> 32), c2 = key.c2, c3 = key.c3;
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk0[i], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk1[i];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
      }
      const uint32_t w4[4] = {c0, c1, c2, c3};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if ((w4[i] << 16) >= thr_hi) word |= 1u << (q * 8 + 2 * i);        // low 16-bit draw  -> element 2i
        if (w4[i] >= thr_hi) word |= 1u << (q * 8 + 2 * i + 1);            // high 16-bit draw -> element 2i + 1
      }
    }
    bits[cw * R + r] = (TW)word;
  }
}

__global__ void dropout_advance_kernel(uint64_t* state) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) state[1] += 1;
}

// dst[i] = src[clamp(idx[i])]  (compacted copy of a per-utterance int64 vector: raw lengths of the surviving utterances)
__global__ void gather_i64_kernel(const int64_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t* __restrict__ dst, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[min(max(idx[i], 0), n - 1)];
}

// out_bf16[r, c] = bf16(in[r, c]) and colsum[c] += scale * sum_r in[r, c] in ONE pass over a [rows, cols] fp32 tensor (cols % 256 == 0):
// the exit heads' backward needs both the bf16 operand copy of d(logits) and the bias gradient (its column sums)
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float* __restrict__ in, int ld, __nv_bfloat16* __restrict__ out, int ldo,
                                                        float* __restrict__ colsum, int rows, int rows_per_block, float scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][256];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int r = r0 + rl; r < r1; r += 8) {
    float v[8];
    ld8<float>(in + (long)r * ld + c0, v);
    st8<__nv_bfloat16>(out + (long)r * ldo + c0, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cg * 8 + k] = acc[k];
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
  atomicAdd(colsum + blockIdx.x * 256 + threadIdx.x, t * scale);
}

__global__ void axpy_kernel(const float* __restrict__ x, float a, float* __restrict__ y, long n) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(a, x[i], y[i]);
}

__global__ void scale_dev_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y, long n) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const float a = *s;
  if (i < n) y[i] = a * x[i];
}

// y[r, :] = s[r] * x[r, :] for `rows` contiguous slabs of `slab` floats (slab % 4 == 0): all exits' CTC gradient slabs scaled by
// their upstream scalars in ONE launch, 128-bit accesses, grid-stride
__global__ void __launch_bounds__(256) scale_rows_dev_kernel(const float4* x, const float* __restrict__ s, float4* y,   // (x == y allowed)
                                                             long slab4, long total4) {
  pdl_trigger();
  pdl_wait();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const float a = s[i / slab4];
    if (a == 1.0f && x == y) continue;   // in place with a unit scale (the usual upstream of loss.sum()): nothing to move -- 294 MB of traffic per step saved
    float4 v = x[i];
    v.x *= a; v.y *= a; v.z *= a; v.w *= a;
    y[i] = v;
  }
}

__global__ void encoder_lengths_kernel(const int64_t* __restrict__ lengths, int32_t* __restrict__ key_len, int B, int T,
                                       int div, int add) {
  pdl_trigger();
  pdl_wait();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // torch: clamp((lengths+add) / div, max=T).to(int): float32 true division, truncation
  float q = (float)(lengths[b] + add) / (float)div;
  q = fminf(q, (float)T);
  key_len[b] = (int32_t)q;
}

// ------------------------------------------------------------------ Splitformer glue (early_exit.py:318-356)
__global__ void stride2_gather_kernel(const float4* __restrict__ x, float4* __restrict__ y, int B, int T, int T2, int D4) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)B * T2 * D4;
  if (i >= total) return;
  int c = (int)(i % D4);
  long r = i / D4;
  int t2 = (int)(r % T2), b = (int)(r / T2);
  int t = 2 * t2;
  y[i] = (t < T) ? x[((long)b * T + t) * D4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void repeat2_add_kernel(const float4* __restrict__ up, float4* __restrict__ y, int B, int T, int T2, int D4) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)B * T * D4;
  if (i >= total) return;
  int c = (int)(i % D4);
  long r = i / D4;
  int t = (int)(r % T), b = (int)(r / T);
  float4 u = up[((long)b * T2 + (t >> 1)) * D4 + c];
  float4 v = y[i];
  y[i] = make_float4(v.x + u.x, v.y + u.y, v.z + u.z, v.w + u.w);
}
__global__ void repeat2_bwd_kernel(const float4* __restrict__ dy, float4* __restrict__ dh, int B, int T, int T2, int D4) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)B * T2 * D4;
  if (i >= total) return;
  int c = (int)(i % D4);
  long r = i / D4;
  int t2 = (int)(r % T2), b = (int)(r / T2);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < 2; ++j) {
    int t = 2 * t2 + j;
    if (t < T) {
      float4 v = dy[((long)b * T + t) * D4 + c];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  dh[i] = a;
}
__global__ void stride2_scatter_add_kernel(const float4* __restrict__ dh, float4* __restrict__ dx, int B, int T, int T2, int D4) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)B * T2 * D4;
  if (i >= total) return;
  int c = (int)(i % D4);
  long r = i / D4;
  int t2 = (int)(r % T2), b = (int)(r / T2);
  int t = 2 * t2;
  if (t >= T) return;
  long o = ((long)b * T + t) * D4 + c;
  float4 v = dx[o], u = dh[i];
  dx[o] = make_float4(v.x + u.x, v.y + u.y, v.z + u.z, v.w + u.w);
}

}  // namespace eec

using namespace eec;

extern "C" int eec_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out, int out_dtype,
                                 float* mean, float* rstd, int rows, int d, eec_stream_t stream) {
  EEC_CHECK_ARG(d == 256, "layernorm_fwd: d must be 256 (got %d)", d);
  if (rows == 0) return 0;
  int blocks = cdiv(rows, 8);
  if (out_dtype == EEC_F32)
    launch_pdl(layernorm_fwd_kernel<float>, dim3(blocks), dim3(256), 0, S(stream), x, gamma, beta, (float*)out, mean, rstd, rows, active_items(S(stream)));
  else
    launch_pdl(layernorm_fwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, S(stream), x, gamma, beta, (__nv_bfloat16*)out, mean, rstd, rows, active_items(S(stream)));
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_layernorm_bwd_dy(const void* dy_, int dy_dtype, const float* x, const float* mean, const float* rstd,
                                    const float* gamma, float* dx, int dx_accumulate, float* dgamma, float* dbeta,
                                    void* dx_copy, int dx_copy_dtype, float* dx_colsum, float colsum_scale,
                                    const uint64_t* drop_state, float drop_p, uint32_t drop_site, int rows, int d,
                                    eec_stream_t stream) {
  EEC_CHECK_ARG(d == 256, "layernorm_bwd: d must be 256 (got %d)", d);
  if (dy_dtype == EEC_BF16) {
    // upstream gradient in bf16 (the data-gradient GEMM that produced it rounds its fp32 accumulator once): 12 MB less to write and to read
    EEC_CHECK_ARG(!(dx_copy && dx_copy_dtype == EEC_F32), "layernorm_bwd: a bf16 upstream gradient implies the bf16 path (dx_copy bf16)");
    if (rows == 0) return 0;
    const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(dy_);
    static int lnb_blocks = 0;
    if (!lnb_blocks) { const char* e = getenv("EEC_LNB_BLOCKS"); lnb_blocks = e ? atoi(e) : 148 * 2; }
    const int blocks = min(cdiv(rows, 8), lnb_blocks);   // 128 registers: two resident blocks per SM
    const DropArgs drop = make_drop(drop_state, drop_p, drop_site);
    EEC_CHECK_ARG(!drop.state || dx_copy || dx_colsum, "layernorm_bwd: dropout only affects dx_copy / dx_colsum, and neither was requested");
    if (drop.state) launch_pdl(layernorm_bwd_kernel<__nv_bfloat16, true, __nv_bfloat16>, dim3(blocks), dim3(256), 0, S(stream), dy, x, mean, rstd, gamma, dx, dx_accumulate, dgamma, dbeta, (__nv_bfloat16*)dx_copy, dx_colsum, colsum_scale, rows, drop);
    else launch_pdl(layernorm_bwd_kernel<__nv_bfloat16, false, __nv_bfloat16>, dim3(blocks), dim3(256), 0, S(stream), dy, x, mean, rstd, gamma, dx, dx_accumulate, dgamma, dbeta, (__nv_bfloat16*)dx_copy, dx_colsum, colsum_scale, rows, drop);
    EEC_LAUNCH_CHECK();
    return 0;
  }
  const float* dy = reinterpret_cast<const float*>(dy_);
  if (rows == 0) return 0;
  int blocks = min(cdiv(rows, 8), 148 * 4);
  const DropArgs drop = make_drop(drop_state, drop_p, drop_site);
  EEC_CHECK_ARG(!drop.state || dx_copy || dx_colsum, "layernorm_bwd: dropout only affects dx_copy / dx_colsum, and neither was requested");
#define EEC_LNB(TC, DROP) launch_pdl(layernorm_bwd_kernel<TC, DROP>, dim3(blocks), dim3(256), 0, S(stream), dy, x, mean, rstd, gamma, dx, dx_accumulate, dgamma, dbeta, (TC*)dx_copy, dx_colsum, colsum_scale, rows, drop)
  if (dx_copy && dx_copy_dtype == EEC_F32) { if (drop.state) EEC_LNB(float, true); else EEC_LNB(float, false); }
  else { if (drop.state) EEC_LNB(__nv_bfloat16, true); else EEC_LNB(__nv_bfloat16, false); }
#undef EEC_LNB
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                                 const float* gamma, float* dx, int dx_accumulate, float* dgamma, float* dbeta,
                                 void* dx_copy, int dx_copy_dtype, float* dx_colsum, float colsum_scale,
                                 const uint64_t* drop_state, float drop_p, uint32_t drop_site, int rows, int d,
                                 eec_stream_t stream) {
  return eec_layernorm_bwd_dy(dy, EEC_F32, x, mean, rstd, gamma, dx, dx_accumulate, dgamma, dbeta, dx_copy, dx_copy_dtype, dx_colsum,
                              colsum_scale, drop_state, drop_p, drop_site, rows, d, stream);
}

extern "C" int eec_dropout(const void* x, int in_dtype, void* y, int out_dtype, int64_t n, const uint64_t* state, float p,
                           uint32_t site, eec_stream_t stream) {
  if (n == 0) return 0;
  EEC_CHECK_ARG(state != nullptr, "dropout: state is NULL");
  EEC_CHECK_ARG(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1) (got %f)", p);
  DropArgs drop = make_drop(state, p, site);
  if (!drop.state) { drop.state = state; drop.site = site; drop.thr = 0; drop.scale = 1.f; }   // p == 0: identity through the same kernel
  const int blocks = (int)min((long)148 * 16, cdiv64(cdiv64(n, 8), 256));
#define EEC_DROPK(TI, TO) launch_pdl(dropout_kernel<TI, TO>, dim3(blocks), dim3(256), 0, S(stream), (const TI*)x, (TO*)y, (long)n, drop)
  if (in_dtype == EEC_F32 && out_dtype == EEC_F32) EEC_DROPK(float, float);
  else if (in_dtype == EEC_F32 && out_dtype == EEC_BF16) EEC_DROPK(float, __nv_bfloat16);
  else if (in_dtype == EEC_BF16 && out_dtype == EEC_BF16) EEC_DROPK(__nv_bfloat16, __nv_bfloat16);
  else EEC_DROPK(__nv_bfloat16, float);
#undef EEC_DROPK
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_dropout_bits(const uint64_t* state, float p, uint32_t site, int64_t R, int C, int64_t Cs, int W, void* bits,
                                eec_stream_t stream) {
  EEC_CHECK_ARG(state != nullptr && bits != nullptr, "dropout_bits: NULL state / bits");
  EEC_CHECK_ARG(W == 16 || W == 32, "dropout_bits: word width must be 16 or 32 (got %d)", W);
  EEC_CHECK_ARG(Cs % 8 == 0 && Cs >= C, "dropout_bits: row stride (%lld) must be a multiple of 8 and >= C", (long long)Cs);
  if (R == 0 || C == 0) return 0;
  DropArgs drop = make_drop(state, p, site);
  if (!drop.state) { drop.state = state; drop.site = site; drop.thr = 0; drop.scale = 1.f; }   // p == 0: all kept
  const int ncw = (C + W - 1) / W;
  const dim3 grid((unsigned)min((long)cdiv64(R, 256), (long)max(1, 148 * 32 / ncw)), (unsigned)ncw);
  if (W == 16) launch_pdl(dropout_bits_kernel<16, uint16_t>, grid, dim3(256), 0, S(stream), (uint16_t*)bits, (long)R, C, (long)Cs, drop);
  else launch_pdl(dropout_bits_kernel<32, uint32_t>, grid, dim3(256), 0, S(stream), (uint32_t*)bits, (long)R, C, (long)Cs, drop);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_dropout_advance(uint64_t* state, eec_stream_t stream) {
  EEC_CHECK_ARG(state != nullptr, "dropout_advance: state is NULL");
  launch_pdl(dropout_advance_kernel, dim3(1), dim3(32), 0, S(stream), state);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_logsoftmax_fwd(const float* logits, float* out, int32_t* argmax, float* entropy, int rows, int V,
                                  eec_stream_t stream) {
  EEC_CHECK_ARG(V == 256, "logsoftmax: V must be 256 (got %d)", V);
  if (rows == 0) return 0;
  launch_pdl(logsoftmax_kernel, dim3(cdiv(rows, 8)), dim3(256), 0, S(stream), logits, out, argmax, entropy, rows);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_logsoftmax_bwd(const float* g, const float* lp, float* dlogits, int rows, int V, eec_stream_t stream) {
  EEC_CHECK_ARG(V == 256, "logsoftmax_bwd: V must be 256 (got %d)", V);
  if (rows == 0) return 0;
  launch_pdl(logsoftmax_bwd_kernel, dim3(cdiv(rows, 8)), dim3(256), 0, S(stream), g, lp, dlogits, rows);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_cast(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, eec_stream_t stream) {
  if (n == 0) return 0;
  int blocks = (int)cdiv64(cdiv64(n, 8), 256);
  if (in_dtype == EEC_F32 && out_dtype == EEC_BF16)
    launch_pdl(cast_kernel<float, __nv_bfloat16>, dim3(blocks), dim3(256), 0, S(stream), (const float*)in, (__nv_bfloat16*)out, n);
  else if (in_dtype == EEC_BF16 && out_dtype == EEC_F32)
    launch_pdl(cast_kernel<__nv_bfloat16, float>, dim3(blocks), dim3(256), 0, S(stream), (const __nv_bfloat16*)in, (float*)out, n);
  else if (in_dtype == EEC_F32 && out_dtype == EEC_F32)
    launch_pdl(cast_kernel<float, float>, dim3(blocks), dim3(256), 0, S(stream), (const float*)in, (float*)out, n);
  else
    launch_pdl(cast_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(blocks), dim3(256), 0, S(stream), (const __nv_bfloat16*)in, (__nv_bfloat16*)out, n);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_colsum(const void* in, int dtype, int ld, float* out, float scale, int rows, int cols,
                          eec_stream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  if (cols % 256 == 0 && ld % 8 == 0) {
    const int rpbv = 128;
    dim3 gv(cols / 256, cdiv(rows, rpbv));
    if (dtype == EEC_F32) launch_pdl(colsum_vec_kernel<float>, dim3(gv), dim3(256), 0, S(stream), (const float*)in, ld, out, rows, rpbv, scale);
    else launch_pdl(colsum_vec_kernel<__nv_bfloat16>, dim3(gv), dim3(256), 0, S(stream), (const __nv_bfloat16*)in, ld, out, rows, rpbv, scale);
    EEC_LAUNCH_CHECK();
    return 0;
  }
  int rpb = 256;
  dim3 grid(cdiv(cols, 32), cdiv(rows, rpb));
  if (dtype == EEC_F32) launch_pdl(colsum_kernel<float>, dim3(grid), dim3(256), 0, S(stream), (const float*)in, ld, out, rows, cols, rpb, scale);
  else launch_pdl(colsum_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, S(stream), (const __nv_bfloat16*)in, ld, out, rows, cols, rpb, scale);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_gather_i64(const int64_t* src, const int32_t* idx, int64_t* dst, int n, eec_stream_t stream) {
  if (n == 0) return 0;
  launch_pdl(gather_i64_kernel, dim3(cdiv(n, 128)), dim3(128), 0, S(stream), src, idx, dst, n);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_cast_colsum(const float* in, int ld, void* out_bf16, int ldo, float* colsum, float scale, int rows, int cols,
                               eec_stream_t stream) {
  EEC_CHECK_ARG(cols % 256 == 0 && ld % 8 == 0 && ldo % 8 == 0, "cast_colsum: cols %% 256, ld %% 8, ldo %% 8 required");
  if (rows == 0 || cols == 0) return 0;
  const int rpb = 128;
  launch_pdl(cast_colsum_kernel, dim3(cols / 256, cdiv(rows, rpb)), dim3(256), 0, S(stream), in, ld, (__nv_bfloat16*)out_bf16, ldo, colsum, rows, rpb, scale);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_axpy(const float* x, float a, float* y, int64_t n, eec_stream_t stream) {
  if (n == 0) return 0;
  launch_pdl(axpy_kernel, dim3((int)cdiv64(n, 256)), dim3(256), 0, S(stream), x, a, y, n);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_scale_dev(const float* x, const float* s_dev, float* y, int64_t n, eec_stream_t stream) {
  if (n == 0) return 0;
  launch_pdl(scale_dev_kernel, dim3((int)cdiv64(n, 256)), dim3(256), 0, S(stream), x, s_dev, y, n);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_scale_rows_dev(const float* x, const float* s_dev, float* y, int rows, int64_t slab, eec_stream_t stream) {
  if (rows == 0 || slab == 0) return 0;
  EEC_CHECK_ARG(slab % 4 == 0, "scale_rows_dev: slab (%lld) must be a multiple of 4", (long long)slab);
  const long total4 = (long)rows * slab / 4;
  const int blocks = (int)min((long)148 * 16, cdiv64(total4, 256));
  launch_pdl(scale_rows_dev_kernel, dim3(blocks), dim3(256), 0, S(stream), (const float4*)x, s_dev, (float4*)y, (long)(slab / 4), total4);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_encoder_lengths(const int64_t* lengths, int32_t* key_len, int B, int T, int div, int add,
                                   eec_stream_t stream) {
  if (B == 0) return 0;
  launch_pdl(encoder_lengths_kernel, dim3(cdiv(B, 128)), dim3(128), 0, S(stream), lengths, key_len, B, T, div, add);
  EEC_LAUNCH_CHECK();
  return 0;
}

#define EEC_SPLIT_LAUNCH(kernel, total, ...)                                        \
  do {                                                                              \
    long tot_ = (total);                                                            \
    if (tot_ > 0) {                                                                 \
      launch_pdl(kernel, dim3((int)cdiv64(tot_, 256)), dim3(256), 0, S(stream), __VA_ARGS__);           \
      EEC_LAUNCH_CHECK();                                                           \
    }                                                                               \
  } while (0)

extern "C" int eec_stride2_gather(const float* x, float* y, int B, int T, int D, eec_stream_t stream) {
  EEC_CHECK_ARG(D % 4 == 0, "stride2_gather: D %% 4");
  int T2 = (T + 1) / 2;
  EEC_SPLIT_LAUNCH(stride2_gather_kernel, (long)B * T2 * (D / 4), (const float4*)x, (float4*)y, B, T, T2, D / 4);
  return 0;
}
extern "C" int eec_repeat2_add(const float* up, float* y, int B, int T, int D, eec_stream_t stream) {
  EEC_CHECK_ARG(D % 4 == 0, "repeat2_add: D %% 4");
  int T2 = (T + 1) / 2;
  EEC_SPLIT_LAUNCH(repeat2_add_kernel, (long)B * T * (D / 4), (const float4*)up, (float4*)y, B, T, T2, D / 4);
  return 0;
}
extern "C" int eec_repeat2_bwd(const float* dy, float* dhalf, int B, int T, int D, eec_stream_t stream) {
  EEC_CHECK_ARG(D % 4 == 0, "repeat2_bwd: D %% 4");
  int T2 = (T + 1) / 2;
  EEC_SPLIT_LAUNCH(repeat2_bwd_kernel, (long)B * T2 * (D / 4), (const float4*)dy, (float4*)dhalf, B, T, T2, D / 4);
  return 0;
}
extern "C" int eec_stride2_scatter_add(const float* dhalf, float* dx, int B, int T, int D, eec_stream_t stream) {
  EEC_CHECK_ARG(D % 4 == 0, "stride2_scatter_add: D %% 4");
  int T2 = (T + 1) / 2;
  EEC_SPLIT_LAUNCH(stride2_scatter_add_kernel, (long)B * T2 * (D / 4), (const float4*)dhalf, (float4*)dx, B, T, T2, D / 4);
  return 0;
}
