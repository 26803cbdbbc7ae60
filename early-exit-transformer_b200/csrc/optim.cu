// optim.cu -- the optimiser step of the reference's training loop (train.py:69-70) as three launches over FLAT buffers:
//   torch.nn.utils.clip_grad_norm_(model.parameters(), clip)            (train.py:69)
//   NoamOpt.step(): lr = d_model^-0.5 * min(t^-0.5, t * warmup^-1.5)    (util/noam_opt.py:26-40)
//   torch.optim.AdamW(lr, betas, eps, weight_decay).step()              (train.py:261-262)
// The engine already leaves every gradient in one flat fp32 buffer; eec/optim.py lays the parameters and the two Adam
// moments out the same way, so the whole update is ONE bandwidth-bound pass (16 B read + 12 B written per parameter,
// + 2 B for the bf16 GEMM-operand shadow the next forward consumes: no per-tensor cast launches).  The step counter and
// the gradient norm stay on the device: no host sync, and the three launches can sit inside the training step's CUDA graph.
#include "common.cuh"

namespace eec {
namespace {

// state[0] = step counter (as double), state[1] = sum of squares of the gradient, state[2] = last lr, state[3] = last clip coefficient
__global__ void optim_begin_kernel(double* state) {
  pdl_trigger();
  pdl_wait();
  state[0] += 1.0;
  state[1] = 0.0;
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n4, long n, double* __restrict__ state) {
  pdl_trigger();
  pdl_wait();
  float acc = 0.f;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float t = g[n4 * 4 + threadIdx.x]; acc = fmaf(t, t, acc); }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4); t += __shfl_xor_sync(0xffu, t, 2); t += __shfl_xor_sync(0xffu, t, 1);
    if (threadIdx.x == 0) atomicAdd(state + 1, (double)t);
  }
}

struct OptP {
  float model_size, warmup, beta1, beta2, eps, weight_decay, clip, lr_fixed;
};

__device__ __forceinline__ void adamw1(float& p, float g, float& m, float& v, float cg, float decay, float b1, float b2, float step_size,
                                       float inv_bc2_sqrt, float eps) {
  g *= cg;
  p *= decay;
  m = fmaf(b1, m, (1.f - b1) * g);
  v = fmaf(b2, v, (1.f - b2) * g * g);
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) noam_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, long n4, long n,
                                                         double* __restrict__ state, const OptP o) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[4];
  if (threadIdx.x == 0) {
    const double t = state[0];
    const double lr = (o.lr_fixed >= 0.f) ? (double)o.lr_fixed
                                          : pow((double)o.model_size, -0.5) * fmin(pow(t, -0.5), t * pow((double)o.warmup, -1.5));
    const double norm = sqrt(state[1]);
    double cg = (o.clip > 0.f) ? (double)o.clip / (norm + 1e-6) : 1.0;     // clip_grad_norm_: coefficient clamped to 1
    if (cg > 1.0) cg = 1.0;
    const double bc1 = 1.0 - pow((double)o.beta1, t), bc2 = 1.0 - pow((double)o.beta2, t);
    sh[0] = (float)cg;
    sh[1] = (float)(1.0 - lr * (double)o.weight_decay);
    sh[2] = (float)(lr / bc1);
    sh[3] = (float)(1.0 / sqrt(bc2));
    if (blockIdx.x == 0) { state[2] = lr; state[3] = cg; }
  }
  __syncthreads();
  const float cg = sh[0], decay = sh[1], step_size = sh[2], ibc2 = sh[3];
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw1(pp.x, gg.x, mm.x, vv.x, cg, decay, o.beta1, o.beta2, step_size, ibc2, o.eps);
    adamw1(pp.y, gg.y, mm.y, vv.y, cg, decay, o.beta1, o.beta2, step_size, ibc2, o.eps);
    adamw1(pp.z, gg.z, mm.z, vv.z, cg, decay, o.beta1, o.beta2, step_size, ibc2, o.eps);
    adamw1(pp.w, gg.w, mm.w, vv.w, cg, decay, o.beta1, o.beta2, step_size, ibc2, o.eps);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      uint2 u;
      *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(pp.x, pp.y);
      *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(pp.z, pp.w);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {   // tail (n not a multiple of 4)
    const long i = n4 * 4 + threadIdx.x;
    float pp = p[i], mm = m[i], vv = v[i];
    adamw1(pp, g[i], mm, vv, cg, decay, o.beta1, o.beta2, step_size, ibc2, o.eps);
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (shadow) shadow[i] = __float2bfloat16_rn(pp);
  }
}

}  // namespace
}  // namespace eec

using namespace eec;

extern "C" int eec_noam_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* bf16_shadow,
                                   int64_t n, double* state, float model_size, float warmup, float beta1, float beta2, float eps,
                                   float weight_decay, float clip, float lr_fixed, eec_stream_t stream) {
  EEC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state, "noam_adamw_step: NULL argument");
  EEC_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "noam_adamw_step: flat buffers must be 16-byte aligned");
  EEC_CHECK_ARG(lr_fixed >= 0.f || (model_size > 0.f && warmup > 0.f), "noam_adamw_step: Noam schedule needs model_size > 0 and warmup > 0");
  if (n == 0) return 0;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    EEC_CUDA(cudaGetDevice(&dev));
    EEC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long n4 = n / 4;
  const int grid = (int)max(1L, min((long)sms * 8, (n4 + 255) / 256));
  launch_pdl(optim_begin_kernel, dim3(1), dim3(1), 0, S(stream), state);
  EEC_LAUNCH_CHECK();
  launch_pdl(sumsq_kernel, dim3(grid), dim3(256), 0, S(stream), grads, n4, n, state);
  EEC_LAUNCH_CHECK();
  OptP o{model_size, warmup, beta1, beta2, eps, weight_decay, clip, lr_fixed};
  launch_pdl(noam_adamw_kernel, dim3(grid), dim3(256), 0, S(stream), params, grads, exp_avg, exp_avg_sq, reinterpret_cast<__nv_bfloat16*>(bf16_shadow), n4, n, state, o);
  EEC_LAUNCH_CHECK();
  return 0;
}
