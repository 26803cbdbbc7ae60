// mufu_bw.cu -- micro-benchmark: MUFU (tanh / ex2 / rcp) issue rate of ONE warp vs several warps per SM sub-partition, independent operands.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mufu_bw tools/mufu_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)
template <int OP> __device__ __forceinline__ float f(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int OP, int ILP>
__global__ void __launch_bounds__(1024, 1) k(int iters, long long* out, float* sink) {
  float x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = 0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = f<OP>(x[j]);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
template <int OP, int ILP> int run(const char* name, long long* d, float* s) {
  const int iters = 2000;
  for (int nw : {4, 8, 16, 32}) {
    k<OP, ILP><<<148, nw * 32>>>(iters, d, s);
    CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    printf("%-5s ILP %2d warps/SM %2d: %6.2f clk per warp-instruction per warp, %6.2f lanes/clk/SM\n", name, ILP, nw, (double)h / iters / ILP, (double)nw * 32 * ILP * iters / h);
  }
  return 0;
}
int main() {
  long long* d; float* s; CK(cudaMalloc(&d, 8 * 148)); CK(cudaMalloc(&s, 4));
  run<0, 16>("tanh", d, s); run<0, 4>("tanh", d, s); run<1, 16>("ex2", d, s); run<2, 16>("rcp", d, s); run<3, 16>("ffma", d, s);
  return 0;
}
