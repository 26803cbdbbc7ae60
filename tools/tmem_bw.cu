// tmem_bw.cu -- micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM as a function of the number of reading warps and of
// the load width, MUFU.TANH throughput, and both together (what the activation epilogues of the tcgen05 GEMMs are bound by).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_bw tools/tmem_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld16(uint32_t a, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ldwait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// mode bit 0: TMEM loads; bit 1: 16 tanh per 16 loaded values.  Every warp reads 64 columns (4 x16 loads) of its lane quarter per "tile".
__global__ void __launch_bounds__(1024, 1) k(int mode, int tiles, long long* out, float* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) * 64) % 512;
  float acc = 0.f;
  uint32_t ra[16], rb[16];
  for (int j = 0; j < 16; ++j) { ra[j] = threadIdx.x + j; rb[j] = j; }
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    if (mode & 1) ld16(base, ra);
#pragma unroll
    for (int ss = 0; ss < 4; ++ss) {
      uint32_t(&cur)[16] = (ss & 1) ? rb : ra;
      uint32_t(&nxt)[16] = (ss & 1) ? ra : rb;
      if (mode & 1) { ldwait(cur); if (ss < 3) ld16(base + (ss + 1) * 16, nxt); }
      if (mode & 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += tanh_fast(__uint_as_float(cur[j]) + acc * 1e-30f * (j == 0));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += __uint_as_float(cur[j]);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory");
  (void)nw;
}
int main() {
  long long* d; float* s; CK(cudaMalloc(&d, 8 * 148)); CK(cudaMalloc(&s, 4));
  const int tiles = 2000;
  for (int mode = 1; mode <= 3; ++mode)
    for (int nw : {4, 8, 16, 32}) {
      k<<<148, nw * 32>>>(mode, tiles, d, s);
      CK(cudaDeviceSynchronize());
      long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      const double clk = (double)h / tiles;                         // per "tile" = every warp reads 32 lanes x 64 columns
      const double bytes = (double)nw * 32 * 64 * 4;
      printf("mode %d (%s%s) warps %2d: %7.0f clk per round  -> %6.1f B/clk/SM TMEM read, %5.2f tanh/clk/SM\n", mode, (mode & 1) ? "ld " : "", (mode & 2) ? "tanh" : "",
             nw, clk, (mode & 1) ? bytes / clk : 0.0, (mode & 2) ? nw * 32 * 64 / clk : 0.0);
    }
  return 0;
}
