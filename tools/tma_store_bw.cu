// tma_store_bw.cu -- micro-benchmark: per-SM throughput of cp.async.bulk.tensor STORES (smem -> global) as a function of the box shape
// and of the number of issuing warps, for the output pattern of the K = 256 GEMMs (every CTA writes 128 x 256 bf16 tiles of a
// [23936, 2048] matrix).  Decides the staging-box shape of the tcgen05 GEMM epilogues.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tma_store_bw tools/tma_store_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store(const CUtensorMap* m, const void* src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m), "r"(smem_u32(src)), "r"(x), "r"(y) : "memory");
}
// box = bc cols x br rows; a 128 x 256 tile = (256/bc) x (128/br) boxes, dealt round-robin to the issuing warps; `fence`: each warp also
// rewrites its staging buffer and runs fence.proxy.async before every store (what an epilogue does)
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ CUtensorMap tm, int bc, int br, int tiles, int m_tiles, int fence, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int bx = 256 / bc, by = 128 / br, nbox = bx * by, bytes = bc * br * 2;
  uint8_t* buf = smem + warp * bytes;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    const int tile = (blockIdx.x + t * gridDim.x) % (m_tiles * 8);
    const int m0 = (tile / 8) * 128, n0 = (tile % 8) * 256;
    for (int b = warp; b < nbox; b += nw) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      if (fence) {
        for (int i = lane * 16; i < bytes; i += 32 * 16) *reinterpret_cast<uint4*>(buf + i) = make_uint4(t, b, i, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
      }
      if (lane == 0) {
        tma_store(&tm, buf, n0 + (b % bx) * bc, m0 + (b / bx) * br);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
// the same output pattern written straight from registers: warp (q, cg) of 16 owns rows q*32 + lane, columns cg*64 .. +64 of the tile and
// writes them as four 256-bit stores per thread (whole 32-byte sectors, 32 different lines per warp instruction)
__global__ void __launch_bounds__(512, 1) kst(uint16_t* c, int tiles, int m_tiles, long long* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, cg = warp >> 2;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    const int tile = (blockIdx.x + t * gridDim.x) % (m_tiles * 8);
    const int m0 = (tile / 8) * 128, n0 = (tile % 8) * 256;
    uint16_t* p = c + (long)(m0 + q * 32 + lane) * 2048 + n0 + cg * 64;
#pragma unroll
    for (int ss = 0; ss < 4; ++ss)
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + ss * 16), "r"(t), "r"(ss), "r"(lane), "r"(warp), "r"(t), "r"(ss), "r"(lane), "r"(warp) : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  CK(cudaSetDevice(0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeFn enc = (EncodeFn)fn;
  const uint64_t rows = 23936, cols = 2048;
  void* buf; CK(cudaMalloc(&buf, rows * cols * 2));
  long long* d; CK(cudaMalloc(&d, 8 * 148));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      kst<<<148, 512>>>((uint16_t*)buf, 10, 187, d);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    printf("st.global.v8.b32 from registers, 16 warps: %6.0f clk per 128x256 tile on CTA 0 (issue side), kernel %6.1f us = %6.1f GB/s\n", (double)h / 10, ms * 1e3,
           148.0 * 10 * 65536 / (ms * 1e-3) / 1e9);
  }
  struct Cfg { int bc, br, nw; } cfgs[] = {{32, 32, 16}, {32, 32, 8}, {64, 32, 16}, {64, 32, 8}, {64, 32, 4}, {32, 128, 8}, {32, 128, 4}, {64, 128, 4}, {64, 128, 1}, {16, 32, 16}};
  for (auto c : cfgs)
    for (int fence = 0; fence < 2; ++fence) {
      CUtensorMap tm;
      cuuint64_t gd[2] = {cols, rows}, gs[1] = {cols * 2};
      cuuint32_t bx[2] = {(cuuint32_t)c.bc, (cuuint32_t)c.br}, es[2] = {1, 1};
      const CUtensorMapSwizzle sw = c.bc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : c.bc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
      if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
      const int tiles = 10;   // ~1496 tiles over 148 CTAs, like the GEMM
      const size_t sm = (size_t)c.nw * c.bc * c.br * 2;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      k<<<148, c.nw * 32, sm>>>(tm, c.bc, c.br, tiles, 187, fence, d);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      k<<<148, c.nw * 32, sm>>>(tm, c.bc, c.br, tiles, 187, fence, d);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      printf("box %2d cols x %3d rows (%5d B) warps %2d %s: %6.0f clk per 128x256 tile on CTA 0, kernel %6.1f us = %6.1f GB/s\n", c.bc, c.br, c.bc * c.br * 2, c.nw,
             fence ? "sts+fence" : "issue only", (double)h / tiles, ms * 1e3, 148.0 * tiles * 65536 / (ms * 1e-3) / 1e9);
    }
  return 0;
}
