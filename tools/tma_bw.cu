// tma_bw.cu -- micro-benchmark: per-SM TMA (cp.async.bulk.tensor) load throughput from an L2-resident or
// HBM-resident bf16 matrix as a function of the bytes in flight (ring depth) and of the number of active CTAs.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tma_bw tools/tma_bw.cu -lcuda
// Decides the smem ring depth of the tcgen05 kernels.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

// matrix [rows, 256] bf16; a box is {64 cols, 128 rows} = 16 KB.  Each CTA loads `n_boxes` boxes round-robin over the matrix.
__global__ void __launch_bounds__(64, 1) bw_kernel(const __grid_constant__ CUtensorMap tm, int ring, int n_boxes, int row_tiles, int stride_ctas) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + ring * 16384);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ring; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = row_tiles * 4;
    int box = (int)(((long)blockIdx.x * stride_ctas) % total);
    for (int it = 0, s = 0, ph = 1; it < n_boxes; ++it) {
      mbar_wait(&empty[s], ph);
      mbar_expect(&full[s], 16384);
      // stride_ctas = 0: every CTA reads the same boxes (weights); else CTA b starts b*stride_ctas boxes into the matrix (distinct data)
      tma2d(smem + s * 16384, &tm, &full[s], (box & 3) * 64, (box >> 2) * 128);
      if (++box == total) box = 0;
      if (++s == ring) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0, s = 0, ph = 0; it < n_boxes; ++it) {
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
      if (++s == ring) { s = 0; ph ^= 1; }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  CK(cudaSetDevice(0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeFn enc = (EncodeFn)fn;
  CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 14 * 16384 + 512));
  for (int big = 0; big < 3; ++big) {
    // 0: 2048 x 256 bf16 = 1 MB, every CTA reads the same boxes (L2-resident weights)
    // 1: 128K rows = 64 MB, L2-resident, every CTA reads a distinct region (true L2 -> SM aggregate bandwidth)
    // 2: 8M rows = 4 GB, distinct regions, streams from HBM
    const uint64_t rows = big == 2 ? (8ull << 20) : big == 1 ? (128ull << 10) : 2048;
    void* buf;
    CK(cudaMalloc(&buf, rows * 512));
    CK(cudaMemset(buf, 0, rows * 512));
    CUtensorMap tm;
    cuuint64_t gd[2] = {256, rows}, gs[1] = {512};
    cuuint32_t bx[2] = {64, 128}, es[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const int n_boxes = 4096;   // 64 MB per CTA (wraps around the matrix)
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grids[3] = {148, 74, 16};
    const int rings[7] = {1, 2, 3, 4, 6, 8, 12};
    for (int gi = 0; gi < 3; ++gi)
      for (int ri = 0; ri < 7; ++ri) {
        const int ring = rings[ri], grid = grids[gi];
        const int stride = big ? (int)(rows / 128 * 4 / grid) : 0;
        bw_kernel<<<grid, 64, ring * 16384 + 512>>>(tm, ring, n_boxes, (int)(rows / 128), stride);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        bw_kernel<<<grid, 64, ring * 16384 + 512>>>(tm, ring, n_boxes, (int)(rows / 128), stride);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double per_sm = (double)n_boxes * 16384 / (ms * 1e-3) / 1e9;
        printf("%s grid %3d ring %2d (%3d KB in flight): %7.1f GB/s per SM  %8.1f GB/s total   %.0f ns per box\n", big == 2 ? "HBM    " : big == 1 ? "L2 dist" : "L2 same", grid, ring,
               ring * 16, per_sm, per_sm * grid, ms * 1e6 / n_boxes);
      }
    CK(cudaFree(buf));
  }
  return 0;
}
