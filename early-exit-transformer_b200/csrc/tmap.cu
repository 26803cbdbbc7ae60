// tmap.cu -- host-side cache of TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point) shared by the
// tcgen05 GEMM, attention and fused-FFN kernels.
#include <mutex>
#include <unordered_map>
#include <string>
#include <cstring>

#include "tc_common.cuh"

namespace eec {

// ------------------------------------------------------------------ tensor-map cache (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_tm_mu;
static std::unordered_map<std::string, CUtensorMap> g_tm_cache;

static int load_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return 3;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

static int get_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                    const uint32_t* box, int swizzle, bool f32 = false) {
  struct Key { const void* b; int r; int sw; int f32; uint64_t d[3]; uint64_t s[2]; uint32_t x[3]; } k;
  memset(&k, 0, sizeof(k));
  k.b = base; k.r = rank; k.sw = swizzle; k.f32 = f32 ? 1 : 0;
  for (int i = 0; i < rank; ++i) { k.d[i] = dims[i]; k.x[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.s[i] = strides[i];
  std::string key(reinterpret_cast<const char*>(&k), sizeof(k));
  std::lock_guard<std::mutex> lk(g_tm_mu);
  auto it = g_tm_cache.find(key);
  if (it != g_tm_cache.end()) { *out = it->second; return 0; }
  if (int r = load_encode()) return r;
  EEC_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    EEC_CHECK_ARG(strides[i] % 16 == 0, "tensor map: stride %llu not a multiple of 16 bytes", (unsigned long long)strides[i]);
  cuuint64_t gd[3]; cuuint64_t gs[2]; cuuint32_t bx[3]; cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  alignas(64) CUtensorMap m;
  CUresult r = g_encode(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu stride %llu box %u,%u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides[0] : 0), box[0],
              rank > 1 ? box[1] : 0);
    return 4;
  }
  if (g_tm_cache.size() > 8192) g_tm_cache.clear();
  g_tm_cache.emplace(key, m);
  *out = m;
  return 0;
}

int get_tmap_2d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
                uint32_t box1, int swizzle) {
  uint64_t d[2] = {dim0, dim1}, s[1] = {stride1_bytes};
  uint32_t b[2] = {box0, box1};
  return get_tmap(out, base, 2, d, s, b, swizzle);
}
// [rows x cols] output tensor stored 32 columns x 128 rows at a time (gemm_tc2 epilogue staging tiles)
int get_tmap_store(CUtensorMap* out, const void* base, bool bf16, uint64_t cols, uint64_t rows, uint64_t ld_elems) {
  uint64_t d[2] = {cols, rows}, s[1] = {ld_elems * (bf16 ? 2 : 4)};
  uint32_t b[2] = {32, 128};
  return get_tmap(out, base, 2, d, s, b, bf16 ? 2 /*SWIZZLE_64B*/ : 3 /*SWIZZLE_128B*/, !bf16);
}
int get_tmap_box32(CUtensorMap* out, const void* base, bool bf16, uint64_t cols, uint64_t rows, uint64_t ld_elems) {
  uint64_t d[2] = {cols, rows}, s[1] = {ld_elems * (bf16 ? 2 : 4)};
  uint32_t b[2] = {32, 32};
  return get_tmap(out, base, 2, d, s, b, bf16 ? 2 /*SWIZZLE_64B*/ : 3 /*SWIZZLE_128B*/, !bf16);
}
int get_tmap_box(CUtensorMap* out, const void* base, bool bf16, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_cols,
                 uint32_t box_rows, int swizzle) {
  uint64_t d[2] = {cols, rows}, s[1] = {ld_elems * (bf16 ? 2 : 4)};
  uint32_t b[2] = {box_cols, box_rows};
  return get_tmap(out, base, 2, d, s, b, swizzle, !bf16);
}
int get_tmap_3d(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1_bytes,
                uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  uint64_t d[3] = {dim0, dim1, dim2}, s[2] = {stride1_bytes, stride2_bytes};
  uint32_t b[3] = {box0, box1, box2};
  return get_tmap(out, base, 3, d, s, b, 3);
}

}  // namespace eec
