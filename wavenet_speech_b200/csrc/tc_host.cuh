// Host-side helpers of the tensor-core path: TMA tensor-map construction through the driver entry point
// (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace wnb {

constexpr int RB_TILE = 128;   // frames per CTA tile (TMA box rows)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn rb_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static inline int rb_map_2d(CUtensorMap* m, const void* ptr, int rows, int cols, int boxrows) {
  EncodeTiledFn enc = rb_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 5; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d) failed: %d", (int)r); return 5; }
  return 0;
}

// NLC tensor [B][T][C] of `esize`-byte elements -> boxes [1][128 frames][128 bytes], 128B swizzle
static inline int rb_map_nlc(CUtensorMap* m, const void* ptr, int B, int T, int C, int esize, int box_rows = RB_TILE) {
  EncodeTiledFn enc = rb_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 5; }
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * esize, (cuuint64_t)T * C * esize};
  cuuint32_t box[3] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(nlc esize %d) failed: %d", esize, (int)r); return 5; }
  return 0;
}


// NCL tensor [B][C][T] of `esize`-byte elements -> boxes [1][crows channels][tbox frames], no swizzle (output only)
static inline int rb_map_ncl(CUtensorMap* m, const void* ptr, int B, int C, int T, int esize, int tbox, int crows) {
  EncodeTiledFn enc = rb_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 5; }
  cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)T * esize, (cuuint64_t)T * C * esize};
  cuuint32_t box[3] = {(cuuint32_t)tbox, (cuuint32_t)crows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(ncl esize %d) failed: %d", esize, (int)r); return 5; }
  return 0;
}

}  // namespace wnb
