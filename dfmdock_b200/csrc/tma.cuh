// TMA (cp.async.bulk.tensor) helpers: host-side tensor-map encoding through the driver entry point (no -lcuda link
// dependency) and the device-side issue / barrier wrappers.  Row-major [rows, cols] fp16 activations are described as
// rank-2 tensors {cols (innermost), rows}; a box of 64 columns x R rows with SWIZZLE_128B lands in shared memory exactly
// in the K-major SWIZZLE_128B operand layout the tcgen05 descriptors expect (16-byte chunk c of row r at chunk
// c ^ (r & 7)), rows beyond the tensor are zero-filled by the hardware.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

typedef CUresult (*dfm_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline dfm_encode_tiled_fn dfm_tma_encoder() {
  static dfm_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<dfm_encode_tiled_fn>(p);
  }
  return fn;
}

// [rows, cols] fp16 row-major at `base` (row stride = cols * 2 bytes, 16-byte aligned); box = 64 columns x box_rows rows.
inline int dfm_make_tmap_f16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  dfm_encode_tiled_fn enc = dfm_tma_encoder();
  if (!enc) { dfm_set_error("cuTensorMapEncodeTiled is not available from this driver"); return DFM_ECUDA; }
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { dfm_set_error("cuTensorMapEncodeTiled failed (%d) for a [%llu, %llu] fp16 tensor", (int)r,
                                         (unsigned long long)rows, (unsigned long long)cols); return DFM_ECUDA; }
  return 0;
}

#ifdef __CUDACC__
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// one box: coordinates {c0 = first column, c1 = first row}; completes `bytes of the box` on the mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// tile::gather4: four rows {r0..r3} of the rank-2 tensor (box = 64 columns x 1 row), columns c0..c0+63, land as four
// consecutive 128-byte rows at smem_dst (512-byte aligned inside a 1024-byte aligned SWIZZLE_128B tile); completes
// 512 bytes on the mbarrier
__device__ __forceinline__ void tma_gather4_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3,
                                               uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#endif
