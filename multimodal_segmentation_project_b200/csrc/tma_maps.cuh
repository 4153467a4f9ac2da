// tma_maps.cuh — host-side construction of TMA tensor maps over NDHWC bf16 activations and the
// device-side cp.async.bulk.tensor wrapper.  cuTensorMapEncodeTiled is fetched through the runtime
// (cudaGetDriverEntryPoint), so the library does not link against libcuda.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// [N, D, H, W, C] bf16, box = 16 channels x box_w x box_h x 1 x 1, SWIZZLE_32B ([voxel][16 ch] rows of 32 bytes);
// out-of-range coordinates are zero-filled
inline int make_ndhwc_map(CUtensorMap* tm, const void* base, int C, int N, int D, int H, int W, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  B200_REQUIRE(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {16, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for C=%d N=%d D=%d H=%d W=%d", (int)r, C, N, D, H, W);
  return B200_OK;
}

// wide rows: box = box_c (32 / 64) channels x box_w x box_h x 1 x 1 with SWIZZLE_64B / SWIZZLE_128B ([voxel][box_c ch] rows)
inline int make_ndhwc_map_wide(CUtensorMap* tm, const void* base, int C, int N, int D, int H, int W, int box_c, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  B200_REQUIRE(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  B200_REQUIRE(box_c == 32 || box_c == 64, B200_ERR_UNSUPPORTED, "make_ndhwc_map_wide: box_c must be 32 or 64");
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(wide) failed (%d) for C=%d N=%d D=%d H=%d W=%d", (int)r, C, N, D, H, W);
  return B200_OK;
}

// same, but the box walks W and H with element stride 2 (every other voxel): box_w / box_h count LOADED voxels.
// Used to fetch one of the eight child planes of a 2x finer grid (ConvTranspose3d k=2,s=2 backward).
inline int make_ndhwc_map_stride2(CUtensorMap* tm, const void* base, int C, int N, int D, int H, int W, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  B200_REQUIRE(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {16, (cuuint32_t)(2 * box_w), (cuuint32_t)(2 * box_h), 1, 1};  // traversed extent; ceil(box/stride) elements are loaded
  cuuint32_t estr[5] = {1, 2, 2, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(stride 2) failed (%d) for C=%d N=%d D=%d H=%d W=%d", (int)r, C, N, D, H, W);
  return B200_OK;
}

__device__ __forceinline__ void load_5d(uint32_t dst, const CUtensorMap* tm, int c, int w, int h, int d, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

}  // namespace tma
