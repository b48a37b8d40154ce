// Host-side helper shared by the kernels that stage tiles with TMA: the driver's cuTensorMapEncodeTiled is fetched through
// the runtime (cudaGetDriverEntryPoint), so the library links against libcudart only.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ftc {

typedef CUresult (*TmaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmaEncodeTiledFn tma_encoder() {
  static TmaEncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (TmaEncodeTiledFn)p;
  }
  return fn;
}

// 4-D NHWC activation map: dims {C, W, H, B} (elements), pixel stride in elements, box {bc, bw, bh, 1}; zero fill out of bounds
inline int tma_encode_nhwc(CUtensorMap* tm, const void* base, int dtype, uint64_t C, uint64_t pix_stride, uint64_t W, uint64_t H,
                           uint64_t B, uint32_t bc, uint32_t bw, uint32_t bh, CUtensorMapSwizzle swz) {
  TmaEncodeTiledFn enc = tma_encoder();
  FTC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available");
  const uint64_t es = dtype == DT_F32 ? 4 : 2;
  FTC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (pix_stride * es) % 16 == 0, "TMA tensor alignment");
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {pix_stride * es, pix_stride * es * W, pix_stride * es * W * H};
  cuuint32_t box[4] = {bc, bw, bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, dtype == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return -2;
  }
  return 0;
}

}  // namespace ftc
