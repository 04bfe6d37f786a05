// Host side of the TMA operand descriptors: tensor-map encoding through the driver entry points (no -lcuda), shared by the
// tcgen05 kernels (gemm_tc.cu, stem_pool.cu). Each translation unit gets its own copy of the (idempotent) entry-point cache.
#pragma once

#include "common.cuh"

namespace {

constexpr int BM = 128;      // rows per tile  (UMMA M)
constexpr int BK = 64;       // fp16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;   // fixed for 16-bit inputs

// ---------------------------------------------------------------------------
// Host side: tensor-map encoding through the driver entry points (no -lcuda).
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                     cuuint32_t, cuuint32_t, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled g_encode_tiled = nullptr;
PFN_encodeIm2col g_encode_im2col = nullptr;

int load_driver_entry_points() {
  if (g_encode_tiled && g_encode_im2col) return MAUV_OK;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return mauv_set_error(MAUV_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  g_encode_tiled = reinterpret_cast<PFN_encodeTiled>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return mauv_set_error(MAUV_ERR_DRIVER, "cuTensorMapEncodeIm2col not available from the driver");
  g_encode_im2col = reinterpret_cast<PFN_encodeIm2col>(fn);
  return MAUV_OK;
}

// [G][rows][K] fp16 row-major, box = 64 x box_rows x 1, 128B swizzle.
int make_tiled_map(CUtensorMap* tm, const void* base, int64_t K, int64_t rows, int64_t G,
                   int64_t sample_stride_elems, int box_rows) {
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(G)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(K) * 2,
                           static_cast<cuuint64_t>(sample_stride_elems) * 2};
  if (G == 1) strides[1] = static_cast<cuuint64_t>(K) * 2 * static_cast<cuuint64_t>(rows);
  cuuint32_t box[3] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return mauv_set_error(MAUV_ERR_DRIVER,
                          "cuTensorMapEncodeTiled failed (%d) K=%lld rows=%lld G=%lld stride=%lld",
                          (int)r, (long long)K, (long long)rows, (long long)G,
                          (long long)sample_stride_elems);
  return MAUV_OK;
}

// NHWC fp16 activations seen as (C, W, H, N) in im2col mode: 64 channels x 128 output pixels.
int make_im2col_map(CUtensorMap* tm, const void* base, int64_t C, int64_t W, int64_t H, int64_t N,
                    int kh, int kw, int stride, int pad, int pixel_box = BM) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W),
                        static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(C) * 2 * W,
                           static_cast<cuuint64_t>(C) * 2 * W * H};
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (kw - 1), pad - (kh - 1)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = g_encode_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims,
                               strides, lower, upper, BK, pixel_box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return mauv_set_error(MAUV_ERR_DRIVER,
                          "cuTensorMapEncodeIm2col failed (%d) C=%lld W=%lld H=%lld N=%lld k=%dx%d s=%d p=%d",
                          (int)r, (long long)C, (long long)W, (long long)H, (long long)N, kh, kw,
                          stride, pad);
  return MAUV_OK;
}

}  // namespace
