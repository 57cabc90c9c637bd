// Host-side construction of TMA tensor maps.  libcuda is NOT linked: the driver entry
// point is resolved at run time through the (statically linked) runtime, so that
// librz_b200.so still loads on a machine without a GPU driver (the CPU test box).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rz {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn != nullptr) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      p == nullptr)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D fp16/bf16 row-major tensor [rows, cols] (cols contiguous, row pitch `pitch_bytes`),
// box = [box_rows, 64 cols] = box_rows rows of 128 B, SWIZZLE_128B: the shared-memory image
// is exactly the K-major UMMA operand chunk (rows 128 B apart, 16-byte units XOR row%8).
// Out-of-bounds rows/cols are zero-filled.
inline bool make_map_2d_sw128(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t pitch_bytes, uint32_t box_rows, uint32_t box_cols = 64) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 3-D variant [batch, rows, cols]: coordinates (col, row, batch); rows beyond `rows` of a
// batch entry are zero-filled instead of running into the next entry.
inline bool make_map_3d_sw128(CUtensorMap* map, const void* base, uint64_t batch, uint64_t rows,
                              uint64_t cols, uint64_t pitch_bytes, uint64_t batch_pitch_bytes,
                              uint32_t box_rows, uint32_t box_cols = 64,
                              CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {pitch_bytes, batch_pitch_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, dtype, 3, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
// fp32 variant for TMA STORES of [batch, rows, cols] fp32 tensors: box = [box_rows, 32 cols]
// (32 floats = one 128-byte swizzle row).  pitch_bytes must be a multiple of 16.
inline bool make_map_3d_f32_sw128(CUtensorMap* map, const void* base, uint64_t batch, uint64_t rows,
                                  uint64_t cols, uint64_t pitch_bytes, uint64_t batch_pitch_bytes,
                                  uint32_t box_rows) {
  return make_map_3d_sw128(map, base, batch, rows, cols, pitch_bytes, batch_pitch_bytes, box_rows, 32,
                           CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

}  // namespace rz
