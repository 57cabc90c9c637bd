// Shared host/device helpers for the radzero_b200 CUDA kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/rz_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "radzero_b200 kernels are written for sm_100a (B200) only"
#endif

#define RZ_HIDDEN 768  // hidden_dim of RadZeroLoss (exp/cxr_pt/configs/radzero.yaml:40)

#define RZ_CUDA_OK(expr)                                  \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) { rz_note_cuda_error(_e); return RZ_ERR_CUDA; } \
  } while (0)

#define RZ_LAUNCH_OK()                                    \
  do {                                                    \
    cudaError_t _e = cudaGetLastError();                  \
    if (_e != cudaSuccess) { rz_note_cuda_error(_e); return RZ_ERR_CUDA; } \
  } while (0)

void rz_note_cuda_error(cudaError_t e);
int rz_sm_count();
void rz_count_launch(int n = 1);

namespace rz {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// streaming 128-bit global load (read once, do not pollute L1)
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
               : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// Row loaders: one warp owns one row of RZ_HIDDEN elements; lane holds 24 of them as
// 6 groups of 4 consecutive elements, group j at element offset 4*(lane + 32*j).
template <typename T> struct RowLoad;
template <> struct RowLoad<float> {
  static __device__ __forceinline__ void load(const float* row, int lane, float (&x)[24]) {
    const float4* p = reinterpret_cast<const float4*>(row) + lane;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float4 v = ldg_stream_f4(p + 32 * j);
      x[4 * j + 0] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    }
  }
};
template <> struct RowLoad<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* row, int lane, float (&x)[24]) {
    const uint2* p = reinterpret_cast<const uint2*>(row) + lane;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      uint2 v = ldg_stream_u2(p + 32 * j);
      x[4 * j + 0] = __uint_as_float(v.x << 16);
      x[4 * j + 1] = __uint_as_float(v.x & 0xffff0000u);
      x[4 * j + 2] = __uint_as_float(v.y << 16);
      x[4 * j + 3] = __uint_as_float(v.y & 0xffff0000u);
    }
  }
};
template <> struct RowLoad<__half> {
  static __device__ __forceinline__ void load(const __half* row, int lane, float (&x)[24]) {
    const uint2* p = reinterpret_cast<const uint2*>(row) + lane;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      uint2 v = ldg_stream_u2(p + 32 * j);
      float2 a = __half22float2(*reinterpret_cast<__half2*>(&v.x));
      float2 b = __half22float2(*reinterpret_cast<__half2*>(&v.y));
      x[4 * j + 0] = a.x; x[4 * j + 1] = a.y; x[4 * j + 2] = b.x; x[4 * j + 3] = b.y;
    }
  }
};

struct RowStats { float mean, rstd, inv_norm; };

// LayerNorm (biased variance, eps_ln) then L2 normalisation (x / max(|x|, eps_l2)) of the
// 24 values a lane holds, in place.  gamma/beta are read as float4 at the lane's offsets
// (pass nullptr to skip LayerNorm).  losses.py:90-91,163-164 (LN) and :212-213 (normalize).
__device__ __forceinline__ RowStats ln_l2_row(float (&x)[24], const float* __restrict__ gamma,
                                              const float* __restrict__ beta, int lane,
                                              float eps_ln, float eps_l2, bool do_l2 = true) {
  RowStats st;
  st.mean = 0.f; st.rstd = 1.f; st.inv_norm = 1.f;
  if (gamma != nullptr) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) s += x[i];
    s = warp_sum(s);
    const float mu = s * (1.0f / RZ_HIDDEN);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { x[i] -= mu; v = fmaf(x[i], x[i], v); }
    v = warp_sum(v);
    const float rstd = rsqrtf(v * (1.0f / RZ_HIDDEN) + eps_ln);
    st.mean = mu; st.rstd = rstd;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * j));
      const float4 b = *reinterpret_cast<const float4*>(beta + 4 * (lane + 32 * j));
      x[4 * j + 0] = fmaf(x[4 * j + 0] * rstd, g.x, b.x);
      x[4 * j + 1] = fmaf(x[4 * j + 1] * rstd, g.y, b.y);
      x[4 * j + 2] = fmaf(x[4 * j + 2] * rstd, g.z, b.z);
      x[4 * j + 3] = fmaf(x[4 * j + 3] * rstd, g.w, b.w);
    }
  }
  if (do_l2) {
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) n2 = fmaf(x[i], x[i], n2);
    n2 = warp_sum(n2);
    const float inv = 1.0f / fmaxf(sqrtf(n2), eps_l2);
    st.inv_norm = inv;
#pragma unroll
    for (int i = 0; i < 24; ++i) x[i] *= inv;
  }
  return st;
}

// Byte offset of (row, byte-in-128B-row) inside a 128B-swizzled chunk whose base is
// 1024-byte aligned: rows are 128 B apart, the 16-byte unit index is XOR-ed with row%8.
// This is the layout TMA SWIZZLE_128B produces and UMMA LayoutType::SWIZZLE_128B reads.
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t byte_in_row) {
  return row * 128u + ((((byte_in_row >> 4) ^ (row & 7u)) << 4) | (byte_in_row & 15u));
}

}  // namespace rz
