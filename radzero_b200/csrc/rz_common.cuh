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

// four floats -> four 16-bit values of type T (bf16 / fp16), packed
template <typename T> __device__ __forceinline__ uint2 pack4(float a, float b, float c, float d);
template <> __device__ __forceinline__ uint2 pack4<__half>(float a, float b, float c, float d) {
  return make_uint2(pack_half2(a, b), pack_half2(c, d));
}
template <> __device__ __forceinline__ uint2 pack4<__nv_bfloat16>(float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}
template <> __device__ __forceinline__ uint2 pack4<float>(float, float, float, float) { return make_uint2(0u, 0u); }

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

// Packed fp32 pairs (Blackwell fma.rn.f32x2 / mul.rn.f32x2): one issue slot for two lanes of epilogue
// arithmetic.  The persistent GEMMs run ONE epilogue warp per scheduler, so their epilogues are bound by
// instruction issue latency, not by any pipe: halving the instruction count is a direct speed-up.
struct P2 { uint64_t v; };
__device__ __forceinline__ P2 p2(float a, float b) {
  P2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ P2 p2(float a) { return p2(a, a); }
__device__ __forceinline__ void p2_unpack(P2 x, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
}
__device__ __forceinline__ P2 p2_fma(P2 a, P2 b, P2 c) {
  P2 r;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ P2 p2_mul(P2 a, P2 b) {
  P2 r;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ P2 p2_add(P2 a, P2 b) {
  P2 r;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ uint32_t p2_pack_h2(P2 x) {
  float a, b;
  p2_unpack(x, a, b);
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ P2 p2_sub(P2 a, P2 b) {            // a - b  (fma with -1: there is no sub.f32x2)
  P2 r;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(b.v), "l"(p2(-1.0f).v), "l"(a.v));
  return r;
}

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

// ln_l2_row with packed fp32 pairs: the same operations in the same order on two neighbouring elements
// per instruction (FADD2 / FFMA2 / FMUL2), i.e. bit-identical results at about half the FP32 issue slots.
// For kernels whose row loop is bound by instruction issue (the converter warps of sim_small_kernel).
__device__ __forceinline__ RowStats ln_l2_row_packed(float (&x)[24], const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, int lane,
                                                     float eps_ln, float eps_l2, bool do_l2 = true) {
  RowStats st;
  st.mean = 0.f; st.rstd = 1.f; st.inv_norm = 1.f;
  P2 v[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) v[i] = p2(x[2 * i], x[2 * i + 1]);
  if (gamma != nullptr) {
    P2 s2 = v[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) s2 = p2_add(s2, v[i]);
    float s0, s1;
    p2_unpack(s2, s0, s1);
    const float mu = warp_sum(s0 + s1) * (1.0f / RZ_HIDDEN);
    const P2 mu2 = p2(mu);
    P2 q2 = p2(0.f);
#pragma unroll
    for (int i = 0; i < 12; ++i) { v[i] = p2_sub(v[i], mu2); q2 = p2_fma(v[i], v[i], q2); }
    float q0, q1;
    p2_unpack(q2, q0, q1);
    const float rstd = rsqrtf(warp_sum(q0 + q1) * (1.0f / RZ_HIDDEN) + eps_ln);
    st.mean = mu; st.rstd = rstd;
    const P2 r2 = p2(rstd);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * j));
      const float4 b = *reinterpret_cast<const float4*>(beta + 4 * (lane + 32 * j));
      v[2 * j] = p2_fma(p2_mul(v[2 * j], r2), p2(g.x, g.y), p2(b.x, b.y));
      v[2 * j + 1] = p2_fma(p2_mul(v[2 * j + 1], r2), p2(g.z, g.w), p2(b.z, b.w));
    }
  }
  if (do_l2) {
    P2 n2 = p2(0.f);
#pragma unroll
    for (int i = 0; i < 12; ++i) n2 = p2_fma(v[i], v[i], n2);
    float n0, n1;
    p2_unpack(n2, n0, n1);
    const float inv = 1.0f / fmaxf(sqrtf(warp_sum(n0 + n1)), eps_l2);
    st.inv_norm = inv;
    const P2 i2 = p2(inv);
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = p2_mul(v[i], i2);
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) p2_unpack(v[i], x[2 * i], x[2 * i + 1]);
  return st;
}

// SEVERAL rows per warp with the lane's gamma / beta resident in registers as packed pairs (g2 / b2: the pairs
// of load_lane_pairs): the operations and their order per row are those of ln_l2_row_packed (bit-identical
// results); the two rows' dependent chains (three shuffle reductions each) are interleaved explicitly, and
// the per-row parameter loads are gone.
__device__ __forceinline__ void load_lane_pairs(const float* __restrict__ p, int lane, float fill, P2 (&r)[12]) {
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float4 v = make_float4(fill, fill, fill, fill);
    if (p != nullptr) v = *reinterpret_cast<const float4*>(p + 4 * (lane + 32 * j));
    r[2 * j] = p2(v.x, v.y);
    r[2 * j + 1] = p2(v.z, v.w);
  }
}
// R rows per warp, interleaved: reductions of R values share each shuffle stage's latency.
template <int R>
__device__ __forceinline__ void warp_sum_rows(float (&a)[R]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
  }
}
template <int R>
__device__ __forceinline__ void ln_l2_rows_packed(float (&x)[R][24], const P2 (&g2)[12], const P2 (&b2)[12],
                                                  bool do_ln, float eps_ln, float eps_l2, bool do_l2) {
  P2 v[R][12];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < 12; ++i) v[r][i] = p2(x[r][2 * i], x[r][2 * i + 1]);
  if (do_ln) {
    float mu[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      P2 s = v[r][0];
#pragma unroll
      for (int i = 1; i < 12; ++i) s = p2_add(s, v[r][i]);
      float a0, a1;
      p2_unpack(s, a0, a1);
      mu[r] = a0 + a1;
    }
    warp_sum_rows<R>(mu);
    float var[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const P2 m = p2(mu[r] * (1.0f / RZ_HIDDEN));
      P2 q = p2(0.f);
#pragma unroll
      for (int i = 0; i < 12; ++i) { v[r][i] = p2_sub(v[r][i], m); q = p2_fma(v[r][i], v[r][i], q); }
      float a0, a1;
      p2_unpack(q, a0, a1);
      var[r] = a0 + a1;
    }
    warp_sum_rows<R>(var);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const P2 rs = p2(rsqrtf(var[r] * (1.0f / RZ_HIDDEN) + eps_ln));
#pragma unroll
      for (int i = 0; i < 12; ++i) v[r][i] = p2_fma(p2_mul(v[r][i], rs), g2[i], b2[i]);
    }
  }
  if (do_l2) {
    float n[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      P2 q = p2(0.f);
#pragma unroll
      for (int i = 0; i < 12; ++i) q = p2_fma(v[r][i], v[r][i], q);
      float a0, a1;
      p2_unpack(q, a0, a1);
      n[r] = a0 + a1;
    }
    warp_sum_rows<R>(n);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const P2 inv = p2(1.0f / fmaxf(sqrtf(n[r]), eps_l2));
#pragma unroll
      for (int i = 0; i < 12; ++i) v[r][i] = p2_mul(v[r][i], inv);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < 12; ++i) p2_unpack(v[r][i], x[r][2 * i], x[r][2 * i + 1]);
}

// Row constants of the LayerNorm parameters used by ln_l2_row_onepass.
struct LnConsts { float sum_g2, sum_gb, sum_b2; };

// Same result as ln_l2_row (LayerNorm then L2 normalisation) with ONE warp reduction stage
// instead of three dependent ones: with x' = x - K (K = the row's first element; LayerNorm is
// shift-invariant and the shift keeps the one-pass moments well conditioned) the five sums
//   S1 = sum x', S2 = sum x'^2, G2 = sum g^2 x'^2, G1 = sum g^2 x', GB = sum g b x'
// give mean' = S1/D, var = S2/D - mean'^2 and
//   |LN(x)|^2 = rstd^2 (G2 - 2 mean' G1 + mean'^2 sum g^2) + 2 rstd (GB - mean' sum g b) + sum b^2,
// after which each element is one pair of FMAs: k = x' (a g) + (b inv - mean' a g), a = rstd inv.
__device__ __forceinline__ void load_lane_params(const float* __restrict__ p, int lane, float (&r)[24]) {
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * (lane + 32 * j));
    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
  }
}
// g / b: this lane's 24 gamma / beta values (load_lane_params), loop-invariant in the callers.
__device__ __forceinline__ RowStats ln_l2_row_onepass(float (&x)[24], const float (&g)[24],
                                                      const float (&b)[24], const LnConsts c,
                                                      float eps_ln, float eps_l2, bool do_l2 = true) {
  const float shift = __shfl_sync(0xffffffffu, x[0], 0);
  float s1 = 0.f, s2 = 0.f, g2 = 0.f, g1 = 0.f, gb = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    x[i] -= shift;
    const float gg = g[i] * g[i];
    const float xx = x[i] * x[i];
    s1 += x[i];
    s2 += xx;
    g2 = fmaf(gg, xx, g2);
    g1 = fmaf(gg, x[i], g1);
    gb = fmaf(g[i] * b[i], x[i], gb);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    g2 += __shfl_xor_sync(0xffffffffu, g2, o);
    g1 += __shfl_xor_sync(0xffffffffu, g1, o);
    gb += __shfl_xor_sync(0xffffffffu, gb, o);
  }
  const float mu = s1 * (1.0f / RZ_HIDDEN);
  const float var = fmaxf(s2 * (1.0f / RZ_HIDDEN) - mu * mu, 0.0f);
  const float rstd = rsqrtf(var + eps_ln);
  float inv = 1.0f;
  if (do_l2) {
    const float n2 = rstd * rstd * (g2 - 2.0f * mu * g1 + mu * mu * c.sum_g2) +
                     2.0f * rstd * (gb - mu * c.sum_gb) + c.sum_b2;
    inv = 1.0f / fmaxf(sqrtf(fmaxf(n2, 0.0f)), eps_l2);
  }
  const float a = rstd * inv;
  const float ma = -mu * a;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const float t = a * g[i];
    x[i] = fmaf(x[i], t, fmaf(ma, g[i], b[i] * inv));
  }
  RowStats st;
  st.mean = mu + shift; st.rstd = rstd; st.inv_norm = inv;
  return st;
}

// ln_l2_row_onepass for register-starved callers: gamma / beta are re-read (L1-resident float4
// loads) in both sweeps instead of living in 48 registers.
__device__ __forceinline__ RowStats ln_l2_row_onepass_mem(float (&x)[24], const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int lane,
                                                          const LnConsts c, float eps_ln, float eps_l2,
                                                          bool do_l2 = true) {
  const float shift = __shfl_sync(0xffffffffu, x[0], 0);
  float s1 = 0.f, s2 = 0.f, g2 = 0.f, g1 = 0.f, gb = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float& xv = x[4 * j + i];
      xv -= shift;
      const float g2v = gg[i] * gg[i], xx = xv * xv;
      s1 += xv;
      s2 += xx;
      g2 = fmaf(g2v, xx, g2);
      g1 = fmaf(g2v, xv, g1);
      gb = fmaf(gg[i] * bb[i], xv, gb);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    g2 += __shfl_xor_sync(0xffffffffu, g2, o);
    g1 += __shfl_xor_sync(0xffffffffu, g1, o);
    gb += __shfl_xor_sync(0xffffffffu, gb, o);
  }
  const float mu = s1 * (1.0f / RZ_HIDDEN);
  const float var = fmaxf(s2 * (1.0f / RZ_HIDDEN) - mu * mu, 0.0f);
  const float rstd = rsqrtf(var + eps_ln);
  float inv = 1.0f;
  if (do_l2) {
    const float n2 = rstd * rstd * (g2 - 2.0f * mu * g1 + mu * mu * c.sum_g2) +
                     2.0f * rstd * (gb - mu * c.sum_gb) + c.sum_b2;
    inv = 1.0f / fmaxf(sqrtf(fmaxf(n2, 0.0f)), eps_l2);
  }
  const float a = rstd * inv;
  const float ma = -mu * a;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      x[4 * j + i] = fmaf(x[4 * j + i], a * gg[i], fmaf(ma, gg[i], bb[i] * inv));
  }
  RowStats st;
  st.mean = mu + shift; st.rstd = rstd; st.inv_norm = inv;
  return st;
}

// sum g^2, sum g b, sum b^2 over the 768 features; one warp, result in every lane
__device__ __forceinline__ LnConsts ln_consts_warp(const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, int lane) {
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = lane; i < RZ_HIDDEN; i += 32) {
    const float g = gamma[i], bb = beta[i];
    a = fmaf(g, g, a); b = fmaf(g, bb, b); c = fmaf(bb, bb, c);
  }
  LnConsts r;
  r.sum_g2 = warp_sum(a); r.sum_gb = warp_sum(b); r.sum_b2 = warp_sum(c);
  return r;
}

// Byte offset of (row, byte-in-128B-row) inside a 128B-swizzled chunk whose base is
// 1024-byte aligned: rows are 128 B apart, the 16-byte unit index is XOR-ed with row%8.
// This is the layout TMA SWIZZLE_128B produces and UMMA LayoutType::SWIZZLE_128B reads.
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t byte_in_row) {
  return row * 128u + ((((byte_in_row >> 4) ^ (row & 7u)) << 4) | (byte_in_row & 15u));
}

}  // namespace rz
