// Backward of K1+K2 (row LayerNorm + L2 normalisation, losses.py:90-91, 163-164, 212-213):
// given dL/dk for the normalised rows k = y/|y|, y = LN(x) = xhat*gamma + beta, produce dL/dx
// and the LayerNorm parameter gradients.  One warp per row, the row statistics are recomputed
// from x (cheaper than storing them), gamma/beta gradients are accumulated in registers per
// warp, reduced per CTA in shared memory and summed by a second tiny kernel in a fixed order
// (deterministic: no float atomics).
#include "rz_common.cuh"

namespace {

constexpr int kWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kWarps * 32, 2)
prep_rows_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, long long rows, int rows_per_group,
                     int rows_per_group_padded, const float* __restrict__ dnorm, int l2,
                     void* __restrict__ dx_out, int dx_native, float* __restrict__ part /* [grid][2][768] */) {
  __shared__ float red[kWarps][2 * RZ_HIDDEN / 4];   // reduced in 4 passes of 384 floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * kWarps + warp;
  const long long nwarps = (long long)gridDim.x * kWarps;
  rz::P2 dg2[12], db2[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) { dg2[i] = rz::p2(0.f); db2[i] = rz::p2(0.f); }
  const bool has_ln = gamma != nullptr;
  for (long long row = warp0; row < rows; row += nwarps) {
    const long long g = row / rows_per_group;
    const long long slot = g * rows_per_group_padded + (row - g * rows_per_group);
    float v[24], d[24];
    rz::RowLoad<T>::load(x + row * RZ_HIDDEN, lane, v);
    {
      const float4* p = reinterpret_cast<const float4*>(dnorm + slot * RZ_HIDDEN) + lane;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float4 t = rz::ldg_stream_f4(p + 32 * j);
        d[4 * j] = t.x; d[4 * j + 1] = t.y; d[4 * j + 2] = t.z; d[4 * j + 3] = t.w;
      }
    }
    // packed fp32 pairs (fma.rn.f32x2): the row is ~400 dependent FP32 operations per lane between six
    // shuffle reductions, and 16 warps per SM cannot hide that -- half the instructions, same order of
    // operations per element
    using rz::P2; using rz::p2; using rz::p2_fma; using rz::p2_mul; using rz::p2_add; using rz::p2_sub;
    using rz::p2_unpack;
    P2 v2[12], d2[12], gm2[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { v2[i] = p2(v[2 * i], v[2 * i + 1]); d2[i] = p2(d[2 * i], d[2 * i + 1]); }
    auto hsum = [](P2 a) { float x0, x1; p2_unpack(a, x0, x1); return x0 + x1; };
    float rstd = 1.f;
    if (has_ln) {
      P2 s2 = v2[0];
#pragma unroll
      for (int i = 1; i < 12; ++i) s2 = p2_add(s2, v2[i]);
      const float mu = rz::warp_sum(hsum(s2)) * (1.0f / RZ_HIDDEN);
      const P2 mu2 = p2(mu);
      P2 q2 = p2(0.f);
#pragma unroll
      for (int i = 0; i < 12; ++i) { v2[i] = p2_sub(v2[i], mu2); q2 = p2_fma(v2[i], v2[i], q2); }
      rstd = rsqrtf(rz::warp_sum(hsum(q2)) * (1.0f / RZ_HIDDEN) + RZ_LN_EPS);
      const P2 r2 = p2(rstd);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * j));
        gm2[2 * j] = p2(t.x, t.y); gm2[2 * j + 1] = p2(t.z, t.w);
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) v2[i] = p2_mul(v2[i], r2);            // v = xhat
    }
    if (l2) {
      // y (LayerNorm output) -> k = y/|y|;  dy = (d - k <k,d>) / |y|
      P2 y2[12];
      if (has_ln) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const float4 t = *reinterpret_cast<const float4*>(beta + 4 * (lane + 32 * j));
          y2[2 * j] = p2_fma(v2[2 * j], gm2[2 * j], p2(t.x, t.y));
          y2[2 * j + 1] = p2_fma(v2[2 * j + 1], gm2[2 * j + 1], p2(t.z, t.w));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) y2[i] = v2[i];
      }
      P2 n22 = p2(0.f), yd2 = p2(0.f);
#pragma unroll
      for (int i = 0; i < 12; ++i) { n22 = p2_fma(y2[i], y2[i], n22); yd2 = p2_fma(y2[i], d2[i], yd2); }
      const float n2 = rz::warp_sum(hsum(n22));
      const float yd = rz::warp_sum(hsum(yd2));
      const float inv = 1.0f / fmaxf(sqrtf(n2), RZ_L2_EPS);
      const P2 nc2 = p2(-yd * inv * inv), inv2 = p2(inv);             // <k,d>/|y| = <y,d>/|y|^2
#pragma unroll
      for (int i = 0; i < 12; ++i) d2[i] = p2_mul(p2_fma(y2[i], nc2, d2[i]), inv2);
    }
    if (has_ln) {
      P2 m12 = p2(0.f), m22 = p2(0.f);
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        dg2[i] = p2_fma(d2[i], v2[i], dg2[i]);
        db2[i] = p2_add(db2[i], d2[i]);
        d2[i] = p2_mul(d2[i], gm2[i]);                                  // d = dxhat
        m12 = p2_add(m12, d2[i]);
        m22 = p2_fma(d2[i], v2[i], m22);
      }
      const float m1 = rz::warp_sum(hsum(m12)) * (1.0f / RZ_HIDDEN);
      const float m2 = rz::warp_sum(hsum(m22)) * (1.0f / RZ_HIDDEN);
      const P2 nm1 = p2(-m1), nm2 = p2(-m2), r2 = p2(rstd);
#pragma unroll
      for (int i = 0; i < 12; ++i) d2[i] = p2_mul(r2, p2_fma(v2[i], nm2, p2_add(d2[i], nm1)));
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) p2_unpack(d2[i], d[2 * i], d[2 * i + 1]);
    if (dx_native && sizeof(T) == 2) {
      // dL/dx in the input's own 16-bit type: no fp32 round trip + conversion pass afterwards
      uint2* o = reinterpret_cast<uint2*>(static_cast<T*>(dx_out) + row * RZ_HIDDEN) + lane;
#pragma unroll
      for (int j = 0; j < 6; ++j) o[32 * j] = rz::pack4<T>(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]);
    } else {
      float4* o = reinterpret_cast<float4*>(static_cast<float*>(dx_out) + row * RZ_HIDDEN) + lane;
#pragma unroll
      for (int j = 0; j < 6; ++j) o[32 * j] = make_float4(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]);
    }
  }
  if (part == nullptr) return;
  float dg[24], db[24];
#pragma unroll
  for (int i = 0; i < 12; ++i) { rz::p2_unpack(dg2[i], dg[2 * i], dg[2 * i + 1]); rz::p2_unpack(db2[i], db[2 * i], db[2 * i + 1]); }
  // CTA reduction of the parameter gradients, 384 floats (= 12 of the 48 per-lane values) a pass
  float* out = part + (long long)blockIdx.x * 2 * RZ_HIDDEN;
  for (int pass = 0; pass < 4; ++pass) {
    // pass 0,1: dgamma groups j = 0..2 / 3..5 ; pass 2,3: dbeta
    const float* src = pass < 2 ? dg : db;
    const int j0 = (pass & 1) * 3;
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        red[warp][jj * 128 + lane * 4 + e] = (pass < 2 ? dg : db)[4 * (j0 + jj) + e];
    }
    (void)src;
    __syncthreads();
    for (int i = threadIdx.x; i < 384; i += kWarps * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += red[w][i];
      // element index: group j = j0 + i/128, lane = (i%128)/4, e = i%4 -> feature 4*(lane+32*j)+e
      const int jj = i >> 7, ln = (i & 127) >> 2, e = i & 3;
      out[(pass < 2 ? 0 : RZ_HIDDEN) + 4 * (ln + 32 * (j0 + jj)) + e] = s;
    }
    __syncthreads();
  }
}

__global__ void param_grad_sum_kernel(const float* __restrict__ part, int n_parts,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                      int accumulate, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * RZ_HIDDEN) return;
  float s = 0.f;
  for (int k = 0; k < n_parts; ++k) s += part[(long long)k * 2 * RZ_HIDDEN + i];
  s *= scale;
  float* dst = i < RZ_HIDDEN ? dgamma + i : dbeta + (i - RZ_HIDDEN);
  *dst = accumulate ? *dst + s : s;
}

}  // namespace

extern "C" int rz_prep_rows_bwd_blocks(long long rows) {
  if (rows <= 0) return 0;
  long long blocks = (rows + kWarps - 1) / kWarps;
  const long long cap = (long long)rz_sm_count() * 2;
  return (int)(blocks > cap ? cap : blocks);
}

extern "C" int rz_prep_rows_bwd(const void* x, int dtype, const float* gamma, const float* beta,
                                long long rows, int rows_per_group, int rows_per_group_padded,
                                const float* dnorm, int l2, void* dx, int dx_native, float* partials,
                                float* dgamma, float* dbeta, int accumulate, float grad_scale,
                                void* stream) {
  if (x == nullptr || dnorm == nullptr || dx == nullptr || rows < 0) return RZ_ERR_INVALID;
  if (rows_per_group <= 0 || rows_per_group_padded < rows_per_group) return RZ_ERR_INVALID;
  if ((gamma == nullptr) != (beta == nullptr)) return RZ_ERR_INVALID;
  if (gamma != nullptr && (partials == nullptr || dgamma == nullptr || dbeta == nullptr)) return RZ_ERR_INVALID;
  if (rows == 0) return RZ_OK;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dnorm) & 15) ||
      (reinterpret_cast<uintptr_t>(dx) & 15))
    return RZ_ERR_ALIGNMENT;
  const int blocks = rz_prep_rows_bwd_blocks(rows);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* part = gamma != nullptr ? partials : nullptr;
  switch (dtype) {
    case RZ_F32:
      prep_rows_bwd_kernel<float><<<blocks, kWarps * 32, 0, s>>>(
          static_cast<const float*>(x), gamma, beta, rows, rows_per_group, rows_per_group_padded, dnorm,
          l2, dx, dx_native, part);
      break;
    case RZ_BF16:
      prep_rows_bwd_kernel<__nv_bfloat16><<<blocks, kWarps * 32, 0, s>>>(
          static_cast<const __nv_bfloat16*>(x), gamma, beta, rows, rows_per_group, rows_per_group_padded,
          dnorm, l2, dx, dx_native, part);
      break;
    case RZ_F16:
      prep_rows_bwd_kernel<__half><<<blocks, kWarps * 32, 0, s>>>(
          static_cast<const __half*>(x), gamma, beta, rows, rows_per_group, rows_per_group_padded, dnorm,
          l2, dx, dx_native, part);
      break;
    default:
      return RZ_ERR_INVALID;
  }
  RZ_LAUNCH_OK();
  rz_count_launch();
  if (gamma != nullptr) {
    param_grad_sum_kernel<<<(2 * RZ_HIDDEN + 255) / 256, 256, 0, s>>>(partials, blocks, dgamma, dbeta,
                                                                      accumulate, grad_scale);
    RZ_LAUNCH_OK();
    rz_count_launch();
  }
  return RZ_OK;
}
