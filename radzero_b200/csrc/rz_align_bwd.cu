// A3: backward of the AlignTransformer layers (transformers Dinov2Layer, called from
// exp/cxr_pt/model/align_transformers.py:37-45) -- the only trainable vision compute of RadZero
// (radzero.yaml `module_to_update: [align_transformer, ...]`).
//
// The nine GEMM-shaped products of a layer's backward (dX = dY W and dW = dY^T X for the four linear
// layers, plus the recomputation of fc1's pre-activation) run on the tcgen05 GEMM of rz_align_linear.cu
// (rz_linear); this file holds what sits between them:
//
//   rz_grad_scale      one power-of-two scale for the whole fp16 gradient chain, from max|dL/dtokens|,
//                      computed on the device (no host sync): every fp16 gradient carries 2^k, every fp32
//                      result is multiplied by 2^-k (the chain is linear in the incoming gradient)
//   rz_ls_cast_bwd     do16 = fp16(2^k ls dy) (ls optional), dls += sum_rows dy o (optional)
//   rz_ls_weight_bwd   Dinov2LayerScale backward from the weight gradient of the unscaled product: no recomputation
//   rz_transpose_pad   [rows, cols] -> [cols, rows_padded] fp16 (the K-major operands of the dW GEMMs,
//                      K = rows) + the bias gradients as column sums of the same read
//   rz_gelu_bwd        du16 = dg16 gelu_erf'(u16)
//   rz_ln_rows_bwd     nn.LayerNorm backward of one row per warp + the residual-path gradient (+ the fp16
//                      operand of the next product from the same registers)
//   rz_attention_bwd   softmax(q k^T) v backward per (image, head), head dim 64, warp-level
//                      mma.sync.m16n8k16 (fp16 in, fp32 accumulate), recomputing the probabilities:
//                        kernel 1 (64 query rows / CTA): delta = rowsum(dO o), dQ against a running maximum
//                                 (one pass over the keys, the log-sum-exp falls out at the end)
//                        kernel 2 (64 key rows / CTA):   dK, dV
//                      No atomics: every output element has one owner, the result is run-to-run identical.
//                      This is the one kernel of the repo on the legacy tensor-core path; a tcgen05 version
//                      on the forward kernel's skeleton (rz_align_attn.cu) is the known next step (DESIGN).
#ifndef RZ_ATTN_BWD_DQ_CTAS
#define RZ_ATTN_BWD_DQ_CTAS 4      // resident CTAs per SM the two attention-backward kernels are compiled for
#endif
#ifndef RZ_ATTN_BWD_DKV_CTAS
#define RZ_ATTN_BWD_DKV_CTAS 3
#endif
#include <algorithm>

#include "rz_common.cuh"

namespace {

constexpr int kD = RZ_HIDDEN;
constexpr int kScaleVec = 3072;          // widest GEMM output the 2^-k vector serves (mlp.fc1)
constexpr int kScaleFloats = 4 + kScaleVec;

// ------------------------------------------------------------------------------------------ scale
__global__ void amax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ amax_bits) {
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n & ~3ll; i < n; ++i) m = fmaxf(m, fabsf(x[i]));
  m = rz::warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(amax_bits, __float_as_uint(m));   // non-negative floats order as integers
}

// sc[0] = 2^k, sc[1] = 2^-k, sc[4 .. 4 + 3072) = 2^-k (the `scale` vector of rz_linear's fp32 epilogue);
// k puts max|g| 2^k in [32, 64): ten binades of head room below the fp16 maximum for the growth along the chain
__global__ void scale_kernel(const unsigned* __restrict__ amax_bits, float* __restrict__ sc) {
  const float a = __uint_as_float(*amax_bits);
  int k = 0;
  if (a > 0.f && a < 3.0e38f) {
    const int e = (int)((__float_as_uint(a) >> 23) & 0xff) - 126;   // a = f 2^e, f in [0.5, 1) (subnormals: e = -126)
    k = max(-60, min(60, 6 - e));
  }
  const float up = __uint_as_float((unsigned)(127 + k) << 23), down = __uint_as_float((unsigned)(127 - k) << 23);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) { sc[0] = up; sc[1] = down; sc[2] = a; sc[3] = (float)k; }
  if (t < kScaleVec) sc[4 + t] = down;
}

// ------------------------------------------------------------------------------------------ LayerScale
// one thread = four columns of a row; the block walks rows with stride gridDim.x
__global__ void __launch_bounds__(192) ls_cast_kernel(const float* __restrict__ dy, const float* __restrict__ ls,
                                                      const __half* __restrict__ o16, const float* __restrict__ sc,
                                                      long long rows, __half* __restrict__ do16,
                                                      float* __restrict__ dls) {
  const int c = threadIdx.x * 4;
  const float up = sc[0];
  const float4 l = ls != nullptr ? *reinterpret_cast<const float4*>(ls + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4 g = *reinterpret_cast<const float4*>(dy + r * kD + c);
    *reinterpret_cast<uint2*>(do16 + r * kD + c) =
        rz::pack4<__half>(g.x * l.x * up, g.y * l.y * up, g.z * l.z * up, g.w * l.w * up);
    if (o16 != nullptr) {
      const uint2 ov = *reinterpret_cast<const uint2*>(o16 + r * kD + c);
      const float2 o01 = __half22float2(*reinterpret_cast<const __half2*>(&ov.x));
      const float2 o23 = __half22float2(*reinterpret_cast<const __half2*>(&ov.y));
      a0 += g.x * o01.x; a1 += g.y * o01.y; a2 += g.z * o23.x; a3 += g.w * o23.y;
    }
  }
  if (dls != nullptr && o16 != nullptr) {
    atomicAdd(dls + c, a0); atomicAdd(dls + c + 1, a1); atomicAdd(dls + c + 2, a2); atomicAdd(dls + c + 3, a3);
  }
}

// LayerScale without touching the activations: with G = dy^T x (the weight gradient of the UNSCALED
// product, one row per output feature n), o = x W^T + b and y = ls * o,
//   dW[n, :] = ls[n] G[n, :],   db[n] = ls[n] c[n],   dls[n] = sum_k W[n, k] G[n, k] + b[n] c[n],   c = colsum(dy),
// so the backward never recomputes o: one block per output feature finishes all three from G in place.
__global__ void __launch_bounds__(256) ls_weight_kernel(float* __restrict__ g, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ cs,
                                                        const float* __restrict__ ls, int k, float* __restrict__ dls) {
  __shared__ float red[8];
  const int n = blockIdx.x;
  const float l = ls[n];
  float s = 0.f;
  for (int i = threadIdx.x * 4; i < k; i += 1024) {
    float4 gv = *reinterpret_cast<float4*>(g + (long long)n * k + i);
    const float4 wv = *reinterpret_cast<const float4*>(w + (long long)n * k + i);
    s += gv.x * wv.x + gv.y * wv.y + gv.z * wv.z + gv.w * wv.w;
    gv.x *= l; gv.y *= l; gv.z *= l; gv.w *= l;
    *reinterpret_cast<float4*>(g + (long long)n * k + i) = gv;
  }
  s = rz::warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    const float c = cs[n];
    dls[n] = t + (bias != nullptr ? bias[n] * c : 0.f);
    cs[n] = l * c;
  }
}

// ------------------------------------------------------------------------------------------ transpose
// 64 x 64 tiles; rows >= `rows` read as zero (rows_padded % 64 == 0, cols % 64 == 0).  The tile's 16-byte
// column chunks are XOR-swizzled by the row's octet, so that the transposed read -- eight lanes walking rows
// rh, rh + 8, ... of one column -- hits eight different bank groups instead of one (pitch 72 halfs alone left
// an 8-way conflict there).
__global__ void __launch_bounds__(256) transpose_kernel(const __half* __restrict__ in, long long rows, int cols,
                                                        long long rows_padded, __half* __restrict__ out,
                                                        float* __restrict__ colsum, const float* __restrict__ sc) {
  __shared__ __align__(16) __half tile[64][72];
  const long long r0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64, t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (t >> 3) + 32 * i, ch = t & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r0 + r < rows) v = *reinterpret_cast<const uint4*>(in + (r0 + r) * cols + c0 + ch * 8);
    *reinterpret_cast<uint4*>(&tile[r][(ch ^ ((r >> 3) & 7)) * 8]) = v;
  }
  __syncthreads();
  if (colsum != nullptr && t < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < 64; ++r) s += __half2float(tile[r][(((t >> 3) ^ ((r >> 3) & 7)) << 3) + (t & 7)]);
    atomicAdd(colsum + c0 + t, s * sc[1]);
  }
  if (out != nullptr) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = (t >> 3) + 32 * i, ro = t & 7;       // ro: the row octet this thread gathers (rows 8 ro .. 8 ro + 7)
      const int pc = (((c >> 3) ^ ro) << 3) + (c & 7);
      __align__(16) __half v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tile[8 * ro + j][pc];
      *reinterpret_cast<uint4*>(out + (long long)(c0 + c) * rows_padded + r0 + 8 * ro) = *reinterpret_cast<uint4*>(v);
    }
  }
}

// ------------------------------------------------------------------------------------------ GELU'
__device__ __forceinline__ float gelu_erf_grad(float u) {
  // d/du [u Phi(u)] = Phi(u) + u phi(u)   (transformers GELUActivation = erf form, Dinov2MLP)
  const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * u * u);
  return cdf + u * pdf;
}

__global__ void gelu_bwd_kernel(const __half* __restrict__ dg, const __half* __restrict__ u, long long n8,
                                __half* __restrict__ du) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 gv = reinterpret_cast<const uint4*>(dg)[i], uv = reinterpret_cast<const uint4*>(u)[i];
    const uint32_t gs[4] = {gv.x, gv.y, gv.z, gv.w}, us[4] = {uv.x, uv.y, uv.z, uv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g2 = __half22float2(*reinterpret_cast<const __half2*>(&gs[j]));
      const float2 u2 = __half22float2(*reinterpret_cast<const __half2*>(&us[j]));
      o[j] = rz::pack_half2(g2.x * gelu_erf_grad(u2.x), g2.y * gelu_erf_grad(u2.y));
    }
    reinterpret_cast<uint4*>(du)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm
// one warp per row, a lane owns columns lane * 4 + 128 i (i < 6).  dx = dres + 2^-k LN'(dh); the
// parameter gradients are kept per lane over the warp's rows, folded across the CTA's warps through
// shared memory and added to the global vectors with one atomic per column and CTA.
constexpr int kLnWarps = 8;

__global__ void __launch_bounds__(32 * kLnWarps) ln_bwd_kernel(const float* __restrict__ x, const __half* __restrict__ dh,
                                                              const float* __restrict__ gamma, float eps,
                                                              const float* __restrict__ dres, const float* __restrict__ sc,
                                                              long long rows, float* __restrict__ dx,
                                                              __half* __restrict__ dx16,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float fold[kLnWarps][kD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float down = sc[1], up = sc[0];
  float g[24], ag[24], ab[24];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(gamma + lane * 4 + 128 * i);
    g[4 * i] = v.x; g[4 * i + 1] = v.y; g[4 * i + 2] = v.z; g[4 * i + 3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) ag[i] = ab[i] = 0.f;
  for (long long r = (long long)blockIdx.x * kLnWarps + warp; r < rows; r += (long long)gridDim.x * kLnWarps) {
    float xv[24], dv[24];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(x + r * kD + lane * 4 + 128 * i);
      xv[4 * i] = v.x; xv[4 * i + 1] = v.y; xv[4 * i + 2] = v.z; xv[4 * i + 3] = v.w;
      s += (v.x + v.y) + (v.z + v.w);
      const uint2 h = *reinterpret_cast<const uint2*>(dh + r * kD + lane * 4 + 128 * i);
      const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
      const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
      dv[4 * i] = h01.x; dv[4 * i + 1] = h01.y; dv[4 * i + 2] = h23.x; dv[4 * i + 3] = h23.y;
    }
    const float mean = rz::warp_sum(s) * (1.f / kD);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { xv[i] -= mean; q += xv[i] * xv[i]; }
    const float rstd = rsqrtf(rz::warp_sum(q) * (1.f / kD) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      xv[i] *= rstd;                       // x_hat
      ag[i] += dv[i] * xv[i];
      ab[i] += dv[i];
      dv[i] *= g[i];                       // dL/dx_hat
      s1 += dv[i];
      s2 += dv[i] * xv[i];
    }
    s1 = rz::warp_sum(s1) * (1.f / kD);
    s2 = rz::warp_sum(s2) * (1.f / kD);
    const float k = rstd * down;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dres != nullptr) o = *reinterpret_cast<const float4*>(dres + r * kD + lane * 4 + 128 * i);
      o.x += k * (dv[4 * i] - s1 - xv[4 * i] * s2);
      o.y += k * (dv[4 * i + 1] - s1 - xv[4 * i + 1] * s2);
      o.z += k * (dv[4 * i + 2] - s1 - xv[4 * i + 2] * s2);
      o.w += k * (dv[4 * i + 3] - s1 - xv[4 * i + 3] * s2);
      *reinterpret_cast<float4*>(dx + r * kD + lane * 4 + 128 * i) = o;
      if (dx16 != nullptr)      // the next product's fp16 operand (2^k dx) from the same registers
        *reinterpret_cast<uint2*>(dx16 + r * kD + lane * 4 + 128 * i) = rz::pack4<__half>(o.x * up, o.y * up, o.z * up, o.w * up);
    }
  }
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? dgamma : dbeta;
    if (dst == nullptr) continue;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) fold[warp][lane * 4 + 128 * i + j] = pass == 0 ? ag[4 * i + j] : ab[4 * i + j];
    __syncthreads();
    for (int c = threadIdx.x; c < kD; c += 32 * kLnWarps) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kLnWarps; ++w) s += fold[w][c];
      atomicAdd(dst + c, s * down);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ attention
constexpr int kT = 64;        // rows per tile (query tile of kernel 1, key tile of kernel 2) = streamed tile
constexpr int kLd = 72;       // shared-memory row pitch in halfs: 144 B, ldmatrix rows fall in distinct 16-byte lanes

struct AttnBwd {
  const __half* qkv; const __half* o; const __half* dout;
  __half* dqkv; float* lse; float* delta;
  int B, L, H, Lp; float q_scale;      // Lp = L rounded up to the tile: row pitch of lse / delta
};

__device__ __forceinline__ void ldsm_x4(const __half* p, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(const __half* p, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 64 rows x 64 halfs from `base + (r0 + r) * ld` into a [64][72] tile; rows >= L are zero
__device__ __forceinline__ void load_tile(__half* tile, const __half* base, long long ld, int r0, int L) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = threadIdx.x + 128 * i, r = ch >> 3, c = (ch & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r0 + r < L) v = *reinterpret_cast<const uint4*>(base + (long long)(r0 + r) * ld + c);
    *reinterpret_cast<uint4*>(tile + r * kLd + c) = v;
  }
}

// the 16 x 64 A operand of this warp (rows w16 .. w16 + 15 of a tile, K = the 64 columns) as 4 k-steps
__device__ __forceinline__ void load_a_frags(const __half* tile, int w16, uint32_t (&f)[4][4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    ldsm_x4(tile + (w16 + (lane & 15)) * kLd + 16 * t + (lane >> 4) * 8, f[t][0], f[t][1], f[t][2], f[t][3]);
}

__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
}

// ---- the warp-level products, on HALF of the streamed tile at a time (32 of its 64 rows): the score /
// probability blocks of a warp are 16 x 32, which keeps the kernels at 3 - 4 resident CTAs per SM (16 x 64
// blocks: 2).  One ldmatrix.x4 = the B fragments of TWO column blocks for one k-step, so that consecutive
// MMAs write different accumulators.
// c[nb] (nb < 4) += A (16 x 64) . T[32 hb .. 32 hb + 31]^T, T = a [64][72] tile holding [n][k]
__device__ __forceinline__ void gemm_nt_half(float (&c)[4][4], const uint32_t (&a)[4][4], const __half* tile, int hb) {
  const int lane = threadIdx.x & 31;
  const __half* base = tile + (32 * hb + 8 * (lane >> 4) + (lane & 7)) * kLd + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(base + 16 * np * kLd + 16 * t, b0, b1, b2, b3);
      mma16816(c[2 * np], a[t], b0, b1);
      mma16816(c[2 * np + 1], a[t], b2, b3);
    }
}
// c[nb] (nb < 8) += A (16 x 32, fragments of two k-steps) . T[32 hb .. 32 hb + 31], T holding [k][n]
__device__ __forceinline__ void gemm_nn_half(float (&c)[8][4], const uint32_t (&a)[2][4], const __half* tile, int hb) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile + (32 * hb + 16 * t + (lane & 15)) * kLd + 16 * np + (lane >> 4) * 8, b0, b1, b2, b3);
      mma16816(c[2 * np], a[t], b0, b1);
      mma16816(c[2 * np + 1], a[t], b2, b3);
    }
}
__device__ __forceinline__ void acc_to_a_half(const float (&c)[4][4], uint32_t (&a)[2][4]) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    a[t][0] = rz::pack_half2(c[2 * t][0], c[2 * t][1]);
    a[t][1] = rz::pack_half2(c[2 * t][2], c[2 * t][3]);
    a[t][2] = rz::pack_half2(c[2 * t + 1][0], c[2 * t + 1][1]);
    a[t][3] = rz::pack_half2(c[2 * t + 1][2], c[2 * t + 1][3]);
  }
}
__device__ __forceinline__ void zero_half(float (&c)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
}

// rows w16 + g and w16 + g + 8 of the tile starting at global row r0, 64 columns at `dst` (row pitch ld)
__device__ __forceinline__ void store_acc(__half* dst, long long ld, int r0, int w16, int L, const float (&c)[8][4],
                                          float mul_a, float mul_b) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int ra = r0 + w16 + g, rb = ra + 8;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    if (ra < L) *reinterpret_cast<uint32_t*>(dst + (long long)ra * ld + 8 * nb + 2 * tig) = rz::pack_half2(c[nb][0] * mul_a, c[nb][1] * mul_a);
    if (rb < L) *reinterpret_cast<uint32_t*>(dst + (long long)rb * ld + 8 * nb + 2 * tig) = rz::pack_half2(c[nb][2] * mul_b, c[nb][3] * mul_b);
  }
}

// 16-byte asynchronous copy global -> shared; `valid` false zero-fills (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16(__half* dst, const __half* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// load_tile without waiting: the streamed tiles are double buffered under the MMAs of the previous tile
__device__ __forceinline__ void prefetch_tile(__half* tile, const __half* base, long long ld, int r0, int L) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = threadIdx.x + 128 * i, r = ch >> 3, c = (ch & 7) * 8;
    const bool in = r0 + r < L;
    cp_async16(tile + r * kLd + c, base + (long long)(in ? r0 + r : 0) * ld + c, in);
  }
}

// kernel 1: grid (q tiles, B * H), 4 warps x 16 query rows.  ONE pass over the keys: the softmax
// normaliser is not known yet, so dQ is accumulated against a running maximum exactly as the forward
// accumulates O,   dQ_i = (1 / l_i) sum_j e^{s_ij - m_i} (dP_ij - delta_i) k_j,
// rescaled when the maximum moves (delta_i = dO_i . o_i does not depend on the normaliser); the row's
// log-sum-exp m_i + ln l_i falls out at the end and is what kernel 2 reads.
__global__ void __launch_bounds__(128, RZ_ATTN_BWD_DQ_CTAS) attn_bwd_dq_kernel(const AttnBwd p) {
  __shared__ __align__(16) __half Ta[kT * kLd], Tb[kT * kLd], Tc[kT * kLd], Td[kT * kLd];
  __shared__ float delta_s[kT];
  const int tid = threadIdx.x, lane = tid & 31, w16 = (tid >> 5) * 16, g = lane >> 2, tig = lane & 3;
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H, q0 = blockIdx.x * kT, L = p.L;
  const long long ld = 3ll * p.H * 64, ldo = (long long)p.H * 64;
  const __half* qb = p.qkv + (long long)b * L * ld + h * 64;
  const __half* kb = qb + ldo;
  const __half* vb = qb + 2 * ldo;
  const __half* ob = p.o + (long long)b * L * ldo + h * 64;
  const __half* gb = p.dout + (long long)b * L * ldo + h * 64;

  load_tile(Tc, qb, ld, q0, L);
  load_tile(Td, gb, ldo, q0, L);
  load_tile(Ta, ob, ldo, q0, L);                       // the forward output, only for delta
  __syncthreads();
  {
    const int r = tid >> 1, c0 = (tid & 1) * 32;
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < 32; ++c) s += __half2float(Td[r * kLd + c0 + c]) * __half2float(Ta[r * kLd + c0 + c]);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if ((tid & 1) == 0) {
      delta_s[r] = s;
      p.delta[(long long)bh * p.Lp + q0 + r] = s;       // rows >= L: zero-filled tiles give 0
    }
  }
  uint32_t qf[4][4], gf[4][4];
  load_a_frags(Tc, w16, qf);
  load_a_frags(Td, w16, gf);
  __syncthreads();                                     // Q / dO / O tiles are free: all four tiles stream K, V
  const float d0 = delta_s[w16 + g], d1 = delta_s[w16 + g + 8];
  auto kbuf = [&](int i) { return i ? Tc : Ta; };
  auto vbuf = [&](int i) { return i ? Td : Tb; };
  const int nt = (L + kT - 1) / kT;
  prefetch_tile(kbuf(0), kb, ld, 0, L);
  prefetch_tile(vbuf(0), vb, ld, 0, L);
  cp_async_commit();

  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float acc[8][4];
  zero_acc(acc);
  for (int it = 0; it < nt; ++it) {
    const int kv0 = it * kT;
    if (it + 1 < nt) {
      prefetch_tile(kbuf((it + 1) & 1), kb, ld, kv0 + kT, L);
      prefetch_tile(vbuf((it + 1) & 1), vb, ld, kv0 + kT, L);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __half* Ks = kbuf(it & 1);
    const __half* Vs = vbuf(it & 1);
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {                   // two halves of 32 keys, each its own online-softmax step
      const int c0 = kv0 + 32 * hb;
      if (c0 >= L) break;                              // a half of nothing but padding (block-uniform)
      float s[4][4], dp[4][4];
      zero_half(s);
      zero_half(dp);
      gemm_nt_half(s, qf, Ks, hb);
      gemm_nt_half(dp, gf, Vs, hb);
      float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (c0 + 8 * nb + 2 * tig + (e & 1) >= L) s[nb][e] = -INFINITY;
          if (e < 2) x0 = fmaxf(x0, s[nb][e]); else x1 = fmaxf(x1, s[nb][e]);
        }
      x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
      x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
      const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);
      if (n0 != m0 || n1 != m1) {                      // warp-divergent only in the first few tiles
        const float r0 = __expf(m0 - n0), r1 = __expf(m1 - n1);
        l0 *= r0; l1 *= r1;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) { acc[nb][0] *= r0; acc[nb][1] *= r0; acc[nb][2] *= r1; acc[nb][3] *= r1; }
        m0 = n0; m1 = n1;
      }
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = __expf(s[nb][e] - (e < 2 ? m0 : m1));     // masked keys: e^{-inf} = 0
          if (e < 2) l0 += pv; else l1 += pv;
          s[nb][e] = pv * (dp[nb][e] - (e < 2 ? d0 : d1));
        }
      uint32_t af[2][4];
      acc_to_a_half(s, af);
      gemm_nn_half(acc, af, Ks, hb);
    }
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  if (tig == 0) {                                      // padded queries get +inf: e^{s - inf} = 0 drops them in kernel 2
    p.lse[(long long)bh * p.Lp + q0 + w16 + g] = q0 + w16 + g < L ? m0 + __logf(l0) : INFINITY;
    p.lse[(long long)bh * p.Lp + q0 + w16 + g + 8] = q0 + w16 + g + 8 < L ? m1 + __logf(l1) : INFINITY;
  }
  store_acc(p.dqkv + (long long)b * L * ld + h * 64, ld, q0, w16, L, acc, p.q_scale / l0, p.q_scale / l1);
}

// kernel 2: grid (key tiles, B * H), 4 warps x 16 key rows; everything is held transposed (keys are rows)
__global__ void __launch_bounds__(128, RZ_ATTN_BWD_DKV_CTAS) attn_bwd_dkv_kernel(const AttnBwd p) {
  __shared__ __align__(16) __half Ta[kT * kLd], Tb[kT * kLd], Tc[kT * kLd], Td[kT * kLd];
  __shared__ __align__(16) float lse_s[2][kT], delta_s[2][kT];
  const int tid = threadIdx.x, lane = tid & 31, w16 = (tid >> 5) * 16, tig = lane & 3;
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H, k0 = blockIdx.x * kT, L = p.L;
  const long long ld = 3ll * p.H * 64, ldo = (long long)p.H * 64;
  const __half* qb = p.qkv + (long long)b * L * ld + h * 64;
  const __half* kb = qb + ldo;
  const __half* vb = qb + 2 * ldo;
  const __half* gb = p.dout + (long long)b * L * ldo + h * 64;
  const float* lse = p.lse + (long long)bh * p.Lp;
  const float* delta = p.delta + (long long)bh * p.Lp;
  // 64 floats of lse (threads 0-15) and of delta (threads 16-31) ride in the tile's cp.async group
  auto prefetch_stats = [&](int buf, int q0) {
    if (tid < 32) {
      const float* src = (tid < 16 ? lse : delta) + q0 + (tid & 15) * 4;
      float* dst = (tid < 16 ? lse_s[buf] : delta_s[buf]) + (tid & 15) * 4;
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    }
  };

  load_tile(Tc, kb, ld, k0, L);
  load_tile(Td, vb, ld, k0, L);
  auto qbuf = [&](int i) { return i ? Tc : Ta; };
  auto gbuf = [&](int i) { return i ? Td : Tb; };
  const int nt = (L + kT - 1) / kT;
  prefetch_tile(qbuf(0), qb, ld, 0, L);
  prefetch_tile(gbuf(0), gb, ldo, 0, L);
  prefetch_stats(0, 0);
  cp_async_commit();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_a_frags(Tc, w16, kf);
  load_a_frags(Td, w16, vf);
  __syncthreads();                                     // K / V tiles are free: second buffer of the stream
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  for (int it = 0; it < nt; ++it) {
    const int q0 = it * kT, cur = it & 1;
    if (it + 1 < nt) {
      prefetch_tile(qbuf(cur ^ 1), qb, ld, q0 + kT, L);
      prefetch_tile(gbuf(cur ^ 1), gb, ldo, q0 + kT, L);
      prefetch_stats(cur ^ 1, q0 + kT);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __half* Qs = qbuf(cur);
    const __half* Gs = gbuf(cur);
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {                   // two halves of 32 queries
      if (q0 + 32 * hb >= L) break;                    // nothing but padded queries (block-uniform)
      float s[4][4], dp[4][4];
      zero_half(s);
      zero_half(dp);
      gemm_nt_half(s, kf, Qs, hb);                     // S^T = K Q^T
      gemm_nt_half(dp, vf, Gs, hb);                    // dP^T = V dO^T
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 32 * hb + 8 * nb + 2 * tig + (e & 1);   // query index inside the tile = column
          const float pv = __expf(s[nb][e] - lse_s[cur][c]);
          dp[nb][e] = pv * (dp[nb][e] - delta_s[cur][c]);       // dS^T
          s[nb][e] = pv;                                        // P^T
        }
      uint32_t af[2][4];
      acc_to_a_half(s, af);
      gemm_nn_half(dv, af, Gs, hb);                    // dV += P^T dO
      acc_to_a_half(dp, af);
      gemm_nn_half(dk, af, Qs, hb);                    // dK += dS^T Q
    }
    __syncthreads();
  }
  // q carries the folded 1/sqrt(64): dK = dS^T q_packed is already the gradient of the true key projection
  store_acc(p.dqkv + (long long)b * L * ld + ldo + h * 64, ld, k0, w16, L, dk, 1.f, 1.f);
  store_acc(p.dqkv + (long long)b * L * ld + 2 * ldo + h * 64, ld, k0, w16, L, dv, 1.f, 1.f);
}

inline bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

}  // namespace

extern "C" size_t rz_grad_scale_floats(void) { return kScaleFloats; }

extern "C" int rz_grad_scale(const float* grad, long long n, float* sc, void* stream) {
  if (!grad || !sc || n < 0) return RZ_ERR_INVALID;
  if (misaligned(grad) || misaligned(sc)) return RZ_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RZ_CUDA_OK(cudaMemsetAsync(sc, 0, 16, s));
  if (n > 0) {
    const int blocks = (int)std::min<long long>((n / 4 + 255) / 256 + 1, (long long)rz_sm_count() * 8);
    amax_kernel<<<blocks, 256, 0, s>>>(grad, n, reinterpret_cast<unsigned*>(sc) + 2);
    RZ_LAUNCH_OK();
  }
  // the maximum is accumulated in sc[2] (as bits) and replaced by its float value by scale_kernel
  scale_kernel<<<kScaleVec / 256, 256, 0, s>>>(reinterpret_cast<unsigned*>(sc) + 2, sc);
  RZ_LAUNCH_OK();
  rz_count_launch(2);
  return RZ_OK;
}

extern "C" int rz_ls_cast_bwd(const float* dy, const float* ls, const void* o_f16, const float* sc, long long rows,
                              void* do_f16, float* dls, void* stream) {
  if (!dy || !sc || !do_f16 || rows < 0) return RZ_ERR_INVALID;
  if (misaligned(dy) || misaligned(ls) || misaligned(o_f16) || misaligned(do_f16)) return RZ_ERR_ALIGNMENT;
  if (rows == 0) return RZ_OK;
  const int blocks = (int)std::min<long long>(rows, (long long)rz_sm_count() * 8);
  ls_cast_kernel<<<blocks, 192, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, ls, static_cast<const __half*>(o_f16), sc, rows, static_cast<__half*>(do_f16), dls);
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

extern "C" int rz_ls_weight_bwd(float* g, const float* w, const float* bias, float* colsum, const float* ls,
                               int n, int k, float* dls, void* stream) {
  if (!g || !w || !colsum || !ls || !dls || n <= 0 || k <= 0) return RZ_ERR_INVALID;
  if (k % 4 != 0) return RZ_ERR_UNSUPPORTED;
  if (misaligned(g) || misaligned(w)) return RZ_ERR_ALIGNMENT;
  ls_weight_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, w, bias, colsum, ls, k, dls);
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

extern "C" int rz_transpose_pad(const void* in_f16, long long rows, int cols, long long rows_padded, void* out_f16,
                                float* colsum, const float* sc, void* stream) {
  if (!in_f16 || rows < 0 || cols <= 0 || (!out_f16 && !colsum) || (colsum && !sc)) return RZ_ERR_INVALID;
  if (cols % 64 != 0 || rows_padded % 64 != 0 || rows_padded < rows) return RZ_ERR_UNSUPPORTED;
  if (misaligned(in_f16) || misaligned(out_f16)) return RZ_ERR_ALIGNMENT;
  if (rows_padded == 0) return RZ_OK;
  if (rows_padded / 64 >= (1ll << 31)) return RZ_ERR_UNSUPPORTED;
  const dim3 grid((unsigned)(rows_padded / 64), (unsigned)(cols / 64));
  transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(in_f16), rows, cols, rows_padded, static_cast<__half*>(out_f16), colsum, sc);
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

extern "C" int rz_gelu_bwd(const void* dg_f16, const void* u_f16, long long n, void* du_f16, void* stream) {
  if (!dg_f16 || !u_f16 || !du_f16 || n < 0) return RZ_ERR_INVALID;
  if (n % 8 != 0) return RZ_ERR_UNSUPPORTED;
  if (misaligned(dg_f16) || misaligned(u_f16) || misaligned(du_f16)) return RZ_ERR_ALIGNMENT;
  if (n == 0) return RZ_OK;
  const int blocks = (int)std::min<long long>((n / 8 + 255) / 256, (long long)rz_sm_count() * 16);
  gelu_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dg_f16), static_cast<const __half*>(u_f16), n / 8, static_cast<__half*>(du_f16));
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

extern "C" int rz_ln_rows_bwd(const float* x, const void* dh_f16, const float* gamma, float eps, const float* dres,
                              const float* sc, long long rows, float* dx, void* dx_f16, float* dgamma, float* dbeta,
                              void* stream) {
  if (!x || !dh_f16 || !gamma || !sc || !dx || rows < 0) return RZ_ERR_INVALID;
  if (misaligned(x) || misaligned(dh_f16) || misaligned(gamma) || misaligned(dres) || misaligned(dx) ||
      misaligned(dx_f16))
    return RZ_ERR_ALIGNMENT;
  if (rows == 0) return RZ_OK;
  const int blocks = (int)std::min<long long>((rows + kLnWarps - 1) / kLnWarps, (long long)rz_sm_count() * 2);
  ln_bwd_kernel<<<blocks, 32 * kLnWarps, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<const __half*>(dh_f16), gamma, eps, dres, sc, rows, dx, static_cast<__half*>(dx_f16), dgamma, dbeta);
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

extern "C" int rz_attention_bwd(const void* qkv_f16, const void* out_f16, const void* dout_f16, int n_images,
                                int tokens, int heads, float q_scale, float* lse, float* delta, void* dqkv_f16,
                                void* stream) {
  if (!qkv_f16 || !out_f16 || !dout_f16 || !lse || !delta || !dqkv_f16 || n_images < 0 || tokens < 0 || heads <= 0)
    return RZ_ERR_INVALID;
  if (misaligned(qkv_f16) || misaligned(out_f16) || misaligned(dout_f16) || misaligned(dqkv_f16) ||
      misaligned(lse) || misaligned(delta))
    return RZ_ERR_ALIGNMENT;
  if (n_images == 0 || tokens == 0) return RZ_OK;
  if ((long long)n_images * heads > 65535) return RZ_ERR_UNSUPPORTED;
  AttnBwd p;
  p.qkv = static_cast<const __half*>(qkv_f16); p.o = static_cast<const __half*>(out_f16);
  p.dout = static_cast<const __half*>(dout_f16); p.dqkv = static_cast<__half*>(dqkv_f16);
  p.lse = lse; p.delta = delta; p.B = n_images; p.L = tokens; p.H = heads; p.q_scale = q_scale;
  p.Lp = (tokens + kT - 1) / kT * kT;
  const dim3 grid((unsigned)((tokens + kT - 1) / kT), (unsigned)(n_images * heads));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  attn_bwd_dq_kernel<<<grid, 128, 0, s>>>(p);
  RZ_LAUNCH_OK();
  attn_bwd_dkv_kernel<<<grid, 128, 0, s>>>(p);
  RZ_LAUNCH_OK();
  rz_count_launch(2);
  return RZ_OK;
}
