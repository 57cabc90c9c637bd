// K1+K2: row LayerNorm + L2 normalisation, fp16 / fp32 outputs and backward statistics.
// HBM-bound: one warp per 768-wide row, 128-bit streaming loads, two-pass variance in
// registers, warp-shuffle reductions, packed fp16 stores.  Replaces nn.LayerNorm
// (exp/cxr_pt/model/losses.py:90-91,163-164) + F.normalize (losses.py:212-213).
#include "rz_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
prep_rows_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, long long rows, int rows_per_group,
                 int rows_per_group_padded, __half* __restrict__ out_h, float* __restrict__ out_f,
                 float* __restrict__ stats, int l2, long long padded_rows_total, float eps_ln) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  // iterate over PADDED row slots so that padding rows get zero-filled by the same grid
  for (long long slot = warp0; slot < padded_rows_total; slot += nwarps) {
    const unsigned g = (unsigned)slot / (unsigned)rows_per_group_padded;   // slots < 2^31 (checked on host)
    const int r = (int)((unsigned)slot - g * (unsigned)rows_per_group_padded);
    if (r >= rows_per_group) {
      if (out_h != nullptr) {
        uint2* o = reinterpret_cast<uint2*>(out_h + slot * RZ_HIDDEN) + lane;
#pragma unroll
        for (int j = 0; j < 6; ++j) o[32 * j] = make_uint2(0u, 0u);
      }
      continue;
    }
    const long long row = (long long)g * rows_per_group + r;
    if (row >= rows) continue;
    float v[24];
    rz::RowLoad<T>::load(x + row * RZ_HIDDEN, lane, v);
    const rz::RowStats st = rz::ln_l2_row(v, gamma, beta, lane, eps_ln, RZ_L2_EPS, l2 != 0);
    if (out_h != nullptr) {
      uint2* o = reinterpret_cast<uint2*>(out_h + slot * RZ_HIDDEN) + lane;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        o[32 * j] = make_uint2(rz::pack_half2(v[4 * j], v[4 * j + 1]),
                               rz::pack_half2(v[4 * j + 2], v[4 * j + 3]));
    }
    if (out_f != nullptr) {
      float4* o = reinterpret_cast<float4*>(out_f + row * RZ_HIDDEN) + lane;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        o[32 * j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    if (stats != nullptr && lane == 0) {
      stats[row * 3 + 0] = st.mean;
      stats[row * 3 + 1] = st.rstd;
      stats[row * 3 + 2] = st.inv_norm;
    }
  }
}

// T0: masked mean pooling of the text encoder's token embeddings (one warp per sentence; masked-out
// tokens are never read) fused with the path's LayerNorm + L2 normalisation of the pooled vector.
//   feats[i] = sum_t m[i,t] h[i,t,:] / max(sum_t m[i,t], 1e-9)          modeling.py:147-156
//   q16[i]   = L2(LN(feats[i]))  as fp16                                  losses.py:163-164, 212-213
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
text_pool_kernel(const T* __restrict__ hidden, const long long* __restrict__ mask, int n, int tokens,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int l2,
                 float* __restrict__ feats, __half* __restrict__ q16) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= n) return;
  float acc[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) acc[k] = 0.f;
  float msum = 0.f;
  for (int t = 0; t < tokens; ++t) {
    const float m = (float)__ldg(mask + (long long)i * tokens + t);
    if (m == 0.f) continue;                                  // warp-uniform: padding tokens cost nothing
    float v[24];
    rz::RowLoad<T>::load(hidden + ((long long)i * tokens + t) * RZ_HIDDEN, lane, v);
#pragma unroll
    for (int k = 0; k < 24; ++k) acc[k] = fmaf(m, v[k], acc[k]);
    msum += m;
  }
  const float inv = 1.0f / fmaxf(msum, 1e-9f);
#pragma unroll
  for (int k = 0; k < 24; ++k) acc[k] *= inv;
  if (feats != nullptr) {
    float4* o = reinterpret_cast<float4*>(feats + (long long)i * RZ_HIDDEN) + lane;
#pragma unroll
    for (int j = 0; j < 6; ++j) o[32 * j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  }
  if (q16 != nullptr) {
    rz::ln_l2_row(acc, gamma, beta, lane, RZ_LN_EPS, RZ_L2_EPS, l2 != 0);
    uint2* o = reinterpret_cast<uint2*>(q16 + (long long)i * RZ_HIDDEN) + lane;
#pragma unroll
    for (int j = 0; j < 6; ++j)
      o[32 * j] = make_uint2(rz::pack_half2(acc[4 * j], acc[4 * j + 1]), rz::pack_half2(acc[4 * j + 2], acc[4 * j + 3]));
  }
}

}  // namespace

static int prep_rows_impl(const void* x, int dtype, const float* gamma, const float* beta,
                          long long rows, int rows_per_group, int rows_per_group_padded,
                          void* out_f16, float* out_f32, float* stats, int l2, float eps_ln,
                          void* stream) {
  if (x == nullptr || rows < 0 || rows_per_group <= 0 || rows_per_group_padded < rows_per_group)
    return RZ_ERR_INVALID;
  if ((gamma == nullptr) != (beta == nullptr)) return RZ_ERR_INVALID;
  if (rows == 0) return RZ_OK;
  if (rows % rows_per_group != 0) return RZ_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out_f16) & 15) ||
      (reinterpret_cast<uintptr_t>(out_f32) & 15) || (reinterpret_cast<uintptr_t>(gamma) & 15) ||
      (reinterpret_cast<uintptr_t>(beta) & 15))
    return RZ_ERR_ALIGNMENT;
  const long long groups = rows / rows_per_group;
  const long long slots = groups * rows_per_group_padded;
  if (slots >= (1ll << 31)) return RZ_ERR_UNSUPPORTED;
  long long blocks = (slots + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)rz_sm_count() * 3;  // 3 resident CTAs per SM (launch bounds)
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)blocks), block(kWarpsPerBlock * 32);
  __half* oh = static_cast<__half*>(out_f16);
  switch (dtype) {
    case RZ_F32:
      prep_rows_kernel<float><<<grid, block, 0, s>>>(static_cast<const float*>(x), gamma, beta, rows,
                                                     rows_per_group, rows_per_group_padded, oh,
                                                     out_f32, stats, l2, slots, eps_ln);
      break;
    case RZ_BF16:
      prep_rows_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(
          static_cast<const __nv_bfloat16*>(x), gamma, beta, rows, rows_per_group,
          rows_per_group_padded, oh, out_f32, stats, l2, slots, eps_ln);
      break;
    case RZ_F16:
      prep_rows_kernel<__half><<<grid, block, 0, s>>>(static_cast<const __half*>(x), gamma, beta,
                                                      rows, rows_per_group, rows_per_group_padded,
                                                      oh, out_f32, stats, l2, slots, eps_ln);
      break;
    default:
      return RZ_ERR_INVALID;
  }
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

extern "C" int rz_prep_rows(const void* x, int dtype, const float* gamma, const float* beta,
                            long long rows, int rows_per_group, int rows_per_group_padded,
                            void* out_f16, float* out_f32, float* stats, int l2, void* stream) {
  return prep_rows_impl(x, dtype, gamma, beta, rows, rows_per_group, rows_per_group_padded, out_f16,
                        out_f32, stats, l2, RZ_LN_EPS, stream);
}

// A0: plain LayerNorm with the caller's eps -> fp16 GEMM operand rows (Dinov2Layer.norm1 / norm2,
// layer_norm_eps = 1e-6)
extern "C" int rz_ln_rows(const void* x, int dtype, const float* gamma, const float* beta, float eps,
                          long long rows, void* out_f16, void* stream) {
  if (gamma == nullptr || beta == nullptr || out_f16 == nullptr || !(eps > 0.f)) return RZ_ERR_INVALID;
  if (rows == 0) return RZ_OK;
  const int rpg = rows < (1 << 30) ? (int)rows : 0;
  if (rpg == 0) return RZ_ERR_UNSUPPORTED;
  return prep_rows_impl(x, dtype, gamma, beta, rows, rpg, rpg, out_f16, nullptr, nullptr, 0, eps, stream);
}

extern "C" int rz_text_pool(const void* hidden, int dtype, const long long* attention_mask, int n_sentences,
                            int tokens, const float* gamma, const float* beta, int l2, float* feats_f32,
                            void* q_f16, void* stream) {
  if (hidden == nullptr || attention_mask == nullptr || n_sentences < 0 || tokens <= 0) return RZ_ERR_INVALID;
  if ((gamma == nullptr) != (beta == nullptr)) return RZ_ERR_INVALID;
  if (feats_f32 == nullptr && q_f16 == nullptr) return RZ_ERR_INVALID;
  if (n_sentences == 0) return RZ_OK;
  if ((reinterpret_cast<uintptr_t>(hidden) & 15) || (reinterpret_cast<uintptr_t>(feats_f32) & 15) ||
      (reinterpret_cast<uintptr_t>(q_f16) & 15) || (reinterpret_cast<uintptr_t>(gamma) & 15) ||
      (reinterpret_cast<uintptr_t>(beta) & 15) || (reinterpret_cast<uintptr_t>(attention_mask) & 7))
    return RZ_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((n_sentences + kWarpsPerBlock - 1) / kWarpsPerBlock)), block(kWarpsPerBlock * 32);
  __half* q = static_cast<__half*>(q_f16);
  switch (dtype) {
    case RZ_F32:
      text_pool_kernel<float><<<grid, block, 0, s>>>(static_cast<const float*>(hidden), attention_mask, n_sentences,
                                                     tokens, gamma, beta, l2, feats_f32, q);
      break;
    case RZ_BF16:
      text_pool_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(static_cast<const __nv_bfloat16*>(hidden), attention_mask,
                                                             n_sentences, tokens, gamma, beta, l2, feats_f32, q);
      break;
    case RZ_F16:
      text_pool_kernel<__half><<<grid, block, 0, s>>>(static_cast<const __half*>(hidden), attention_mask, n_sentences,
                                                      tokens, gamma, beta, l2, feats_f32, q);
      break;
    default:
      return RZ_ERR_INVALID;
  }
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

// ---- group_map without a host-to-device copy ---------------------------------------------------
// group_map[j] = global image index of sentence j (losses.py:131-151).  The per-image sentence counts
// live on the host; they travel to the device as KERNEL PARAMETERS (up to 1024 per launch), not through
// the copy engine: a 48 KB cudaMemcpyAsync on the compute stream queues behind whatever bulk upload a
// data loader has in flight on that engine (measured: 27 ms per step behind a 2 GB prefetch).
namespace {
struct CountsParam { unsigned short c[1024]; };
__global__ void __launch_bounds__(64)
group_map_kernel(const __grid_constant__ CountsParam cp, int n_images, long long first_image,
                 long long base, long long* __restrict__ out) {
  const int i = blockIdx.x;
  if (i >= n_images) return;
  long long off = base;
  for (int k = 0; k < i; ++k) off += cp.c[k];
  const int n = cp.c[i];
  for (int j = threadIdx.x; j < n; j += 64) out[off + j] = first_image + i;
}
}  // namespace

extern "C" int rz_group_map(const int* counts_host, int n_images, long long first_image, long long* out,
                            void* stream) {
  if (n_images < 0 || (n_images > 0 && (!counts_host || !out))) return RZ_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long base = 0;
  for (int i0 = 0; i0 < n_images; i0 += 1024) {
    const int n = n_images - i0 < 1024 ? n_images - i0 : 1024;
    CountsParam cp;
    long long sum = 0;
    for (int k = 0; k < n; ++k) {
      const int c = counts_host[i0 + k];
      if (c < 0 || c > 65535) return RZ_ERR_UNSUPPORTED;
      cp.c[k] = (unsigned short)c;
      sum += c;
    }
    for (int k = n; k < 1024; ++k) cp.c[k] = 0;
    group_map_kernel<<<n, 64, 0, s>>>(cp, n, first_image + i0, base, out);
    RZ_LAUNCH_OK();
    rz_count_launch();
    base += sum;
  }
  return RZ_OK;
}
