// A1: the linear layers of the AlignTransformer (two DINOv2 encoder layers in front of the VL-CABS
// path, exp/cxr_pt/model/align_transformers.py:23-45 -> transformers Dinov2Layer) as policies of the
// persistent tcgen05 GEMM skeleton (rz_gemm.cuh):
//
//   out = epilogue(A[M, K] . W[N, K]^T + bias)        A, W fp16 (K-major, TMA SWIZZLE_128B), fp32 accumulate
//
//   RZ_LIN_BIAS      out fp16 [M, N] = acc + bias                       (fused q/k/v projection)
//   RZ_LIN_GELU      out fp16 [M, N] = gelu_erf(acc + bias)             (mlp.fc1 + activation)
//   RZ_LIN_RESIDUAL  out fp32 [M, N] = residual + scale * (acc + bias)  (attention.output.dense / mlp.fc2
//                                       + LayerScale + the residual add; out may alias residual)
//
// Tiles are 128 x 256; a CTA walks the N tiles of one 128-row block back to back (the A block comes
// from HBM once, the weights stay in L2), and two CTAs of a cluster pair (cta_group::2) take row
// blocks (2i, 2i+1) so that each loads half of the weight tile.  Rows >= M are zero-filled by the TMA
// loads and clipped by the TMA stores.
#include "rz_gemm.cuh"

namespace {

using namespace rz::gemm;

struct LinParams {
  int M, N, K, m_tiles, n_tiles;
  const float* bias;        // [N] or NULL
  const float* scale;       // [N] or NULL (RZ_LIN_RESIDUAL)
  const float* residual;    // [M, N] fp32 (RZ_LIN_RESIDUAL)
};

// erf-GELU (torch.nn.functional.gelu default, Dinov2MLP hidden_act = "gelu")
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int EPI, int C>
struct Lin : PolicyBase {
  static constexpr int kCluster = C;
  using Params = LinParams;
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = false, kBMn = false, kTwoPhase = false;
  // per warp: fp16 outputs one [32 rows x 64 cols] box; fp32 outputs two [32 rows x 32 cols] boxes
  static constexpr int kWarpStage = EPI == RZ_LIN_RESIDUAL ? 8192 : 4096;
  static constexpr int kEpiSmem = 4 * kWarpStage;
  __host__ __device__ static int num_tiles(const Params& p) { return p.m_tiles * p.n_tiles; }
  __host__ __device__ static int inner(const Params& p) { return p.n_tiles; }
  __host__ __device__ static int k_steps(const Params& p) { return p.K / kBK; }
  __device__ static void load(const Params& p, const Maps& m, int tile, int ks, uint8_t* a, uint8_t*,
                              uint8_t* bsm, uint64_t* bar, int rank) {
    const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
    load_kmajor<C>(&m.a, bar, a, ks * kBK, mt * kBM, 0);                            // A [M, K]
    load_kmajor_shared<C>(&m.b, bar, bsm, ks * kBK, nt * kBN, 0, kBN, rank);        // W [N, K]
  }

  __device__ static __forceinline__ void chunk_f16(const Params& p, const Maps& maps, Cols64& v, int col0,
                                                   int row0, int lane, uint32_t stg) {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t* src = j < 4 ? &v.lo[8 * j] : &v.hi[8 * (j - 4)];
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(src[i]);
      if (p.bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * j));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * j + 4));
        x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
        x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
      }
      if (EPI == RZ_LIN_GELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = gelu_erf(x[i]);
      }
      sts_v4(stg + stage_off(lane, j), pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]),
             pack_h2(x[6], x[7]));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&maps.c, stg, col0, row0, 0);
      tma_store_commit();
    }
  }

  // one 32-column half of a chunk: out = residual + scale * (acc + bias), fp32
  __device__ static __forceinline__ void half_res(const Params& p, const Maps& maps, const uint32_t (&a)[32],
                                                  int col0, int row0, int lane, bool row_ok,
                                                  const float* rrow, uint32_t stg) {
    float4 r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      r[j] = row_ok ? __ldcs(reinterpret_cast<const float4*>(rrow + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane == 0) tma_store_wait_read_n<1>();       // the box written two stores ago is free again
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f), s = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
      if (p.scale != nullptr) s = __ldg(reinterpret_cast<const float4*>(p.scale + col0) + j);
      const float o0 = fmaf(s.x, __uint_as_float(a[4 * j]) + b.x, r[j].x);
      const float o1 = fmaf(s.y, __uint_as_float(a[4 * j + 1]) + b.y, r[j].y);
      const float o2 = fmaf(s.z, __uint_as_float(a[4 * j + 2]) + b.z, r[j].z);
      const float o3 = fmaf(s.w, __uint_as_float(a[4 * j + 3]) + b.w, r[j].w);
      sts_v4(stg + stage_off(lane, j), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2),
             __float_as_uint(o3));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&maps.c, stg, col0, row0, 0);
      tma_store_commit();
    }
  }

  __device__ static void epilogue(const Params& p, const Maps& maps, int tile, int, uint32_t tmem, int warp,
                                  int lane, uint64_t*, State&, uint8_t* epi_smem) {
    const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
    const int row0 = mt * kBM + warp * 32;
    const int row = row0 + lane;
    const bool row_ok = row < p.M;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t stg = smem_u32(epi_smem) + warp * kWarpStage;
    const float* rrow = EPI == RZ_LIN_RESIDUAL ? p.residual + (long long)(row_ok ? row : 0) * p.N : nullptr;
    Cols64 va, vb;
    ld64(taddr, va);
#pragma unroll 1
    for (int c = 0; c < kBN / 64; c += 2) {
      const int col = nt * kBN + c * 64;
      wait64(va);
      ld64(taddr + (c + 1) * 64, vb);
      if (EPI == RZ_LIN_RESIDUAL) {
        half_res(p, maps, va.lo, col, row0, lane, row_ok, rrow, stg);
        half_res(p, maps, va.hi, col + 32, row0, lane, row_ok, rrow, stg + 4096);
      } else {
        chunk_f16(p, maps, va, col, row0, lane, stg);
      }
      wait64(vb);
      if (c + 2 < kBN / 64) ld64(taddr + (c + 2) * 64, va);
      if (EPI == RZ_LIN_RESIDUAL) {
        half_res(p, maps, vb.lo, col + 64, row0, lane, row_ok, rrow, stg);
        half_res(p, maps, vb.hi, col + 96, row0, lane, row_ok, rrow, stg + 4096);
      } else {
        chunk_f16(p, maps, vb, col + 64, row0, lane, stg);
      }
    }
  }
};

template <int EPI>
int launch_lin(const Maps& m, const LinParams& p, bool pair, cudaStream_t s) {
  return pair ? launch<Lin<EPI, 2>>(m, p, s) : launch<Lin<EPI, 1>>(m, p, s);
}

}  // namespace

extern "C" int rz_linear(const void* a_f16, long long m, int k, const void* w_f16, int n,
                         const float* bias, int epilogue, const float* scale, const float* residual,
                         void* out, void* stream) {
  if (!a_f16 || !w_f16 || !out || m < 0 || k <= 0 || n <= 0) return RZ_ERR_INVALID;
  if (k % kBK != 0 || n % 256 != 0) return RZ_ERR_UNSUPPORTED;
  if (epilogue != RZ_LIN_BIAS && epilogue != RZ_LIN_GELU && epilogue != RZ_LIN_RESIDUAL) return RZ_ERR_INVALID;
  if (epilogue == RZ_LIN_RESIDUAL && residual == nullptr) return RZ_ERR_INVALID;
  if (m == 0) return RZ_OK;
  if (m >= (1ll << 31) - 256) return RZ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a_f16) & 15) || (reinterpret_cast<uintptr_t>(w_f16) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15) ||
      (reinterpret_cast<uintptr_t>(scale) & 15) || (reinterpret_cast<uintptr_t>(residual) & 15))
    return RZ_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LinParams p;
  p.M = (int)m; p.N = n; p.K = k;
  p.n_tiles = n / 256;
  const int mt = (int)((m + kBM - 1) / kBM);
  const bool pair = mt >= 2;
  p.m_tiles = pair ? (mt + 1) / 2 * 2 : mt;       // an odd last row block gets an all-padding partner
  p.bias = bias; p.scale = scale; p.residual = residual;
  Maps mp = {};
  if (!rz::make_map_3d_sw128(&mp.a, a_f16, 1, (uint64_t)m, (uint64_t)k, (uint64_t)k * 2, (uint64_t)m * k * 2, kBM))
    return RZ_ERR_CUDA;
  if (!rz::make_map_3d_sw128(&mp.b, w_f16, 1, (uint64_t)n, (uint64_t)k, (uint64_t)k * 2, (uint64_t)n * k * 2,
                             pair ? 128 : 256))
    return RZ_ERR_CUDA;
  mp.a2 = mp.a; mp.b2 = mp.b;
  if (epilogue == RZ_LIN_RESIDUAL) {
    if (!rz::make_map_3d_f32_sw128(&mp.c, out, 1, (uint64_t)m, (uint64_t)n, (uint64_t)n * 4, (uint64_t)m * n * 4, 32))
      return RZ_ERR_CUDA;
  } else {
    if (!rz::make_map_3d_sw128(&mp.c, out, 1, (uint64_t)m, (uint64_t)n, (uint64_t)n * 2, (uint64_t)m * n * 2, 32))
      return RZ_ERR_CUDA;
  }
  mp.c2 = mp.c;
  switch (epilogue) {
    case RZ_LIN_BIAS: return launch_lin<RZ_LIN_BIAS>(mp, p, pair, s);
    case RZ_LIN_GELU: return launch_lin<RZ_LIN_GELU>(mp, p, pair, s);
    default: return launch_lin<RZ_LIN_RESIDUAL>(mp, p, pair, s);
  }
}
