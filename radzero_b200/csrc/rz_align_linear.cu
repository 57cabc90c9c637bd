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
//   RZ_LIN_RESIDUAL_F16  the same value, written ONLY as fp16 [M, N]: the hand-off of the last layer's
//                                       tokens to the similarity kernel at half the bytes
//
// Tiles are 128 x 256; the N tiles of one 128-row block run on neighbouring CTAs at the same time (the
// A block comes from HBM once, the weights stay in L2), and two CTAs of a cluster pair
// (cta_group::2) take row blocks (2i, 2i+1) so that each loads half of the weight tile.  Rows >= M are zero-filled by the TMA
// loads and clipped by the TMA stores.
#include "rz_gemm.cuh"

namespace {

using namespace rz::gemm;

struct LinParams {
  int M, N, K, m_tiles, n_tiles;
  const float* bias;        // [N] or NULL
  const float* scale;       // [N] or NULL (RZ_LIN_RESIDUAL)
  const float* residual;    // [M, N] fp32 (RZ_LIN_RESIDUAL; read through maps.c2)
};

// packed fp32 pairs (sm_100 FFMA2 / FMUL2): the epilogues are bound by the FP32 instruction count of
// their four warps, so everything that can is done two columns per instruction
struct F2 { uint64_t v; };
__device__ __forceinline__ F2 f2(float a, float b) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ F2 f2(float a) { return f2(a, a); }
__device__ __forceinline__ void unpack(F2 x, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
  F2 r;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
  F2 r;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// erf-GELU (torch.nn.functional.gelu default, Dinov2MLP hidden_act = "gelu") of two columns:
//   gelu(x) = relu(x) - 0.5 |x| erfc(|x| / sqrt 2),   erfc(z) = poly5(t) exp(-z^2),  t = 1 / (1 + p z)
// (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7: far below the fp16 output ulp); 8 packed FP32 + 4
// scalar + 4 MUFU instructions per pair.  libdevice erff costs ~3x as many and made the fc1 epilogue
// several times slower than its MMAs.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const float kS = 0.70710678118654752f;
  const float z0 = fabsf(x0) * kS, z1 = fabsf(x1) * kS;
  const F2 z = f2(z0, z1), x = f2(x0, x1);
  float d0, d1;
  unpack(fma2(z, f2(0.3275911f), f2(1.0f)), d0, d1);
  const F2 t = f2(rcp_approx(d0), rcp_approx(d1));
  F2 q = fma2(t, f2(1.061405429f), f2(-1.453152027f));
  q = fma2(q, t, f2(1.421413741f));
  q = fma2(q, t, f2(-0.284496736f));
  q = fma2(q, t, f2(0.254829592f));
  q = mul2(q, t);
  float u0, u1;
  unpack(mul2(mul2(x, x), f2(-0.72134752044448170f)), u0, u1);   // exp(-z^2) = 2^(-x^2 / 2 * log2 e)
  const F2 e = f2(exp2f(u0), exp2f(u1));
  const F2 w = mul2(mul2(q, e), mul2(z, f2(-kS)));                // -0.5 |x| erfc
  unpack(add2(f2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)), w), x0, x1);
}

template <int EPI, int C>
struct Lin : PolicyBase {
  static constexpr int kCluster = C;
  using Params = LinParams;
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = false, kBMn = false, kTwoPhase = false;
  // per warp: fp16 outputs two alternating [32 rows x 64 cols] boxes; the residual epilogue a ring of four
  // [32 rows x 32 cols] fp32 boxes (residual in by TMA, result out by TMA from the same box) plus
  // one mbarrier per box
  static constexpr bool kRes = EPI == RZ_LIN_RESIDUAL || EPI == RZ_LIN_RESIDUAL_F16;
  static constexpr bool kF16Out = EPI == RZ_LIN_RESIDUAL_F16;
  // (fp16 hand-off: the four fp32 residual boxes + two alternating [32 rows x 64 cols] fp16 output boxes)
  static constexpr int kWarpStage = kF16Out ? 24576 : (kRes ? 16384 : 8192);
  // the GELU epilogue is bound by the instruction issue of its warps: eight epilogue warps, two per
  // TMEM lane quarter, each taking one 128-column half of the tile (fc1 1.51 -> 1.41 ms); the plain
  // bias epilogue is not, and the extra warps cost it tensor-pipe time (90 % -> 81 % active)
  static constexpr int kEpiWarps = EPI == RZ_LIN_GELU ? 8 : 4;
  static constexpr int kEpiSmem = kEpiWarps * kWarpStage + (kRes ? 128 : 0);
  struct State { uint32_t g; int ready; };     // g: 32-column boxes this warp has consumed so far
  __host__ __device__ static int num_tiles(const Params& p) { return p.m_tiles * p.n_tiles; }
  __host__ __device__ static int k_steps(const Params& p) { return p.K / kBK; }
  // Tile order: the N tiles of one row block (pair of row blocks) are consecutive ITEMS, i.e. they run
  // on neighbouring CTAs at the same time, so the A block is fetched from HBM once and hit in L2 by
  // the others while it is hot.  (Walking them back to back on one CTA instead re-read the 786 KB
  // A blocks of fc2 from HBM: 148 of them in flight exceed L2.)
  __device__ static void decode(const Params& p, int tile, int& mt, int& nt) {
    const int item = tile / C, rank = tile % C;
    nt = item % p.n_tiles;
    mt = (item / p.n_tiles) * C + rank;
  }
  __device__ static void load(const Params& p, const Maps& m, int tile, int ks, uint8_t* a, uint8_t*,
                              uint8_t* bsm, uint64_t* bar, int rank) {
    int mt, nt;
    decode(p, tile, mt, nt);
    load_kmajor<C>(&m.a, bar, a, ks * kBK, mt * kBM, 0);                            // A [M, K]
    load_kmajor_shared<C>(&m.b, bar, bsm, ks * kBK, nt * kBN, 0, kBN, rank);        // W [N, K]
  }

  // one 64-column chunk -> fp16.  The arithmetic runs BEFORE the wait for the staging box (two boxes
  // per warp, alternating), so the previous chunk's TMA store drains behind it.
  __device__ static __forceinline__ void chunk_f16(const Params& p, const Maps& maps, Cols64& v, int col0,
                                                   int row0, int lane, uint32_t stg) {
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t* src = j < 4 ? &v.lo[8 * j] : &v.hi[8 * (j - 4)];
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(src[i]);
      if (p.bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * j));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * j + 4));
        unpack(add2(f2(x[0], x[1]), f2(b0.x, b0.y)), x[0], x[1]);
        unpack(add2(f2(x[2], x[3]), f2(b0.z, b0.w)), x[2], x[3]);
        unpack(add2(f2(x[4], x[5]), f2(b1.x, b1.y)), x[4], x[5]);
        unpack(add2(f2(x[6], x[7]), f2(b1.z, b1.w)), x[6], x[7]);
      }
      if (EPI == RZ_LIN_GELU) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) gelu_erf2(x[i], x[i + 1]);
      }
      pk[4 * j] = pack_h2(x[0], x[1]); pk[4 * j + 1] = pack_h2(x[2], x[3]);
      pk[4 * j + 2] = pack_h2(x[4], x[5]); pk[4 * j + 3] = pack_h2(x[6], x[7]);
    }
    if (lane == 0) tma_store_wait_read_n<1>();       // the box written two stores ago is free again
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_v4(stg + stage_off(lane, j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&maps.c, stg, col0, row0, 0);
      tma_store_commit();
    }
  }

  // ---- residual epilogue -------------------------------------------------------------------
  // The residual tile is streamed through shared memory by TMA, two boxes ahead of the arithmetic
  // (a thread-per-row global load would touch 32 different 128-byte lines per instruction and
  // leave the DRAM latency exposed once per box).  Box i of a warp's ring: TMA load (residual) ->
  // in-place update by the 32 lanes -> TMA store (result).  Coordinates of running box index g are
  // those of (tile, g % 8) -- a warp consumes exactly 8 boxes per tile.
  __device__ static __forceinline__ uint64_t* res_bar(uint8_t* epi_smem, int warp, int box) {
    return reinterpret_cast<uint64_t*>(epi_smem + 4 * kWarpStage) + warp * 4 + box;
  }
  __device__ static __forceinline__ void res_load(const Params& p, const Maps& maps, int tile, int h, uint32_t g,
                                                  int warp, uint8_t* epi_smem) {
    int mt, nt;
    decode(p, tile, mt, nt);
    const int box = (int)(g & 3u);
    uint64_t* bar = res_bar(epi_smem, warp, box);
    mbar_arrive_expect_tx(bar, 4096u);
    tma_load_3d(&maps.c2, bar, epi_smem + warp * kWarpStage + box * 4096, nt * kBN + h * 32,
                mt * kBM + warp * 32, 0, kEvictFirst);
  }
  template <class S>
  __device__ static void prologue(const Params& p, const Maps& maps, int tile, int warp, int lane, uint64_t*,
                                  S& st, uint8_t* epi_smem) {
    if (!kRes || st.ready) return;
    st.ready = 1;
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) mbar_init(res_bar(epi_smem, warp, i), 1);
      fence_barrier_init();
      res_load(p, maps, tile, 0, st.g, warp, epi_smem);
      res_load(p, maps, tile, 1, st.g + 1, warp, epi_smem);
    }
    __syncwarp();
  }
  __device__ static __forceinline__ void half_res(const Params& p, const Maps& maps, const uint32_t (&a)[32],
                                                  int tile, int next_tile, int h, int col0, int row0, int warp,
                                                  int lane, State& st, uint8_t* epi_smem) {
    const uint32_t g = st.g;
    if (lane == 0) {
      // prefetch two boxes ahead; its box was last used by the store of box g - 2
      const int h2 = h + 2;
      const int t2 = h2 < 8 ? tile : next_tile;
      if (t2 >= 0) {
        // fp32 mode: the box was last the source of the store of box g - 2.  fp16 hand-off: nothing is
        // stored from the residual boxes; box g - 2 was only read, two halves of arithmetic ago
        if (!kF16Out) tma_store_wait_read_n<1>();
        res_load(p, maps, t2, h2 & 7, g + 2, warp, epi_smem);
      }
    }
    const int box = (int)(g & 3u);
    const uint32_t stg = smem_u32(epi_smem) + warp * kWarpStage + box * 4096;
    mbar_wait(res_bar(epi_smem, warp, box), (g >> 2) & 1u);
    // fp16 hand-off: two halves (32 columns each) fill one [32 rows x 64 cols] fp16 box, two boxes alternate
    const uint32_t out16 = smem_u32(epi_smem) + warp * kWarpStage + 16384 + ((g >> 1) & 1u) * 4096;
    if (kF16Out && (h & 1) == 0) {
      if (lane == 0) tma_store_wait_read_n<1>();     // the store issued from this box two pairs ago
      __syncwarp();
    }
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 r;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(stg + stage_off(lane, j)));
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f), s = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
      if (p.scale != nullptr) s = __ldg(reinterpret_cast<const float4*>(p.scale + col0) + j);
      const float o0 = fmaf(s.x, __uint_as_float(a[4 * j]) + b.x, r.x);
      const float o1 = fmaf(s.y, __uint_as_float(a[4 * j + 1]) + b.y, r.y);
      const float o2 = fmaf(s.z, __uint_as_float(a[4 * j + 2]) + b.z, r.z);
      const float o3 = fmaf(s.w, __uint_as_float(a[4 * j + 3]) + b.w, r.w);
      if (kF16Out) {
        pk[2 * j] = pack_h2(o0, o1);
        pk[2 * j + 1] = pack_h2(o2, o3);
      } else {
        sts_v4(stg + stage_off(lane, j), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2),
               __float_as_uint(o3));
      }
    }
    if (kF16Out) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        sts_v4(out16 + stage_off(lane, 4 * (h & 1) + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
      if (h & 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&maps.c, out16, col0 - 32, row0, 0);
          tma_store_commit();
        }
      }
    } else {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&maps.c, stg, col0, row0, 0);
        tma_store_commit();
      }
    }
    st.g = g + 1;
  }

  __device__ static void epilogue(const Params& p, const Maps& maps, int tile, int next_tile, uint32_t tmem,
                                  int warp, int lane, uint64_t*, State& st, uint8_t* epi_smem) {
    int mt, nt;
    decode(p, tile, mt, nt);
    const int quarter = warp & 3;                     // TMEM lanes 32 * quarter .. + 31 = rows of this warp
    const int row0 = mt * kBM + quarter * 32;
    const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t stg = smem_u32(epi_smem) + warp * kWarpStage;
    Cols64 va, vb;
    if (!kRes) {
      // warp w takes 64-column chunks [c0, c0 + kChunks) of the tile: all four with 4 epilogue warps,
      // one 128-column half with 8
      constexpr int kChunks = 16 / kEpiWarps;
      const int c0 = (warp >> 2) * kChunks;
      ld64(taddr + c0 * 64, va);
#pragma unroll 1
      for (int c = c0; c < c0 + kChunks; c += 2) {
        const int col = nt * kBN + c * 64;
        wait64(va);
        ld64(taddr + (c + 1) * 64, vb);
        chunk_f16(p, maps, va, col, row0, lane, stg);
        wait64(vb);
        if (c + 2 < c0 + kChunks) ld64(taddr + (c + 2) * 64, va);
        chunk_f16(p, maps, vb, col + 64, row0, lane, stg + 4096);
      }
      return;
    }
    ld64(taddr, va);
#pragma unroll 1
    for (int c = 0; c < kBN / 64; c += 2) {
      const int col = nt * kBN + c * 64;
      wait64(va);
      ld64(taddr + (c + 1) * 64, vb);
      half_res(p, maps, va.lo, tile, next_tile, 2 * c, col, row0, warp, lane, st, epi_smem);
      half_res(p, maps, va.hi, tile, next_tile, 2 * c + 1, col + 32, row0, warp, lane, st, epi_smem);
      wait64(vb);
      if (c + 2 < kBN / 64) ld64(taddr + (c + 2) * 64, va);
      half_res(p, maps, vb.lo, tile, next_tile, 2 * c + 2, col + 64, row0, warp, lane, st, epi_smem);
      half_res(p, maps, vb.hi, tile, next_tile, 2 * c + 3, col + 96, row0, warp, lane, st, epi_smem);
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
  if (epilogue != RZ_LIN_BIAS && epilogue != RZ_LIN_GELU && epilogue != RZ_LIN_RESIDUAL &&
      epilogue != RZ_LIN_RESIDUAL_F16)
    return RZ_ERR_INVALID;
  const bool res = epilogue == RZ_LIN_RESIDUAL || epilogue == RZ_LIN_RESIDUAL_F16;
  if (res && residual == nullptr) return RZ_ERR_INVALID;
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
  if (res &&
      !rz::make_map_3d_f32_sw128(&mp.c2, residual, 1, (uint64_t)m, (uint64_t)n, (uint64_t)n * 4, (uint64_t)m * n * 4, 32))
    return RZ_ERR_CUDA;
  switch (epilogue) {
    case RZ_LIN_BIAS: return launch_lin<RZ_LIN_BIAS>(mp, p, pair, s);
    case RZ_LIN_GELU: return launch_lin<RZ_LIN_GELU>(mp, p, pair, s);
    case RZ_LIN_RESIDUAL_F16: return launch_lin<RZ_LIN_RESIDUAL_F16>(mp, p, pair, s);
    default: return launch_lin<RZ_LIN_RESIDUAL>(mp, p, pair, s);
  }
}
