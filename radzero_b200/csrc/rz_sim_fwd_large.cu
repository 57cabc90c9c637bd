// K3-K6 for LARGE prompt sets (open-vocabulary sweeps, the contrastive step): the same
// computation as rz_sim_fwd.cu, restructured as TWO full-rate (M = 128) tcgen05 GEMM passes on
// the shared skeleton of rz_gemm.cuh, because one SM cannot hold the pooled accumulator of more
// than 64 prompts x 768 features in TMEM:
//   pass S   S = q k^T * scale per (image, 128 prompts) item, swept over the image's tokens in
//            tiles of 256.  One epilogue thread owns one prompt row for the whole item and keeps
//            a LAZY running reference maximum m and the sum l of exp(s - m) in registers
//            (FlashAttention-style, but the reference only moves when the maximum grows by more
//            than e^10, which keeps exp(s - m) inside fp16 and makes rescaling rare).  It emits
//            the optional fp32 scores (the similarity map; warp-transposed through shared memory
//            so that every store instruction writes 128 contiguous bytes of one row) and the
//            UNNORMALISED probabilities P~ = exp(s - m) as fp16 [B, N, Lp], then (m, l) per row.
//            If the reference does move, the thread rescales the part of its own row it has
//            already written.
//   pass PK  o = (P~ k) / l per (image, 128 prompts, 256 features), K = tokens; epilogue: pooled
//            vectors (fp16, kept for the backward) + per-tile |o|^2 and <q, o> partials -> Z
// S is computed exactly once: 2 GEMM units of MMA work for the 2 algorithmic ones.
// Replaces SimilarityLogit.forward (exp/cxr_pt/model/losses.py:187-240).
#include "rz_gemm.cuh"

namespace {

using namespace rz::gemm;
constexpr int kD = RZ_HIDDEN;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kGrow = 10.0f;    // the running reference moves when the maximum exceeds it by this

struct S2Params {
  int B, N, L, Lp, m_tiles, n_tiles;
  float scale;
  const float* log_tau_scale;
  __half* p_out;               // [B, N, Lp]  exp(s - mref)   (written through maps.a2)
  float* mref;                 // [B, N]
  float* lsum;                 // [B, N]      sum_l exp(s - mref)
  float* lse;                  // optional [B, N] = mref + log(lsum)
};

// kScores: also emit the fp32 similarity map (maps.b2 = [B, N, L] fp32 store map)
template <bool kScores, int C>
struct PassS2 : PolicyBase {
  static constexpr int kCluster = C;   // CTA pair: prompt tiles (2i, 2i+1) over the same tokens, one 256-row MMA
  using Params = S2Params;
  struct State { float m, l, scale; };
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = false, kBMn = false, kTwoPhase = false;
  // per-warp TMA-store staging (SWIZZLE_128B boxes of 32 rows x 128 B):
  //   [scores, 32 columns fp32] (kScores) [P~, 64 columns fp16]
  static constexpr int kWarpStage = kScores ? 8192 : 4096;
  static constexpr int kEpiSmem = 4 * kWarpStage;
  __host__ __device__ static int num_tiles(const Params& p) { return p.B * p.m_tiles * p.n_tiles; }
  __host__ __device__ static int inner(const Params& p) { return p.n_tiles; }
  __host__ __device__ static int k_steps(const Params&) { return kD / kBK; }
  __host__ __device__ static int tile_n(const Params& p, int tile) {
    const int rem = p.Lp - (tile % p.n_tiles) * kBN;
    return rem < kBN ? rem : kBN;
  }
  __device__ static void decode(const Params& p, int tile, int& b, int& mt, int& nt) {
    nt = tile % p.n_tiles;
    const int r = tile / p.n_tiles;
    mt = r % p.m_tiles;
    b = r / p.m_tiles;
  }
  __device__ static void load(const Params& p, const Maps& m, int tile, int ks, uint8_t* a, uint8_t*,
                              uint8_t* bsm, uint64_t* bar, int rank) {
    int b, mt, nt;
    decode(p, tile, b, mt, nt);
    load_kmajor<C>(&m.a, bar, a, ks * kBK, mt * kBM, 0);     // q [N, 768]
    // k [B, Lp, 768] (rows >= Lp: zero fill): each CTA of a pair loads half of the token tile
    load_kmajor_shared<C>(&m.b, bar, bsm, ks * kBK, nt * kBN, b, tile_n(p, tile), rank);
  }

  // one 64-column chunk of this thread's row: lazy maximum, exp, stage, TMA store.  kMasked: the
  // chunk straddles the last real token (columns >= L are padding); the common case runs mask-free.
  template <bool kMasked>
  __device__ static __forceinline__ void chunk(const Params& p, const Maps& maps, Cols64& v, int l0,
                                               int b, int row0, int lane, bool row_ok, long long pi,
                                               uint32_t stg, State& st) {
    const float scale = st.scale;                    // > 0: the maximum commutes with it
    float amax = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float a0 = __uint_as_float(v.lo[i]), a1 = __uint_as_float(v.hi[i]);
      if (!kMasked) {
        amax = fmaxf(amax, fmaxf(a0, a1));
      } else {
        if (l0 + i < p.L) amax = fmaxf(amax, a0);
        if (l0 + 32 + i < p.L) amax = fmaxf(amax, a1);
      }
    }
    const float cmax = amax * scale;
    const bool grow = cmax > st.m + kGrow;          // true for the first chunk (m = -inf)
    const bool resc = grow && st.l > 0.f;
    if (__any_sync(0xffffffffu, resc)) {
      // rare: some row's maximum grew by more than e^10 -- rescale what that row has accumulated
      // and what it has already written.  The bulk stores of this warp must have landed first.
      if (lane == 0) tma_store_wait_all();
      __syncwarp();
      if (resc) {
        const float alpha = exp2f((st.m - cmax) * kLog2e);
        st.l *= alpha;
        if (row_ok) {
          __half* prow = p.p_out + pi * p.Lp;
          for (int c = 0; c < l0; c += 8) {
            uint4 w = __ldcg(reinterpret_cast<const uint4*>(prow + c));
            __half2* h = reinterpret_cast<__half2*>(&w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(h[j]);
              h[j] = __floats2half2_rn(f.x * alpha, f.y * alpha);
            }
            *reinterpret_cast<uint4*>(prow + c) = w;
          }
        }
      }
      __syncwarp();
    }
    if (grow) st.m = cmax;
    // Staging: one box for the scores (used for columns 0-31, then 32-63) and one for P~; the work
    // is ordered so that every wait for a staging box comes after a long stretch of arithmetic.
    const uint32_t stg_p = stg + (kScores ? 4096 : 0);
    if (kScores) {
      if (lane == 0) tma_store_wait_read();          // previous chunk's stores (issued long ago)
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts_v4(stg + stage_off(lane, j), __float_as_uint(__uint_as_float(v.lo[4 * j]) * scale),
               __float_as_uint(__uint_as_float(v.lo[4 * j + 1]) * scale),
               __float_as_uint(__uint_as_float(v.lo[4 * j + 2]) * scale),
               __float_as_uint(__uint_as_float(v.lo[4 * j + 3]) * scale));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&maps.c2, stg, l0, row0, b);
        tma_store_commit();
      }
    }
    // exponentials into registers (the scores box drains meanwhile)
    const float sl2 = scale * kLog2e;                // exp(s - m) = 2^(a * scale*log2e - m*log2e)
    const float mb = st.m * kLog2e;
    // packed fp32 pairs for the affine map and the row sum (one issue slot per two columns)
    const P2 sl22 = p2(sl2), nmb2 = p2(-mb);
    P2 lacc = p2(0.f);
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {                    // 8 columns -> one 16-byte chunk of the P~ row
      const uint32_t* src = j < 4 ? &v.lo[8 * j] : &v.hi[8 * (j - 4)];
      const int lc = l0 + 8 * j;
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float x0, x1;
        p2_unpack(p2_fma(p2(__uint_as_float(src[i]), __uint_as_float(src[i + 1])), sl22, nmb2), x0, x1);
        float e0 = exp2f(x0), e1 = exp2f(x1);
        if (kMasked && lc + i >= p.L) e0 = 0.f;
        if (kMasked && lc + i + 1 >= p.L) e1 = 0.f;
        const P2 ee = p2(e0, e1);
        lacc = p2_add(lacc, ee);
        pk[4 * j + (i >> 1)] = p2_pack_h2(ee);
      }
    }
    float lacc0, lacc1;
    p2_unpack(lacc, lacc0, lacc1);
    st.l += lacc0 + lacc1;
    if (lane == 0) tma_store_wait_read();            // !kScores: the previous chunk's P~ store
    __syncwarp();
    if (kScores) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts_v4(stg + stage_off(lane, j), __float_as_uint(__uint_as_float(v.hi[4 * j]) * scale),
               __float_as_uint(__uint_as_float(v.hi[4 * j + 1]) * scale),
               __float_as_uint(__uint_as_float(v.hi[4 * j + 2]) * scale),
               __float_as_uint(__uint_as_float(v.hi[4 * j + 3]) * scale));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_v4(stg_p + stage_off(lane, j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (kScores && l0 + 32 < p.L) tma_store_3d(&maps.c2, stg, l0 + 32, row0, b);
      tma_store_3d(&maps.c, stg_p, l0, row0, b);
      tma_store_commit();
    }
  }
  __device__ static __forceinline__ void chunk_any(const Params& p, const Maps& maps, Cols64& v, int l0,
                                                   int b, int row0, int lane, bool row_ok, long long pi,
                                                   uint32_t stg, State& st) {
    if (l0 + 64 <= p.L) chunk<false>(p, maps, v, l0, b, row0, lane, row_ok, pi, stg, st);
    else chunk<true>(p, maps, v, l0, b, row0, lane, row_ok, pi, stg, st);
  }

  __device__ static void epilogue(const Params& p, const Maps& maps, int tile, int, uint32_t tmem,
                                  int warp, int lane, uint64_t*, State& st, uint8_t* epi_smem) {
    int b, mt, nt;
    decode(p, tile, b, mt, nt);
    const int row0 = mt * kBM + warp * 32;
    const int n = row0 + lane;
    const bool row_ok = n < p.N;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const long long pi = (long long)b * p.N + (row_ok ? n : 0);
    const uint32_t stg = smem_u32(epi_smem) + warp * kWarpStage;
    if (nt == 0) {
      st.m = -INFINITY;
      st.l = 0.f;
      st.scale = p.log_tau_scale != nullptr ? __expf(-__ldg(p.log_tau_scale)) : p.scale;
    }
    const int nch = tile_n(p, tile) / 64;            // 4, or 2 for the narrow last tile
    const int tok0 = nt * kBN;
    Cols64 va, vb;
    ld64(taddr, va);
#pragma unroll 1
    for (int c = 0; c < nch; c += 2) {
      wait64(va);
      ld64(taddr + (c + 1) * 64, vb);
      chunk_any(p, maps, va, tok0 + c * 64, b, row0, lane, row_ok, pi, stg, st);
      wait64(vb);
      if (c + 2 < nch) ld64(taddr + (c + 2) * 64, va);
      chunk_any(p, maps, vb, tok0 + (c + 1) * 64, b, row0, lane, row_ok, pi, stg, st);
    }
    if (nt == p.n_tiles - 1 && row_ok) {
      p.mref[pi] = st.m;
      p.lsum[pi] = st.l;
      if (p.lse != nullptr) p.lse[pi] = st.m + __logf(st.l);
    }
  }
};

// ------------------------------------------------------------------------------------------
struct PKParams {
  int B, N, Lp, m_tiles;
  const __half* q;             // [N, 768] (for <q, o>)
  const float* lsum;           // [B, N]: the accumulator is divided by it (P~ is unnormalised)
  __half* pooled;              // optional [B, N, 768]
  float* part;                 // [B, N, 3, 2] (|o|^2, <q,o>) per 256-feature slab
};

template <int C>
struct PassPK : PolicyBase {
  static constexpr int kCluster = C;   // prompt tiles (2i, 2i+1) of one image share the token operand
  using Params = PKParams;
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = false, kBMn = true, kTwoPhase = false;
  // per warp: [q in, buffer 0][q in, buffer 1][pooled out], boxes of 32 rows x 128 B (64 fp16 columns)
  static constexpr int kWarpStage = 3 * 4096;
  static constexpr int kEpiSmem = 4 * kWarpStage;
  struct State { uint32_t g; int primed; };         // q boxes consumed so far by this warp
  __host__ __device__ static int num_tiles(const Params& p) { return p.B * p.m_tiles * (kD / kBN); }
  // the three feature tiles of one (image, prompt tile) run back to back on one CTA: the P~ tile
  // they share as A operand is fetched from HBM once and re-read from L2
  __host__ __device__ static int inner(const Params&) { return kD / kBN; }
  __host__ __device__ static int k_steps(const Params& p) { return p.Lp / kBK; }
  __device__ static void decode(const Params& p, int tile, int& b, int& mt, int& ft) {
    ft = tile % (kD / kBN);
    const int r = tile / (kD / kBN);
    mt = r % p.m_tiles;
    b = r / p.m_tiles;
  }
  __device__ static void load(const Params& p, const Maps& m, int tile, int ks, uint8_t* a, uint8_t*,
                              uint8_t* bsm, uint64_t* bar, int rank) {
    int b, mt, ft;
    decode(p, tile, b, mt, ft);
    load_kmajor<C>(&m.a, bar, a, ks * kBK, mt * kBM, b);                              // P [B, N, Lp]
    load_mnmajor_shared<C>(&m.b, bar, bsm, ft * kBN, ks * kBK, b, kBN / 64, rank);    // k [B, Lp, 768]
  }
  // lane 0: fetch the q box (32 prompts x 64 features) of chunk `c` of `tile` for this warp's rows.
  // (A thread-per-row global load touches 32 different lines per instruction and left the epilogue
  // waiting on the L2: it was the stall that kept this pass at 75 % of the tensor pipe.)
  __device__ static __forceinline__ void fetch(const Params& p, const Maps& maps, int tile, int c, int warp,
                                               uint32_t stg, uint64_t* bars, uint32_t g) {
    int b, mt, ft;
    decode(p, tile, b, mt, ft);
    uint64_t* bar = bars + warp * 2 + (g & 1);
    mbar_arrive_expect_tx(bar, 4096u);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(stg + (g & 1) * 4096u),
        "l"(reinterpret_cast<uint64_t>(&maps.a2)), "r"(smem_u32(bar)), "r"(ft * kBN + c * 64),
        "r"(mt * kBM + warp * 32), "r"(0), "l"(kEvictLast)
        : "memory");
  }
  template <class S>
  __device__ static void prologue(const Params& p, const Maps& maps, int tile, int warp, int lane,
                                  uint64_t* bars, S& st, uint8_t* epi_smem) {
    if (st.primed) return;
    st.primed = 1;
    st.g = 0;
    if (lane == 0) {
      const uint32_t stg = smem_u32(epi_smem) + warp * kWarpStage;
      fetch(p, maps, tile, 0, warp, stg, bars, 0);
      fetch(p, maps, tile, 1, warp, stg, bars, 1);
    }
  }
  __device__ static __forceinline__ void chunk(const Params& p, const Maps& maps, Cols64& v, int tile,
                                               int next_tile, int c, int f0, int b, int row0, int warp, int lane,
                                               float linv, uint32_t stg, uint64_t* bars, State& st, P2& osq,
                                               P2& qo) {
    const bool store = p.pooled != nullptr;
    // this chunk's q box: 64 halves of this thread's prompt row
    const uint32_t g = st.g;
    const uint32_t in = stg + (g & 1) * 4096u;
    mbar_wait(bars + warp * 2 + (g & 1), (g >> 1) & 1);
    uint32_t qw[32];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(qw[4 * j]), "=r"(qw[4 * j + 1]), "=r"(qw[4 * j + 2]), "=r"(qw[4 * j + 3])
                   : "r"(in + stage_off(lane, j)));
    lds_returned(smem_u32(bars + 8) + 4u * (uint32_t)warp, qw[3], qw[7], qw[11], qw[15], qw[19], qw[23], qw[27],
                 qw[31]);                               // before the refill of the same buffer is issued
    __syncwarp();
    if (lane == 0) {
      constexpr int nch = kBN / 64;
      if (c + 2 < nch) fetch(p, maps, tile, c + 2, warp, stg, bars, g + 2);
      else if (next_tile >= 0) fetch(p, maps, next_tile, c + 2 - nch, warp, stg, bars, g + 2);
    }
    st.g = g + 1;
    const uint32_t out = stg + 8192;
    if (store) {
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    }
    const P2 linv2 = p2(linv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t* src = j < 4 ? &v.lo[8 * j] : &v.hi[8 * (j - 4)];
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const P2 a = p2_mul(p2(__uint_as_float(src[i]), __uint_as_float(src[i + 1])), linv2);
        const float2 qq = __half22float2(*reinterpret_cast<const __half2*>(&qw[4 * j + (i >> 1)]));
        osq = p2_fma(a, a, osq);
        qo = p2_fma(a, p2(qq.x, qq.y), qo);
        o[i >> 1] = p2_pack_h2(a);
      }
      if (store) sts_v4(out + stage_off(lane, j), o[0], o[1], o[2], o[3]);
    }
    if (store) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&maps.c, out, f0, row0, b);
        tma_store_commit();
      }
    }
  }
  __device__ static void epilogue(const Params& p, const Maps& maps, int tile, int next_tile, uint32_t tmem,
                                  int warp, int lane, uint64_t* bars, State& st, uint8_t* epi_smem) {
    int b, mt, ft;
    decode(p, tile, b, mt, ft);
    const int row0 = mt * kBM + warp * 32;
    const int n = row0 + lane;
    const bool row_ok = n < p.N;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const long long pi = (long long)b * p.N + (row_ok ? n : 0);
    const uint32_t stg = smem_u32(epi_smem) + warp * kWarpStage;
    const float linv = row_ok ? 1.0f / p.lsum[pi] : 0.f;
    P2 osq = p2(0.f), qo = p2(0.f);
    Cols64 va, vb;
    ld64(taddr, va);
#pragma unroll 1
    for (int c = 0; c < kBN / 64; c += 2) {
      wait64(va);
      ld64(taddr + (c + 1) * 64, vb);
      chunk(p, maps, va, tile, next_tile, c, ft * kBN + c * 64, b, row0, warp, lane, linv, stg, bars, st, osq, qo);
      wait64(vb);
      if (c + 2 < kBN / 64) ld64(taddr + (c + 2) * 64, va);
      chunk(p, maps, vb, tile, next_tile, c + 1, ft * kBN + (c + 1) * 64, b, row0, warp, lane, linv, stg, bars,
            st, osq, qo);
    }
    if (row_ok) {
      float o0, o1, q0, q1;
      p2_unpack(osq, o0, o1);
      p2_unpack(qo, q0, q1);
      float2* d = reinterpret_cast<float2*>(p.part) + (pi * (kD / kBN) + ft);
      *d = make_float2(o0 + o1, q0 + q1);
    }
  }
};

struct FinParams {
  const float* part; int B, N;
  const float* q_inv_norm;
  float* z; long long z_sn, z_sb; float z_scale; const float* log_tau_z; int z_sigmoid;
  float* onorm;
};

__global__ void z_finalize_kernel(FinParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.B * p.N) return;
  const int b = (int)(i / p.N), n = (int)(i - (long long)b * p.N);
  const float2* q = reinterpret_cast<const float2*>(p.part) + i * (kD / 256);
  float osq = 0.f, qo = 0.f;
  for (int t = 0; t < kD / 256; ++t) { osq += q[t].x; qo += q[t].y; }
  const float on = sqrtf(osq);
  float z = qo / fmaxf(on, RZ_L2_EPS);
  if (p.q_inv_norm != nullptr) z *= p.q_inv_norm[n];
  if (p.onorm != nullptr) p.onorm[i] = on;
  if (p.z != nullptr) {
    const float zs = p.log_tau_z != nullptr ? __expf(-__ldg(p.log_tau_z)) : p.z_scale;
    float zo = z * zs;
    if (p.z_sigmoid) zo = 1.0f / (1.0f + __expf(-zo));
    p.z[(long long)n * p.z_sn + (long long)b * p.z_sb] = zo;
  }
}

}  // namespace

extern "C" size_t rz_sim_fwd_large_workspace_bytes(int n_images, int n_text, int tokens_padded) {
  if (n_images <= 0 || n_text <= 0 || tokens_padded <= 0) return 0;
  const size_t pairs = (size_t)n_images * n_text;
  return pairs * tokens_padded * sizeof(__half)                  // P~ (when the caller keeps none)
         + pairs * 3 * 2 * sizeof(float)                         // (|o|^2, <q,o>) partials
         + pairs * 2 * sizeof(float) + 256;                      // mref, lsum
}

extern "C" int rz_sim_fwd_large(const void* k_f16, int n_images, int tokens, int tokens_padded,
                                const void* q_f16, int n_text, float scale,
                                const float* log_tau_scale, const float* q_inv_norm, float* scores,
                                long long scores_stride_image, long long scores_stride_text,
                                int drop_cls, float* z, long long z_stride_text,
                                long long z_stride_image, float z_scale, const float* log_tau_z,
                                int z_sigmoid, float* lse, float* onorm, void* pooled_f16,
                                void* p_f16, float* mref, float* lsum, int want_pool,
                                void* workspace, size_t workspace_bytes, void* stream) {
  if (!k_f16 || !q_f16 || !workspace) return RZ_ERR_INVALID;
  if (n_images <= 0 || n_text <= 0 || tokens <= 0 || tokens_padded < tokens) return RZ_ERR_INVALID;
  if (tokens_padded % 128 != 0 || (drop_cls != 0 && drop_cls != 1)) return RZ_ERR_INVALID;
  {
    size_t need = rz_sim_fwd_large_workspace_bytes(n_images, n_text, tokens_padded);
    if (p_f16 != nullptr) need -= (size_t)n_images * n_text * tokens_padded * sizeof(__half);
    if (workspace_bytes < need) return RZ_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(k_f16) & 15) ||
      (reinterpret_cast<uintptr_t>(q_f16) & 15) || (reinterpret_cast<uintptr_t>(pooled_f16) & 15) ||
      (reinterpret_cast<uintptr_t>(p_f16) & 15))
    return RZ_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = n_images, N = n_text, Lp = tokens_padded;
  const size_t pairs = (size_t)B * N;
  const int m_tiles = (N + 127) / 128;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* part_o = reinterpret_cast<float*>(ws);
  float* mref_ws = part_o + pairs * 6;
  float* lsum_ws = mref_ws + pairs;
  const size_t small = (pairs * 8 * sizeof(float) + 255) & ~size_t(255);
  __half* pbuf = p_f16 != nullptr ? static_cast<__half*>(p_f16) : reinterpret_cast<__half*>(ws + small);
  if (mref == nullptr) mref = mref_ws;
  if (lsum == nullptr) lsum = lsum_ws;
  const bool pool = want_pool != 0 || z != nullptr || onorm != nullptr || pooled_f16 != nullptr;

  const bool tma_scores = scores != nullptr;
  if (tma_scores) {
    // the similarity map is written by TMA: rows hold all `tokens` columns (CLS at column 0; the
    // caller slices it off), 16-byte aligned base and pitches
    if (drop_cls != 0) return RZ_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(scores) & 15) || (scores_stride_text % 4) || (scores_stride_image % 4) ||
        scores_stride_text < tokens)
      return RZ_ERR_ALIGNMENT;
  }
  Maps m = {};
  if (!rz::make_map_3d_sw128(&m.a, q_f16, 1, N, kD, kD * 2, (uint64_t)N * kD * 2, kBM)) return RZ_ERR_CUDA;
  const int C = (m_tiles % 2 == 0) ? 2 : 1;     // cluster pairs need an even number of prompt tiles
  if (!rz::make_map_3d_sw128(&m.b, k_f16, B, Lp, kD, kD * 2, (uint64_t)Lp * kD * 2, 256 / C)) return RZ_ERR_CUDA;
  m.a2 = m.a; m.b2 = m.b;
  if (!rz::make_map_3d_sw128(&m.c, pbuf, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, 32)) return RZ_ERR_CUDA;
  m.c2 = m.c;
  if (tma_scores &&
      !rz::make_map_3d_f32_sw128(&m.c2, scores, B, N, tokens, (uint64_t)scores_stride_text * 4,
                                 (uint64_t)scores_stride_image * 4, 32))
    return RZ_ERR_CUDA;
  S2Params sp;
  sp.B = B; sp.N = N; sp.L = tokens; sp.Lp = Lp; sp.m_tiles = m_tiles;
  sp.n_tiles = (Lp + 255) / 256;
  sp.scale = scale; sp.log_tau_scale = log_tau_scale;
  sp.p_out = pbuf; sp.mref = mref; sp.lsum = lsum; sp.lse = lse;
  {
    int rc = tma_scores ? (C == 2 ? launch<PassS2<true, 2>>(m, sp, s) : launch<PassS2<true, 1>>(m, sp, s))
                        : (C == 2 ? launch<PassS2<false, 2>>(m, sp, s) : launch<PassS2<false, 1>>(m, sp, s));
    if (rc != RZ_OK) return rc;
  }
  if (!pool) return RZ_OK;
  {
    Maps mk = {};
    if (!rz::make_map_3d_sw128(&mk.a, pbuf, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, kBM)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&mk.b, k_f16, B, Lp, kD, kD * 2, (uint64_t)Lp * kD * 2, 64)) return RZ_ERR_CUDA;
    mk.b2 = mk.b; mk.c = mk.a; mk.c2 = mk.a;
    // q boxes of 32 prompts x 64 features for the epilogue warps (rows >= N: zero fill)
    if (!rz::make_map_3d_sw128(&mk.a2, q_f16, 1, N, kD, kD * 2, (uint64_t)N * kD * 2, 32)) return RZ_ERR_CUDA;
    if (pooled_f16 != nullptr &&
        !rz::make_map_3d_sw128(&mk.c, pooled_f16, B, N, kD, kD * 2, (uint64_t)N * kD * 2, 32))
      return RZ_ERR_CUDA;
    PKParams kp;
    kp.B = B; kp.N = N; kp.Lp = Lp; kp.m_tiles = m_tiles; kp.q = static_cast<const __half*>(q_f16);
    kp.lsum = lsum; kp.pooled = static_cast<__half*>(pooled_f16); kp.part = part_o;
    // CTA pairs pay off while the re-read P~ stream still hits L2; on multi-GB streams (the C4 step)
    // the pair kernels were measured to re-read more from HBM (46 vs 27 GB) and to be slower
    const bool pair_pk = C == 2 && pairs * Lp * sizeof(__half) <= ((size_t)2 << 30);
    int rc = pair_pk ? launch<PassPK<2>>(mk, kp, s) : launch<PassPK<1>>(mk, kp, s);
    if (rc != RZ_OK) return rc;
    FinParams fp;
    fp.part = part_o; fp.B = B; fp.N = N; fp.q_inv_norm = q_inv_norm;
    fp.z = z; fp.z_sn = z_stride_text; fp.z_sb = z_stride_image; fp.z_scale = z_scale;
    fp.log_tau_z = log_tau_z; fp.z_sigmoid = z_sigmoid; fp.onorm = onorm;
    z_finalize_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(fp);
    RZ_LAUNCH_OK();
    rz_count_launch();
  }
  return RZ_OK;
}
