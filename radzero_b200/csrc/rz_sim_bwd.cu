// Backward of the VL-CABS similarity (SimilarityLogit, exp/cxr_pt/model/losses.py:187-240)
// with respect to the normalised operands q (N x 768) and k (B x L x 768), in closed form, as
// three tcgen05 GEMM passes over per-image (prompt x token) coefficient matrices.
//
// Per pair (b, n), with s_l = <q_n, k_bl>/tau, p = softmax_l(s), o = sum_l p_l k_bl,
// Z = <q, o/|o|> and g = dL/dZ:
//   a = g/|o|,  r = Z/|o|,   T_l = <o, k_bl>
//   dL/ds_l = a tau p_l (s_l - (r/tau) T_l)                       (since <dL/do, o> = 0)
//   W1_l = a p_l (1 + s_l - (r/tau) T_l),   W2_l = -a r p_l
//   dL/dq_n  = sum_b sum_l W1_l k_bl                               (pass Q)
//   dL/dk_bl = sum_n W1_l q_n + W2_l o_bn                          (pass K)
//   dL/dlog tau_attn = -sum dL/ds_l s_l
// Pass D recomputes S and T for a (128 prompts x 128 tokens) tile with TWO accumulators that
// share the token tile as B operand, and writes W1 / W2 as fp16 (scaled by a power of two
// `*scale` chosen on the device so that the small gradient values stay in fp16's normal range).
#include "rz_gemm.cuh"

namespace {

using namespace rz::gemm;
constexpr int kD = RZ_HIDDEN;
constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------
// per-pair coefficients + the fp16 scale
struct CoefParams {
  const float* g; const float* z; long long ldz;    // [N, ldz] (column = local image)
  const float* onorm;                               // [B, N]
  const float* q_inv_norm;                          // [N] or null: sim_op "dot" (operands not L2-normalised)
  float* coef_a; float* coef_r;                     // [B, N]
  unsigned int* amax_bits;                          // max |a| as float bits
  int B, N;
};

// g and z are [N, ldz] (image fastest), the coefficients [B, N] (prompt fastest): a 32 x 32 tile goes through
// shared memory so that both sides are read and written in 128-byte rows (a thread per pair read the
// inputs with a stride of ldz: 164 us at C4).  grid = (ceil(N / 32), ceil(B / 32)), block = (32, 8).
__global__ void __launch_bounds__(256) pair_coef_kernel(CoefParams p) {
  __shared__ float sg[32][33], sz[32][33];
  const int n0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {          // rows = prompts, columns = images (coalesced along b)
    const int n = n0 + r, b = b0 + threadIdx.x;
    const bool ok = n < p.N && b < p.B;
    sg[r][threadIdx.x] = ok ? p.g[(long long)n * p.ldz + b] : 0.f;
    sz[r][threadIdx.x] = ok ? p.z[(long long)n * p.ldz + b] : 0.f;
  }
  __syncthreads();
  float a_abs = 0.f;
  for (int r = threadIdx.y; r < 32; r += 8) {          // rows = images, columns = prompts (coalesced along n)
    const int b = b0 + r, n = n0 + threadIdx.x;
    if (b < p.B && n < p.N) {
      const long long i = (long long)b * p.N + n;
      const float on = fmaxf(p.onorm[i], RZ_L2_EPS);
      // sim_op "dot": Z = <q/|q|, o/|o|> and s = c <q, k>.  With a' = a / |q| and r' = r |q| the cosine
      // formulas hold unchanged (1/tau' = c |q| per prompt, folded into r'); the radial term of dq is
      // added by the caller.
      const float qin = p.q_inv_norm != nullptr ? p.q_inv_norm[n] : 1.0f;
      const float a = sg[threadIdx.x][r] / on * qin;
      p.coef_a[i] = a;
      p.coef_r[i] = sz[threadIdx.x][r] / (on * qin);
      a_abs = fmaxf(a_abs, fabsf(a));
    }
  }
  a_abs = rz::warp_max(a_abs);
  if (threadIdx.x == 0 && a_abs > 0.f) atomicMax(p.amax_bits, __float_as_uint(a_abs));
}

// scale = 2^k with scale * max|a| in [256, 512)
__global__ void pick_scale_kernel(const unsigned int* amax_bits, float* scale) {
  const float m = __uint_as_float(*amax_bits);
  float s = 1.0f;
  if (m > 0.f && isfinite(m)) {
    int e;
    frexpf(m, &e);                // m = f * 2^e, f in [0.5, 1)
    s = ldexpf(1.0f, 9 - e);      // s*m in [256, 512)
  }
  scale[0] = s;
  scale[1] = 1.0f / s;
}

// ------------------------------------------------------------------------------------------
// pass D: W1, W2 for every (image, prompt tile, token tile)
struct DParams {
  int B, N, L, Lp;
  int m_tiles, n_tiles;
  float inv_tau;                 // 1/tau (host value) ...
  const float* log_tau;          // ... or device log-temperature
  const float* lse;              // [B, N]
  const float* coef_a; const float* coef_r;
  const float* scale;            // [2] = S, 1/S
  __half* w1; __half* w2;        // [B, N, Lp]
  float* dtau_part;              // [tiles] partial sums of -dL/ds * s (unscaled)
};

template <int C>
struct PassD : PolicyBase {
  static constexpr int kCluster = C;   // prompt tiles (2i, 2i+1) share the token tile
  using Params = DParams;
  static constexpr int kBN = 128, kAccs = 2;
  static constexpr bool kAMn = false, kBMn = false, kTwoPhase = false;
  // per-warp TMA-store staging: one 64-column fp16 box (32 rows x 128 B, SWIZZLE_128B) each for W1, W2
  static constexpr int kEpiSmem = 4 * 8192;
  __host__ __device__ static int num_tiles(const Params& p) { return p.B * p.m_tiles * p.n_tiles; }
  __host__ __device__ static int k_steps(const Params&) { return kD / kBK; }
  // prompt tile fastest: consecutive tile ids (a cluster pair) share image and token tile
  __device__ static void decode(const Params& p, int tile, int& b, int& mt, int& nt) {
    mt = tile % p.m_tiles;
    const int r = tile / p.m_tiles;
    nt = r % p.n_tiles;
    b = r / p.n_tiles;
  }
  __device__ static void load(const Params& p, const Maps& m, int tile, int ks, uint8_t* a,
                              uint8_t* a2, uint8_t* bsm, uint64_t* bar, int rank) {
    int b, mt, nt;
    decode(p, tile, b, mt, nt);
    load_kmajor<C>(&m.a, bar, a, ks * kBK, mt * kBM, 0);     // q      [N, 768]
    load_kmajor<C>(&m.a2, bar, a2, ks * kBK, mt * kBM, b);   // pooled [B, N, 768]
    load_kmajor_shared<C>(&m.b, bar, bsm, ks * kBK, nt * kBN, b, kBN, rank);   // k [B, Lp, 768]
  }
  struct Row {
    float inv_tau, lse2, aS, c2, rt;     // lse2 = lse * log2(e); c2 = -aS * r; rt = r / tau
  };
  // 32 columns of S and T of this thread's row -> half of a staged 64-column W1 / W2 box
  template <bool kMasked>
  __device__ static __forceinline__ void chunk(const Params& p, uint32_t (&sv)[32], uint32_t (&tv)[32],
                                               int l0, int half, int lane, const Row& r, uint32_t stg,
                                               float& u) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t o1[4], o2[4];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float w1v[2], w2v[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = 8 * j + i + q;
          const float s = __uint_as_float(sv[c]) * r.inv_tau;
          const float t = __uint_as_float(tv[c]);
          float pr = exp2f(fmaf(s, kLog2e, -r.lse2));
          if (kMasked && l0 + c >= p.L) pr = 0.f;
          const float e = fmaf(-r.rt, t, s);        // s - (r/tau) T
          const float x = r.aS * pr;
          w1v[q] = fmaf(x, e, x);                   // aS p (1 + e)
          w2v[q] = r.c2 * pr;                       // -aS r p
          u = fmaf(pr * e, s, u);                   // dL/dlog tau = -a tau sum p e s
        }
        o1[i >> 1] = pack_h2(w1v[0], w1v[1]);
        o2[i >> 1] = pack_h2(w2v[0], w2v[1]);
      }
      sts_v4(stg + stage_off(lane, 4 * half + j), o1[0], o1[1], o1[2], o1[3]);
      sts_v4(stg + 4096 + stage_off(lane, 4 * half + j), o2[0], o2[1], o2[2], o2[3]);
    }
  }
  __device__ static __forceinline__ void ld2(uint32_t taddr, int c0, uint32_t (&sv)[32], uint32_t (&tv)[32]) {
    tmem_ld_x32(taddr + c0, sv);
    tmem_ld_x32(taddr + kBN + c0, tv);
  }
  __device__ static __forceinline__ void wait2(uint32_t (&sv)[32], uint32_t (&tv)[32]) {
    tmem_ld_wait_x32(sv);
    tmem_ld_wait_x32(tv);
  }
  __device__ static void epilogue(const Params& p, const Maps& maps, int tile, int, uint32_t tmem,
                                  int warp, int lane, uint64_t*, State&, uint8_t* epi_smem) {
    int b, mt, nt;
    decode(p, tile, b, mt, nt);
    const int row0 = mt * kBM + warp * 32;
    const int n = row0 + lane;
    const bool row_ok = n < p.N;
    const long long pi = (long long)b * p.N + (row_ok ? n : 0);
    Row r;
    r.inv_tau = p.log_tau != nullptr ? __expf(-__ldg(p.log_tau)) : p.inv_tau;
    const float tau = 1.0f / r.inv_tau;
    r.lse2 = (row_ok ? p.lse[pi] : 0.f) * kLog2e;
    const float a = row_ok ? p.coef_a[pi] : 0.f;
    const float rr = row_ok ? p.coef_r[pi] : 0.f;
    r.aS = a * p.scale[0];
    r.c2 = -r.aS * rr;
    r.rt = rr * r.inv_tau;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t stg = smem_u32(epi_smem) + warp * 8192;
    const int tok0 = nt * kBN;
    float u = 0.f;
    uint32_t sa[32], ta[32], sb[32], tb[32];
    ld2(taddr, 0, sa, ta);
#pragma unroll 1
    for (int c = 0; c < kBN; c += 64) {
      wait2(sa, ta);
      ld2(taddr, c + 32, sb, tb);
      // the staging boxes were last read by the bulk stores of the previous 64 columns
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
      if (tok0 + c + 32 <= p.L) chunk<false>(p, sa, ta, tok0 + c, 0, lane, r, stg, u);
      else chunk<true>(p, sa, ta, tok0 + c, 0, lane, r, stg, u);
      wait2(sb, tb);
      if (c + 64 < kBN) ld2(taddr, c + 64, sa, ta);
      if (tok0 + c + 64 <= p.L) chunk<false>(p, sb, tb, tok0 + c + 32, 1, lane, r, stg, u);
      else chunk<true>(p, sb, tb, tok0 + c + 32, 1, lane, r, stg, u);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&maps.c, stg, tok0 + c, row0, b);             // W1 [B, N, Lp]
        tma_store_3d(&maps.c2, stg + 4096, tok0 + c, row0, b);     // W2 [B, N, Lp]
        tma_store_commit();
      }
    }
    // deterministic per-(tile, warp) partial of dL/dlog(tau_attn) = -a tau sum_l p e s
    u = rz::warp_sum(-a * tau * u);
    if (lane == 0) p.dtau_part[(long long)tile * 4 + warp] = u;
  }
};

// ------------------------------------------------------------------------------------------
// pass D2: the same W1, W2 from ONE GEMM.  When the forward kept its unnormalised probabilities
// P~ = exp(s - m) (fp16) and the row statistics (m, l), the scores need not be recomputed:
//   p_l = P~_l / l,     s_l = m + ln P~_l     (|error| <= 2^-11 from the fp16 rounding of P~; where
//   P~ underflowed to 0 the weight p_l is 0 and s_l is irrelevant)
// so only T = o k^T is left for the tensor cores: (128 prompts x 256 tokens) tiles like the
// forward's pass S.  The epilogue warps fetch their own P~ boxes by TMA, two chunks ahead.
struct D2Params {
  int B, N, L, Lp;
  int m_tiles, n_tiles;
  float inv_tau;
  const float* log_tau;
  const float* mref; const float* lsum;   // [B, N]
  const float* coef_a; const float* coef_r;
  const float* scale;            // [2] = S, 1/S
  float* dtau_part;              // [tiles * 4]
};

// Timing-only ablations of pass D2 (results WRONG; profiles/experiments/r2_d2_ablate.sh): bit 0 = no W2 bulk
// store, bit 1 = no W1 bulk store, bit 2 = P~ boxes fetched once and reused, bit 3 = no staging stores at all
#ifndef RZ_EXP_D2
#define RZ_EXP_D2 0
#endif
template <int C>
struct PassD2 : PolicyBase {
  static constexpr int kCluster = C;   // prompt tiles (2i, 2i+1) sweep the same tokens
  using Params = D2Params;
  struct State { uint32_t g; int primed; };     // chunks consumed so far by this warp
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = false, kBMn = false, kTwoPhase = false;
  // per warp: [P~ in, buffer 0][P~ in, buffer 1][W1 out][W2 out], boxes of 32 rows x 128 B
  static constexpr int kWarpStage = 16384;
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
    load_kmajor<C>(&m.a, bar, a, ks * kBK, mt * kBM, b);                         // pooled [B, N, 768]
    load_kmajor_shared<C>(&m.b, bar, bsm, ks * kBK, nt * kBN, b, tile_n(p, tile), rank);   // k [B, Lp, 768]
  }
  // lane 0: fetch the P~ box of chunk `c` (64 tokens) of `tile` for this warp's 32 rows
  __device__ static __forceinline__ void fetch(const Params& p, const Maps& maps, int tile, int c, int warp,
                                               uint32_t stg, uint64_t* bars, uint32_t g) {
    int b, mt, nt;
    decode(p, tile, b, mt, nt);
    uint64_t* bar = bars + warp * 2 + (g & 1);
    mbar_arrive_expect_tx(bar, 4096u);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(stg + (g & 1) * 4096u),
        "l"(reinterpret_cast<uint64_t>(&maps.a2)), "r"(smem_u32(bar)), "r"(nt * kBN + c * 64),
        "r"(mt * kBM + warp * 32), "r"(b), "l"(kEvictFirst)
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
  struct Row {
    float m, linv, aS, c2, rt;     // m = reference maximum; c2 = -aS * r; rt = r / tau
  };
  // one 64-token chunk of this thread's row
  __device__ static __forceinline__ void chunk(const Params& p, const Maps& maps, Cols64& v, int tile,
                                               int next_tile, int c, int nch, int tok0, int b, int row0,
                                               int warp, int lane, const Row& r, uint32_t stg,
                                               uint64_t* bars, State& st, float& u) {
    // this chunk's P~ box: 64 halves of this thread's row
    const uint32_t g = st.g;
    const uint32_t in = stg + (g & 1) * 4096u;
    if (!(RZ_EXP_D2 & 4) || g < 2) mbar_wait(bars + warp * 2 + (g & 1), (g >> 1) & 1);
    uint32_t pw[32];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(pw[4 * j]), "=r"(pw[4 * j + 1]), "=r"(pw[4 * j + 2]), "=r"(pw[4 * j + 3])
                   : "r"(in + stage_off(lane, j)));
    // the loads must have returned before the TMA refill of the same buffer is issued (lds_returned)
    lds_returned(smem_u32(bars + 8) + 4u * (uint32_t)warp, pw[3], pw[7], pw[11], pw[15], pw[19], pw[23], pw[27],
                 pw[31]);                               // Ctrl::sink sits right behind epi_bar[8]
    __syncwarp();
    // the buffer is free again: fetch the box two chunks ahead (possibly of the next tile)
    if (lane == 0 && !(RZ_EXP_D2 & 4)) {
      if (c + 2 < nch) fetch(p, maps, tile, c + 2, warp, stg, bars, g + 2);
      else if (next_tile >= 0) fetch(p, maps, next_tile, c + 2 - nch, warp, stg, bars, g + 2);
    }
    st.g = g + 1;
    // packed fp32 pairs: the epilogue warp is alone on its scheduler, so its time is instructions x latency
    const P2 kLn2 = p2(0.6931471805599453f), m2 = p2(r.m), linv2 = p2(r.linv), nrt2 = p2(-r.rt),
             aS2 = p2(r.aS), c22 = p2(r.c2);
    P2 u2 = p2(0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {                      // two halves of 32 columns
      uint32_t o1[16], o2[16];
      const uint32_t* tv = h == 0 ? v.lo : v.hi;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 pp = __half22float2(*reinterpret_cast<const __half2*>(&pw[16 * h + (i >> 1)]));
        const P2 pt = p2(pp.x, pp.y);
        const P2 lg = p2(__log2f(fmaxf(pp.x, 5.9604645e-8f)), __log2f(fmaxf(pp.y, 5.9604645e-8f)));
        const P2 t = p2(__uint_as_float(tv[i]), __uint_as_float(tv[i + 1]));
        const P2 s = p2_fma(lg, kLn2, m2);             // s = m + ln P~
        const P2 pr = p2_mul(pt, linv2);               // p = P~ / l  (exactly 0 where P~ underflowed)
        const P2 e = p2_fma(nrt2, t, s);               // s - (r/tau) T
        const P2 x = p2_mul(aS2, pr);
        o1[i >> 1] = p2_pack_h2(p2_fma(x, e, x));      // aS p (1 + e)
        o2[i >> 1] = p2_pack_h2(p2_mul(c22, pr));      // -aS r p
        u2 = p2_fma(p2_mul(pr, e), s, u2);             // dL/dlog tau = -a tau sum p e s
      }
      if (h == 0) {
        // the output boxes were last read by the bulk stores of the previous chunk
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (RZ_EXP_D2 & 8) {        // keep the arithmetic alive without the staging stores
          if (o1[4 * j] == 0x12345678u && o2[4 * j + 3] == 0x9abcdef0u) u += 1.f;
          continue;
        }
        sts_v4(stg + 8192 + stage_off(lane, 4 * h + j), o1[4 * j], o1[4 * j + 1], o1[4 * j + 2], o1[4 * j + 3]);
        sts_v4(stg + 12288 + stage_off(lane, 4 * h + j), o2[4 * j], o2[4 * j + 1], o2[4 * j + 2], o2[4 * j + 3]);
      }
    }
    {
      float ua, ub;
      p2_unpack(u2, ua, ub);
      u += ua + ub;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (!(RZ_EXP_D2 & 2)) tma_store_3d(&maps.c, stg + 8192, tok0 + c * 64, row0, b);       // W1 [B, N, Lp]
      if (!(RZ_EXP_D2 & 1)) tma_store_3d(&maps.c2, stg + 12288, tok0 + c * 64, row0, b);     // W2 [B, N, Lp]
      tma_store_commit();
    }
  }
  __device__ static void epilogue(const Params& p, const Maps& maps, int tile, int next_tile, uint32_t tmem,
                                  int warp, int lane, uint64_t* bars, State& st, uint8_t* epi_smem) {
    int b, mt, nt;
    decode(p, tile, b, mt, nt);
    const int row0 = mt * kBM + warp * 32;
    const int n = row0 + lane;
    const bool row_ok = n < p.N;
    const long long pi = (long long)b * p.N + (row_ok ? n : 0);
    const float inv_tau = p.log_tau != nullptr ? __expf(-__ldg(p.log_tau)) : p.inv_tau;
    const float tau = 1.0f / inv_tau;
    Row r;
    r.m = row_ok ? p.mref[pi] : 0.f;
    r.linv = row_ok ? 1.0f / p.lsum[pi] : 0.f;
    const float a = row_ok ? p.coef_a[pi] : 0.f;
    const float rr = row_ok ? p.coef_r[pi] : 0.f;
    r.aS = a * p.scale[0];
    r.c2 = -r.aS * rr;
    r.rt = rr * inv_tau;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t stg = smem_u32(epi_smem) + warp * kWarpStage;
    const int nch = tile_n(p, tile) / 64;              // 4, or 2 for the narrow last tile
    const int tok0 = nt * kBN;
    float u = 0.f;
    Cols64 va, vb;
    ld64(taddr, va);
#pragma unroll 1
    for (int c = 0; c < nch; c += 2) {
      wait64(va);
      ld64(taddr + (c + 1) * 64, vb);
      chunk(p, maps, va, tile, next_tile, c, nch, tok0, b, row0, warp, lane, r, stg, bars, st, u);
      wait64(vb);
      if (c + 2 < nch) ld64(taddr + (c + 2) * 64, va);
      chunk(p, maps, vb, tile, next_tile, c + 1, nch, tok0, b, row0, warp, lane, r, stg, bars, st, u);
    }
    // deterministic per-(tile, warp) partial of dL/dlog(tau_attn) = -a tau sum_l p e s
    u = rz::warp_sum(-a * tau * u);
    if (lane == 0) p.dtau_part[(long long)tile * 4 + warp] = u;
  }
};

// ------------------------------------------------------------------------------------------
// pass Q: dq[n, f] = (1/S) sum_b sum_l W1[b, n, l] k[b, l, f]
struct QParams {
  int B, N, Lp;
  int m_tiles;
  const float* scale;
  float* dq;                      // [N, 768] fp32
};

template <int C>
struct PassQ : PolicyBase {
  static constexpr int kCluster = C;   // prompt tiles (2i, 2i+1) share the token operand
  using Params = QParams;
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = false, kBMn = true, kTwoPhase = false;
  __host__ __device__ static int num_tiles(const Params& p) { return p.m_tiles * (kD / kBN); }
  __host__ __device__ static int k_steps(const Params& p) { return p.B * (p.Lp / kBK); }
  __device__ static void load(const Params& p, const Maps& m, int tile, int ks, uint8_t* a, uint8_t*,
                              uint8_t* bsm, uint64_t* bar, int rank) {
    const int mt = tile % p.m_tiles, ft = tile / p.m_tiles;
    const int per = p.Lp / kBK;
    const int b = ks / per, lc = ks - b * per;
    load_kmajor<C>(&m.a, bar, a, lc * kBK, mt * kBM, b);                               // W1 [B, N, Lp]
    load_mnmajor_shared<C>(&m.b, bar, bsm, ft * kBN, lc * kBK, b, kBN / 64, rank);     // k  [B, Lp, 768]
  }
  __device__ static void epilogue(const Params& p, const Maps&, int tile, int, uint32_t tmem, int warp,
                                  int lane, uint64_t*, State&, uint8_t*) {
    const int mt = tile % p.m_tiles, ft = tile / p.m_tiles;
    const int n = mt * kBM + warp * 32 + lane;
    const float inv_s = p.scale[1];
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* dst = p.dq + (long long)n * kD + ft * kBN;
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem + lane_base + c0, v);
      tmem_ld_wait();
      if (n < p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(dst + c0 + i) =
              make_float4(__uint_as_float(v[i]) * inv_s, __uint_as_float(v[i + 1]) * inv_s,
                          __uint_as_float(v[i + 2]) * inv_s, __uint_as_float(v[i + 3]) * inv_s);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// pass K: dk[b, l, f] = (1/S) sum_n ( W1[b, n, l] q[n, f] + W2[b, n, l] o[b, n, f] )
struct KParams {
  int B, N, Lp;
  int l_tiles, n_chunks;
  const float* scale;
  float* dk;                      // [B, Lp, 768] fp32
};

struct PassK : PolicyBase {
  using Params = KParams;
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = true, kBMn = true, kTwoPhase = true;
  __host__ __device__ static int num_tiles(const Params& p) { return p.B * p.l_tiles * (kD / kBN); }
  __host__ __device__ static int k_steps(const Params& p) { return 2 * p.n_chunks; }
  __device__ static void decode(const Params& p, int tile, int& b, int& lt, int& ft) {
    ft = tile % (kD / kBN);
    const int r = tile / (kD / kBN);
    lt = r % p.l_tiles;
    b = r / p.l_tiles;
  }
  __device__ static void load(const Params& p, const Maps& m, int tile, int ks, uint8_t* a, uint8_t*,
                              uint8_t* bsm, uint64_t* bar, int) {
    int b, lt, ft;
    decode(p, tile, b, lt, ft);
    if (ks < p.n_chunks) {
      load_mnmajor(&m.a, bar, a, lt * kBM, ks * kBK, b, kBM / 64);        // W1 [B, N, Lp] (M = l)
      load_mnmajor(&m.b, bar, bsm, ft * kBN, ks * kBK, 0, kBN / 64);      // q  [N, 768]
    } else {
      const int nc = ks - p.n_chunks;
      load_mnmajor(&m.a2, bar, a, lt * kBM, nc * kBK, b, kBM / 64);       // W2
      load_mnmajor(&m.b2, bar, bsm, ft * kBN, nc * kBK, b, kBN / 64);     // pooled [B, N, 768]
    }
  }
  __device__ static void epilogue(const Params& p, const Maps&, int tile, int, uint32_t tmem, int warp,
                                  int lane, uint64_t*, State&, uint8_t*) {
    int b, lt, ft;
    decode(p, tile, b, lt, ft);
    const int l = lt * kBM + warp * 32 + lane;
    const float inv_s = p.scale[1];
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* dst = p.dk + ((long long)b * p.Lp + l) * kD + ft * kBN;
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem + lane_base + c0, v);
      tmem_ld_wait();
      if (l < p.Lp) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(dst + c0 + i) =
              make_float4(__uint_as_float(v[i]) * inv_s, __uint_as_float(v[i + 1]) * inv_s,
                          __uint_as_float(v[i + 2]) * inv_s, __uint_as_float(v[i + 3]) * inv_s);
      }
    }
  }
};

// fixed-order sum of n partials: kSumBlocks blocks sum contiguous slices (each thread a strided sub-sequence,
// then a shared-memory tree), the last block to finish (ticket) adds the kSumBlocks slice sums in index order
constexpr int kSumBlocks = 128;
__global__ void __launch_bounds__(1024) sum_partials_kernel(const float* part, int n, float* out, float* slice,
                                                            unsigned int* ticket) {
  __shared__ float sh[1024];
  __shared__ bool last;
  const int per = (n + kSumBlocks - 1) / kSumBlocks;
  const int i0 = blockIdx.x * per, i1 = min(n, i0 + per);
  float a = 0.f;
  for (int i = i0 + threadIdx.x; i < i1; i += 1024) a += part[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    slice[blockIdx.x] = sh[0];
    __threadfence();
    last = atomicAdd(ticket, 1u) == kSumBlocks - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.f;
    for (int k = 0; k < kSumBlocks; ++k) t += __ldcg(slice + k);
    out[0] = t;
    *ticket = 0u;
  }
}

}  // namespace

extern "C" size_t rz_sim_bwd_workspace_bytes(int n_images, int n_text, int tokens_padded) {
  if (n_images <= 0 || n_text <= 0 || tokens_padded <= 0) return 0;
  const size_t pairs = (size_t)n_images * n_text;
  const size_t w = pairs * (size_t)tokens_padded * sizeof(__half);     // one of W1 / W2
  const size_t tiles = (size_t)n_images * ((n_text + 127) / 128) * (tokens_padded / 128);
  return 2 * w + 2 * pairs * sizeof(float) + 4 * tiles * sizeof(float) + 1024;   // + scale, amax, ticket, 128 slice sums
}

extern "C" int rz_sim_bwd(const void* k_f16, int n_images, int tokens, int tokens_padded,
                          const void* q_f16, int n_text, float inv_tau, const float* log_tau,
                          const float* z, const float* dz, long long ldz, const float* lse,
                          const float* onorm, const void* pooled_f16, const void* p_f16,
                          const float* mref, const float* lsum, const float* q_inv_norm, float* dq, float* dk,
                          float* dlog_tau, void* workspace, size_t workspace_bytes, void* stream) {
  if (!k_f16 || !q_f16 || !z || !dz || !onorm || !pooled_f16 || !dq || !dk || !dlog_tau ||
      !workspace)
    return RZ_ERR_INVALID;
  if (n_images <= 0 || n_text <= 0 || tokens <= 0 || tokens_padded < tokens || ldz < n_images)
    return RZ_ERR_INVALID;
  if (tokens_padded % 128 != 0) return RZ_ERR_INVALID;
  if (workspace_bytes < rz_sim_bwd_workspace_bytes(n_images, n_text, tokens_padded)) return RZ_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(dq) & 15) ||
      (reinterpret_cast<uintptr_t>(dk) & 15))
    return RZ_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = n_images, N = n_text, Lp = tokens_padded;
  const size_t pairs = (size_t)B * N;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* w1 = reinterpret_cast<__half*>(ws);
  __half* w2 = w1 + pairs * Lp;
  float* coef_a = reinterpret_cast<float*>(w2 + pairs * Lp);
  float* coef_r = coef_a + pairs;
  float* dtau_part = coef_r + pairs;
  const int m_tiles = (N + 127) / 128, n_tiles = Lp / 128;
  // CTA pairs need an even number of prompt tiles, and pay off only while the streamed P~ / W1
  // operands are small enough for their re-reads to hit L2 (see rz_sim_fwd_large.cu)
  const int C = (m_tiles % 2 == 0 && pairs * Lp * sizeof(__half) <= ((size_t)2 << 30)) ? 2 : 1;
  const char* exp_pair = getenv("RZ_EXP_PAIR");     // experiment: bit 0 = D2, bit 1 = Q as pairs regardless of size
  const int xp = exp_pair ? atoi(exp_pair) : 0;
  const int Cd = (m_tiles % 2 == 0 && (xp & 1)) ? 2 : C;
  const int Cq = (m_tiles % 2 == 0 && (xp & 2)) ? 2 : C;
  const int d_tiles = B * m_tiles * n_tiles;
  float* scale = dtau_part + 4 * (size_t)d_tiles;          // [2]
  unsigned int* amax = reinterpret_cast<unsigned int*>(scale + 2);
  unsigned int* sum_ticket = amax + 1;
  float* sum_slice = reinterpret_cast<float*>(amax + 2);                 // [kSumBlocks]

  RZ_CUDA_OK(cudaMemsetAsync(amax, 0, 2 * sizeof(unsigned int), s));        // amax and the ticket of sum_partials_kernel
  CoefParams cp;
  cp.g = dz; cp.z = z; cp.ldz = ldz; cp.onorm = onorm; cp.q_inv_norm = q_inv_norm; cp.coef_a = coef_a; cp.coef_r = coef_r;
  cp.amax_bits = amax; cp.B = B; cp.N = N;
  pair_coef_kernel<<<dim3((unsigned)((N + 31) / 32), (unsigned)((B + 31) / 32)), dim3(32, 8), 0, s>>>(cp);
  RZ_LAUNCH_OK();
  pick_scale_kernel<<<1, 1, 0, s>>>(amax, scale);
  RZ_LAUNCH_OK();
  rz_count_launch(2);

  // ---- pass D
  const bool have_p = p_f16 != nullptr && mref != nullptr && lsum != nullptr;
  if (have_p) {
    if (reinterpret_cast<uintptr_t>(p_f16) & 15) return RZ_ERR_ALIGNMENT;
    Maps m = {};
    if (!rz::make_map_3d_sw128(&m.a, pooled_f16, B, N, kD, kD * 2, (uint64_t)N * kD * 2, kBM)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.b, k_f16, B, Lp, kD, kD * 2, (uint64_t)Lp * kD * 2, 256 / Cd)) return RZ_ERR_CUDA;
    m.b2 = m.b;
    if (!rz::make_map_3d_sw128(&m.a2, p_f16, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, 32)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.c, w1, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, 32)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.c2, w2, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, 32)) return RZ_ERR_CUDA;
    D2Params p;
    p.B = B; p.N = N; p.L = tokens; p.Lp = Lp; p.m_tiles = m_tiles; p.n_tiles = (Lp + 255) / 256;
    p.inv_tau = inv_tau; p.log_tau = log_tau; p.mref = mref; p.lsum = lsum; p.coef_a = coef_a;
    p.coef_r = coef_r; p.scale = scale; p.dtau_part = dtau_part;
    const int t2 = B * m_tiles * p.n_tiles;
    int rc = Cd == 2 ? launch<PassD2<2>>(m, p, s) : launch<PassD2<1>>(m, p, s);
    if (rc != RZ_OK) return rc;
    sum_partials_kernel<<<kSumBlocks, 1024, 0, s>>>(dtau_part, 4 * t2, dlog_tau, sum_slice, sum_ticket);
    RZ_LAUNCH_OK();
    rz_count_launch();
  } else {
    if (!lse) return RZ_ERR_INVALID;
    Maps m = {};
    if (!rz::make_map_3d_sw128(&m.a, q_f16, 1, N, kD, kD * 2, (uint64_t)N * kD * 2, kBM)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.a2, pooled_f16, B, N, kD, kD * 2, (uint64_t)N * kD * 2, kBM)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.b, k_f16, B, Lp, kD, kD * 2, (uint64_t)Lp * kD * 2, 128 / C)) return RZ_ERR_CUDA;
    m.b2 = m.b;
    if (!rz::make_map_3d_sw128(&m.c, w1, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, 32)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.c2, w2, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, 32)) return RZ_ERR_CUDA;
    DParams p;
    p.B = B; p.N = N; p.L = tokens; p.Lp = Lp; p.m_tiles = m_tiles; p.n_tiles = n_tiles;
    p.inv_tau = inv_tau; p.log_tau = log_tau; p.lse = lse; p.coef_a = coef_a; p.coef_r = coef_r;
    p.scale = scale; p.w1 = w1; p.w2 = w2; p.dtau_part = dtau_part;
    int rc = C == 2 ? launch<PassD<2>>(m, p, s) : launch<PassD<1>>(m, p, s);
    if (rc != RZ_OK) return rc;
    sum_partials_kernel<<<kSumBlocks, 1024, 0, s>>>(dtau_part, 4 * d_tiles, dlog_tau, sum_slice, sum_ticket);
    RZ_LAUNCH_OK();
    rz_count_launch();
  }
  // ---- pass Q
  {
    Maps m = {};
    if (!rz::make_map_3d_sw128(&m.a, w1, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, kBM)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.b, k_f16, B, Lp, kD, kD * 2, (uint64_t)Lp * kD * 2, 64)) return RZ_ERR_CUDA;
    m.a2 = m.a; m.b2 = m.b;
    QParams p;
    p.B = B; p.N = N; p.Lp = Lp; p.m_tiles = m_tiles; p.scale = scale; p.dq = dq;
    int rc = Cq == 2 ? launch<PassQ<2>>(m, p, s) : launch<PassQ<1>>(m, p, s);
    if (rc != RZ_OK) return rc;
  }
  // ---- pass K
  {
    Maps m = {};
    if (!rz::make_map_3d_sw128(&m.a, w1, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, 64)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.a2, w2, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, 64)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.b, q_f16, 1, N, kD, kD * 2, (uint64_t)N * kD * 2, 64)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&m.b2, pooled_f16, B, N, kD, kD * 2, (uint64_t)N * kD * 2, 64)) return RZ_ERR_CUDA;
    KParams p;
    p.B = B; p.N = N; p.Lp = Lp; p.l_tiles = Lp / 128; p.n_chunks = (N + 63) / 64; p.scale = scale;
    p.dk = dk;
    int rc = launch<PassK>(m, p, s);
    if (rc != RZ_OK) return rc;
  }
  return RZ_OK;
}
