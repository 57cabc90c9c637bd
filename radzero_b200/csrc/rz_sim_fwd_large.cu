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
  float* scores; long long scores_sb, scores_sn; int drop_cls;
  __half* p_out;               // [B, N, Lp]  exp(s - mref)
  float* mref;                 // [B, N]
  float* lsum;                 // [B, N]      sum_l exp(s - mref)
  float* lse;                  // optional [B, N] = mref + log(lsum)
};

struct PassS2 : PolicyBase {
  using Params = S2Params;
  struct State { float m, l; };
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = false, kBMn = false, kTwoPhase = false;
  static constexpr int kEpiSmem = 4 * 32 * 33 * 4;     // one 32x33 fp32 transpose buffer per warp
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
                              uint8_t* bsm, uint64_t* bar) {
    int b, mt, nt;
    decode(p, tile, b, mt, nt);
    load_kmajor(&m.a, bar, a, ks * kBK, mt * kBM, 0);        // q [N, 768]
    load_kmajor(&m.b, bar, bsm, ks * kBK, nt * kBN, b);      // k [B, Lp, 768] (rows >= Lp: zero fill)
  }
  __device__ static void epilogue(const Params& p, int tile, uint32_t tmem, int warp, int lane, float*,
                                  State& st, uint8_t* epi_smem) {
    int b, mt, nt;
    decode(p, tile, b, mt, nt);
    const int n = mt * kBM + warp * 32 + lane;
    const bool row_ok = n < p.N;
    const float scale = p.log_tau_scale != nullptr ? __expf(-__ldg(p.log_tau_scale)) : p.scale;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const long long pi = (long long)b * p.N + (row_ok ? n : 0);
    __half* prow = p.p_out + pi * p.Lp;
    float* tbuf = reinterpret_cast<float*>(epi_smem) + warp * (32 * 33);
    if (nt == 0) { st.m = -INFINITY; st.l = 0.f; }
    const int cols = tile_n(p, tile);
#pragma unroll 1
    for (int c0 = 0; c0 < cols; c0 += 32) {
      const int l0 = nt * kBN + c0;
      uint32_t v[32];
      tmem_ld_x32(tmem + lane_base + c0, v);
      tmem_ld_wait();
      float cmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float s = __uint_as_float(v[i]) * scale;
        v[i] = __float_as_uint(s);
        if (l0 + i < p.L) cmax = fmaxf(cmax, s);
      }
      if (p.scores != nullptr) {
        // warp transpose: lane = row in, lane = token out -> 128-byte coalesced row segments
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) tbuf[lane * 33 + i] = __uint_as_float(v[i]);
        __syncwarp();
        const int l = l0 + lane;
        const bool col_ok = l < p.L && l >= p.drop_cls;
        float* dst = p.scores + (long long)b * p.scores_sb + (long long)(l - p.drop_cls);
        const int row0 = mt * kBM + warp * 32;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          if (col_ok && row0 + rr < p.N) __stcs(dst + (long long)(row0 + rr) * p.scores_sn, tbuf[rr * 33 + lane]);
        }
      }
      if (cmax > st.m + kGrow) {          // always true for the first chunk (m = -inf)
        if (st.l > 0.f) {
          // rare: the maximum grew by more than e^10 -- rescale what this row has accumulated
          // and what it has already written (same thread wrote it: program order suffices)
          const float alpha = exp2f((st.m - cmax) * kLog2e);
          st.l *= alpha;
          if (row_ok) {
            for (int c = 0; c < l0; c += 8) {
              uint4 w = *reinterpret_cast<uint4*>(prow + c);
              __half2* h = reinterpret_cast<__half2*>(&w);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(h[j]);
                h[j] = __floats2half2_rn(f.x * alpha, f.y * alpha);
              }
              *reinterpret_cast<uint4*>(prow + c) = w;
            }
          }
        }
        st.m = cmax;
      }
      const float mb = st.m * kLog2e;
      uint32_t o[16];
      float lacc = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = (l0 + i < p.L) ? exp2f(fmaf(__uint_as_float(v[i]), kLog2e, -mb)) : 0.f;
        const float p1 = (l0 + i + 1 < p.L) ? exp2f(fmaf(__uint_as_float(v[i + 1]), kLog2e, -mb)) : 0.f;
        lacc += p0 + p1;
        o[i >> 1] = pack_h2(p0, p1);
      }
      st.l += lacc;
      if (row_ok) {
        uint4* d = reinterpret_cast<uint4*>(prow + l0);
        d[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d[1] = make_uint4(o[4], o[5], o[6], o[7]);
        d[2] = make_uint4(o[8], o[9], o[10], o[11]);
        d[3] = make_uint4(o[12], o[13], o[14], o[15]);
      }
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

struct PassPK : PolicyBase {
  using Params = PKParams;
  static constexpr int kBN = 256, kAccs = 1;
  static constexpr bool kAMn = false, kBMn = true, kTwoPhase = false;
  __host__ __device__ static int num_tiles(const Params& p) { return p.B * p.m_tiles * (kD / kBN); }
  __host__ __device__ static int k_steps(const Params& p) { return p.Lp / kBK; }
  __device__ static void decode(const Params& p, int tile, int& b, int& mt, int& ft) {
    ft = tile % (kD / kBN);
    const int r = tile / (kD / kBN);
    mt = r % p.m_tiles;
    b = r / p.m_tiles;
  }
  __device__ static void load(const Params& p, const Maps& m, int tile, int ks, uint8_t* a, uint8_t*,
                              uint8_t* bsm, uint64_t* bar) {
    int b, mt, ft;
    decode(p, tile, b, mt, ft);
    load_kmajor(&m.a, bar, a, ks * kBK, mt * kBM, b);                     // P [B, N, Lp]
    load_mnmajor(&m.b, bar, bsm, ft * kBN, ks * kBK, b, kBN / 64);        // k [B, Lp, 768]
  }
  __device__ static void epilogue(const Params& p, int tile, uint32_t tmem, int warp, int lane, float*, State&, uint8_t*) {
    int b, mt, ft;
    decode(p, tile, b, mt, ft);
    const int n = mt * kBM + warp * 32 + lane;
    const bool row_ok = n < p.N;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const long long pi = (long long)b * p.N + (row_ok ? n : 0);
    const __half* qrow = p.q + (long long)(row_ok ? n : 0) * kD + ft * kBN;
    __half* orow = p.pooled != nullptr ? p.pooled + pi * kD + ft * kBN : nullptr;
    const float linv = row_ok ? 1.0f / p.lsum[pi] : 0.f;
    float osq = 0.f, qo = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem + lane_base + c0, v);
      tmem_ld_wait();
      const uint4 q0 = *reinterpret_cast<const uint4*>(qrow + c0);
      const uint4 q1 = *reinterpret_cast<const uint4*>(qrow + c0 + 8);
      const uint32_t qw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float a0 = __uint_as_float(v[i]) * linv, a1 = __uint_as_float(v[i + 1]) * linv;
        const float2 qq = __half22float2(*reinterpret_cast<const __half2*>(&qw[i >> 1]));
        osq = fmaf(a0, a0, fmaf(a1, a1, osq));
        qo = fmaf(a0, qq.x, fmaf(a1, qq.y, qo));
        o[i >> 1] = pack_h2(a0, a1);
      }
      if (row_ok && orow != nullptr) {
        uint4* d = reinterpret_cast<uint4*>(orow + c0);
        d[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    if (row_ok) {
      float2* d = reinterpret_cast<float2*>(p.part) + (pi * (kD / kBN) + ft);
      *d = make_float2(osq, qo);
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

  Maps m;
  if (!rz::make_map_3d_sw128(&m.a, q_f16, 1, N, kD, kD * 2, (uint64_t)N * kD * 2, kBM)) return RZ_ERR_CUDA;
  if (!rz::make_map_3d_sw128(&m.b, k_f16, B, Lp, kD, kD * 2, (uint64_t)Lp * kD * 2, PassS2::kBN)) return RZ_ERR_CUDA;
  m.a2 = m.a; m.b2 = m.b;
  S2Params sp;
  sp.B = B; sp.N = N; sp.L = tokens; sp.Lp = Lp; sp.m_tiles = m_tiles;
  sp.n_tiles = (Lp + PassS2::kBN - 1) / PassS2::kBN;
  sp.scale = scale; sp.log_tau_scale = log_tau_scale;
  sp.scores = scores; sp.scores_sb = scores_stride_image; sp.scores_sn = scores_stride_text;
  sp.drop_cls = drop_cls; sp.p_out = pbuf; sp.mref = mref; sp.lsum = lsum; sp.lse = lse;
  {
    int rc = launch<PassS2>(m, sp, s);
    if (rc != RZ_OK) return rc;
  }
  if (!pool) return RZ_OK;
  {
    Maps mk;
    if (!rz::make_map_3d_sw128(&mk.a, pbuf, B, N, Lp, (uint64_t)Lp * 2, (uint64_t)N * Lp * 2, kBM)) return RZ_ERR_CUDA;
    if (!rz::make_map_3d_sw128(&mk.b, k_f16, B, Lp, kD, kD * 2, (uint64_t)Lp * kD * 2, 64)) return RZ_ERR_CUDA;
    mk.a2 = mk.a; mk.b2 = mk.b;
    PKParams kp;
    kp.B = B; kp.N = N; kp.Lp = Lp; kp.m_tiles = m_tiles; kp.q = static_cast<const __half*>(q_f16);
    kp.lsum = lsum; kp.pooled = static_cast<__half*>(pooled_f16); kp.part = part_o;
    int rc = launch<PassPK>(mk, kp, s);
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
