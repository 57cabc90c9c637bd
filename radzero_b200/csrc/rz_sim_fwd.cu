// K3-K6: the fused VL-CABS similarity forward on tcgen05 tensor cores.
//
//   S[l, n]  = <k_l, q_n> * scale          (similarity GEMM,        losses.py:219-221)
//   P        = softmax over the L tokens    (streamed, never stored, losses.py:222)
//   o_n      = sum_l P[l, n] k_l            (pooling GEMM,           losses.py:224)
//   Z[n, b]  = <q_n, o_n / |o_n|>           (pooled logit,           losses.py:226-233)
//
// One CTA owns (image b, block of NBLK prompts) and streams the image's tokens in tiles of
// 64.  Operands are the fp16 LayerNorm+L2-normalised rows rz_prep_rows wrote (K-major,
// SWIZZLE_128B via TMA).  Per tile:
//   * warp 5 (one elected thread) issues S = k_tile . q^T       (M=64 tokens, N=NBLK, K=768)
//     into a double-buffered TMEM accumulator,
//   * warps 0-3 read S from TMEM (one token per thread), optionally store the scores,
//     take the per-prompt tile maximum (CREDUX + shared memory), exponentiate against a
//     LAZY running maximum (rescaling the pooled accumulator only when the maximum grows by
//     more than e^8, FlashAttention-4 style) and write P as the fp16 K-major B operand,
//   * warp 5 issues O^T[768, NBLK] += k_tile^T . P   (M=128 features x 6, N=NBLK, K=64 tokens)
//     with the SAME shared-memory token tile re-read as an MN-major A operand, accumulating
//     in TMEM across all tiles of the image; each completed feature slab releases its part
//     of the token tile back to the TMA producer (warp 4).
// After the last tile the four epilogue warps reduce |o|^2 and <q, o> out of TMEM
// (warp-shuffle + shared-memory reductions) and emit Z, the log-sum-exp, |o| and
// (for the backward pass) the normalised pooled vectors.
//
// TMEM budget: 6*NBLK columns for O^T + 2*NBLK for S  (NBLK = 64 -> all 512 columns).
#include "rz_common.cuh"
#include "rz_tma.cuh"
#include "rz_umma.cuh"

namespace {

using namespace rz::umma;

constexpr int kD = RZ_HIDDEN;                 // 768
constexpr int kTok = 64;                      // tokens per tile
constexpr int kSlabs = kD / 128;              // 6 feature slabs of 128 (= pairs of 64-wide chunks)
constexpr int kPairBytes = 2 * kTok * 128;    // one slab of a token tile: 2 chunks [64 tok x 128 B]
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThresh = 8.0f;        // natural-log units; p <= e^8 fits fp16 comfortably
constexpr int kSoftmaxThreads = 128;
constexpr int kBaseThreads = 192;             // 4 softmax/epilogue warps + TMA warp + MMA warp
constexpr int kLoaderWarps = 24;              // fused-prep variant: warps that normalise raw rows

struct FwdParams {
  int B, L, Lp, N;
  float scale;
  const float* log_tau_scale;  // optional device scalar: scale = exp(-*log_tau_scale)
  const float* log_tau_z;      // optional device scalar: z_scale = exp(-*log_tau_z)
  const void* raw;             // fused-prep variant: raw tokens [B, L, 768]
  const float* gamma;          // LayerNorm weight / bias (fused-prep variant), may be NULL
  const float* beta;
  int l2;
  const float* q_inv_norm;     // optional [N]: multiplies Z (sim_op "dot": 1/|q|)
  float* scores;               // optional
  long long scores_sb, scores_sn;
  int drop_cls;
  float* z;                    // optional
  long long z_sn, z_sb;
  float z_scale; int z_sigmoid;
  float* lse;                  // optional [B, N]
  float* onorm;                // optional [B, N]
  __half* o_out;               // optional [B, N, 768]
  int n_blocks, items, tiles;
};

template <int NBLK>
struct Ctrl {
  uint64_t q_full;
  uint64_t k_full[2 * kSlabs];
  uint64_t k_empty[2 * kSlabs];
  uint64_t s_full[2];
  uint64_t p_full[2];
  uint64_t o_done[2];
  uint64_t tmem_free;
  uint32_t tmem_slot;
  int rescale_flag[2];
  float m_ref[NBLK];
  float alpha[NBLK];
  float linv[NBLK];
  float smax[4][NBLK];
};

template <int NBLK, int KSTAGES>
struct Cfg {
  static constexpr int kPairs = kSlabs * KSTAGES;
  static constexpr int kQBytes = NBLK * kD * 2;
  static constexpr int kKBytes = kPairs * kPairBytes;
  static constexpr int kPBuf = (NBLK * 128 < 1024) ? 1024 : NBLK * 128;
  // [q][k tiles][P x2 (the item epilogue reuses P as reduction scratch)][Ctrl]
  static constexpr int kSmem = kQBytes + kKBytes + 2 * kPBuf + (int)sizeof(Ctrl<NBLK>);
  static_assert(4 * 3 * NBLK * 4 <= 2 * kPBuf, "reduction scratch must fit in the P buffers");
  static constexpr int kTmemCols = 8 * NBLK;
  static constexpr int kSCol = 6 * NBLK;
};

template <int NBLK, int KSTAGES, bool STATS, int LOADERS, typename TIn>
__global__ void __launch_bounds__(kBaseThreads + 32 * LOADERS, 1)
sim_fwd_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap qmap,
               const FwdParams p_in) {
  FwdParams p = p_in;
  if (p.log_tau_scale != nullptr) p.scale = __expf(-__ldg(p.log_tau_scale));
  if (p.log_tau_z != nullptr) p.z_scale = __expf(-__ldg(p.log_tau_z));
  using C = Cfg<NBLK, KSTAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // no static shared memory in this kernel: the dynamic segment starts 1024-byte aligned
  // (checked, since SWIZZLE_128B operands need it and there is no room for a slack pad)
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* q_s = smem_raw;
  uint8_t* k_s = q_s + C::kQBytes;
  uint8_t* p_s = k_s + C::kKBytes;
  Ctrl<NBLK>* ctl = reinterpret_cast<Ctrl<NBLK>*>(p_s + 2 * C::kPBuf);
  float (*red)[3 * NBLK] = reinterpret_cast<float (*)[3 * NBLK]>(p_s);   // epilogue scratch

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.tiles;

  if (tid == 0) {
    mbar_init(&ctl->q_full, 1);
    for (int i = 0; i < C::kPairs; ++i) {
      mbar_init(&ctl->k_full[i], LOADERS > 0 ? LOADERS : 1);
      mbar_init(&ctl->k_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->s_full[i], 1);
      mbar_init(&ctl->p_full[i], kSoftmaxThreads);
      mbar_init(&ctl->o_done[i], 1);
    }
    mbar_init(&ctl->tmem_free, kSoftmaxThreads);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) { if (LOADERS == 0) prefetch_tmap(&kmap); prefetch_tmap(&qmap); }
    tmem_alloc(&ctl->tmem_slot, C::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_slot;

  if (warp == 4) {
    // ================================================================= TMA producer
    if (elect_one()) {
      int prev_pb = -1;
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int b = item / p.n_blocks, pb = item - b * p.n_blocks;
        if (pb != prev_pb) {
          // the previous item's MMAs and epilogue (which reads q) must be finished
          if (it > 0) mbar_wait(&ctl->tmem_free, (uint32_t)((it - 1) & 1));
          mbar_arrive_expect_tx(&ctl->q_full, (uint32_t)C::kQBytes);
          for (int c = 0; c < kD / 64; ++c)
            tma_load_2d(&qmap, &ctl->q_full, q_s + c * NBLK * 128, c * 64, pb * NBLK, kEvictLast);
          prev_pb = pb;
        }
        for (int j = 0; LOADERS == 0 && j < T; ++j) {
          const long long gt = (long long)it * T + j;
          for (int s = 0; s < kSlabs; ++s) {
            const long long g = gt * kSlabs + s;
            const int slot = (int)(g % C::kPairs);
            mbar_wait(&ctl->k_empty[slot], (uint32_t)(((g / C::kPairs) & 1) ^ 1));
            mbar_arrive_expect_tx(&ctl->k_full[slot], (uint32_t)kPairBytes);
            uint8_t* dst = k_s + slot * kPairBytes;
            const int row = b * p.Lp + j * kTok;
            tma_load_2d(&kmap, &ctl->k_full[slot], dst, (2 * s) * 64, row, kEvictNormal);
            tma_load_2d(&kmap, &ctl->k_full[slot], dst + kTok * 128, (2 * s + 1) * 64, row, kEvictNormal);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ================================================================= MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(64, NBLK, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(128, NBLK, 1, 0);
      const uint32_t q_addr = smem_u32(q_s), k_addr = smem_u32(k_s), p_addr = smem_u32(p_s);
      int prev_pb = -1, qcount = 0, it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int b = item / p.n_blocks, pb = item - b * p.n_blocks;
        (void)b;
        if (pb != prev_pb) {
          mbar_wait(&ctl->q_full, (uint32_t)(qcount & 1));
          ++qcount;
          prev_pb = pb;
        }
        if (it > 0) mbar_wait(&ctl->tmem_free, (uint32_t)((it - 1) & 1));
        tc_fence_after();
        auto issue_s = [&](int j) {
          const long long gt = (long long)it * T + j;
          const int buf = (int)(gt & 1);
          const uint32_t d = tmem_base + C::kSCol + buf * NBLK;
          for (int s = 0; s < kSlabs; ++s) {
            const long long g = gt * kSlabs + s;
            const int slot = (int)(g % C::kPairs);
            mbar_wait(&ctl->k_full[slot], (uint32_t)((g / C::kPairs) & 1));
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t ad = make_smem_desc(k_addr + slot * kPairBytes + h * (kTok * 128) + k4 * 32, 0, 1024);
                const uint64_t bd = make_smem_desc(q_addr + (2 * s + h) * NBLK * 128 + k4 * 32, 0, 1024);
                mma_f16_ss(d, ad, bd, idesc_s, (s | h | k4) ? 1u : 0u);
              }
            }
          }
          mma_commit(&ctl->s_full[buf]);
        };
        auto issue_o = [&](int j) {
          const long long gt = (long long)it * T + j;
          const int buf = (int)(gt & 1);
          mbar_wait(&ctl->p_full[buf], (uint32_t)((gt >> 1) & 1));
          tc_fence_after();
          for (int s = 0; s < kSlabs; ++s) {
            const long long g = gt * kSlabs + s;
            const int slot = (int)(g % C::kPairs);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              // A = k_tile^T: MN-major, two 64-feature blocks 8192 B apart, 16 tokens per step
              const uint64_t ad = make_smem_desc(k_addr + slot * kPairBytes + k4 * 2048, kTok * 128, 1024);
              const uint64_t bd = make_smem_desc(p_addr + buf * C::kPBuf + k4 * 32, 0, 1024);
              mma_f16_ss(tmem_base + s * NBLK, ad, bd, idesc_o, (j > 0 || k4 > 0) ? 1u : 0u);
            }
            mma_commit(&ctl->k_empty[slot]);
          }
          mma_commit(&ctl->o_done[buf]);
        };
        if (KSTAGES >= 2) {
          issue_s(0);
          for (int j = 0; j < T; ++j) {
            if (j + 1 < T) issue_s(j + 1);
            issue_o(j);
          }
        } else {
          for (int j = 0; j < T; ++j) { issue_s(j); issue_o(j); }
        }
      }
    }
    __syncwarp();
  } else if (LOADERS > 0 && warp >= 6) {
    // ================================================================= fused prep: raw rows ->
    // LayerNorm + L2 (fp32, registers) -> fp16 -> the swizzled K-major token tile
    const int lw = warp - 6;
    const TIn* raw = static_cast<const TIn*>(p.raw);
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int b = item / p.n_blocks;
      const TIn* img = raw + (long long)b * p.L * kD;
      for (int j = 0; j < T; ++j) {
        const long long gt = (long long)it * T + j;
        const int slot0 = (int)((gt * kSlabs) % C::kPairs);       // 6 consecutive slots, no wrap
        const uint32_t par = (uint32_t)((((gt * kSlabs) / C::kPairs) & 1) ^ 1);
        for (int sl = 0; sl < kSlabs; ++sl) mbar_wait(&ctl->k_empty[slot0 + sl], par);
        uint8_t* tile = k_s + slot0 * kPairBytes;
        // this lane's 4 consecutive features of slab jj live at: slab jj, half lane/16, byte 8*(lane%16)
        uint8_t* lane_base_ptr = tile + (lane >> 4) * (kTok * 128);
        const uint32_t byte_in_row = (uint32_t)(8 * (lane & 15));
        // one row per warp in flight, many warps: measured on B200, a warp does not overlap the
        // 128-bit loads of several rows, so HBM concurrency has to come from the warp count
#pragma unroll 1
        for (int r = lw; r < kTok; r += LOADERS) {
          float v[24];
          const int t = j * kTok + r;
          const bool ok = t < p.L;
          if (ok) {
            rz::RowLoad<TIn>::load(img + (long long)t * kD, lane, v);
            rz::ln_l2_row(v, p.gamma, p.beta, lane, RZ_LN_EPS, RZ_L2_EPS, p.l2 != 0);
          }
          const uint32_t o = rz::sw128_offset((uint32_t)r, byte_in_row);
#pragma unroll
          for (int jj = 0; jj < 6; ++jj) {
            uint2 w = make_uint2(0u, 0u);
            if (ok)
              w = make_uint2(rz::pack_half2(v[4 * jj], v[4 * jj + 1]),
                             rz::pack_half2(v[4 * jj + 2], v[4 * jj + 3]));
            *reinterpret_cast<uint2*>(lane_base_ptr + jj * kPairBytes + o) = w;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0)
          for (int sl = 0; sl < kSlabs; ++sl) mbar_arrive(&ctl->k_full[slot0 + sl]);
      }
    }
  } else if (warp < 4) {
    // ================================================================= softmax + epilogue warps
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool act = lane < 16;                 // M=64 accumulator: rows live in lanes 0-15 of each quarter
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int b = item / p.n_blocks, pb = item - b * p.n_blocks;
      const int n0 = pb * NBLK;
      if (tid < NBLK) ctl->m_ref[tid] = -INFINITY;
      float lpart[STATS ? NBLK : 1];
      if (STATS) {
#pragma unroll
        for (int c = 0; c < NBLK; ++c) lpart[c] = 0.f;
      }
      for (int j = 0; j < T; ++j) {
        const long long gt = (long long)it * T + j;
        const int buf = (int)(gt & 1);
        mbar_wait(&ctl->s_full[buf], (uint32_t)((gt >> 1) & 1));
        tc_fence_after();
        float s[NBLK];
#pragma unroll
        for (int c0 = 0; c0 < NBLK; c0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tmem_base + lane_base + C::kSCol + buf * NBLK + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) s[c0 + i] = __uint_as_float(r[i]);
        }
        const int tloc = warp * 16 + lane;       // token within the tile (valid when act)
        const int t = j * kTok + tloc;
        const bool valid = act && t < p.L;
#pragma unroll
        for (int c = 0; c < NBLK; ++c) s[c] = valid ? s[c] * p.scale : -INFINITY;
        if (p.scores != nullptr && valid && t >= p.drop_cls) {
          float* dst = p.scores + (long long)b * p.scores_sb + (long long)(t - p.drop_cls);
#pragma unroll
          for (int c = 0; c < NBLK; ++c)
            if (n0 + c < p.N) __stcs(dst + (long long)(n0 + c) * p.scores_sn, s[c]);
        }
#pragma unroll
        for (int c = 0; c < NBLK; ++c) {
          const float mx = warp_redux_max(s[c]);
          if (lane == 0) ctl->smax[warp][c] = mx;
        }
        if (tid == 0) ctl->rescale_flag[buf] = 0;
        named_bar_sync(1, kSoftmaxThreads);
        if (tid < NBLK) {
          const float mt = fmaxf(fmaxf(ctl->smax[0][tid], ctl->smax[1][tid]),
                                 fmaxf(ctl->smax[2][tid], ctl->smax[3][tid]));
          const float mo = ctl->m_ref[tid];
          float a = 1.0f;
          if (mt > mo + kRescaleThresh) {       // also true on the first tile (mo = -inf)
            a = (j == 0) ? 0.0f : exp2f((mo - mt) * kLog2e);
            ctl->m_ref[tid] = mt;
            if (j > 0) ctl->rescale_flag[buf] = 1;
          }
          ctl->alpha[tid] = a;
        }
        named_bar_sync(1, kSoftmaxThreads);
        // the P buffer was last read by the pooling MMAs of tile gt-2
        if (gt >= 2) mbar_wait(&ctl->o_done[buf], (uint32_t)(((gt - 2) >> 1) & 1));
        if (ctl->rescale_flag[buf] != 0) {
          // rare: the running maximum grew; scale the pooled accumulator columns by alpha
          mbar_wait(&ctl->o_done[buf ^ 1], (uint32_t)(((gt - 1) >> 1) & 1));
          tc_fence_after();
          for (int sl = 0; sl < kSlabs; ++sl) {
#pragma unroll
            for (int c0 = 0; c0 < NBLK; c0 += 16) {
              uint32_t r[16];
              const uint32_t ta = tmem_base + lane_base + sl * NBLK + c0;
              tmem_ld_x16(ta, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * ctl->alpha[c0 + i]);
              tmem_st_x16(ta, r);
            }
          }
          tmem_st_wait();
          if (STATS) {
#pragma unroll
            for (int c = 0; c < NBLK; ++c) lpart[c] *= ctl->alpha[c];
          }
        }
        {
          uint8_t* pb_s = p_s + buf * C::kPBuf;
#pragma unroll
          for (int c = 0; c < NBLK; ++c) {
            const float pv = valid ? exp2f((s[c] - ctl->m_ref[c]) * kLog2e) : 0.0f;
            if (STATS) lpart[c] += pv;
            if (act)
              *reinterpret_cast<__half*>(pb_s + rz::sw128_offset((uint32_t)c, (uint32_t)(2 * tloc))) =
                  __float2half_rn(pv);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&ctl->p_full[buf]);
      }
      // ---------------------------------------------------------------- item epilogue
      {
        const long long last = (long long)it * T + (T - 1);
        mbar_wait(&ctl->o_done[last & 1], (uint32_t)((last >> 1) & 1));
        tc_fence_after();
      }
      const int floc = warp * 32 + lane;           // feature within a slab (TMEM lane)
#pragma unroll 1
      for (int c0 = 0; c0 < NBLK; c0 += 16) {
        float osq[16], qo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { osq[i] = 0.f; qo[i] = 0.f; }
        for (int sl = 0; sl < kSlabs; ++sl) {
          uint32_t r[16];
          tmem_ld_x16(tmem_base + lane_base + sl * NBLK + c0, r);
          tmem_ld_wait();
          const int f = sl * 128 + floc;
          const uint8_t* qc = q_s + (f >> 6) * NBLK * 128;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float o = __uint_as_float(r[i]);
            const float qv = __half2float(*reinterpret_cast<const __half*>(
                qc + rz::sw128_offset((uint32_t)(c0 + i), (uint32_t)(2 * (f & 63)))));
            osq[i] = fmaf(o, o, osq[i]);
            qo[i] = fmaf(qv, o, qo[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = rz::warp_sum(osq[i]);
          const float q2 = rz::warp_sum(qo[i]);
          if (lane == 0) { red[warp][c0 + i] = a; red[warp][NBLK + c0 + i] = q2; }
        }
      }
      if (STATS) {
#pragma unroll
        for (int c = 0; c < NBLK; ++c) {
          const float l = rz::warp_sum(lpart[c]);
          if (lane == 0) red[warp][2 * NBLK + c] = l;
        }
      }
      named_bar_sync(1, kSoftmaxThreads);
      if (tid < NBLK) {
        const int n = n0 + tid;
        const float osq = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
        const float qo = red[0][NBLK + tid] + red[1][NBLK + tid] + red[2][NBLK + tid] +
                         red[3][NBLK + tid];
        float l = 1.0f;
        if (STATS)
          l = red[0][2 * NBLK + tid] + red[1][2 * NBLK + tid] + red[2][2 * NBLK + tid] +
              red[3][2 * NBLK + tid];
        const float linv = 1.0f / l;
        ctl->linv[tid] = linv;
        if (n < p.N) {
          const float on = sqrtf(osq);
          // F.normalize(pooled): o / max(|o|, eps) with o = O / l          losses.py:227
          float z = (qo * linv) / fmaxf(on * linv, RZ_L2_EPS);
          if (p.q_inv_norm != nullptr) z *= p.q_inv_norm[n];
          if (p.z != nullptr) {
            float zo = z * p.z_scale;
            if (p.z_sigmoid) zo = 1.0f / (1.0f + __expf(-zo));
            p.z[(long long)n * p.z_sn + (long long)b * p.z_sb] = zo;
          }
          if (STATS) {
            if (p.lse != nullptr) p.lse[(long long)b * p.N + n] = ctl->m_ref[tid] + logf(l);
            if (p.onorm != nullptr) p.onorm[(long long)b * p.N + n] = on * linv;
          }
        }
      }
      named_bar_sync(1, kSoftmaxThreads);
      if (STATS && p.o_out != nullptr) {
        for (int sl = 0; sl < kSlabs; ++sl) {
          const int f = sl * 128 + floc;
#pragma unroll
          for (int c0 = 0; c0 < NBLK; c0 += 16) {
            uint32_t r[16];
            tmem_ld_x16(tmem_base + lane_base + sl * NBLK + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = n0 + c0 + i;
              if (n < p.N)
                p.o_out[((long long)b * p.N + n) * kD + f] =
                    __float2half_rn(__uint_as_float(r[i]) * ctl->linv[c0 + i]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->tmem_free);
      named_bar_sync(1, kSoftmaxThreads);   // m_ref / linv are rewritten by the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int NBLK, int KSTAGES, int LOADERS, typename TIn>
int launch_fwd(const CUtensorMap& kmap, const CUtensorMap& qmap, FwdParams p, bool stats,
               cudaStream_t s) {
  using C = Cfg<NBLK, KSTAGES>;
  p.n_blocks = (p.N + NBLK - 1) / NBLK;
  p.items = p.B * p.n_blocks;
  p.tiles = (p.L + kTok - 1) / kTok;
  const int grid = p.items < rz_sm_count() ? p.items : rz_sm_count();
  const int threads = kBaseThreads + 32 * LOADERS;
  if (stats) {
    RZ_CUDA_OK(cudaFuncSetAttribute(sim_fwd_kernel<NBLK, KSTAGES, true, LOADERS, TIn>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    sim_fwd_kernel<NBLK, KSTAGES, true, LOADERS, TIn><<<grid, threads, C::kSmem, s>>>(kmap, qmap, p);
  } else {
    RZ_CUDA_OK(cudaFuncSetAttribute(sim_fwd_kernel<NBLK, KSTAGES, false, LOADERS, TIn>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    sim_fwd_kernel<NBLK, KSTAGES, false, LOADERS, TIn><<<grid, threads, C::kSmem, s>>>(kmap, qmap, p);
  }
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

int fill_common(FwdParams& p, int n_images, int tokens, int tokens_padded, int n_text, float scale,
                const float* log_tau_scale, const float* q_inv_norm, float* scores,
                long long scores_stride_image, long long scores_stride_text, int drop_cls, float* z,
                long long z_stride_text, long long z_stride_image, float z_scale,
                const float* log_tau_z, int z_sigmoid, float* lse, float* onorm, void* pooled_f16) {
  if (n_images <= 0 || n_text <= 0 || tokens <= 0 || tokens_padded < tokens) return RZ_ERR_INVALID;
  if (tokens_padded % kTok != 0) return RZ_ERR_INVALID;
  if (drop_cls != 0 && drop_cls != 1) return RZ_ERR_INVALID;
  if ((long long)n_images * tokens_padded >= (1ll << 31)) return RZ_ERR_UNSUPPORTED;
  p.B = n_images; p.L = tokens; p.Lp = tokens_padded; p.N = n_text; p.scale = scale;
  p.log_tau_scale = log_tau_scale; p.log_tau_z = log_tau_z;
  p.raw = nullptr; p.gamma = nullptr; p.beta = nullptr; p.l2 = 1;
  p.q_inv_norm = q_inv_norm;
  p.scores = scores; p.scores_sb = scores_stride_image; p.scores_sn = scores_stride_text;
  p.drop_cls = drop_cls;
  p.z = z; p.z_sn = z_stride_text; p.z_sb = z_stride_image;
  p.z_scale = z_scale; p.z_sigmoid = z_sigmoid;
  p.lse = lse; p.onorm = onorm; p.o_out = static_cast<__half*>(pooled_f16);
  p.n_blocks = p.items = p.tiles = 0;
  return RZ_OK;
}

}  // namespace

extern "C" int rz_sim_fwd(const void* k_f16, int n_images, int tokens, int tokens_padded,
                          const void* q_f16, int n_text, float scale, const float* log_tau_scale,
                          const float* q_inv_norm, float* scores, long long scores_stride_image,
                          long long scores_stride_text, int drop_cls, float* z,
                          long long z_stride_text, long long z_stride_image, float z_scale,
                          const float* log_tau_z, int z_sigmoid, float* lse, float* onorm,
                          void* pooled_f16, void* stream) {
  if (k_f16 == nullptr || q_f16 == nullptr) return RZ_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(k_f16) & 15) || (reinterpret_cast<uintptr_t>(q_f16) & 15))
    return RZ_ERR_ALIGNMENT;
  FwdParams p;
  int rc = fill_common(p, n_images, tokens, tokens_padded, n_text, scale, log_tau_scale, q_inv_norm,
                       scores, scores_stride_image, scores_stride_text, drop_cls, z, z_stride_text,
                       z_stride_image, z_scale, log_tau_z, z_sigmoid, lse, onorm, pooled_f16);
  if (rc != RZ_OK) return rc;
  CUtensorMap kmap, qmap;
  if (!rz::make_map_2d_sw128(&kmap, k_f16, (uint64_t)n_images * tokens_padded, kD, kD * 2, kTok))
    return RZ_ERR_CUDA;
  const bool stats = (lse != nullptr || onorm != nullptr || pooled_f16 != nullptr);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_text <= 16) {
    if (!rz::make_map_2d_sw128(&qmap, q_f16, (uint64_t)n_text, kD, kD * 2, 16)) return RZ_ERR_CUDA;
    return launch_fwd<16, 2, 0, float>(kmap, qmap, p, stats, s);
  } else if (n_text <= 32) {
    if (!rz::make_map_2d_sw128(&qmap, q_f16, (uint64_t)n_text, kD, kD * 2, 32)) return RZ_ERR_CUDA;
    return launch_fwd<32, 1, 0, float>(kmap, qmap, p, stats, s);
  }
  if (!rz::make_map_2d_sw128(&qmap, q_f16, (uint64_t)n_text, kD, kD * 2, 64)) return RZ_ERR_CUDA;
  return launch_fwd<64, 1, 0, float>(kmap, qmap, p, stats, s);
}
