// K1-K6 for SMALL prompt sets (n_text <= 16: zero-shot classification, grounding, model_inference)
// in ONE kernel that reads the RAW vision tokens exactly once:
//
//   raw rows --TMA--> fp32/bf16 ring in smem --8 converter warps: LayerNorm + L2 in registers-->
//   fp16 token tile (16 tokens, SWIZZLE_128B) --tcgen05--> S^T[prompt, token] = q k^T in TMEM
//   --softmax warp: one thread per PROMPT row, lazy running maximum--> P (fp16, smem)
//   --tcgen05--> O^T[feature, prompt] += k_tile^T P   (the same smem tile re-read MN-major)
//
// Replaces nn.LayerNorm + F.normalize on the tokens (exp/cxr_pt/model/losses.py:90-91, 213),
// SimilarityLogit.forward (losses.py:187-240) and the glue of compute_logits (modeling.py:311-328).
// HBM-bound: the only large stream is the raw tokens.
//
// Work partition ("stream-K"): the (image, 16-token tile) sequence is cut into one contiguous,
// equally long range per CTA, so 256 images on 148 SMs -- or ONE image (model_inference) on 43 -- keep
// every SM busy.  A range that ends inside an image leaves a partial (reference maximum, unnormalised
// pooled sums) in the workspace; merge_partials_kernel combines the pieces of split images.
#include "rz_common.cuh"
#include "rz_tma.cuh"
#include "rz_umma.cuh"

namespace {

using namespace rz::umma;

constexpr int kD = RZ_HIDDEN;            // 768
constexpr int kNB = 16;                  // prompts per CTA (MMA N of the pooling GEMM)
#ifndef RZ_SMALL_TOKT
#define RZ_SMALL_TOKT 16
#endif
constexpr int kTokT = RZ_SMALL_TOKT;     // tokens per tile.  The S^T GEMM is a chain of 48 dependent MMAs per tile
                                         // whatever its width: 32 tokens per tile halve that chain per token
constexpr int kChunks = kD / 64;         // 12 K-chunks of 64 features (128 B of fp16)
constexpr int kSlabs = kD / 128;         // 6 feature slabs (MMA M = 128) of the pooled accumulator
constexpr int kGroup = 4;                // rows per ring group (one TMA transaction set)
// Ring of raw-token groups: kSlotsPerTeam slots per converter team (slot s belongs to team s % kTeams, so
// every waiter observes every phase of its slots' barriers).  Shapes measured in round 2 at C2 (0.229 ms
// baseline = 6 teams x 1 slot, 16-token tiles, 5 fp16 stages): 5 teams x 2 slots x 3 stages 0.246 ms;
// 6 x 2 x 5 (16-bit tokens only: shared memory) 0.219 ms; 32-token tiles with 2 / 3 stages 0.342 / 0.223 ms.
// Ablations (results wrong, time only): no LayerNorm arithmetic 0.209 ms, no exponentials 0.224 ms, 1/12 of
// the S-GEMM chain 0.228 ms -- no single stage bounds the kernel; it sits ~10 % above what its fixed per-row
// protocol (two barrier waits, 6 LDS, 6 STS, a proxy fence and two arrives per row) costs by itself.
#ifndef RZ_SMALL_TEAMS
#define RZ_SMALL_TEAMS 6
#endif
#ifndef RZ_SMALL_SPT
#define RZ_SMALL_SPT 1
#endif
#ifndef RZ_SMALL_STAGES
#define RZ_SMALL_STAGES 5
#endif
constexpr int kTeams = RZ_SMALL_TEAMS;   // converter teams of kGroup warps; team t takes groups t, t + kTeams, ...
constexpr int kSlotsPerTeam = RZ_SMALL_SPT;
constexpr int kRing = kTeams * kSlotsPerTeam;    // fp32: 10 x 4 x 3 KB = 120 KB
constexpr int kStages = RZ_SMALL_STAGES; // fp16 token tiles (24 KB each): 1-2 in the MMA chain, the rest being filled
static_assert(kTokT % 16 == 0 && kTokT <= 64, "the softmax warp reads its tile in tcgen05.ld.x16 halves");
static_assert(kRing % kTeams == 0, "parity barriers: every waiter must observe every phase of its slots");
// Rows per converter warp.  1: a team is 4 warps, gamma / beta are re-read (through L1) for every row -- 47 %
// of the LSU wavefronts of the kernel.  2: a team is 2 warps that convert two rows each with the lane's
// gamma / beta resident in registers; the CTA has 640 threads, so the register pool gives the converters 128.
#ifndef RZ_SMALL_RPW
#define RZ_SMALL_RPW 2
#endif
constexpr int kRpw = RZ_SMALL_RPW;
static_assert(kRpw == 1 || kRpw == 2 || kRpw == 4, "rows per converter warp");
constexpr int kTeamWarps = kGroup / kRpw;
constexpr int kConv = kTeamWarps * kTeams;   // converter warps: rows in flight hide the row latency
constexpr int kThreads = 256 + 32 * kConv;      // WG0 softmax/epilogue, WG1 TMA + MMA (+2 spare warps), converters
// register budget per warpgroup after setmaxnreg: the pool is what the CTA got at launch
// (kRegsLaunch registers x kThreads), so the sum over warpgroups must stay inside it
// (ptxas allocates the launch-bound maximum for a kernel that contains setmaxnreg: 64 at 1024 threads, 96 at 640)
constexpr int kRegsLaunch = (65536 / kThreads) / 8 * 8;
constexpr int kRegsEpi = 56, kRegsCtl = 24;
constexpr int kRegsConvFit = (kRegsLaunch * kThreads - (kRegsEpi + kRegsCtl) * 128) / (32 * kConv) / 8 * 8;
constexpr int kRegsConv = kRegsConvFit > 32 * (kRpw + 2) ? 32 * (kRpw + 2) : kRegsConvFit;   // 24 x + 12 packed / row, 48 parameters
static_assert(kRegsLaunch * kThreads <= 65536, "register file");
static_assert((kRegsEpi + kRegsCtl) * 128 + kRegsConv * 32 * kConv <= kRegsLaunch * kThreads, "register pool");
static_assert(kConv % 4 == 0, "setmaxnreg works on whole warpgroups");
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kGrow = 8.0f;            // the reference maximum moves when exceeded by this much

constexpr int kQBytes = kChunks * kNB * 128;            // 24 KB
constexpr int kKChunk = kTokT * 128;                    // 2 KB: [16 tokens x 64 features] fp16
constexpr int kKStage = kChunks * kKChunk;              // 24 KB
constexpr int kPBuf = 2048;                             // [16 prompts x 128 B] (32 B used)
constexpr int kTmemCols = kTokT <= 16 ? 128 : 256;      // 96 (O^T) + 2 x kTokT (S^T), a power of two
constexpr int kSCol = kSlabs * kNB;                     // 96
constexpr int kPartFloats = 2 * kNB + kD * kNB;         // m[16], l[16], O[768][16]

struct Params {
  int B, L, N, T;                  // T = tiles per image
  int total_tiles;
  float scale;
  const float* log_tau_scale;
  const float* log_tau_z;
  const float* gamma;
  const float* beta;
  int l2;
  const float* q_inv_norm;
  const __half* q;                 // [N, 768] (global copy, for the merge kernel)
  const float* text_raw;           // optional [N, 768] fp32: the prompts BEFORE LayerNorm + L2 -- normalised in the
  __half* q_out;                   // prologue of every CTA (no prep launch); CTA 0 also writes them here (= q)
  float* scores; long long scores_sb, scores_sn; int drop_cls;
  float* z; long long z_sn, z_sb; float z_scale; int z_sigmoid;
  float* part;                     // [ctas][2][kPartFloats]
  int n_ctas;
};

struct Ctrl {
  uint64_t q_full;
  uint64_t ring_full[kRing], ring_empty[kRing];
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t s_full[2], p_full[2], o_done[2];
  uint64_t o_free;
  uint32_t tmem_slot;
  // double-buffered by tile parity: warp 0 writes tile lt+1's flag / factors while slower warps may still
  // be reading tile lt's (they are separated by one named barrier only)
  int rescale_flag[2];
  float alpha[2][kNB];
  float m_fin[kNB], l_fin[kNB];
  float red[4][2 * kNB];
  uint32_t sink[kConv];            // scratch words of lds_returned (one per converter warp)
};

template <typename TIn>
struct Cfg {
  static constexpr int kRowBytes = kD * (int)sizeof(TIn);
  static constexpr int kBoxBytes = kGroup * 256 * (int)sizeof(TIn);      // [kGroup rows x 256 elements]
  static constexpr int kGroupBytes = 3 * kBoxBytes;
  static constexpr int kRingBytes = kRing * kGroupBytes;
  // [q][k stage 0][k stage 1][P x2][ring][ctrl]; the S-GEMM's A operand is declared 64 rows tall but
  // only 16 are real: rows 16-63 alias the bytes that follow (finite garbage, never read back)
  static constexpr int kSmem = kQBytes + kStages * kKStage + 2 * kPBuf + kRingBytes + (int)sizeof(Ctrl) + 1024;
};

// 32-bit on purpose (64-bit division is a register-hungry subroutine on the device); the host checks
// total * n < 2^31
__host__ __device__ __forceinline__ int range_begin(int total, int n, int c) {
  return (int)(((unsigned)total * (unsigned)c) / (unsigned)n);
}

// ring row -> 24 floats per lane (6 groups of 4 consecutive features at 4*(lane + 32 j))
template <typename TIn> struct RingRow;
template <> struct RingRow<float> {
  static __device__ __forceinline__ void load(uint32_t grp, int r, int lane, float (&x)[24]) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int e = 4 * (lane + 32 * j);
      const uint32_t a = grp + (uint32_t)((e >> 8) * (kGroup * 1024) + r * 1024 + (e & 255) * 4);
      // "memory": the loads must not be scheduled below the release of the ring slot that follows them
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(x[4 * j]), "=f"(x[4 * j + 1]), "=f"(x[4 * j + 2]), "=f"(x[4 * j + 3])
                   : "r"(a)
                   : "memory");
    }
  }
};
template <> struct RingRow<__nv_bfloat16> {
  static __device__ __forceinline__ void load(uint32_t grp, int r, int lane, float (&x)[24]) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int e = 4 * (lane + 32 * j);
      const uint32_t a = grp + (uint32_t)((e >> 8) * (kGroup * 512) + r * 512 + (e & 255) * 2);
      uint32_t u, v;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(u), "=r"(v) : "r"(a) : "memory");
      x[4 * j] = __uint_as_float(u << 16); x[4 * j + 1] = __uint_as_float(u & 0xffff0000u);
      x[4 * j + 2] = __uint_as_float(v << 16); x[4 * j + 3] = __uint_as_float(v & 0xffff0000u);
    }
  }
};
template <> struct RingRow<__half> {
  static __device__ __forceinline__ void load(uint32_t grp, int r, int lane, float (&x)[24]) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int e = 4 * (lane + 32 * j);
      const uint32_t a = grp + (uint32_t)((e >> 8) * (kGroup * 512) + r * 512 + (e & 255) * 2);
      uint32_t u, v;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(u), "=r"(v) : "r"(a) : "memory");
      const float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&u));
      const float2 f1 = __half22float2(*reinterpret_cast<__half2*>(&v));
      x[4 * j] = f0.x; x[4 * j + 1] = f0.y; x[4 * j + 2] = f1.x; x[4 * j + 3] = f1.y;
    }
  }
};

// Z of one (image, prompt) from |O|^2 and <q, O> (O unnormalised: the scale cancels)   losses.py:226-233
__device__ __forceinline__ void emit_z(const Params& p, float z_scale, int b, int n, float osq, float qo) {
  if (p.z == nullptr || n >= p.N) return;
  float z = qo / fmaxf(sqrtf(osq), RZ_L2_EPS);
  if (p.q_inv_norm != nullptr) z *= p.q_inv_norm[n];
  float zo = z * z_scale;
  if (p.z_sigmoid) zo = 1.0f / (1.0f + __expf(-zo));
  p.z[(long long)n * p.z_sn + (long long)b * p.z_sb] = zo;
}

template <typename TIn>
__global__ void __launch_bounds__(kThreads, 1)
sim_small_kernel(const __grid_constant__ CUtensorMap tokmap, const __grid_constant__ CUtensorMap qmap,
                 const __grid_constant__ Params p) {
  using C = Cfg<TIn>;
  // (the parameter block stays in the constant bank: no local copy)
  const float scale = p.log_tau_scale != nullptr ? __expf(-__ldg(p.log_tau_scale)) : p.scale;
  const float z_scale = p.log_tau_z != nullptr ? __expf(-__ldg(p.log_tau_z)) : p.z_scale;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = base;
  uint8_t* k_s = q_s + kQBytes;
  uint8_t* p_s = k_s + kStages * kKStage;
  uint8_t* ring = p_s + 2 * kPBuf;
  Ctrl* ctl = reinterpret_cast<Ctrl*>(ring + C::kRingBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g_begin = range_begin(p.total_tiles, (int)gridDim.x, (int)blockIdx.x);
  const int g_end = range_begin(p.total_tiles, (int)gridDim.x, (int)blockIdx.x + 1);
  const int T = p.T;

  if (tid == 0) {
    mbar_init(&ctl->q_full, p.text_raw != nullptr ? kConv : 1);
    for (int i = 0; i < kRing; ++i) { mbar_init(&ctl->ring_full[i], 1); mbar_init(&ctl->ring_empty[i], kTeamWarps); }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&ctl->k_full[i], kTokT / kRpw);                // one arrival per converter warp and group
      mbar_init(&ctl->k_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->s_full[i], 1);
      mbar_init(&ctl->p_full[i], 1);
      mbar_init(&ctl->o_done[i], 1);
    }
    mbar_init(&ctl->o_free, 128);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) { prefetch_tmap(&tokmap); prefetch_tmap(&qmap); }
    tmem_alloc(&ctl->tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_slot;

  if (warp >= 4 && warp < 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
  if (warp == 4) {
    // ================================================================= TMA producer
    if (elect_one()) {
      if (p.text_raw == nullptr) {
        mbar_arrive_expect_tx(&ctl->q_full, (uint32_t)kQBytes);
        for (int c = 0; c < kChunks; ++c)
          tma_load_2d(&qmap, &ctl->q_full, q_s + c * (kNB * 128), c * 64, 0, kEvictLast);
      }
      int rg = 0;                                          // ring groups issued by this CTA
      for (int g = g_begin; g < g_end; ++g) {
        const int b = g / T, j = g - b * T;
        for (int q4 = 0; q4 < kTokT / kGroup; ++q4, ++rg) {
          const int slot = (int)(rg % kRing);
          mbar_wait(&ctl->ring_empty[slot], (uint32_t)(((rg / kRing) & 1) ^ 1));
          mbar_arrive_expect_tx(&ctl->ring_full[slot], (uint32_t)C::kGroupBytes);
          uint8_t* dst = ring + slot * C::kGroupBytes;
          const int row = j * kTokT + q4 * kGroup;         // rows >= L of the image: zero fill
          for (int c = 0; c < 3; ++c)
            tma_load_3d(&tokmap, &ctl->ring_full[slot], dst + c * C::kBoxBytes, c * 256, row, b, kEvictFirst);
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ================================================================= MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(64, kTokT, 0, 0);     // S^T[prompt, token]
      constexpr uint32_t idesc_o = make_idesc_f16(128, kNB, 1, 0);      // O^T[feature, prompt]
      const uint32_t q_addr = smem_u32(q_s), k_addr = smem_u32(k_s), p_addr = smem_u32(p_s);
      mbar_wait(&ctl->q_full, 0);
      const int n_local = g_end - g_begin;
      // One thread issues 54 MMAs per 16-token tile: descriptors advance by 32-bit adds on the low
      // word (the high word -- stride, version, swizzle -- never changes).
      const uint64_t qd0 = make_smem_desc(q_addr, 0, 1024);
      const uint32_t qd_lo = (uint32_t)qd0, d_hi = (uint32_t)(qd0 >> 32);   // high word: SBO, version, swizzle
      auto mma = [](uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                    uint32_t acc) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d),
            "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
            : "memory");
      };
      auto issue_s = [&](int lt) {
        const int st = lt % kStages, sb = lt & 1;
        mbar_wait(&ctl->k_full[st], (uint32_t)((lt / kStages) & 1));
        tc_fence_after();
        const uint32_t d = tmem_base + kSCol + sb * kTokT;
        const uint32_t kd_lo = ((k_addr + st * kKStage) & 0x3FFFFu) >> 4;
        uint32_t alo = qd_lo, blo = kd_lo;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c, alo += (kNB * 128) >> 4, blo += kKChunk >> 4) {
          mma(d, alo, d_hi, blo, d_hi, idesc_s, c > 0 ? 1u : 0u);
          mma(d, alo + 2, d_hi, blo + 2, d_hi, idesc_s, 1u);      // +32 B = 16 features per K step
          mma(d, alo + 4, d_hi, blo + 4, d_hi, idesc_s, 1u);
          mma(d, alo + 6, d_hi, blo + 6, d_hi, idesc_s, 1u);
        }
        mma_commit(&ctl->s_full[sb]);
      };
      int seg = 0;
      int lt0 = 0;                                         // local index of the segment's first tile
      for (int g = g_begin; g < g_end;) {
        const int b = g / T, tb = g - b * T;
        (void)b;
        const int left = g_end - g;
        const int te = (T - tb) < left ? T : tb + left;
        const int nt = te - tb;
        if (seg > 0) {                                     // the previous segment's O has been read out
          mbar_wait(&ctl->o_free, (uint32_t)((seg - 1) & 1));
          tc_fence_after();
        }
        if (lt0 == 0) issue_s(0);
        for (int i = 0; i < nt; ++i) {
          const int lt = lt0 + i;
          if (lt + 1 < n_local) issue_s(lt + 1);           // next tile's scores (possibly next segment's)
          const int st = lt % kStages, sb = lt & 1;
          mbar_wait(&ctl->p_full[sb], (uint32_t)((lt >> 1) & 1));
          tc_fence_after();
          {
            // A = k_tile^T (MN-major): two 64-feature blocks kKChunk apart; the whole tile is one K step
            // (the leading-dimension offset = kKChunk sits in bits 16-29 of the LOW descriptor word)
            const uint32_t ka_lo = (((k_addr + st * kKStage) & 0x3FFFFu) >> 4) | ((uint32_t)(kKChunk >> 4) << 16);
            const uint32_t pb_lo = ((p_addr + sb * kPBuf) & 0x3FFFFu) >> 4;
            // K = the tile's tokens, 16 per MMA: A advances 16 token rows (2 KB) inside each chunk, B 32 bytes
#pragma unroll
            for (int kk = 0; kk < kTokT / 16; ++kk)
#pragma unroll
              for (int s = 0; s < kSlabs; ++s)
                mma(tmem_base + s * kNB, ka_lo + ((2 * s * kKChunk) >> 4) + kk * (2048 >> 4), d_hi, pb_lo + 2 * kk, d_hi,
                    idesc_o, (i > 0 || kk > 0) ? 1u : 0u);
          }
          mma_commit(&ctl->k_empty[st]);
          mma_commit(&ctl->o_done[sb]);
        }
        lt0 += nt;
        g += nt;
        ++seg;
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ================================================================= converters: raw row ->
    // LayerNorm + L2 (fp32 registers, ONE shuffle-reduction stage) -> fp16 -> swizzled K-major tile.
    // 24 warps in 6 teams of 4: team t converts ring groups t, t+6, ... (warp = row of the group).
    // A row is ~300 instructions of mostly dependent latency (~3k cycles); 24 rows in flight keep up
    // with HBM exactly as the stand-alone rz_prep_rows kernel does with its 24 resident warps.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsConv));
    const int team = (warp - 8) / kTeamWarps, w = (warp - 8) % kTeamWarps;
    const uint32_t ring_addr = smem_u32(ring);
    const bool ln = p.gamma != nullptr;
    rz::LnConsts lc;
    lc.sum_g2 = lc.sum_gb = lc.sum_b2 = 0.f;
    (void)lc;
#if RZ_SMALL_RPW >= 2
    rz::P2 g2[12], b2[12];
    rz::load_lane_pairs(ln ? p.gamma : nullptr, lane, 1.f, g2);
    rz::load_lane_pairs(ln ? p.beta : nullptr, lane, 0.f, b2);
#endif
    if (p.text_raw != nullptr) {
      // The prompts' own LayerNorm + L2 (compute_text_features + F.normalize, losses.py:163-164, 212): the 16
      // rows of the S-GEMM's A operand are written straight into its swizzled K-major tile -- the same layout
      // as a converted token tile (kNB = kTokT rows, chunks kKChunk apart) -- instead of being prepared by a
      // launch of their own and fetched by TMA.  Rows >= N are zero.
      static_assert(kNB * 128 == kKChunk || kTokT != 16, "q tile = one 16-token tile");
#if RZ_SMALL_RPW >= 2
      // rows (warp, warp + kConv) together: both rows' loads are in flight before the arithmetic starts, the
      // two dependent chains interleave, and gamma / beta are the register copies
      static_assert(2 * kConv >= kNB, "two prompt rows per converter warp cover the tile");
      {
        const int r0 = warp - 8, r1 = r0 + kConv;
        float v[2][24];
#pragma unroll
        for (int i = 0; i < 24; ++i) { v[0][i] = 0.f; v[1][i] = 0.f; }
        if (r0 < p.N) rz::RowLoad<float>::load(p.text_raw + (long long)r0 * kD, lane, v[0]);
        if (r1 < p.N) rz::RowLoad<float>::load(p.text_raw + (long long)r1 * kD, lane, v[1]);
        rz::ln_l2_rows_packed<2>(v, g2, b2, ln, RZ_LN_EPS, RZ_L2_EPS, p.l2 != 0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = h == 0 ? r0 : r1;
          if (r >= kNB) continue;
          uint8_t* tile = q_s + rz::sw128_offset((uint32_t)r, (uint32_t)(8 * (lane & 15))) + (lane >> 4) * (kNB * 128);
#pragma unroll
          for (int jj = 0; jj < 6; ++jj) {
            uint2 o = make_uint2(0u, 0u);                   // rows >= N stay zero (not LN(0) = beta)
            if (r < p.N) o = make_uint2(rz::pack_half2(v[h][4 * jj], v[h][4 * jj + 1]),
                                        rz::pack_half2(v[h][4 * jj + 2], v[h][4 * jj + 3]));
            *reinterpret_cast<uint2*>(tile + jj * (2 * kNB * 128)) = o;
            if (blockIdx.x == 0 && r < p.N)                // the merge kernel reads the rows from global memory
              *reinterpret_cast<uint2*>(p.q_out + (long long)r * kD + 4 * (lane + 32 * jj)) = o;
          }
        }
      }
#else
      for (int r = warp - 8; r < kNB; r += kConv) {
        float v[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) v[i] = 0.f;
        if (r < p.N) {
          rz::RowLoad<float>::load(p.text_raw + (long long)r * kD, lane, v);
          rz::ln_l2_row_packed(v, ln ? p.gamma : nullptr, ln ? p.beta : nullptr, lane, RZ_LN_EPS, RZ_L2_EPS, p.l2 != 0);
        }
        uint8_t* tile = q_s + rz::sw128_offset((uint32_t)r, (uint32_t)(8 * (lane & 15))) + (lane >> 4) * (kNB * 128);
#pragma unroll
        for (int jj = 0; jj < 6; ++jj) {
          const uint2 o = make_uint2(rz::pack_half2(v[4 * jj], v[4 * jj + 1]), rz::pack_half2(v[4 * jj + 2], v[4 * jj + 3]));
          *reinterpret_cast<uint2*>(tile + jj * (2 * kNB * 128)) = o;
          if (blockIdx.x == 0 && r < p.N)                  // the merge kernel reads the rows from global memory
            *reinterpret_cast<uint2*>(p.q_out + (long long)r * kD + 4 * (lane + 32 * jj)) = o;
        }
      }
#endif
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->q_full);
    }
    constexpr int kGpt = kTokT / kGroup;                   // ring groups per tile
    const int n_groups = (g_end - g_begin) * kGpt;
    for (int rg = team; rg < n_groups; rg += kTeams) {
      const int lt = rg / kGpt;                            // local tile
      const int q4 = rg - lt * kGpt;
      const int j = (g_begin + lt) % T;
      const int st = lt % kStages;
      mbar_wait(&ctl->k_empty[st], (uint32_t)(((lt / kStages) & 1) ^ 1));
      const int slot = (int)(rg % kRing);
      mbar_wait(&ctl->ring_full[slot], (uint32_t)((rg / kRing) & 1));
#if RZ_SMALL_RPW >= 2
      // kRpw rows of the group per warp, gamma / beta in registers
      const int r = q4 * kGroup + kRpw * w;                // first of this warp's token rows within the tile
      float v[kRpw][24];
#pragma unroll
      for (int h = 0; h < kRpw; ++h) RingRow<TIn>::load(ring_addr + slot * C::kGroupBytes, kRpw * w + h, lane, v[h]);
      // every lane's loads must have RETURNED before the slot goes back to the TMA producer (see
      // lds_returned in rz_umma.cuh: the arrive can overtake loads still queued in the LSU)
#pragma unroll
      for (int h = 0; h < kRpw; ++h)                       // one register of EVERY load of the row
        lds_returned(smem_u32(&ctl->sink[warp - 8]), __float_as_uint(v[h][3]), __float_as_uint(v[h][7]),
                     __float_as_uint(v[h][11]), __float_as_uint(v[h][15]), __float_as_uint(v[h][19]),
                     __float_as_uint(v[h][23]));
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->ring_empty[slot]);
      rz::ln_l2_rows_packed<kRpw>(v, g2, b2, ln, RZ_LN_EPS, RZ_L2_EPS, p.l2 != 0);
#pragma unroll
      for (int h = 0; h < kRpw; ++h) {
        const bool ok = j * kTokT + r + h < p.L;
        uint8_t* tile = k_s + st * kKStage + rz::sw128_offset((uint32_t)(r + h), (uint32_t)(8 * (lane & 15))) +
                        (lane >> 4) * kKChunk;
#pragma unroll
        for (int jj = 0; jj < 6; ++jj) {
          uint2 o = make_uint2(0u, 0u);
          if (ok) o = make_uint2(rz::pack_half2(v[h][4 * jj], v[h][4 * jj + 1]),
                                 rz::pack_half2(v[h][4 * jj + 2], v[h][4 * jj + 3]));
          *reinterpret_cast<uint2*>(tile + jj * (2 * kKChunk)) = o;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->k_full[st]);
#else
      const int r = q4 * kGroup + w;                       // token row within the tile
      float v[24];
      RingRow<TIn>::load(ring_addr + slot * C::kGroupBytes, w, lane, v);
#ifndef RZ_EXP_LATE_RELEASE
      // every lane's loads must have RETURNED before the slot goes back to the TMA producer (see
      // lds_returned in rz_umma.cuh: the arrive can overtake loads still queued in the LSU)
      lds_returned(smem_u32(&ctl->sink[warp - 8]), __float_as_uint(v[3]), __float_as_uint(v[7]),
                   __float_as_uint(v[11]), __float_as_uint(v[15]), __float_as_uint(v[19]), __float_as_uint(v[23]));
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->ring_empty[slot]);
#endif
      const bool ok = j * kTokT + r < p.L;
      rz::ln_l2_row_packed(v, ln ? p.gamma : nullptr, ln ? p.beta : nullptr, lane, RZ_LN_EPS, RZ_L2_EPS, p.l2 != 0);
      // lane's 4 features of group jj: feature 4*(lane + 32 jj) -> chunk (lane + 32 jj) / 16,
      // byte 8 * ((lane + 32 jj) % 16) of the token's 128-byte row
      uint8_t* tile = k_s + st * kKStage + rz::sw128_offset((uint32_t)r, (uint32_t)(8 * (lane & 15))) +
                      (lane >> 4) * kKChunk;
#pragma unroll
      for (int jj = 0; jj < 6; ++jj) {
        uint2 o = make_uint2(0u, 0u);
        if (ok) o = make_uint2(rz::pack_half2(v[4 * jj], v[4 * jj + 1]), rz::pack_half2(v[4 * jj + 2], v[4 * jj + 3]));
        *reinterpret_cast<uint2*>(tile + jj * (2 * kKChunk)) = o;   // unit = lane + 32 jj: chunk += 2
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->k_full[st]);
#ifdef RZ_EXP_LATE_RELEASE
      if (lane == 0) mbar_arrive(&ctl->ring_empty[slot]);
#endif
#endif
    }
  } else if (warp < 4) {
    // ================================================================= softmax (warp 0) + epilogue (warps 0-3)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    // Warp 0, lanes 0-15: one thread per prompt row of S^T keeps (m) in registers.  Warps 1-3 follow
    // the tile sequence through one named barrier per tile so that they can take part in the rare
    // rescale of the pooled accumulator, and read it out at the end of a segment.
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool row = warp == 0 && lane < kNB;              // this thread owns prompt `lane`
    int seg = 0;
    int lt0 = 0;
    for (int g = g_begin; g < g_end;) {
      const int b = g / T, tb = g - b * T;
      const int left = g_end - g;
      const int te = (T - tb) < left ? T : tb + left;
      const int nt = te - tb;
      float m = -INFINITY;
      for (int i = 0; i < nt; ++i) {
        const int lt = lt0 + i;
        const int st = lt & 1;
        const int l0 = (tb + i) * kTokT;
        const bool full = l0 + kTokT <= p.L;
        if (warp == 0) {
          mbar_wait(&ctl->s_full[st], (uint32_t)((lt >> 1) & 1));
          tc_fence_after();
          // pass A over the scores: maximum (+ the optional similarity map); the values are read
          // again from TMEM for the exponentials, which keeps this kernel inside 64 registers
          float cmax = -INFINITY;
#pragma unroll
          for (int hh = 0; hh < kTokT; hh += 16) {         // 16 columns at a time (register budget)
            uint32_t v[16];
            tmem_ld_x16(tmem_base + lane_base + kSCol + st * kTokT + hh, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float s = __uint_as_float(v[c]) * scale;
              if (full || l0 + hh + c < p.L) cmax = fmaxf(cmax, s);
              if (p.scores != nullptr) v[c] = __float_as_uint(s);
            }
            if (row && p.scores != nullptr && lane < p.N) {
              float* dst = p.scores + (long long)b * p.scores_sb + (long long)lane * p.scores_sn - p.drop_cls;
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const int tkn = l0 + hh + c;
                if (tkn < p.L && tkn >= p.drop_cls) __stcs(dst + tkn, __uint_as_float(v[c]));
              }
            }
          }
          float a = 1.0f;
          bool grow = false;
          if (row && cmax > m + kGrow) {                   // always true on a segment's first tile
            grow = i > 0;                                  // (whose pooling MMAs overwrite: no rescale)
            a = exp2f((m - cmax) * kLog2e);
            m = cmax;
          }
          const bool any = __any_sync(0xffffffffu, grow);
          if (any && lane < kNB) ctl->alpha[lt & 1][lane] = a;
          if (lane == 0) ctl->rescale_flag[lt & 1] = any ? 1 : 0;
        }
        named_bar_sync(1, 128);
        if (ctl->rescale_flag[lt & 1] != 0) {
          // rare: scale the pooled accumulator columns; it is quiescent once tile lt-1 has pooled
          mbar_wait(&ctl->o_done[st ^ 1], (uint32_t)(((lt - 1) >> 1) & 1));
          tc_fence_after();
          for (int sl = 0; sl < kSlabs; ++sl) {
            uint32_t r[16];
            const uint32_t ta = tmem_base + lane_base + sl * kNB;
            tmem_ld_x16(ta, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < kNB; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * ctl->alpha[lt & 1][c]);
            tmem_st_x16(ta, r);
          }
          tmem_st_wait();
          tc_fence_before();
          named_bar_sync(1, 128);
        }
        if (warp == 0) {
          // P (fp16) for the pooling GEMM; its buffer was last read by the pooling MMAs of tile lt-2
          if (lt >= 2) mbar_wait(&ctl->o_done[st], (uint32_t)(((lt - 2) >> 1) & 1));
          const float sl2 = scale * kLog2e, mb = m * kLog2e;
#pragma unroll
          for (int hh = 0; hh < kTokT; hh += 16) {
            uint32_t pk[8];
            uint32_t v[16];
            tmem_ld_x16(tmem_base + lane_base + kSCol + st * kTokT + hh, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              float e0 = exp2f(fmaf(__uint_as_float(v[c]), sl2, -mb));
              float e1 = exp2f(fmaf(__uint_as_float(v[c + 1]), sl2, -mb));
              if (!full) {
                if (l0 + hh + c >= p.L) e0 = 0.f;
                if (l0 + hh + c + 1 >= p.L) e1 = 0.f;
              }
              pk[c >> 1] = rz::pack_half2(e0, e1);
            }
            if (row) {
              uint8_t* prow = p_s + st * kPBuf;
#pragma unroll
              for (int u = 0; u < 2; ++u)
                *reinterpret_cast<uint4*>(prow + rz::sw128_offset((uint32_t)lane, (uint32_t)(2 * hh + 16 * u))) =
                    make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tc_fence_before(); mbar_arrive(&ctl->p_full[st]); }
        }
      }
      // ---------------------------------------------------------------- end of the segment
      {
        const int last = lt0 + nt - 1;
        mbar_wait(&ctl->o_done[last & 1], (uint32_t)((last >> 1) & 1));
        tc_fence_after();
      }
      const int floc = warp * 32 + lane;                   // feature within a slab (TMEM lane)
      const bool complete = tb == 0 && te == T;
      if (complete) {
#pragma unroll 1
        for (int h = 0; h < kNB; h += 8) {                 // 8 prompts at a time (register budget)
          float osq[8], qo[8];
#pragma unroll
          for (int n = 0; n < 8; ++n) { osq[n] = 0.f; qo[n] = 0.f; }
          for (int sl = 0; sl < kSlabs; ++sl) {
            uint32_t r[8];
            tmem_ld_x8(tmem_base + lane_base + sl * kNB + h, r);
            tmem_ld_wait();
            const int f = sl * 128 + floc;
            const uint8_t* qc = q_s + (f >> 6) * (kNB * 128);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              const float o = __uint_as_float(r[n]);
              const float qv = __half2float(*reinterpret_cast<const __half*>(
                  qc + rz::sw128_offset((uint32_t)(h + n), (uint32_t)(2 * (f & 63)))));
              osq[n] = fmaf(o, o, osq[n]);
              qo[n] = fmaf(qv, o, qo[n]);
            }
          }
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float a2 = rz::warp_sum(osq[n]);
            const float q2 = rz::warp_sum(qo[n]);
            if (lane == 0) { ctl->red[warp][h + n] = a2; ctl->red[warp][kNB + h + n] = q2; }
          }
        }
        named_bar_sync(1, 128);
        if (tid < kNB)
          emit_z(p, z_scale, b, tid, ctl->red[0][tid] + ctl->red[1][tid] + ctl->red[2][tid] + ctl->red[3][tid],
                 ctl->red[0][kNB + tid] + ctl->red[1][kNB + tid] + ctl->red[2][kNB + tid] +
                     ctl->red[3][kNB + tid]);
      } else {
        // a piece of a split image: (m, unnormalised O) for merge_partials_kernel
        float* part = p.part + ((long long)blockIdx.x * 2 + (seg == 0 ? 0 : 1)) * kPartFloats;
        if (row) part[lane] = m;
        for (int sl = 0; sl < kSlabs; ++sl) {
          uint32_t r[16];
          tmem_ld_x16(tmem_base + lane_base + sl * kNB, r);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(part + 2 * kNB + (long long)(sl * 128 + floc) * kNB);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            dst[u] = make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]),
                                 __uint_as_float(r[4 * u + 2]), __uint_as_float(r[4 * u + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->o_free);
      named_bar_sync(1, 128);          // ctl->red / alpha are rewritten by the next segment
      lt0 += nt;
      g += nt;
      ++seg;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}


// Pieces of images whose tiles were split over several CTAs -> Z.  One block per image.
__global__ void __launch_bounds__(kD)
merge_partials_kernel(const __grid_constant__ Params p) {
  const float z_scale = p.log_tau_z != nullptr ? __expf(-__ldg(p.log_tau_z)) : p.z_scale;
  const int b = blockIdx.x, n_ctas = p.n_ctas, T = p.T;
  const int G = p.total_tiles, gb = b * T, ge = gb + T;
  int c = (int)(((unsigned)gb * (unsigned)n_ctas) / (unsigned)G);
  while (c + 1 < n_ctas && range_begin(G, n_ctas, c + 1) <= gb) ++c;
  while (c > 0 && range_begin(G, n_ctas, c) > gb) --c;
  if (range_begin(G, n_ctas, c + 1) >= ge) return;        // the image lived in one CTA: already final
  __shared__ float m_s[kNB];
  __shared__ float red[kD / 32][2 * kNB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < kNB) {
    float m = -INFINITY;
    for (int cc = c; cc < n_ctas && range_begin(G, n_ctas, cc) < ge; ++cc) {
      const int k = range_begin(G, n_ctas, cc) >= gb ? 0 : 1;
      m = fmaxf(m, p.part[((long long)cc * 2 + k) * kPartFloats + tid]);
    }
    m_s[tid] = m;
  }
  __syncthreads();
  float osq[kNB], qo[kNB];
#pragma unroll
  for (int n = 0; n < kNB; ++n) { osq[n] = 0.f; qo[n] = 0.f; }
  {
    const int f = tid;                                     // one feature per thread
    float o[kNB];
#pragma unroll
    for (int n = 0; n < kNB; ++n) o[n] = 0.f;
    for (int cc = c; cc < n_ctas && range_begin(G, n_ctas, cc) < ge; ++cc) {
      const int k = range_begin(G, n_ctas, cc) >= gb ? 0 : 1;
      const float* part = p.part + ((long long)cc * 2 + k) * kPartFloats;
      const float4* src = reinterpret_cast<const float4*>(part + 2 * kNB + (long long)f * kNB);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 v = src[u];
        o[4 * u] = fmaf(v.x, exp2f((part[4 * u] - m_s[4 * u]) * kLog2e), o[4 * u]);
        o[4 * u + 1] = fmaf(v.y, exp2f((part[4 * u + 1] - m_s[4 * u + 1]) * kLog2e), o[4 * u + 1]);
        o[4 * u + 2] = fmaf(v.z, exp2f((part[4 * u + 2] - m_s[4 * u + 2]) * kLog2e), o[4 * u + 2]);
        o[4 * u + 3] = fmaf(v.w, exp2f((part[4 * u + 3] - m_s[4 * u + 3]) * kLog2e), o[4 * u + 3]);
      }
    }
#pragma unroll
    for (int n = 0; n < kNB; ++n) {
      const float qv = n < p.N ? __half2float(p.q[(long long)n * kD + f]) : 0.f;
      osq[n] = fmaf(o[n], o[n], osq[n]);
      qo[n] = fmaf(qv, o[n], qo[n]);
    }
  }
#pragma unroll
  for (int n = 0; n < kNB; ++n) {
    const float a2 = rz::warp_sum(osq[n]);
    const float q2 = rz::warp_sum(qo[n]);
    if (lane == 0) { red[warp][n] = a2; red[warp][kNB + n] = q2; }
  }
  __syncthreads();
  if (tid < kNB) {
    float a2 = 0.f, q2 = 0.f;
    for (int w = 0; w < kD / 32; ++w) { a2 += red[w][tid]; q2 += red[w][kNB + tid]; }
    emit_z(p, z_scale, b, tid, a2, q2);
  }
}

template <typename TIn>
int launch_small(const void* tokens_raw, CUtensorMapDataType dt, const CUtensorMap& qmap, Params p,
                 cudaStream_t s) {
  using C = Cfg<TIn>;
  rz::EncodeTiledFn fn = rz::encode_tiled_fn();
  if (fn == nullptr) return RZ_ERR_CUDA;
  CUtensorMap tokmap;
  {
    cuuint64_t dims[3] = {(cuuint64_t)kD, (cuuint64_t)p.L, (cuuint64_t)p.B};
    cuuint64_t strides[2] = {(cuuint64_t)kD * sizeof(TIn), (cuuint64_t)p.L * kD * sizeof(TIn)};
    cuuint32_t box[3] = {256, (cuuint32_t)kGroup, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&tokmap, dt, 3, const_cast<void*>(tokens_raw), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return RZ_ERR_CUDA;
  }
  RZ_CUDA_OK(cudaFuncSetAttribute(sim_small_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
  sim_small_kernel<TIn><<<p.n_ctas, kThreads, C::kSmem, s>>>(tokmap, qmap, p);
  RZ_LAUNCH_OK();
  rz_count_launch();
  if (p.z != nullptr && p.n_ctas > 1) {
    merge_partials_kernel<<<p.B, kD, 0, s>>>(p);
    RZ_LAUNCH_OK();
    rz_count_launch();
  }
  return RZ_OK;
}

}  // namespace

extern "C" size_t rz_sim_fwd_tokens_workspace_bytes(int n_images, int n_text) {
  if (n_images <= 0 || n_text <= 0) return 0;
  return (size_t)rz_sm_count() * 2 * kPartFloats * sizeof(float) + 256 + (size_t)kNB * kD * sizeof(__half);
}

extern "C" int rz_sim_fwd_tokens(const void* tokens_raw, int dtype, const float* gamma,
                                 const float* beta, int l2, int n_images, int tokens,
                                 const void* q_f16, int n_text, float scale,
                                 const float* log_tau_scale, const float* q_inv_norm, float* scores,
                                 long long scores_stride_image, long long scores_stride_text,
                                 int drop_cls, float* z, long long z_stride_text,
                                 long long z_stride_image, float z_scale, const float* log_tau_z,
                                 int z_sigmoid, const float* text_f32, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  if (tokens_raw == nullptr || (q_f16 == nullptr && text_f32 == nullptr) || workspace == nullptr) return RZ_ERR_INVALID;
  if (text_f32 != nullptr && (q_inv_norm != nullptr || (reinterpret_cast<uintptr_t>(text_f32) & 15))) return RZ_ERR_UNSUPPORTED;
  if ((gamma == nullptr) != (beta == nullptr)) return RZ_ERR_INVALID;
  if (n_images <= 0 || n_text <= 0 || tokens <= 0) return RZ_ERR_INVALID;
  if (drop_cls != 0 && drop_cls != 1) return RZ_ERR_INVALID;
  if (n_text > kNB) return RZ_ERR_UNSUPPORTED;   // larger prompt sets: rz_prep_rows + rz_sim_fwd
  if (workspace_bytes < rz_sim_fwd_tokens_workspace_bytes(n_images, n_text)) return RZ_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(tokens_raw) & 15) || (reinterpret_cast<uintptr_t>(q_f16) & 15) ||
      (reinterpret_cast<uintptr_t>(gamma) & 15) || (reinterpret_cast<uintptr_t>(beta) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15))
    return RZ_ERR_ALIGNMENT;
  Params p;
  p.B = n_images; p.L = tokens; p.N = n_text; p.T = (tokens + kTokT - 1) / kTokT;
  if ((long long)p.B * p.T * rz_sm_count() >= (1ll << 31)) return RZ_ERR_UNSUPPORTED;
  p.total_tiles = p.B * p.T;
  p.scale = scale; p.log_tau_scale = log_tau_scale; p.log_tau_z = log_tau_z;
  p.gamma = gamma; p.beta = beta; p.l2 = l2; p.q_inv_norm = q_inv_norm;
  p.q = static_cast<const __half*>(q_f16);
  p.text_raw = text_f32;
  p.q_out = nullptr;
  if (text_f32 != nullptr) {       // the normalised rows live behind the partials in the workspace
    p.q_out = reinterpret_cast<__half*>(static_cast<uint8_t*>(workspace) +
                                        (size_t)rz_sm_count() * 2 * kPartFloats * sizeof(float) + 256);
    p.q = p.q_out;
    q_f16 = p.q_out;
  }
  p.scores = scores; p.scores_sb = scores_stride_image; p.scores_sn = scores_stride_text;
  p.drop_cls = drop_cls;
  p.z = z; p.z_sn = z_stride_text; p.z_sb = z_stride_image; p.z_scale = z_scale; p.z_sigmoid = z_sigmoid;
  p.part = static_cast<float*>(workspace);
  // at least 4 tiles per CTA: amortises the prompt load and bounds the pieces per split image
  const int sms = rz_sm_count(), want = (p.total_tiles + 3) / 4;
  p.n_ctas = want < sms ? want : sms;
  CUtensorMap qmap;
  if (!rz::make_map_2d_sw128(&qmap, q_f16, (uint64_t)n_text, kD, kD * 2, kNB)) return RZ_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case RZ_F32: return launch_small<float>(tokens_raw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, qmap, p, s);
    case RZ_BF16: return launch_small<__nv_bfloat16>(tokens_raw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qmap, p, s);
    case RZ_F16: return launch_small<__half>(tokens_raw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, qmap, p, s);
    default: return RZ_ERR_INVALID;
  }
}
