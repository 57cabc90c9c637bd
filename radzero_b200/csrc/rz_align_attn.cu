// A2: multi-head self-attention of the AlignTransformer layers (transformers Dinov2SelfAttention,
// called from exp/cxr_pt/model/align_transformers.py:40) on tcgen05 tensor cores:
//
//   O[b, l, h] = softmax_l'(Q[b, l, h] . K[b, l', h]^T) V[b, l', h]       head dim 64, no mask
//
// Q, K, V are column blocks of the fused projection output qkv [B, L, 3 * H * 64] fp16 (Q already
// carries the 1/sqrt(64) scale: it is folded into the projection weights on the host, exactly, since
// it is a power of two, up to fp16 subnormals).  A work item is (image, head, 128 query rows): its CTA streams the image's keys in
// tiles of 128.  The kernel is persistent, two CTAs resident per SM (256 TMEM columns and ~113 KB of shared memory each) so
// that one CTA's exponentials overlap the other's MMAs.
//
//   producer warp: lane 0 loads Q and the K tiles, lane 1 the V tiles, into two-stage rings
//   MMA warp (one elected lane): S = Q K^T (M 128, N 128, K 64) into TMEM, and O += P_j V_j (M 128,
//           N 64, K 128; P read from TMEM as the A operand, the V tile as an MN-major B operand)
//           accumulated in TMEM over all key tiles -- S of tile j+1 is issued before waiting for P_j
//   8 softmax warps, TWO threads per query row (warps w and w + 4 share the TMEM lane quarter w % 4
//           and split the row's 128 keys 64 / 64).  A thread reads its 64 scores from TMEM ONCE (S is
//           released to the MMA warp right away), reduces them with FMNMX3, exchanges the half maximum
//           with its partner through shared memory, exponentiates in the log2 domain with packed
//           FFMA2 / FADD2, packs the fp16 probabilities in place and stores them to TMEM.  The running
//           reference maximum is LAZY (FlashAttention-4 style): it only moves when the tile maximum
//           exceeds it by more than 2^8, so P <= 256 stays comfortably inside fp16 and the accumulator
//           in TMEM is rescaled (tcgen05.ld / st by the row's own threads) only on those rare tiles.
// Keys >= L of the last tile are masked to probability 0; rows >= L are zero-filled by the TMA
// loads and clipped by the TMA store.
#ifndef RZ_ATTN_PINGPONG
#define RZ_ATTN_PINGPONG 0
#endif
#ifndef RZ_ATTN_EMU
#define RZ_ATTN_EMU 2
#endif
#include "rz_common.cuh"
#include "rz_tma.cuh"
#include "rz_umma.cuh"

namespace {

using namespace rz::umma;

constexpr int kHd = 64;                    // head dim
constexpr int kTile = 128;                 // query rows per CTA = keys per tile
constexpr int kTileBytes = kTile * 128;    // 128 rows x 64 halfs
constexpr int kSoftmaxWarps = 8;            // two per TMEM lane quarter: warp w and w + 4 split a row's 128 keys
constexpr int kGroupThreads = 32 * (kSoftmaxWarps + 2);   // one pipeline: 8 softmax warps + producer + MMA warp
constexpr int kThreads = 2 * kGroupThreads;              // TWO independent pipelines (groups) per CTA, see below
constexpr int kTmemCols = 256;             // per group: S [0,128)  O [128,192)  P [192,256) (fp16 pairs)
constexpr int kEmuPeriod = 8, kEmuCount = RZ_ATTN_EMU;   // RZ_ATTN_EMU of every 8 pairs: exp2 on the FMA pipe
constexpr float kLazy = 8.0f;              // log2 units: the reference maximum moves when exceeded by 2^8
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  int B, L, H, q_tiles, kv_tiles, items;     // items = B * H * q_tiles (image, head, query tile)
};

struct Ctrl {
  uint64_t q_full, q_empty;
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t s_full, s_empty, p_full;
  uint64_t o_full;
  uint32_t tmem_slot;
  float xmax[2][2][kTile];     // [tile parity][half][row]: the two threads of a row exchange their half maxima
  float xsum[2][kTile];        // [half][row]: ... and their half row sums at the end of an item
};

constexpr int kGroupSmem = 5 * kTileBytes + kTileBytes + 4096;   // Q, K x2, V x2, output staging (4 x 4 KB), Ctrl
constexpr int kSmem = 2 * kGroupSmem;
constexpr int kPingPongBar = 9;             // named barriers 9, 10: the groups take turns in the exponential phase
static_assert(sizeof(Ctrl) <= 4096, "control block");

__device__ __forceinline__ uint32_t p_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// packed fp32 pairs (sm_100 FFMA2 / FADD2): d = a * b + c, d += a
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  uint64_t A, B, C, D;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(C) : "f"(c));
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(C));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
  uint64_t A, D;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(D) : "f"(d0), "f"(d1));
  asm("add.rn.ftz.f32x2 %0, %0, %1;" : "+l"(D) : "l"(A));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (here the fp16 probabilities, two K elements per 32-bit
// column, row = TMEM lane) is read from tensor memory, so P never goes through shared memory
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 2^x for a PAIR of arguments (x <= 8) on the FMA / ALU pipes instead of the MUFU: Cody-Waite split
// x = n + f (round to nearest through the 1.5 * 2^23 magic add, f in [-0.5, 0.5]), a degree-3 minimax
// polynomial for 2^f (relative error 7.5e-5, below the fp16 rounding of the probabilities) and n added
// into the exponent field with one integer multiply-add.  The MUFU-bound exponential phase also floods
// the MIO queue that the TMEM loads / stores and barrier polls of the other warps go through, so moving
// part of the exponentials off the MUFU shortens both phases (FlashAttention-4 does the same).
__device__ __forceinline__ void exp2_pair_fma(float& x0, float& x1) {
  const float kMagic = 12582912.0f;                   // 1.5 * 2^23
  x0 = fmaxf(x0, -126.0f);                            // masked keys carry -inf
  x1 = fmaxf(x1, -126.0f);
  float t0 = x0, t1 = x1, n0, n1;
  fadd2(t0, t1, kMagic, kMagic);                      // integer part in the low mantissa bits
  n0 = t0; n1 = t1;
  fadd2(n0, n1, -kMagic, -kMagic);
  float f0, f1;
  ffma2(f0, f1, n0, n1, -1.0f, 0.f);                  // -n
  fadd2(f0, f1, x0, x1);                              // f = x - n
  float p0, p1;
  ffma2(p0, p1, f0, f1, 0.05517083778977394f, 0.24260935187339783f);
  {
    uint64_t P, F, C, D;
    asm("mov.b64 %0, {%1, %2};" : "=l"(P) : "f"(p0), "f"(p1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(F) : "f"(f0), "f"(f1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(C) : "f"(0.6932609677314758f));
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(P), "l"(F), "l"(C));
    asm("mov.b64 %0, {%1, %1};" : "=l"(C) : "f"(0.9999281764030457f));
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(D), "l"(F), "l"(C));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(P));
  }
  x0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  x1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}
__device__ __forceinline__ void tmem_ld_wait_x16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15])
               :
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
attn_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ CUtensorMap out_map,
            const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  // One CTA per SM holds TWO complete pipelines ("groups" of 10 warps, each with its own shared memory,
  // barriers and 256 TMEM columns) working on two neighbouring items; 16 softmax warps per SM keep the
  // MUFU / FMA pipes busier than two independent 2-CTA/SM launches of one pipeline did.  Named barriers
  // 9 / 10 can make the groups take turns in the exponential phase (the FlashAttention-3 ping-pong,
  // RZ_ATTN_PINGPONG=1); measured on B200 the free-running groups are ~5 % faster, so it is off.
  const int grp = (int)threadIdx.x / kGroupThreads;
  const int tid = (int)threadIdx.x % kGroupThreads, warp = tid >> 5, lane = tid & 31;
  uint8_t* gbase = smem + grp * kGroupSmem;
  uint8_t* q_s = gbase;
  uint8_t* k_s = q_s + kTileBytes;
  uint8_t* v_s = k_s + 2 * kTileBytes;
  uint8_t* p_s = v_s + 2 * kTileBytes;
  Ctrl* ctl = reinterpret_cast<Ctrl*>(p_s + kTileBytes);
  Ctrl* ctl0 = reinterpret_cast<Ctrl*>(smem + 6 * kTileBytes);

  const int T = p.kv_tiles;
  const int width = p.H * kHd;
  // PERSISTENT: CTA c walks item PAIRS c, c + gridDim.x, ...; group g takes item 2 * pair + g (the two
  // items are neighbouring query tiles, so they share their K / V in L2).  Both groups always run the
  // same number of items -- an odd item count gives group 1 a duplicate of the last item whose output
  // is not stored -- which keeps the ping-pong barriers balanced.  Every pipeline barrier keeps counting
  // across items (G = running key-tile index, it = running item index), so the next item's Q load
  // and first S = Q K^T overlap the softmax of this item's last tile and its output store.
  const int n_pairs = (p.items + 1) / 2;
  const int my_items = (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto decode = [&](int it, int& b, int& h, int& qt) {
    int item = 2 * ((int)blockIdx.x + it * (int)gridDim.x) + grp;
    const bool real = item < p.items;
    if (!real) item = p.items - 1;
    qt = item % p.q_tiles;
    const int bh = item / p.q_tiles;
    h = bh % p.H;
    b = bh / p.H;
    return real;
  };

  if (tid == 0) {
    mbar_init(&ctl->q_full, 1);
    mbar_init(&ctl->q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->k_full[i], 1); mbar_init(&ctl->k_empty[i], 1);
      mbar_init(&ctl->v_full[i], 1); mbar_init(&ctl->v_empty[i], 1);
    }
    mbar_init(&ctl->o_full, 1);
    mbar_init(&ctl->s_full, 1);
    mbar_init(&ctl->s_empty, 32 * kSoftmaxWarps);
    mbar_init(&ctl->p_full, 32 * kSoftmaxWarps);
    fence_barrier_init();
  }
  if (warp == kSoftmaxWarps && grp == 0) {
    if (lane == 0) { prefetch_tmap(&qkv_map); prefetch_tmap(&out_map); }
    tmem_alloc(&ctl0->tmem_slot, 2 * kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_all = ctl0->tmem_slot;
  const uint32_t tmem_base = tmem_all + grp * kTmemCols;

  if (warp == kSoftmaxWarps) {
    // Two producer lanes with their own loops: lane 0 loads Q and the K tiles, lane 1 the V tiles.  (One
    // thread doing K(j), V(j), K(j+1), ... in order would hold K(j+1) back until the P V product of
    // tile j-2 has released its V stage, which leaves the next S = Q K^T about one TMA latency short.)
    if (lane == 0) {
      int G = 0;
      for (int it = 0; it < my_items; ++it) {
        int b, h, qt;
        decode(it, b, h, qt);
        mbar_wait_relaxed(&ctl->q_empty, (uint32_t)((it & 1) ^ 1));   // all S MMAs of the previous item issued + done
        mbar_arrive_expect_tx(&ctl->q_full, kTileBytes);
        tma_load_3d(&qkv_map, &ctl->q_full, q_s, h * kHd, qt * kTile, b, kEvictNormal);
        for (int j = 0; j < T; ++j, ++G) {
          const int st = G & 1;
          mbar_wait_relaxed(&ctl->k_empty[st], (uint32_t)(((G >> 1) & 1) ^ 1));
          mbar_arrive_expect_tx(&ctl->k_full[st], kTileBytes);
          tma_load_3d(&qkv_map, &ctl->k_full[st], k_s + st * kTileBytes, width + h * kHd, j * kTile, b, kEvictNormal);
        }
      }
    } else if (lane == 1) {
      int G = 0;
      for (int it = 0; it < my_items; ++it) {
        int b, h, qt;
        decode(it, b, h, qt);
        for (int j = 0; j < T; ++j, ++G) {
          const int st = G & 1;
          mbar_wait_relaxed(&ctl->v_empty[st], (uint32_t)(((G >> 1) & 1) ^ 1));
          mbar_arrive_expect_tx(&ctl->v_full[st], kTileBytes);
          tma_load_3d(&qkv_map, &ctl->v_full[st], v_s + st * kTileBytes, 2 * width + h * kHd, j * kTile, b, kEvictNormal);
        }
      }
    }
    __syncwarp();
  } else if (warp == kSoftmaxWarps + 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(kTile, kTile, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(kTile, kHd, 0, 1);
      const uint32_t qa = smem_u32(q_s), ka = smem_u32(k_s), va = smem_u32(v_s);
      const int total = my_items * T;
      auto issue_s = [&](int G) {
        const int st = G & 1, j = G % T, it = G / T;
        if (j == 0) mbar_wait(&ctl->q_full, (uint32_t)(it & 1));
        mbar_wait(&ctl->k_full[st], (uint32_t)((G >> 1) & 1));
        mbar_wait(&ctl->s_empty, (uint32_t)((G & 1) ^ 1));        // softmax has drained S of tile G-1
        tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          mma_f16_ss(tmem_base, make_smem_desc(qa + k4 * 32, 0, 1024),
                     make_smem_desc(ka + st * kTileBytes + k4 * 32, 0, 1024), idesc_s, k4 ? 1u : 0u);
        mma_commit(&ctl->s_full);
        mma_commit(&ctl->k_empty[st]);
        if (j == T - 1) mma_commit(&ctl->q_empty);     // Q may be replaced by the next item's
      };
      if (total > 0) issue_s(0);
      for (int G = 0; G < total; ++G) {
        // S of the next tile goes first (it overlaps this tile's softmax) -- except across an item
        // boundary, where it has to wait for the next Q anyway and would hold back these MMAs
        const bool next_same_item = (G + 1) % T != 0;
        if (G + 1 < total && next_same_item) issue_s(G + 1);
        const int st = G & 1, j = G % T;
        mbar_wait(&ctl->v_full[st], (uint32_t)((G >> 1) & 1));
        // P of this tile is ready; this also orders the rescale of O (and, for j == 0, the read-out of
        // the previous item's O by the softmax warps) before these MMAs
        mbar_wait(&ctl->p_full, (uint32_t)(G & 1));
        tc_fence_after();
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8)
          mma_f16_ts(tmem_base + 128, tmem_base + 192 + k8 * 8,
                     make_smem_desc(va + st * kTileBytes + k8 * 2048, 8192, 1024), idesc_o, (j | k8) ? 1u : 0u);
        mma_commit(&ctl->o_full);
        mma_commit(&ctl->v_empty[st]);
        if (G + 1 < total && !next_same_item) issue_s(G + 1);
      }
    }
    __syncwarp();
  } else {
    // Two threads per query row: warp w (half 0) and warp w + 4 (half 1) share the TMEM lane quarter
    // q = w % 4 and split the row's 128 keys of every tile 64 / 64 (and the 64 output columns 32 / 32).
    // Four softmax warps per scheduler instead of two is what keeps the MUFU pipe fed: with one
    // thread per row the exponential phase of one warp could not cover the latency-bound phase
    // (TMEM load, maximum, P store) of the only other one.
    // the TMEM lane quarter a warp may touch is fixed by its CTA-level warp id (group 1 starts at warp 10)
    const int q = ((int)threadIdx.x >> 5) & 3, hf = warp >> 2;
    const int r = q * 32 + lane;                        // query row of this thread = TMEM lane
    const int pair_bar = 1 + q + 4 * grp;               // named barrier of the row's two warps (64 threads)
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_s = t_lane + hf * 64;              // this thread's 64 score columns
    const uint32_t t_o = t_lane + 128 + hf * 32;        // ... 32 accumulator columns
    const uint32_t t_p = t_lane + 192 + hf * 32;        // ... 32 packed probability columns (64 keys)
    const uint32_t stg = smem_u32(p_s) + (uint32_t)(q * 4096);   // output staging box of the quarter
    const int n_phases = my_items * T;                  // the same in both groups
    if (RZ_ATTN_PINGPONG && grp == 1 && n_phases > 0) named_bar_arrive(kPingPongBar, 512);     // group 0 goes first
    int G = 0;
    for (int it = 0; it < my_items; ++it) {
      int b, h, qt;
      const bool real = decode(it, b, h, qt);
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < T; ++j, ++G) {
        const int valid = min(kTile, p.L - j * kTile) - hf * 64;    // real keys among this thread's 64
        uint32_t v[2][32];
        mbar_wait_soft(&ctl->s_full, (uint32_t)(G & 1));
        tc_fence_after();
        tmem_ld_x32(t_s, v[0]);
        tmem_ld_x32(t_s + 32, v[1]);
        tmem_ld_wait_x32(v[0]);
        tmem_ld_wait_x32(v[1]);
        tc_fence_before();
        mbar_arrive(&ctl->s_empty);                      // S is in registers: the next S = Q K^T may start
        if (valid < 64) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= valid) v[c][i] = 0xff800000u;   // -inf: probability 0
        }
        // half-row maximum: four independent FMNMX3 chains, then the exchange with the other half
        float mq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) mq[k] = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            mq[(i >> 1) & 3] = fmaxf(mq[(i >> 1) & 3], fmaxf(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])));
        float mx = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
        ctl->xmax[G & 1][hf][r] = mx;
        named_bar_sync(pair_bar, 64);
        mx = fmaxf(mx, ctl->xmax[G & 1][hf ^ 1][r]);
        const float mxs = mx * kLog2e;
        const bool need = mxs > m_run + kLazy;           // always true for the first tile; same in both halves
        float alpha = 1.f;
        if (need) {
          alpha = exp2f(m_run - mxs);                    // 0 for the first tile
          m_run = mxs;
        }
        // exponentials -> fp16 pairs, packed IN PLACE (pair i of chunk c lands in v[c][i / 2]).
        // Only one group at a time is in this MUFU-bound phase.
        if (RZ_ATTN_PINGPONG) named_bar_sync(kPingPongBar + grp, 512);
        float rs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) rs[k] = 0.f;
        const float nm = -m_run;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0, e1;
            ffma2(e0, e1, __uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1]), kLog2e, nm);
            if (((i >> 1) % kEmuPeriod) < kEmuCount) {   // this share of the pairs bypasses the MUFU
              exp2_pair_fma(e0, e1);
            } else {
              e0 = exp2f(e0);
              e1 = exp2f(e1);
            }
            fadd2(rs[i & 2], rs[(i & 2) + 1], e0, e1);
            v[c][i >> 1] = pack_h2(e0, e1);
          }
        }
        if (RZ_ATTN_PINGPONG && !(grp == 1 && G == n_phases - 1)) named_bar_arrive(kPingPongBar + (grp ^ 1), 512);   // the other group's turn
        if (j > 0) {
          // the MMAs of tile j-1 must be done before P (single buffer) is overwritten / O is rescaled
          mbar_wait_soft(&ctl->o_full, (uint32_t)((G - 1) & 1));
          tc_fence_after();
        }
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          // rare: rescale this row's half of the accumulator in TMEM
#pragma unroll 1
          for (int k = 0; k < 2; ++k) {
            uint32_t o[16];
            tmem_ld_x16(t_o + k * 16, o);
            tmem_ld_wait_x16(o);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x16(t_o + k * 16, o);
          }
        }
        // P -> TMEM as the A operand of O += P V: row = this thread's lane, keys (2c, 2c + 1) in column c
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tmem_st_x16(t_p + c * 16, *reinterpret_cast<const uint32_t(*)[16]>(&v[c][0]));
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ctl->p_full);
        l_run = fmaf(l_run, alpha, (rs[0] + rs[1]) + (rs[2] + rs[3]));
      }
      // all MMAs of the item done: normalise, fp16, TMA store through the quarter's staging box
      mbar_wait_soft(&ctl->o_full, (uint32_t)((G - 1) & 1));
      tc_fence_after();
      if (it > 0 && hf == 0) {
        // the previous item's output store (issued by this warp) must have read the staging box
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      ctl->xsum[hf][r] = l_run;
      named_bar_sync(pair_bar, 64);
      const float inv = 1.0f / (l_run + ctl->xsum[hf ^ 1][r]);
      {
        uint32_t o[32];
        tmem_ld_x32(t_o, o);
        tmem_ld_wait_x32(o);
        tc_fence_before();
#pragma unroll
        for (int g = 0; g < 4; ++g)
          sts_v4(stg + p_off(lane, hf * 4 + g),
                 pack_h2(__uint_as_float(o[8 * g]) * inv, __uint_as_float(o[8 * g + 1]) * inv),
                 pack_h2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv),
                 pack_h2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv),
                 pack_h2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv));
      }
      fence_proxy_async_smem();
      named_bar_sync(pair_bar, 64);                      // both halves of the box are written (and xsum is read)
      if (hf == 0 && lane == 0 && real) {
        tma_store_3d(&out_map, stg, h * kHd, qt * kTile + q * 32, b);
        tma_store_commit();
      }
    }
    if (hf == 0 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps && grp == 0) tmem_dealloc(tmem_all, 2 * kTmemCols);
}

}  // namespace

extern "C" int rz_attention(const void* qkv_f16, int n_images, int tokens, int heads, void* out_f16,
                            void* stream) {
  if (!qkv_f16 || !out_f16 || n_images < 0 || tokens <= 0 || heads <= 0) return RZ_ERR_INVALID;
  if (n_images == 0) return RZ_OK;
  if ((reinterpret_cast<uintptr_t>(qkv_f16) & 15) || (reinterpret_cast<uintptr_t>(out_f16) & 15))
    return RZ_ERR_ALIGNMENT;
  const int width = heads * kHd;
  AttnParams p;
  p.B = n_images; p.L = tokens; p.H = heads;
  p.q_tiles = (tokens + kTile - 1) / kTile;
  p.kv_tiles = p.q_tiles;
  const long long items = (long long)n_images * heads * p.q_tiles;
  if (items * p.kv_tiles >= (1ll << 31)) return RZ_ERR_UNSUPPORTED;
  p.items = (int)items;
  const long long pairs = (items + 1) / 2;              // a CTA works on two items at a time (one per group)
  const long long ctas = pairs < rz_sm_count() ? pairs : rz_sm_count();
  CUtensorMap qkv_map, out_map;
  if (!rz::make_map_3d_sw128(&qkv_map, qkv_f16, (uint64_t)n_images, (uint64_t)tokens, (uint64_t)3 * width,
                             (uint64_t)3 * width * 2, (uint64_t)tokens * 3 * width * 2, kTile))
    return RZ_ERR_CUDA;
  if (!rz::make_map_3d_sw128(&out_map, out_f16, (uint64_t)n_images, (uint64_t)tokens, (uint64_t)width,
                             (uint64_t)width * 2, (uint64_t)tokens * width * 2, 32))
    return RZ_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RZ_CUDA_OK(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  attn_kernel<<<(unsigned)ctas, kThreads, kSmem, s>>>(qkv_map, out_map, p);
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}
