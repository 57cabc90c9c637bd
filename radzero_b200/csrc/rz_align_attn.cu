// A2: multi-head self-attention of the AlignTransformer layers (transformers Dinov2SelfAttention,
// called from exp/cxr_pt/model/align_transformers.py:40) on tcgen05 tensor cores:
//
//   O[b, l, h] = softmax_l'(Q[b, l, h] . K[b, l', h]^T) V[b, l', h]       head dim 64, no mask
//
// Q, K, V are column blocks of the fused projection output qkv [B, L, 3 * H * 64] fp16 (Q already
// carries the 1/sqrt(64) scale: it is folded into the projection weights on the host, exactly, since
// it is a power of two).  One CTA owns (image, head, 128 query rows) and streams the image's keys in
// tiles of 128; two CTAs are resident per SM (256 TMEM columns and ~113 KB of shared memory each) so
// that one CTA's exponentials overlap the other's MMAs.
//
//   warp 4  TMA producer: Q once, then K / V tiles into two-stage rings (separate barriers: S can
//           start as soon as K has landed)
//   warp 5  MMA issuer (one elected lane): S = Q K^T (M 128, N 128, K 64) into TMEM, and
//           O_j = P_j V_j (M 128, N 64, K 128; V tile read as an MN-major B operand) into one of two
//           TMEM buffers -- S of tile j+1 is issued before waiting for P of tile j
//   warps 0-3  softmax, one query row per thread (TMEM lane = row, no shuffles): exact online
//           softmax in the log2 domain; S is read from TMEM twice (maximum, then exponentials) to
//           keep the register file small enough for two CTAs per SM; P goes to shared memory as the
//           fp16 K-major A operand; the per-tile products O_j are accumulated in REGISTERS with the
//           running rescale (so TMEM is never read-modify-written), one tile behind the MMAs.
// Keys >= L of the last tile are masked to probability 0; rows >= L are zero-filled by the TMA
// loads and clipped by the TMA store.
#include "rz_common.cuh"
#include "rz_tma.cuh"
#include "rz_umma.cuh"

namespace {

using namespace rz::umma;

constexpr int kHd = 64;                    // head dim
constexpr int kTile = 128;                 // query rows per CTA = keys per tile
constexpr int kTileBytes = kTile * 128;    // 128 rows x 64 halfs
constexpr int kThreads = 192;
constexpr int kTmemCols = 256;             // S [0,128)  O0 [128,192)  O1 [192,256)
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  int B, L, H, q_tiles, kv_tiles;
};

struct Ctrl {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t s_full, s_empty, p_full;
  uint64_t o_full[2], o_empty[2];
  uint32_t tmem_slot;
};

constexpr int kSmem = 5 * kTileBytes + 2 * kTileBytes + 256;   // Q, K x2, V x2, P (two 64-key boxes), Ctrl
static_assert(sizeof(Ctrl) <= 256, "control block");

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

__global__ void __launch_bounds__(kThreads, 2)
attn_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ CUtensorMap out_map,
            const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + kTileBytes;
  uint8_t* v_s = k_s + 2 * kTileBytes;
  uint8_t* p_s = v_s + 2 * kTileBytes;
  Ctrl* ctl = reinterpret_cast<Ctrl*>(p_s + 2 * kTileBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = (int)blockIdx.x % p.q_tiles;
  const int bh = (int)blockIdx.x / p.q_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int T = p.kv_tiles;
  const int width = p.H * kHd;

  if (tid == 0) {
    mbar_init(&ctl->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->k_full[i], 1); mbar_init(&ctl->k_empty[i], 1);
      mbar_init(&ctl->v_full[i], 1); mbar_init(&ctl->v_empty[i], 1);
      mbar_init(&ctl->o_full[i], 1); mbar_init(&ctl->o_empty[i], 128);
    }
    mbar_init(&ctl->s_full, 1);
    mbar_init(&ctl->s_empty, 128);
    mbar_init(&ctl->p_full, 128);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) { prefetch_tmap(&qkv_map); prefetch_tmap(&out_map); }
    tmem_alloc(&ctl->tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&ctl->q_full, kTileBytes);
      tma_load_3d(&qkv_map, &ctl->q_full, q_s, h * kHd, qt * kTile, b, kEvictNormal);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        mbar_wait(&ctl->k_empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&ctl->k_full[st], kTileBytes);
        tma_load_3d(&qkv_map, &ctl->k_full[st], k_s + st * kTileBytes, width + h * kHd, j * kTile, b, kEvictNormal);
        mbar_wait(&ctl->v_empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&ctl->v_full[st], kTileBytes);
        tma_load_3d(&qkv_map, &ctl->v_full[st], v_s + st * kTileBytes, 2 * width + h * kHd, j * kTile, b, kEvictNormal);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(kTile, kTile, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(kTile, kHd, 0, 1);
      const uint32_t qa = smem_u32(q_s), ka = smem_u32(k_s), va = smem_u32(v_s), pa = smem_u32(p_s);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&ctl->k_full[st], (uint32_t)((j >> 1) & 1));
        mbar_wait(&ctl->s_empty, (uint32_t)((j & 1) ^ 1));        // softmax has drained S of tile j-1
        tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          mma_f16_ss(tmem_base, make_smem_desc(qa + k4 * 32, 0, 1024),
                     make_smem_desc(ka + st * kTileBytes + k4 * 32, 0, 1024), idesc_s, k4 ? 1u : 0u);
        mma_commit(&ctl->s_full);
        mma_commit(&ctl->k_empty[st]);
      };
      mbar_wait(&ctl->q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int st = j & 1, ob = j & 1;
        mbar_wait(&ctl->v_full[st], (uint32_t)((j >> 1) & 1));
        mbar_wait(&ctl->o_empty[ob], (uint32_t)(((j >> 1) & 1) ^ 1));
        mbar_wait(&ctl->p_full, (uint32_t)(j & 1));
        tc_fence_after();
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8)
          mma_f16_ss(tmem_base + 128 + ob * kHd,
                     make_smem_desc(pa + (k8 >> 2) * kTileBytes + (k8 & 3) * 32, 0, 1024),
                     make_smem_desc(va + st * kTileBytes + k8 * 2048, 8192, 1024), idesc_o, k8 ? 1u : 0u);
        mma_commit(&ctl->o_full[ob]);
        mma_commit(&ctl->v_empty[st]);
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;                     // query row of this thread = TMEM lane
    const uint32_t t_s = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t t_o = t_s + 128;
    const uint32_t prow = smem_u32(p_s);
    float acc[kHd];
#pragma unroll
    for (int i = 0; i < kHd; ++i) acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;

    auto add_o = [&](int jo, float alpha) {
      // acc = acc * alpha + O_jo  (O_jo is at the scale of the running maximum after tile jo)
      const int ob = jo & 1;
      mbar_wait(&ctl->o_full[ob], (uint32_t)((jo >> 1) & 1));
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_x32(t_o + ob * kHd, o0);
      tmem_ld_x32(t_o + ob * kHd + 32, o1);
      tmem_ld_wait_x32(o0);
      tmem_ld_wait_x32(o1);
      tc_fence_before();
      mbar_arrive(&ctl->o_empty[ob]);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        acc[i] = fmaf(acc[i], alpha, __uint_as_float(o0[i]));
        acc[32 + i] = fmaf(acc[32 + i], alpha, __uint_as_float(o1[i]));
      }
    };

    for (int j = 0; j < T; ++j) {
      const int valid = min(kTile, p.L - j * kTile);    // real keys in this tile
      mbar_wait(&ctl->s_full, (uint32_t)(j & 1));
      tc_fence_after();
      // pass 1: row maximum of the tile
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(t_s + c * 32, v);
        tmem_ld_wait_x32(v);
        if (c * 32 + 32 <= valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx * kLog2e);
      const float alpha = exp2f(m_run - m_new);          // 0 for the first tile
      // the product of the previous tile: its MMAs are done long before; this also frees the P buffer
      if (j > 0) add_o(j - 1, alpha_prev);
      // pass 2: exponentials -> P (fp16, K-major SWIZZLE_128B A operand: two boxes of 64 keys)
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(t_s + c * 32, v);
        tmem_ld_wait_x32(v);
        if (c == 3) {                                    // S of this tile is fully drained
          tc_fence_before();
          mbar_arrive(&ctl->s_empty);
        }
        const uint32_t box = prow + (uint32_t)((c >> 1) * kTileBytes);
        const bool full = c * 32 + 32 <= valid;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = exp2f(fmaf(__uint_as_float(v[8 * g + i]), kLog2e, -m_new));
            if (!full && c * 32 + 8 * g + i >= valid) e[i] = 0.f;
          }
          rs0 += (e[0] + e[1]) + (e[2] + e[3]);
          rs1 += (e[4] + e[5]) + (e[6] + e[7]);
          sts_v4(box + p_off(r, (c & 1) * 4 + g), pack_h2(e[0], e[1]), pack_h2(e[2], e[3]),
                 pack_h2(e[4], e[5]), pack_h2(e[6], e[7]));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&ctl->p_full);
      l_run = fmaf(l_run, alpha, rs0 + rs1);
      m_run = m_new;
      alpha_prev = alpha;
    }
    add_o(T - 1, alpha_prev);
    // normalise, fp16, TMA store through this warp's staging box (the P buffer is free now)
    const float inv = 1.0f / l_run;
    const uint32_t stg = prow + (uint32_t)(warp * 4096);
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 8; ++g)
      sts_v4(stg + p_off(lane, g), pack_h2(acc[8 * g] * inv, acc[8 * g + 1] * inv),
             pack_h2(acc[8 * g + 2] * inv, acc[8 * g + 3] * inv),
             pack_h2(acc[8 * g + 4] * inv, acc[8 * g + 5] * inv),
             pack_h2(acc[8 * g + 6] * inv, acc[8 * g + 7] * inv));
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&out_map, stg, h * kHd, qt * kTile + warp * 32, b);
      tma_store_commit();
      tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
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
  const long long ctas = (long long)n_images * heads * p.q_tiles;
  if (ctas >= (1ll << 31)) return RZ_ERR_UNSUPPORTED;
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
