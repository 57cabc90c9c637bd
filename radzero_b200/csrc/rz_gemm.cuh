// Persistent, warp-specialised tcgen05 GEMM skeleton shared by the large-N forward and the
// backward kernels of the VL-CABS path.
//
//   warp EW     : TMA producer   (one elected lane; SWIZZLE_128B boxes into a kStages-deep ring)
//   warp EW + 1 : MMA issuer     (one elected lane; tcgen05.mma kind::f16, fp32 accumulate in TMEM)
//   warps 0..EW-1 : epilogue     (thread = accumulator row; tcgen05.ld -> policy functor -> global)
// EW = V::kEpiWarps = 4, or 8 for epilogues that are bound by their own instruction issue (warps w and
// w + 4 share the TMEM lane quarter w % 4 and split the tile's columns)
//
// Output tile: 128 rows x (kAccs * kBN) fp32 TMEM columns, double buffered (2 * kAccs * kBN
// <= 512) so that the epilogue of tile i overlaps the MMAs of tile i+1.  K advances 64
// elements (one 128-byte swizzle row) per pipeline stage = 4 MMAs of K=16.
// A policy class V supplies the problem decomposition, the TMA loads and the epilogue:
//   V::kBN, V::kAccs (1 or 2 A operands sharing one B), V::kAMn / V::kBMn (operand is
//   MN-major in shared memory), V::Params,
//   V::num_tiles(p), V::k_steps(p), V::load(p, maps, tile, ks, a_smem, a2_smem, b_smem, bar, rank),
//   V::epilogue(p, maps, tile, next_tile, tmem_acc, warp, lane, epi_bar, state, epi_smem)
//   V::inner(p): tiles are handed to a CTA in runs of `inner` consecutive ids (one "item"), so an
//     epilogue thread can carry V::State (registers) across the tiles of an item;
//   V::tile_n(p, tile): MMA N of this tile (<= kBN; a narrower last tile of a run);
//   V::kEpiSmem: bytes of shared memory reserved for the epilogue warps (staging / transposes).
#pragma once

#include "rz_common.cuh"
#include "rz_tma.cuh"
#include "rz_umma.cuh"

namespace rz {
namespace gemm {

using namespace rz::umma;

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kABytes = kBM * 128;       // one A stage: 128 rows x 128 B (or 2 MN blocks of 8 KB)
constexpr int kMnBlock = 64 * 128;       // MN-major block: 64 K-rows x 128 B
constexpr int kThreads = 192;

struct Maps {
  CUtensorMap a, a2, b, b2;   // operand loads
  CUtensorMap c, c2;          // epilogue TMA stores
};

// defaults a policy inherits: one tile per item, full-width tiles, no epilogue state / scratch,
// no cluster
struct PolicyBase {
  struct State {};
  static constexpr int kEpiSmem = 0;
  static constexpr int kCluster = 1;
  static constexpr int kEpiWarps = 4;
  template <class P> __host__ __device__ static int inner(const P&) { return 1; }
  template <class P> __host__ __device__ static int tile_n(const P&, int) { return 0; }
  // called by the epilogue warps before they wait for the tile's accumulator
  template <class P, class S>
  __device__ static void prologue(const P&, const Maps&, int, int, int, uint64_t*, S&, uint8_t*) {}
};

template <class V>
struct Layout {
  static constexpr int kBBytes = V::kBN * 128 / V::kCluster;      // a CTA pair splits B between its CTAs
  static constexpr int kStageBytes = V::kAccs * kABytes + kBBytes;
  // 227 KB of dynamic shared memory per CTA, minus alignment slack + control block
  static constexpr int kBudget = 232448 - 2048 - V::kEpiSmem;
  static constexpr int kStages = kBudget / kStageBytes > 6 ? 6 : kBudget / kStageBytes;
  static constexpr int kTmemCols = 2 * V::kAccs * V::kBN;
  static constexpr int kSmem = 1024 + kStages * kStageBytes + V::kEpiSmem + 1024;
  static_assert(kTmemCols <= 512, "accumulators exceed TMEM");
  static_assert(kStages >= 2, "pipeline too shallow");
};

struct Ctrl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t epi_bar[8];      // two per epilogue warp: TMA loads issued by the epilogue itself
  uint32_t sink[8];         // scratch words of lds_returned (one per epilogue warp), right behind epi_bar
  uint32_t tmem_slot;
};

// V::kCluster == 2: CTA-PAIR mode (tcgen05 cta_group::2).  The two CTAs of a cluster run tiles
// (2i, 2i+1) of an item -- two 128-row A tiles that share their B operand -- as ONE 256 x kBN MMA:
// each CTA loads its own A tile and HALF of B into its own shared memory, the leader's elected thread
// issues the MMAs for both, and each CTA's tensor core reads the other half of B from its partner.
// Per CTA and k-step this moves 16 + kBN/4 KB instead of 16 + kBN/2 KB into shared memory (6 stages
// instead of 4 at kBN = 256), which is what the single-CTA main loop was short of.
//   full[st]      lives in the leader: its own arrive.expect_tx (bytes of BOTH CTAs) + the partner's
//                 remote arrive; both CTAs' TMA loads complete_tx on it
//   empty[st], acc_full[a]   per CTA, signalled by the leader's multicast commits
//   acc_empty[a]  lives in the leader: one arrival per epilogue warp of BOTH CTAs (4 local + 4 remote)
template <class V>
__global__ void __launch_bounds__(32 * (V::kEpiWarps + 2), 1)
gemm_kernel(const __grid_constant__ Maps maps, const typename V::Params p) {
  using L = Layout<V>;
  constexpr int C = V::kCluster;
  constexpr int EW = V::kEpiWarps;
  extern __shared__ uint8_t smem_raw[];
  // the dynamic segment starts at the same offset in every CTA, so the aligned base does too
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = base + L::kStages * L::kStageBytes;
  Ctrl* ctl = reinterpret_cast<Ctrl*>(epi_smem + V::kEpiSmem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = C > 1 ? (int)cluster_ctarank() : 0;
  const int cid = (int)blockIdx.x / C, ncl = (int)gridDim.x / C;
  const int inner = V::inner(p);
  const int n_items = V::num_tiles(p) / (inner * C);
  const int ksteps = V::k_steps(p);
  constexpr uint32_t kTmem = L::kTmemCols < 32 ? 32 : L::kTmemCols;

  if (tid == 0) {
    for (int i = 0; i < L::kStages; ++i) { mbar_init(&ctl->full[i], C); mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&ctl->acc_full[i], 1); mbar_init(&ctl->acc_empty[i], EW * C); }
    for (int i = 0; i < 8; ++i) mbar_init(&ctl->epi_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == EW) {
    if (lane == 0) {
      prefetch_tmap(&maps.a); prefetch_tmap(&maps.b);
      if (V::kAccs > 1 || V::kTwoPhase) { prefetch_tmap(&maps.a2); prefetch_tmap(&maps.b2); }
    }
    if (C > 1) tmem_alloc_pair(&ctl->tmem_slot, kTmem); else tmem_alloc(&ctl->tmem_slot, kTmem);
  }
  tc_fence_before();
  if (C > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_slot;

  if (warp == EW) {
    if (elect_one()) {
      long long g = 0;
      for (int item = cid; item < n_items; item += ncl) {
        for (int sub = 0; sub < inner; ++sub) {
          const int tile = (item * C + rank) * inner + sub;
          for (int ks = 0; ks < ksteps; ++ks, ++g) {
            const int st = (int)(g % L::kStages);
            mbar_wait(&ctl->empty[st], (uint32_t)(((g / L::kStages) & 1) ^ 1));
            if (rank == 0) mbar_arrive_expect_tx(&ctl->full[st], (uint32_t)(L::kStageBytes * C));
            else mbar_arrive_cluster(mapa_u32(smem_u32(&ctl->full[st]), 0));
            uint8_t* sa = base + st * L::kStageBytes;
            V::load(p, maps, tile, ks, sa, sa + kABytes, sa + V::kAccs * kABytes, &ctl->full[st], rank);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == EW + 1) {
    if (rank == 0 && elect_one()) {
      long long g = 0;
      int it = 0;
      for (int item = cid; item < n_items; item += ncl)
      for (int sub = 0; sub < inner; ++sub, ++it) {
        const int tile = (item * C) * inner + sub;
        const int tn = V::tile_n(p, tile);
        const uint32_t idesc = make_idesc_f16(kBM * C, (uint32_t)(tn > 0 ? tn : V::kBN), V::kAMn ? 1 : 0, V::kBMn ? 1 : 0);
        const int acc = it & 1;
        mbar_wait(&ctl->acc_empty[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * (V::kAccs * V::kBN);
        for (int ks = 0; ks < ksteps; ++ks, ++g) {
          const int st = (int)(g % L::kStages);
          mbar_wait(&ctl->full[st], (uint32_t)((g / L::kStages) & 1));
          tc_fence_after();
          const uint32_t sa = smem_u32(base + st * L::kStageBytes);
          const uint32_t sb = sa + V::kAccs * kABytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t bd = V::kBMn ? make_smem_desc(sb + k4 * 2048, kMnBlock, 1024)
                                        : make_smem_desc(sb + k4 * 32, 0, 1024);
#pragma unroll
            for (int a = 0; a < V::kAccs; ++a) {
              const uint32_t aa = sa + a * kABytes;
              const uint64_t ad = V::kAMn ? make_smem_desc(aa + k4 * 2048, kMnBlock, 1024)
                                          : make_smem_desc(aa + k4 * 32, 0, 1024);
              if (C > 1) mma_f16_ss_pair(d0 + a * V::kBN, ad, bd, idesc, (ks | k4) ? 1u : 0u);
              else mma_f16_ss(d0 + a * V::kBN, ad, bd, idesc, (ks | k4) ? 1u : 0u);
            }
          }
          if (C > 1) mma_commit_pair(&ctl->empty[st]); else mma_commit(&ctl->empty[st]);
        }
        if (C > 1) mma_commit_pair(&ctl->acc_full[acc]); else mma_commit(&ctl->acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    int it = 0;
    typename V::State state{};
    const uint32_t acc_empty_leader[2] = {C > 1 ? mapa_u32(smem_u32(&ctl->acc_empty[0]), 0) : 0u,
                                          C > 1 ? mapa_u32(smem_u32(&ctl->acc_empty[1]), 0) : 0u};
    for (int item = cid; item < n_items; item += ncl)
    for (int sub = 0; sub < inner; ++sub, ++it) {
      const int tile = (item * C + rank) * inner + sub;
      // the tile this CTA handles next (-1: none), for epilogues that prefetch their own inputs
      const int next_tile = sub + 1 < inner ? tile + 1
                            : (item + ncl < n_items ? ((item + ncl) * C + rank) * inner : -1);
      const int acc = it & 1;
      V::prologue(p, maps, tile, warp, lane, ctl->epi_bar, state, epi_smem);
      mbar_wait(&ctl->acc_full[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      V::epilogue(p, maps, tile, next_tile, tmem_base + acc * (V::kAccs * V::kBN), warp, lane, ctl->epi_bar,
                  state, epi_smem);
      tc_fence_before();
      __syncwarp();                       // one arrival per epilogue warp
      if (lane == 0) {
        if (C > 1) mbar_arrive_cluster(acc_empty_leader[acc]); else mbar_arrive(&ctl->acc_empty[acc]);
      }
    }
    // bulk stores issued by epilogue lanes must have read their staging buffers before exit
    if (V::kEpiSmem > 0 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  // (pair) nobody leaves while the partner may still read this CTA's operands or signal its barriers
  if (C > 1) cluster_sync_all(); else __syncthreads();
  if (warp == EW) { if (C > 1) tmem_dealloc_pair(tmem_base, kTmem); else tmem_dealloc(tmem_base, kTmem); }
}

// resident clusters of gemm_kernel<V> on this device (persistent grid size / kCluster)
template <class V>
int max_clusters(cudaLaunchConfig_t* cfg) {
  static int cached_all[64] = {0};           // per device (a process may drive several GPUs)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  int& cached = cached_all[dev];
  if (cached > 0) return cached;
  int n = 0;
  if (V::kCluster > 1) {
    if (cudaOccupancyMaxActiveClusters(&n, gemm_kernel<V>, cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = rz_sm_count() / V::kCluster;
    }
  } else {
    n = rz_sm_count();
  }
  cached = n;
  return n;
}

template <class V>
int launch(const Maps& maps, const typename V::Params& p, cudaStream_t s) {
  using L = Layout<V>;
  const int items = V::num_tiles(p) / (V::inner(p) * V::kCluster);
  if (items <= 0) return RZ_OK;
  RZ_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmem));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = V::kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(32 * (V::kEpiWarps + 2)); cfg.dynamicSmemBytes = L::kSmem; cfg.stream = s;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)(rz_sm_count() / V::kCluster * V::kCluster));
  const int cap = max_clusters<V>(&cfg);
  cfg.gridDim = dim3((unsigned)((items < cap ? items : cap) * V::kCluster));
  RZ_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_kernel<V>, maps, p));
  rz_count_launch();
  return RZ_OK;
}

// ---- helpers for policies ---------------------------------------------------------------
// One TMA box into this CTA's shared memory.  C == 2 (CTA pair): the completion is signalled on the
// LEADER's copy of `bar` (the MMA-issuing thread waits there for both CTAs' halves).
template <int C>
__device__ __forceinline__ void tma_ld(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                       uint64_t hint) {
  if (C == 1) tma_load_3d(m, bar, dst, c0, c1, c2, hint);
  else tma_load_3d_pair(m, mapa_u32(smem_u32(bar), 0), dst, c0, c1, c2, hint);
}
// K-major operand tile: `rows` rows x 64 K-elements, one TMA box
template <int C = 1>
__device__ __forceinline__ void load_kmajor(const CUtensorMap* m, uint64_t* bar, void* dst, int k0,
                                            int row0, int batch, uint64_t hint = kEvictNormal) {
  tma_ld<C>(m, bar, dst, k0, row0, batch, hint);
}
// MN-major operand tile: `blocks` blocks of [64 K-rows x 64 MN-elements]
template <int C = 1>
__device__ __forceinline__ void load_mnmajor(const CUtensorMap* m, uint64_t* bar, uint8_t* dst,
                                             int mn0, int k0, int batch, int blocks,
                                             uint64_t hint = kEvictNormal) {
  for (int i = 0; i < blocks; ++i)
    tma_ld<C>(m, bar, dst + i * kMnBlock, mn0 + i * 64, k0, batch, hint);
}
// The B operand of a CTA pair: CTA `rank` loads ITS half -- rows [rank * rows / C, ...) of a K-major
// tile (the tensor map's box holds kBN / C rows) or blocks [rank * blocks / C, ...) of an MN-major one
// -- to the start of its own B buffer; the pair's MMA reads both halves.
template <int C>
__device__ __forceinline__ void load_kmajor_shared(const CUtensorMap* m, uint64_t* bar, uint8_t* dst,
                                                   int k0, int row0, int batch, int rows, int rank) {
  tma_ld<C>(m, bar, dst, k0, row0 + rank * (rows / C), batch, kEvictNormal);
}
template <int C>
__device__ __forceinline__ void load_mnmajor_shared(const CUtensorMap* m, uint64_t* bar, uint8_t* dst,
                                                    int mn0, int k0, int batch, int blocks, int rank) {
  const int part = blocks / C;
  for (int i = 0; i < part; ++i)
    tma_ld<C>(m, bar, dst + i * kMnBlock, mn0 + (rank * part + i) * 64, k0, batch, kEvictNormal);
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a SWIZZLE_128B staging box
// (rows of 128 B; the box base is 1024-byte aligned)
__device__ __forceinline__ uint32_t stage_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// 64 accumulator columns of this thread's row: two tcgen05.ld.x32 in flight
struct Cols64 {
  uint32_t lo[32], hi[32];
};
__device__ __forceinline__ void ld64(uint32_t taddr, Cols64& v) {
  tmem_ld_x32(taddr, v.lo);
  tmem_ld_x32(taddr + 32, v.hi);
}
__device__ __forceinline__ void wait64(Cols64& v) {
  tmem_ld_wait_x32(v.lo);
  tmem_ld_wait_x32(v.hi);
}

// packed fp32 pairs (rz_common.cuh), visible to the policies that `use namespace rz::gemm`
using rz::P2; using rz::p2; using rz::p2_unpack; using rz::p2_fma; using rz::p2_mul; using rz::p2_add;
using rz::p2_sub; using rz::p2_pack_h2;

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace gemm
}  // namespace rz
