// K10: multi-positive NCE loss (UniCLIP MP-NCE + MIL-NCE variants), forward and backward,
// for a column block [col0, col0+b_local) of the (n_total x b_global) logit matrix.
// Restates multi_positive_nce_loss / get_row_loss / get_col_loss
// (exp/cxr_pt/model/losses.py:243-344) and their autograd in closed form.
//
// TWO launches, one on each side of the cross-rank reduction of the row sums:
//   rz_mpnce_partials  E = exp(Z/tau): row sums, positives, column sums               (read Z)
//   rz_mpnce_finish    coefficients, dZ, loss terms, sum dZ*Z (= -dL/dlog tau share)  (read Z, write dZ)
// Each is ONE persistent cooperative kernel (one or two grid-wide barriers inside) instead of the
// seven small launches of round 1.  The temperature is read from the parameter on the device
// (log_tau), so the host never synchronises.  HBM/L2-bound: Z is read twice and dZ written once.
// All sums use a fixed order -- no float atomics -- so the 1-GPU and N-GPU results agree bit for
// bit on the local block.
#include <cstdio>
#include "rz_common.cuh"

namespace {

constexpr int kRowChunk = 32;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// Grid-wide barrier for a cooperative (co-resident) launch.  `counter` is zeroed before the launch
// and only ever grows: barrier k completes when it reaches k * gridDim.x.
__device__ __forceinline__ void grid_barrier(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int nb = gridDim.x;
    const unsigned int old = atomicAdd(counter, 1u);
    const unsigned int target = (old / nb + 1u) * nb;
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// The sentence -> image map as int32 in shared memory (cooperative, coalesced, all loads in flight at once):
// the per-column phases scan it once per column, and a scan over global memory is a chain of ~50 L2
// latencies per warp.  Values outside int32 can match no column and become -1.
constexpr int kGroupCap = 8192;
__device__ __forceinline__ void stage_group_map(const long long* __restrict__ gm, int n, int* sm) {
  // eight independent loads per thread in flight (a plain loop is one L2 latency per element)
  for (int i0 = threadIdx.x; i0 < n; i0 += 8 * kThreads) {
    long long g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) g[u] = i0 + u * kThreads < n ? __ldg(gm + i0 + u * kThreads) : -1;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i0 + u * kThreads < n) sm[i0 + u * kThreads] = (g[u] >= 0 && g[u] <= 0x7fffffffLL) ? (int)g[u] : -1;
  }
  __syncthreads();
}

// sum over the sentences i of column `gcol` of f(i), lanes striding the staged map four entries at a time
template <class F>
__device__ __forceinline__ void scan_column(const int* gsm, int n, int gcol, int lane, F&& f) {
  for (int i0 = lane; i0 < n; i0 += 128) {
    int g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) g[u] = i0 + 32 * u < n ? gsm[i0 + 32 * u] : -1;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (g[u] == gcol) f(i0 + 32 * u);
  }
}

__device__ __forceinline__ float inv_temperature(const float* log_tau, float inv_tau) {
  return log_tau != nullptr ? expf(-__ldg(log_tau)) : inv_tau;
}

// ---- launch 1 --------------------------------------------------------------------------
// Phase A: a CTA takes chunks of 32 rows; warp w owns rows 4w..4w+3 and walks them across the local
// columns in blocks of 1024: a lane reads eight float4 per row (all loads of a row are issued before
// the first exponential: memory-level parallelism), adds them to its 32 column accumulators and to
// the row sum, and the lane that holds the positive column records pos[i].  The eight warps' column
// accumulators are then combined through shared memory in a fixed order -> colpart[chunk][column].
// Phase B (after the grid barrier): one warp per column adds the chunk partials and the positives of
// that column (lanes stride the chunks / sentences, then an xor butterfly: a fixed order).
struct PartialsParams {
  const float* z; long long ldz; int n_total, b_local;
  const long long* group_map; int col0; float inv_tau; const float* log_tau;
  float *rowsum, *pos, *colneg, *colpos;
  float *acol, *apos, *lcol; // per local column: negatives / positives coefficient, column loss terms
  float eps, inv_2ncol; int col_sum;
  float* colpart;            // [chunks][b_local]: sum over the chunk's rows of E (negatives AND positives)
  unsigned int* barrier;
};

constexpr int kColBlock = 1024;      // columns per pass: 32 lanes x 8 groups x 4

template <bool VEC>
__device__ __forceinline__ float4 load4(const float* row, int c, int b_local) {
  if (VEC) {
    return c < b_local ? __ldg(reinterpret_cast<const float4*>(row + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 v;
  v.x = c + 0 < b_local ? __ldg(row + c + 0) : 0.f;
  v.y = c + 1 < b_local ? __ldg(row + c + 1) : 0.f;
  v.z = c + 2 < b_local ? __ldg(row + c + 2) : 0.f;
  v.w = c + 3 < b_local ? __ldg(row + c + 3) : 0.f;
  return v;
}

#ifdef RZ_EXP_MPNCE_TIMING
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
  return v;
}
#define RZ_T(i) if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) tt[i] = gtime();
#else
#define RZ_T(i)
#endif

template <bool VEC>
__global__ void __launch_bounds__(kThreads)
mpnce_partials_kernel(PartialsParams p) {
  __shared__ float colbuf[kWarps][kColBlock];
#ifdef RZ_EXP_MPNCE_TIMING
  unsigned long long tt[6];
#endif
  RZ_T(0)
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float inv_tau = inv_temperature(p.log_tau, p.inv_tau);
  const int chunks = (p.n_total + kRowChunk - 1) / kRowChunk;
  const int b_local = p.b_local;
  for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    const int r0 = chunk * kRowChunk + warp * 4;
    float racc[4] = {0.f, 0.f, 0.f, 0.f};
    int gloc[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      gloc[rr] = -1;
      if (r0 + rr < p.n_total) {
        const long long gg = __ldg(p.group_map + r0 + rr) - (long long)p.col0;
        gloc[rr] = (gg >= 0 && gg < b_local) ? (int)gg : -1;
        if (gloc[rr] < 0 && lane == 0) p.pos[r0 + rr] = 0.f;
      }
    }
    for (int cb = 0; cb < b_local; cb += kColBlock) {
      float cacc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) { cacc[j][0] = cacc[j][1] = cacc[j][2] = cacc[j][3] = 0.f; }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        if (r0 + rr >= p.n_total) continue;                 // warp-uniform
        const float* row = p.z + (long long)(r0 + rr) * p.ldz;
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = load4<VEC>(row, cb + 128 * j + 4 * lane, b_local);
        float rs = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = cb + 128 * j + 4 * lane;
          float e[4] = {__expf(v[j].x * inv_tau), __expf(v[j].y * inv_tau), __expf(v[j].z * inv_tau),
                        __expf(v[j].w * inv_tau)};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c + u >= b_local) e[u] = 0.f;
            cacc[j][u] += e[u];
            rs += e[u];
            if (c + u == gloc[rr]) p.pos[r0 + rr] = e[u];
          }
        }
        racc[rr] += rz::warp_sum(rs);
      }
      __syncthreads();                                       // colbuf free (previous block consumed)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(&colbuf[warp][128 * j + 4 * lane]) =
            make_float4(cacc[j][0], cacc[j][1], cacc[j][2], cacc[j][3]);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kColBlock / kThreads; ++k) {
        const int c = t + k * kThreads;
        if (cb + c < b_local) {
          float sum = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) sum += colbuf[w][c];
          p.colpart[(long long)chunk * b_local + cb + c] = sum;
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        if (r0 + rr < p.n_total) p.rowsum[r0 + rr] = racc[rr];
    }
  }
  RZ_T(1)
  grid_barrier(p.barrier);
  RZ_T(2)
  // phase B.  colbuf is free now: it holds the int32 sentence -> image map for the column scans
  int* gsm = reinterpret_cast<int*>(&colbuf[0][0]);
  const bool staged = p.n_total <= kGroupCap;
  const bool any_col = (int)blockIdx.x * kWarps < b_local;
  if (staged && any_col) stage_group_map(p.group_map, p.n_total, gsm);
  RZ_T(3)
  for (int c = blockIdx.x * kWarps + warp; c < b_local; c += gridDim.x * kWarps) {
    float all = 0.f, cp = 0.f;
    for (int k = lane; k < chunks; k += 32) all += __ldcg(p.colpart + (long long)k * b_local + c);
    const long long gcol = (long long)p.col0 + c;
    if (staged) {
      scan_column(gsm, p.n_total, (int)gcol, lane, [&](int i) { cp += __ldcg(p.pos + i); });
    } else {
      for (int i0 = lane; i0 < p.n_total; i0 += 128) {
        long long g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = i0 + 32 * u < p.n_total ? __ldg(p.group_map + i0 + 32 * u) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (g[u] == gcol) cp += __ldcg(p.pos + i0 + 32 * u);
      }
    }
    all = rz::warp_sum(all);
    cp = rz::warp_sum(cp);
    const float cn = all - cp;
    // the coefficient applied to the negatives / positives of this column and the column loss terms it
    // owns: they depend on local quantities only (every sentence is present for a local column), so they
    // are computed here, before the cross-rank reduction, and launch 2 starts with the rows right away.
    // The lanes stride over the sentences and are combined in a fixed order (bit-reproducible).
    float a = 0.f, ap = 0.f, l = 0.f;
    if (p.col_sum) {
      // MIL-NCE column term: p = colpos / (colneg + colpos + eps)        losses.py:331-336
      const float cs = cn + cp + p.eps;
      const float pc = cp / cs;
      const float u = 1.0f / (pc + p.eps);
      l = -logf(pc + p.eps);
      a = u * cp / (cs * cs) * p.inv_2ncol;
      ap = -u * (1.0f / cs - cp / (cs * cs)) * p.inv_2ncol;
    } else {
      // MP-NCE: one term per sentence i of this image: p = pos_i/(pos_i + Cneg + eps)  :337-342
      auto term = [&](int i) {
        const float ps = __ldcg(p.pos + i);
        const float den = ps + cn + p.eps;
        const float pc = ps / den;
        const float w = 1.0f / (pc + p.eps);
        l += -logf(pc + p.eps);
        a += w * ps / (den * den);
      };
      if (staged) {
        scan_column(gsm, p.n_total, (int)gcol, lane, term);
      } else {
        for (int i = lane; i < p.n_total; i += 32)
          if (__ldg(p.group_map + i) == gcol) term(i);
      }
      a = rz::warp_sum(a);
      l = rz::warp_sum(l);
      a *= p.inv_2ncol;
    }
    if (lane == 0) {
      p.colneg[c] = cn; p.colpos[c] = cp;
      p.acol[c] = a; p.apos[c] = ap; p.lcol[c] = l;
    }
  }
  RZ_T(4)
#ifdef RZ_EXP_MPNCE_TIMING
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
    printf("partials grid %d: phaseA %llu ns, barrier %llu, stage %llu, phaseB %llu\n", (int)gridDim.x,
           tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], tt[4] - tt[3]);
#endif
}

// ---- launch 2 --------------------------------------------------------------------------
struct FinishParams {
  const float* z; long long ldz; int n_total, b_local, b_global;
  const long long* group_map; int col0; float inv_tau, eps; int row_sum, col_sum;
  const float* log_tau;
  const float* rowsum; const float* pos; const float* colneg; const float* colpos;
  const float *acol, *apos, *lcol; // per local column (from launch 1)
  float *img_rs, *img_ps;    // per global image (row_sum)
  float *lrow, *dzz_part;    // per row: loss term, sum_b dZ*Z
  float* dz; float* loss_terms;
  float inv_2nrow, inv_2ncol;
  unsigned int* barrier;
};

// dZ_ib = E_ib/tau * (ca_i + [b==g_i](cb_i + apos_b) + [b!=g_i] acol_b)
template <bool VEC>
__global__ void __launch_bounds__(kThreads)
mpnce_finish_kernel(FinishParams p) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int gw = blockIdx.x * kWarps + warp, nw = gridDim.x * kWarps;
  const float inv_tau = inv_temperature(p.log_tau, p.inv_tau);
  __shared__ int gsm[kGroupCap];
  const bool staged = p.n_total <= kGroupCap;
  // ---- phase 0b (row_sum only): Rs_b = sum_{i in b} R_i, Ps_b = sum_{i in b} pos_i, every GLOBAL image
  if (p.row_sum) {
    if (staged && (int)blockIdx.x * kWarps < p.b_global) stage_group_map(p.group_map, p.n_total, gsm);
    for (int b = gw; b < p.b_global; b += nw) {
      float rs = 0.f, ps = 0.f;
      for (int i = lane; i < p.n_total; i += 32) {
        const bool hit = staged ? gsm[i] == b : __ldg(p.group_map + i) == (long long)b;
        if (hit) { rs += p.rowsum[i]; ps += p.pos[i]; }
      }
      rs = rz::warp_sum(rs);
      ps = rz::warp_sum(ps);
      if (lane == 0) { p.img_rs[b] = rs; p.img_ps[b] = ps; }
    }
    grid_barrier(p.barrier);           // (the only grid-wide barrier of this launch, MIL-NCE rows only)
  }
  // ---- phase 1: one warp per row: row coefficients, dZ, sum dZ*Z
  float l_acc = 0.f, z_acc = 0.f;
  for (int r = gw; r < p.n_total; r += nw) {
    const long long g = __ldg(p.group_map + r);
    const int gl = (g >= p.col0 && g - p.col0 < p.b_local) ? (int)(g - p.col0) : -1;
    float ca, cb, l = 0.f;
    if (p.row_sum) {
      // losses.py:303-315: one term per image
      const float rs = __ldcg(p.img_rs + g) + p.eps, ps = __ldcg(p.img_ps + g);
      const float pr = ps / rs;
      const float w = 1.0f / (pr + p.eps);
      ca = w * ps / (rs * rs) * p.inv_2nrow;
      cb = -w / rs * p.inv_2nrow;
    } else {
      // losses.py:316-320
      const float rr = p.rowsum[r] + p.eps, ps = p.pos[r];
      const float pr = ps / rr;
      const float w = 1.0f / (pr + p.eps);
      ca = w * ps / (rr * rr) * p.inv_2nrow;
      cb = -w / rr * p.inv_2nrow;
      if (gl >= 0) l = -logf(pr + p.eps);
    }
    if (!p.col_sum && gl >= 0) {
      const float ps = p.pos[r];
      const float den = ps + p.colneg[gl] + p.eps;
      const float pc = ps / den;
      const float u = 1.0f / (pc + p.eps);
      cb += -u * (den - ps) / (den * den) * p.inv_2ncol;
    }
    const float pos_coef = gl >= 0 ? ca + cb + __ldcg(p.apos + gl) : 0.f;   // the positive's own column
    const float* zr = p.z + (long long)r * p.ldz;
    float* dr = p.dz != nullptr ? p.dz + (long long)r * p.ldz : nullptr;
    float acc = 0.f;
    for (int c0 = 0; c0 < p.b_local; c0 += kColBlock) {
      float4 v[8], ac[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + 128 * j + 4 * lane;
        v[j] = load4<VEC>(zr, c, p.b_local);
        if (VEC) {
          ac[j] = c < p.b_local ? __ldcg(reinterpret_cast<const float4*>(p.acol + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          ac[j].x = c + 0 < p.b_local ? __ldcg(p.acol + c + 0) : 0.f;
          ac[j].y = c + 1 < p.b_local ? __ldcg(p.acol + c + 1) : 0.f;
          ac[j].z = c + 2 < p.b_local ? __ldcg(p.acol + c + 2) : 0.f;
          ac[j].w = c + 3 < p.b_local ? __ldcg(p.acol + c + 3) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + 128 * j + 4 * lane;
        if (c >= p.b_local) continue;
        const float zz[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        const float aa[4] = {ac[j].x, ac[j].y, ac[j].z, ac[j].w};
        float d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float e = __expf(zz[u] * inv_tau);
          const float coef = (c + u == gl) ? pos_coef : ca + aa[u];
          d[u] = c + u < p.b_local ? e * inv_tau * coef : 0.f;
          acc = fmaf(d[u], zz[u], acc);
        }
        if (dr != nullptr) {
          if (VEC) {
            *reinterpret_cast<float4*>(dr + c) = make_float4(d[0], d[1], d[2], d[3]);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (c + u < p.b_local) dr[c + u] = d[u];
          }
        }
      }
    }
    acc = rz::warp_sum(acc);
    l_acc += l;                      // this warp's rows, in row order
    z_acc += acc;
  }
  // per-CTA partials (fixed warp order), so that the final reduction reads gridDim.x values, not n_total
  __shared__ float wl[kWarps], wz[kWarps];
  if (lane == 0) { wl[warp] = l_acc; wz[warp] = z_acc; }
  __syncthreads();
  if (t == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) { a += wl[k]; b += wz[k]; }
    p.lrow[blockIdx.x] = a;
    p.dzz_part[blockIdx.x] = b;
  }
  // ---- phase 2: the LAST CTA to get here reduces the per-CTA partials in a fixed order (a ticket instead of
  // a grid barrier: nobody waits): loss_terms = {sum row terms, sum col terms, sum dZ*Z, this rank's share}
  __shared__ unsigned int ticket;
  if (t == 0) {
    __threadfence();
    ticket = atomicAdd(p.barrier + 1, 1u);
  }
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  __shared__ float sh[3][kThreads];
  float a = 0.f, b = 0.f, c = 0.f;
  if (p.row_sum) {
    // image terms owned by this rank = its local columns
    for (int i = t; i < p.b_local; i += kThreads) {
      const float rs = __ldcg(p.img_rs + p.col0 + i) + p.eps;
      a += -logf(__ldcg(p.img_ps + p.col0 + i) / rs + p.eps);
    }
  } else {
    for (int i = t; i < (int)gridDim.x; i += kThreads) a += __ldcg(p.lrow + i);
  }
  for (int i = t; i < p.b_local; i += kThreads) b += __ldcg(p.lcol + i);
  for (int i = t; i < (int)gridDim.x; i += kThreads) c += __ldcg(p.dzz_part + i);
  sh[0][t] = a; sh[1][t] = b; sh[2][t] = c;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if (t < s) { sh[0][t] += sh[0][t + s]; sh[1][t] += sh[1][t + s]; sh[2][t] += sh[2][t + s]; }
    __syncthreads();
  }
  if (t == 0) {
    p.loss_terms[0] = sh[0][0]; p.loss_terms[1] = sh[1][0]; p.loss_terms[2] = sh[2][0];
    p.loss_terms[3] = sh[0][0] * p.inv_2nrow + sh[1][0] * p.inv_2ncol;
  }
}

// persistent grid of a cooperative launch: every CTA is resident, so the barrier cannot deadlock
template <typename K>
int coop_grid(K kernel, int wanted) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const int cap = rz_sm_count() * (per_sm > 4 ? 4 : per_sm);
  return wanted < cap ? (wanted < 1 ? 1 : wanted) : cap;
}

}  // namespace

extern "C" size_t rz_mpnce_partials_scratch_floats(int n_total, int b_local) {
  const long long chunks = (n_total + kRowChunk - 1) / kRowChunk;
  return (size_t)(chunks * (long long)b_local + 4);
}

extern "C" size_t rz_mpnce_finish_scratch_floats(int n_total, int b_local, int b_global) {
  (void)b_local;
  return (size_t)(2LL * n_total + 2LL * b_global + 4);
}

extern "C" int rz_mpnce_partials(const float* z, long long ldz, int n_total, int b_local, int b_global,
                                 const long long* group_map, int col0, float inv_tau,
                                 const float* log_tau, float eps, int col_sum, float* rowsum, float* pos,
                                 float* colstate, float* scratch1, void* stream) {
  if (!z || !group_map || !rowsum || !pos || !colstate || !scratch1) return RZ_ERR_INVALID;
  if (n_total <= 0 || b_local <= 0 || ldz < b_local || b_global < b_local) return RZ_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(colstate) % 16 != 0) return RZ_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int chunks = (n_total + kRowChunk - 1) / kRowChunk;
  PartialsParams p;
  p.z = z; p.ldz = ldz; p.n_total = n_total; p.b_local = b_local; p.group_map = group_map;
  p.col0 = col0; p.inv_tau = inv_tau; p.log_tau = log_tau;
  const long long bl4 = (b_local + 3) / 4 * 4;            // rows of colstate are 16-byte aligned
  p.rowsum = rowsum; p.pos = pos;
  p.colneg = colstate; p.colpos = colstate + bl4; p.acol = colstate + 2 * bl4; p.apos = colstate + 3 * bl4;
  p.lcol = colstate + 4 * bl4;
  p.eps = eps; p.col_sum = col_sum; p.inv_2ncol = 0.5f / (float)(col_sum ? b_global : n_total);
  p.colpart = scratch1;
  p.barrier = reinterpret_cast<unsigned int*>(scratch1 + (long long)chunks * b_local);
  RZ_CUDA_OK(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), s));
  // float4 path: 16-byte aligned rows (the usual case: b_local a multiple of 4)
  const bool vec = (reinterpret_cast<uintptr_t>(z) % 16 == 0) && (ldz % 4 == 0);
  const void* kern = vec ? (const void*)mpnce_partials_kernel<true> : (const void*)mpnce_partials_kernel<false>;
  const int grid = vec ? coop_grid(mpnce_partials_kernel<true>, chunks) : coop_grid(mpnce_partials_kernel<false>, chunks);
  void* args[] = {&p};
  if (getenv("RZ_EXP_MPNCE_PLAIN")) RZ_CUDA_OK(cudaLaunchKernel(kern, dim3(grid), dim3(kThreads), args, 0, s));
  else RZ_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kThreads), args, 0, s));
  rz_count_launch(1);
  return RZ_OK;
}

extern "C" int rz_mpnce_finish(const float* z, long long ldz, int n_total, int b_local,
                               int b_global, const long long* group_map, int col0, float inv_tau,
                               const float* log_tau, float eps, int row_sum, int col_sum,
                               const float* rowsum, const float* pos, const float* colstate,
                               float* scratch2, float* dz, float* loss_terms, void* stream) {
  if (!z || !group_map || !rowsum || !pos || !colstate || !scratch2 || !loss_terms)
    return RZ_ERR_INVALID;
  if (n_total <= 0 || b_local <= 0 || b_global < b_local || col0 < 0 || col0 + b_local > b_global ||
      ldz < b_local)
    return RZ_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FinishParams p;
  p.z = z; p.ldz = ldz; p.n_total = n_total; p.b_local = b_local; p.b_global = b_global;
  p.group_map = group_map; p.col0 = col0; p.inv_tau = inv_tau; p.eps = eps; p.log_tau = log_tau;
  p.row_sum = row_sum; p.col_sum = col_sum;
  const long long bl4 = (b_local + 3) / 4 * 4;
  p.rowsum = rowsum; p.pos = pos;
  p.colneg = colstate; p.colpos = colstate + bl4; p.acol = colstate + 2 * bl4; p.apos = colstate + 3 * bl4;
  p.lcol = colstate + 4 * bl4;
  float* w = scratch2;
  p.img_rs = w; w += b_global;
  p.img_ps = w; w += b_global;
  p.lrow = w; w += n_total;
  p.dzz_part = w; w += n_total;
  p.barrier = reinterpret_cast<unsigned int*>(w);
  p.dz = dz; p.loss_terms = loss_terms;
  p.inv_2nrow = 0.5f / (float)(row_sum ? b_global : n_total);
  p.inv_2ncol = 0.5f / (float)(col_sum ? b_global : n_total);
  RZ_CUDA_OK(cudaMemsetAsync(p.barrier, 0, 2 * sizeof(unsigned int), s));     // [barrier counter, ticket]
  const bool vec = (reinterpret_cast<uintptr_t>(z) % 16 == 0) && (ldz % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(colstate) % 16 == 0) &&
                   (dz == nullptr || reinterpret_cast<uintptr_t>(dz) % 16 == 0);
  const int want = (n_total + kWarps - 1) / kWarps;
  const void* kern = vec ? (const void*)mpnce_finish_kernel<true> : (const void*)mpnce_finish_kernel<false>;
  const int grid = vec ? coop_grid(mpnce_finish_kernel<true>, want) : coop_grid(mpnce_finish_kernel<false>, want);
  void* args[] = {&p};
  RZ_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kThreads), args, 0, s));
  rz_count_launch(1);
  return RZ_OK;
}
