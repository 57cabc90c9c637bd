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

__device__ __forceinline__ float inv_temperature(const float* log_tau, float inv_tau) {
  return log_tau != nullptr ? expf(-__ldg(log_tau)) : inv_tau;
}

// ---- launch 1 --------------------------------------------------------------------------
// Phase A: a CTA takes chunks of 32 rows x all local columns (256 at a time).  Thread t owns column
// c0+t: it walks the 32 rows, accumulating the column partials; E is parked in smem so that warp w
// can then reduce rows 4w..4w+3 across the tile.  Phase B (after the grid barrier): one warp per
// column adds the chunk partials in a fixed order.
struct PartialsParams {
  const float* z; long long ldz; int n_total, b_local;
  const long long* group_map; int col0; float inv_tau; const float* log_tau;
  float *rowsum, *pos, *colneg, *colpos;
  float* colpart;            // [chunks][2][b_local]
  unsigned int* barrier;
};

__global__ void __launch_bounds__(kThreads)
mpnce_partials_kernel(PartialsParams p) {
  __shared__ float tile[kRowChunk][kThreads + 1];
  __shared__ int gcol[kRowChunk];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float inv_tau = inv_temperature(p.log_tau, p.inv_tau);
  const int chunks = (p.n_total + kRowChunk - 1) / kRowChunk;
  const int b_local = p.b_local;
  for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    const int r0 = chunk * kRowChunk;
    const int nrows = min(kRowChunk, p.n_total - r0);
    __syncthreads();
    if (t < kRowChunk) {
      int g = -1;
      if (t < nrows) {
        const long long gg = p.group_map[r0 + t] - (long long)p.col0;
        g = (gg >= 0 && gg < b_local) ? (int)gg : -1;
        if (g < 0) p.pos[r0 + t] = 0.f;
      }
      gcol[t] = g;
    }
    __syncthreads();
    float racc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < b_local; c0 += kThreads) {
      const int c = c0 + t;
      float cneg = 0.f, cpos = 0.f;
      // loads of 8 rows are issued before their exponentials (memory-level parallelism)
      for (int rb = 0; rb < nrows; rb += 8) {
        float zz[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          zz[u] = (c < b_local && rb + u < nrows) ? __ldg(p.z + (long long)(r0 + rb + u) * p.ldz + c) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = rb + u;
          if (r >= nrows) break;
          float e = 0.f;
          if (c < b_local) {
            e = __expf(zz[u] * inv_tau);
            if (gcol[r] == c) { cpos += e; p.pos[r0 + r] = e; } else { cneg += e; }
          }
          tile[r][t] = e;
        }
      }
      if (c < b_local) {
        p.colpart[((long long)chunk * 2 + 0) * b_local + c] = cneg;
        p.colpart[((long long)chunk * 2 + 1) * b_local + c] = cpos;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = warp * 4 + k;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kThreads / 32; ++j) s += tile[r][lane + 32 * j];
        racc[k] += rz::warp_sum(s);
      }
      __syncthreads();
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = warp * 4 + k;
        if (r < nrows) p.rowsum[r0 + r] = racc[k];
      }
    }
  }
  grid_barrier(p.barrier);
  // column sums: lanes stride the chunks, then an xor butterfly -- the same order on every run
  for (int c = blockIdx.x * kWarps + warp; c < b_local; c += gridDim.x * kWarps) {
    float n = 0.f, q = 0.f;
    for (int k = lane; k < chunks; k += 32) {
      n += __ldcg(p.colpart + ((long long)k * 2 + 0) * b_local + c);
      q += __ldcg(p.colpart + ((long long)k * 2 + 1) * b_local + c);
    }
    n = rz::warp_sum(n);
    q = rz::warp_sum(q);
    if (lane == 0) { p.colneg[c] = n; p.colpos[c] = q; }
  }
}

// ---- launch 2 --------------------------------------------------------------------------
struct FinishParams {
  const float* z; long long ldz; int n_total, b_local, b_global;
  const long long* group_map; int col0; float inv_tau, eps; int row_sum, col_sum;
  const float* log_tau;
  const float* rowsum; const float* pos; const float* colneg; const float* colpos;
  float *acol, *apos, *lcol; // per local column: negatives / positives coefficient, loss part
  float *img_rs, *img_ps;    // per global image (row_sum)
  float *lrow, *dzz_part;    // per row: loss term, sum_b dZ*Z
  float* dz; float* loss_terms;
  float inv_2nrow, inv_2ncol;
  unsigned int* barrier;
};

// dZ_ib = E_ib/tau * (ca_i + [b==g_i](cb_i + apos_b) + [b!=g_i] acol_b)
__global__ void __launch_bounds__(kThreads)
mpnce_finish_kernel(FinishParams p) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int gw = blockIdx.x * kWarps + warp, nw = gridDim.x * kWarps;
  const float inv_tau = inv_temperature(p.log_tau, p.inv_tau);
  // ---- phase 0a: per local column, the coefficient applied to the negatives / positives of that
  // column and the column loss terms it owns.  One WARP per column: the lanes stride over the
  // sentences and are combined in a fixed order (bit-reproducible, independent of the rank count).
  for (int c = gw; c < p.b_local; c += nw) {
    const long long gcol = (long long)p.col0 + c;
    float a = 0.f, ap = 0.f, l = 0.f;
    if (p.col_sum) {
      // MIL-NCE column term: p = colpos / (colneg + colpos + eps)        losses.py:331-336
      const float cs = p.colneg[c] + p.colpos[c] + p.eps;
      const float pc = p.colpos[c] / cs;
      const float u = 1.0f / (pc + p.eps);
      l = -logf(pc + p.eps);
      a = u * p.colpos[c] / (cs * cs) * p.inv_2ncol;
      ap = -u * (1.0f / cs - p.colpos[c] / (cs * cs)) * p.inv_2ncol;
    } else {
      // MP-NCE: one term per sentence i of this image: p = pos_i/(pos_i + Cneg + eps)  :337-342
      const float cn = p.colneg[c];
      for (int i = lane; i < p.n_total; i += 32) {
        if (__ldg(p.group_map + i) != gcol) continue;
        const float ps = p.pos[i];
        const float den = ps + cn + p.eps;
        const float pc = ps / den;
        const float u = 1.0f / (pc + p.eps);
        l += -logf(pc + p.eps);
        a += u * ps / (den * den);
      }
      a = rz::warp_sum(a);
      l = rz::warp_sum(l);
      a *= p.inv_2ncol;
    }
    if (lane == 0) { p.acol[c] = a; p.apos[c] = ap; p.lcol[c] = l; }
  }
  // ---- phase 0b (row_sum only): Rs_b = sum_{i in b} R_i, Ps_b = sum_{i in b} pos_i, every GLOBAL image
  if (p.row_sum) {
    for (int b = gw; b < p.b_global; b += nw) {
      float rs = 0.f, ps = 0.f;
      for (int i = lane; i < p.n_total; i += 32)
        if (__ldg(p.group_map + i) == (long long)b) { rs += p.rowsum[i]; ps += p.pos[i]; }
      rs = rz::warp_sum(rs);
      ps = rz::warp_sum(ps);
      if (lane == 0) { p.img_rs[b] = rs; p.img_ps[b] = ps; }
    }
  }
  grid_barrier(p.barrier);
  // ---- phase 1: one warp per row: row coefficients, dZ, sum dZ*Z
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
    const float* zr = p.z + (long long)r * p.ldz;
    float* dr = p.dz != nullptr ? p.dz + (long long)r * p.ldz : nullptr;
    float acc = 0.f;
    for (int c0 = 0; c0 < p.b_local; c0 += 128) {
      float zz[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 32 + lane;
        zz[u] = c < p.b_local ? __ldg(zr + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 32 + lane;
        if (c >= p.b_local) continue;
        const float e = __expf(zz[u] * inv_tau);
        const float coef = ca + (c == gl ? cb + __ldcg(p.apos + c) : __ldcg(p.acol + c));
        const float d = e * inv_tau * coef;
        if (dr != nullptr) dr[c] = d;
        acc = fmaf(d, zz[u], acc);
      }
    }
    acc = rz::warp_sum(acc);
    if (lane == 0) { p.lrow[r] = l; p.dzz_part[r] = acc; }
  }
  grid_barrier(p.barrier);
  // ---- phase 2 (CTA 0): fixed-order tree: loss_terms = {sum row terms, sum col terms, sum dZ*Z,
  // this rank's share of the loss}
  if (blockIdx.x != 0) return;
  __shared__ float sh[3][kThreads];
  float a = 0.f, b = 0.f, c = 0.f;
  if (p.row_sum) {
    // image terms owned by this rank = its local columns
    for (int i = t; i < p.b_local; i += kThreads) {
      const float rs = __ldcg(p.img_rs + p.col0 + i) + p.eps;
      a += -logf(__ldcg(p.img_ps + p.col0 + i) / rs + p.eps);
    }
  } else {
    for (int i = t; i < p.n_total; i += kThreads) a += __ldcg(p.lrow + i);
  }
  for (int i = t; i < p.b_local; i += kThreads) b += __ldcg(p.lcol + i);
  for (int i = t; i < p.n_total; i += kThreads) c += __ldcg(p.dzz_part + i);
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
  return (size_t)(2 * chunks * (long long)b_local + 4);
}

extern "C" size_t rz_mpnce_finish_scratch_floats(int n_total, int b_local, int b_global) {
  return (size_t)(2LL * n_total + 3LL * b_local + 2LL * b_global + 4);
}

extern "C" int rz_mpnce_partials(const float* z, long long ldz, int n_total, int b_local,
                                 const long long* group_map, int col0, float inv_tau,
                                 const float* log_tau, float* rowsum, float* pos, float* colneg,
                                 float* colpos, float* scratch1, void* stream) {
  if (!z || !group_map || !rowsum || !pos || !colneg || !colpos || !scratch1) return RZ_ERR_INVALID;
  if (n_total <= 0 || b_local <= 0 || ldz < b_local) return RZ_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int chunks = (n_total + kRowChunk - 1) / kRowChunk;
  PartialsParams p;
  p.z = z; p.ldz = ldz; p.n_total = n_total; p.b_local = b_local; p.group_map = group_map;
  p.col0 = col0; p.inv_tau = inv_tau; p.log_tau = log_tau;
  p.rowsum = rowsum; p.pos = pos; p.colneg = colneg; p.colpos = colpos;
  p.colpart = scratch1;
  p.barrier = reinterpret_cast<unsigned int*>(scratch1 + 2LL * chunks * b_local);
  RZ_CUDA_OK(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), s));
  const int grid = coop_grid(mpnce_partials_kernel, chunks);
  void* args[] = {&p};
  RZ_CUDA_OK(cudaLaunchCooperativeKernel((const void*)mpnce_partials_kernel, dim3(grid), dim3(kThreads), args, 0, s));
  rz_count_launch(1);
  return RZ_OK;
}

extern "C" int rz_mpnce_finish(const float* z, long long ldz, int n_total, int b_local,
                               int b_global, const long long* group_map, int col0, float inv_tau,
                               const float* log_tau, float eps, int row_sum, int col_sum,
                               const float* rowsum, const float* pos, const float* colneg,
                               const float* colpos, float* scratch2, float* dz, float* loss_terms,
                               void* stream) {
  if (!z || !group_map || !rowsum || !pos || !colneg || !colpos || !scratch2 || !loss_terms)
    return RZ_ERR_INVALID;
  if (n_total <= 0 || b_local <= 0 || b_global < b_local || col0 < 0 || col0 + b_local > b_global ||
      ldz < b_local)
    return RZ_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FinishParams p;
  p.z = z; p.ldz = ldz; p.n_total = n_total; p.b_local = b_local; p.b_global = b_global;
  p.group_map = group_map; p.col0 = col0; p.inv_tau = inv_tau; p.eps = eps; p.log_tau = log_tau;
  p.row_sum = row_sum; p.col_sum = col_sum;
  p.rowsum = rowsum; p.pos = pos; p.colneg = colneg; p.colpos = colpos;
  float* w = scratch2;
  p.acol = w; w += b_local;
  p.apos = w; w += b_local;
  p.lcol = w; w += b_local;
  p.img_rs = w; w += b_global;
  p.img_ps = w; w += b_global;
  p.lrow = w; w += n_total;
  p.dzz_part = w; w += n_total;
  p.barrier = reinterpret_cast<unsigned int*>(w);
  p.dz = dz; p.loss_terms = loss_terms;
  p.inv_2nrow = 0.5f / (float)(row_sum ? b_global : n_total);
  p.inv_2ncol = 0.5f / (float)(col_sum ? b_global : n_total);
  RZ_CUDA_OK(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), s));
  const int grid = coop_grid(mpnce_finish_kernel, (n_total + kWarps - 1) / kWarps);
  void* args[] = {&p};
  RZ_CUDA_OK(cudaLaunchCooperativeKernel((const void*)mpnce_finish_kernel, dim3(grid), dim3(kThreads), args, 0, s));
  rz_count_launch(1);
  return RZ_OK;
}
