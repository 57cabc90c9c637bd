// K10: multi-positive NCE loss (UniCLIP MP-NCE + MIL-NCE variants), forward and backward,
// for a column block [col0, col0+b_local) of the (n_total x b_global) logit matrix.
// Restates multi_positive_nce_loss / get_row_loss / get_col_loss
// (exp/cxr_pt/model/losses.py:243-344) and their autograd in closed form.  HBM-bound
// (read Z twice, write dZ once); all sums use a fixed order -- no float atomics -- so the
// 1-GPU and N-GPU results agree bit for bit on the local block.
#include "rz_common.cuh"

namespace {

constexpr int kRowChunk = 32;
constexpr int kThreads = 256;

// ---- phase 1 ---------------------------------------------------------------------------
// CTA = 32 rows x all local columns (256 at a time).  Thread t owns column c0+t: it walks
// the 32 rows, accumulating the column partials; E is parked in smem so that warp w can
// then reduce rows 4w..4w+3 across the tile.
__global__ void __launch_bounds__(kThreads)
mpnce_partials_kernel(const float* __restrict__ z, long long ldz, int n_total, int b_local,
                      const long long* __restrict__ group_map, int col0, float inv_tau,
                      float* __restrict__ rowsum, float* __restrict__ pos,
                      float* __restrict__ colpart /* [chunks][2][b_local] */) {
  __shared__ float tile[kRowChunk][kThreads + 1];
  __shared__ int gcol[kRowChunk];
  const int chunk = blockIdx.x;
  const int r0 = chunk * kRowChunk;
  const int nrows = min(kRowChunk, n_total - r0);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t < kRowChunk) {
    int g = -1;
    if (t < nrows) {
      const long long gg = group_map[r0 + t] - (long long)col0;
      g = (gg >= 0 && gg < b_local) ? (int)gg : -1;
      if (g < 0) pos[r0 + t] = 0.f;
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
        zz[u] = (c < b_local && rb + u < nrows) ? __ldg(z + (long long)(r0 + rb + u) * ldz + c) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = rb + u;
        if (r >= nrows) break;
        float e = 0.f;
        if (c < b_local) {
          e = __expf(zz[u] * inv_tau);
          if (gcol[r] == c) { cpos += e; pos[r0 + r] = e; } else { cneg += e; }
        }
        tile[r][t] = e;
      }
    }
    if (c < b_local) {
      colpart[((long long)chunk * 2 + 0) * b_local + c] = cneg;
      colpart[((long long)chunk * 2 + 1) * b_local + c] = cpos;
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
      if (r < nrows) rowsum[r0 + r] = racc[k];
    }
  }
}

__global__ void mpnce_colreduce_kernel(const float* __restrict__ colpart, int chunks, int b_local,
                                       float* __restrict__ colneg, float* __restrict__ colpos) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= b_local) return;
  float n = 0.f, p = 0.f;
  for (int k = 0; k < chunks; ++k) {
    n += colpart[((long long)k * 2 + 0) * b_local + c];
    p += colpart[((long long)k * 2 + 1) * b_local + c];
  }
  colneg[c] = n;
  colpos[c] = p;
}

// ---- phase 2 ---------------------------------------------------------------------------
struct FinishParams {
  const float* z; long long ldz; int n_total, b_local, b_global;
  const long long* group_map; int col0; float inv_tau, eps; int row_sum, col_sum;
  const float* rowsum; const float* pos; const float* colneg; const float* colpos;
  float *ca, *cb;            // per row: coefficient on every E_ib / extra on the positive
  float *acol, *apos, *lcol; // per local column: negatives / positives coefficient, loss part
  float *img_rs, *img_ps;    // per global image (row_sum)
  float* dz; float* loss_terms;
  float inv_2nrow, inv_2ncol;
};

// row_sum only: Rs_b = sum_{i in b} R_i, Ps_b = sum_{i in b} pos_i for every GLOBAL image b
__global__ void mpnce_image_sums_kernel(FinishParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.b_global) return;
  float rs = 0.f, ps = 0.f;
  for (int i = 0; i < p.n_total; ++i)
    if (p.group_map[i] == (long long)b) { rs += p.rowsum[i]; ps += p.pos[i]; }
  p.img_rs[b] = rs;
  p.img_ps[b] = ps;
}

// per local column: coefficient applied to the negatives / positives of that column and the
// column loss terms it owns.  One WARP per column: the lanes stride over the sentences and are
// combined in a fixed order (bit-reproducible, independent of the number of ranks).
__global__ void __launch_bounds__(256)
mpnce_col_coeff_kernel(FinishParams p) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= p.b_local) return;
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
      if (p.group_map[i] != gcol) continue;
      const float ps = p.pos[i];
      const float den = ps + cn + p.eps;
      const float pc = ps / den;
      const float u = 1.0f / (pc + p.eps);
      l += -logf(pc + p.eps);
      a += u * ps / (den * den);
    }
    a = rz::warp_sum(a);          // xor-butterfly: the same order on every lane and every run
    l = rz::warp_sum(l);
    a *= p.inv_2ncol;
  }
  if (lane == 0) {
    p.acol[c] = a;
    p.apos[c] = ap;
    p.lcol[c] = l;
  }
}

// per row: coefficients of the row term (and, for MP-NCE columns, the positive's own
// column-term derivative).  Also the row loss terms, reduced deterministically later.
__global__ void mpnce_row_coeff_kernel(FinishParams p, float* __restrict__ lrow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_total) return;
  const long long g = p.group_map[i];
  const int gl = (int)(g - p.col0);
  const bool local = (g >= p.col0 && gl < p.b_local);
  float ca, cb, l = 0.f;
  if (p.row_sum) {
    // losses.py:303-315: one term per image
    const float rs = p.img_rs[g] + p.eps, ps = p.img_ps[g];
    const float pr = ps / rs;
    const float w = 1.0f / (pr + p.eps);
    ca = w * ps / (rs * rs) * p.inv_2nrow;
    cb = -w / rs * p.inv_2nrow;
  } else {
    // losses.py:316-320
    const float r = p.rowsum[i] + p.eps, ps = p.pos[i];
    const float pr = ps / r;
    const float w = 1.0f / (pr + p.eps);
    ca = w * ps / (r * r) * p.inv_2nrow;
    cb = -w / r * p.inv_2nrow;
    if (local) l = -logf(pr + p.eps);
  }
  if (!p.col_sum && local) {
    const float ps = p.pos[i];
    const float den = ps + p.colneg[gl] + p.eps;
    const float pc = ps / den;
    const float u = 1.0f / (pc + p.eps);
    cb += -u * (den - ps) / (den * den) * p.inv_2ncol;
  }
  p.ca[i] = ca;
  p.cb[i] = cb;
  lrow[i] = l;
}

// dZ_ib = E_ib/tau * (ca_i + [b==g_i](cb_i + apos_b) + [b!=g_i] acol_b); partial sum dZ*Z
__global__ void __launch_bounds__(kThreads)
mpnce_dz_kernel(FinishParams p, float* __restrict__ dzz_part) {
  const int r = blockIdx.x;  // one CTA per row
  const long long g = p.group_map[r];
  const int gl = (g >= p.col0 && g - p.col0 < p.b_local) ? (int)(g - p.col0) : -1;
  const float ca = p.ca[r], cb = p.cb[r];
  float acc = 0.f;
  for (int c = threadIdx.x; c < p.b_local; c += kThreads) {
    const float zz = p.z[(long long)r * p.ldz + c];
    const float e = __expf(zz * p.inv_tau);
    const float coef = ca + (c == gl ? cb + p.apos[c] : p.acol[c]);
    const float d = e * p.inv_tau * coef;
    if (p.dz != nullptr) p.dz[(long long)r * p.ldz + c] = d;
    acc = fmaf(d, zz, acc);
  }
  acc = rz::warp_sum(acc);
  __shared__ float w[kThreads / 32];
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < kThreads / 32; ++k) s += w[k];
    dzz_part[r] = s;
  }
}

// single CTA, fixed-order tree: loss_terms = {sum row terms, sum col terms, sum dZ*Z, 0}
__global__ void __launch_bounds__(1024)
mpnce_terms_kernel(FinishParams p, const float* __restrict__ lrow, const float* __restrict__ dzz_part) {
  __shared__ float sh[3][1024];
  const int t = threadIdx.x;
  float a = 0.f, b = 0.f, c = 0.f;
  if (p.row_sum) {
    // image terms owned by this rank = its local columns
    for (int i = t; i < p.b_local; i += 1024) {
      const float rs = p.img_rs[p.col0 + i] + p.eps;
      a += -logf(p.img_ps[p.col0 + i] / rs + p.eps);
    }
  } else {
    for (int i = t; i < p.n_total; i += 1024) a += lrow[i];
  }
  for (int i = t; i < p.b_local; i += 1024) b += p.lcol[i];
  for (int i = t; i < p.n_total; i += 1024) c += dzz_part[i];
  sh[0][t] = a; sh[1][t] = b; sh[2][t] = c;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (t < s) { sh[0][t] += sh[0][t + s]; sh[1][t] += sh[1][t + s]; sh[2][t] += sh[2][t + s]; }
    __syncthreads();
  }
  if (t == 0) {
    p.loss_terms[0] = sh[0][0]; p.loss_terms[1] = sh[1][0];
    p.loss_terms[2] = sh[2][0]; p.loss_terms[3] = 0.f;
  }
}

}  // namespace

extern "C" int rz_mpnce_partials(const float* z, long long ldz, int n_total, int b_local,
                                 const long long* group_map, int col0, float inv_tau,
                                 float* rowsum, float* pos, float* colneg, float* colpos,
                                 float* scratch1, void* stream) {
  if (!z || !group_map || !rowsum || !pos || !colneg || !colpos || !scratch1) return RZ_ERR_INVALID;
  if (n_total <= 0 || b_local <= 0 || ldz < b_local) return RZ_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int chunks = (n_total + kRowChunk - 1) / kRowChunk;
  mpnce_partials_kernel<<<chunks, kThreads, 0, s>>>(z, ldz, n_total, b_local, group_map, col0,
                                                    inv_tau, rowsum, pos, scratch1);
  RZ_LAUNCH_OK();
  mpnce_colreduce_kernel<<<(b_local + 255) / 256, 256, 0, s>>>(scratch1, chunks, b_local, colneg, colpos);
  RZ_LAUNCH_OK();
  rz_count_launch(2);
  return RZ_OK;
}

extern "C" int rz_mpnce_finish(const float* z, long long ldz, int n_total, int b_local,
                               int b_global, const long long* group_map, int col0, float inv_tau,
                               float eps, int row_sum, int col_sum, const float* rowsum,
                               const float* pos, const float* colneg, const float* colpos,
                               float* scratch2, float* dz, float* loss_terms, void* stream) {
  if (!z || !group_map || !rowsum || !pos || !colneg || !colpos || !scratch2 || !loss_terms)
    return RZ_ERR_INVALID;
  if (n_total <= 0 || b_local <= 0 || b_global < b_local || col0 < 0 || col0 + b_local > b_global ||
      ldz < b_local)
    return RZ_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FinishParams p;
  p.z = z; p.ldz = ldz; p.n_total = n_total; p.b_local = b_local; p.b_global = b_global;
  p.group_map = group_map; p.col0 = col0; p.inv_tau = inv_tau; p.eps = eps;
  p.row_sum = row_sum; p.col_sum = col_sum;
  p.rowsum = rowsum; p.pos = pos; p.colneg = colneg; p.colpos = colpos;
  float* w = scratch2;
  p.ca = w; w += n_total;
  p.cb = w; w += n_total;
  p.acol = w; w += b_local;
  p.apos = w; w += b_local;
  p.lcol = w; w += b_local;
  p.img_rs = w; w += b_global;
  p.img_ps = w; w += b_global;
  float* lrow = w; w += n_total;
  float* dzz_part = w; w += n_total;
  p.dz = dz; p.loss_terms = loss_terms;
  p.inv_2nrow = 0.5f / (float)(row_sum ? b_global : n_total);
  p.inv_2ncol = 0.5f / (float)(col_sum ? b_global : n_total);
  int launches = 0;
  if (row_sum) {
    mpnce_image_sums_kernel<<<(b_global + 127) / 128, 128, 0, s>>>(p);
    RZ_LAUNCH_OK(); ++launches;
  }
  mpnce_col_coeff_kernel<<<(b_local + 7) / 8, 256, 0, s>>>(p);
  RZ_LAUNCH_OK(); ++launches;
  mpnce_row_coeff_kernel<<<(n_total + 127) / 128, 128, 0, s>>>(p, lrow);
  RZ_LAUNCH_OK(); ++launches;
  mpnce_dz_kernel<<<n_total, kThreads, 0, s>>>(p, dzz_part);
  RZ_LAUNCH_OK(); ++launches;
  mpnce_terms_kernel<<<1, 1024, 0, s>>>(p, lrow, dzz_part);
  RZ_LAUNCH_OK(); ++launches;
  rz_count_launch(launches);
  return RZ_OK;
}
