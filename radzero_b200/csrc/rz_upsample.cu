// K8+K9: bilinear upsample (align_corners=False) of patch-grid similarity maps to pixel
// resolution, fused with sigmoid / threshold / global argmax.  Pure write bandwidth:
// the grid rows a band needs are staged in shared memory, interpolated vertically once per
// output row, and every thread emits V horizontally adjacent pixels as one vector store.
// Replaces F.interpolate + consumers in exp/cxr_pt/inference/segmentation_utils.py:36-122,
// :225, :258 and exp/cxr_pt/inference/grounding_utils.py:166-261.
#include "rz_common.cuh"

namespace {

constexpr int kBand = 64;      // canvas rows per CTA (amortises the x-table and the row interpolation)
constexpr int kThreads = 256;
constexpr int kMaxGrid = 64;
constexpr int kRowStride = kMaxGrid + 1;  // rowbuf pitch in float2 pairs (compile-time: immediate LDS offsets)
constexpr int kRowGroup = 4;               // rows per work item
constexpr int kMaxOutW = 16384; // x-table lives in shared memory (8 B per canvas column)

struct UpParams {
  const float* scores;
  long long map_stride;
  int grid, out_h, out_w, interp_h, interp_w, off_y, off_x;
  float fill, scale_h, scale_w, thr_logit;
  void* out;
  unsigned long long* keys;
  // threshold statistics (rz_map_threshold_stats)
  const unsigned char* gt; const float* thr; int n_thr;
  unsigned int* hist_all; unsigned int* hist_gt; unsigned int* vmax_bits;
};

__device__ __forceinline__ void src_index(float scale, int dst, int n_in, int& i0, float& w1) {
  // ATen upsample_bilinear2d, align_corners=False: src = max(0, scale*(dst+0.5)-0.5);
  // i1 = i0 + (i0 < n_in-1) is realised by a duplicated last column / row in shared memory
  float src = fmaxf(fmaf(scale, (float)dst + 0.5f, -0.5f), 0.0f);
  i0 = min((int)src, n_in - 1);
  w1 = src - (float)i0;
}

// sigmoid(v) = 0.5 + 0.5 tanh(v / 2), given h = v / 2: ONE MUFU op (the kernel is otherwise bound by
// the 16 lanes/clk MUFU pipe: ex2 + rcp would be two).  tanh.approx has 2^-11 relative error, i.e.
// <= 2.5e-4 absolute on the probability -- inside the 2e-3 map tolerance; thresholded masks never
// go through it (RZ_UP_MASK compares in the score domain).
__device__ __forceinline__ float sigmoid_from_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(t, 0.5f, 0.5f);
}

__device__ __forceinline__ unsigned long long argmax_key(float v, unsigned int flat) {
  unsigned int u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
}

// Stage one band: the grid rows it touches, the x-table and the vertically interpolated pair rows.
template <bool kHalf>
__device__ __forceinline__ void band_setup(const UpParams& p, const float* __restrict__ g, int G, int wpad,
                                           int y_first, int rows_here, float2* xtab, float2* rowbuf,
                                           float* gbuf) {
  // grid rows this band touches (vertical source index is monotone in y)
  int ylo = G, yhi = -1;
  {
    const int iy0 = max(y_first - p.off_y, 0), iy1 = min(y_first + rows_here - 1 - p.off_y, p.interp_h - 1);
    if (iy0 <= iy1) {
      int a; float w;
      src_index(p.scale_h, iy0, G, a, w); ylo = a;
      src_index(p.scale_h, iy1, G, a, w); yhi = min(a + 1, G - 1);
    }
  }
  if (yhi >= ylo) {
    const int n = (yhi - ylo + 1) * G;
    for (int i = threadIdx.x; i < n; i += kThreads) gbuf[i] = __ldg(g + ylo * G + i);
  }
  for (int x = threadIdx.x; x < wpad; x += kThreads) {
    const int ix = x - p.off_x;
    int a = 0; float w = 0.f;
    if (ix >= 0 && ix < p.interp_w && x < p.out_w) src_index(p.scale_w, ix, G, a, w);
    else a = (int)0x80000000u;
    xtab[x] = make_float2(__int_as_float(a), w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows_here * G; i += kThreads) {
    const int r = i / G, c = i - r * G, c1 = min(c + 1, G - 1);
    const int iy = y_first + r - p.off_y;
    float v0 = 0.f, v1 = 0.f;
    if (iy >= 0 && iy < p.interp_h) {
      int a; float w;
      src_index(p.scale_h, iy, G, a, w);
      const int b = min(a + 1, G - 1);
      const float* ga = gbuf + (a - ylo) * G;
      const float* gb = gbuf + (b - ylo) * G;
      v0 = (1.0f - w) * ga[c] + w * gb[c];
      v1 = (1.0f - w) * ga[c1] + w * gb[c1];
    }
    if (kHalf) { v0 *= 0.5f; v1 *= 0.5f; }                     // the consumer wants v / 2
    rowbuf[r * kRowStride + c] = make_float2(v0, v1 - v0);
  }
  __syncthreads();

}

// x-table entry: source column a (bit 31 set = outside the pasted area -> fill) and weight of
// column a+1.  One CTA = one band of kBand canvas rows of one map:
//   1. the grid rows the band touches are staged in shared memory,
//   2. each canvas row is interpolated vertically ONCE into rowbuf[r][0..G) as PAIRS
//      (v[a], v[a+1] - v[a]) (a+1 clamped to the last column),
//   3. the x-table (a, w) of every canvas column is computed once per CTA,
//   4. threads sweep the band's V-pixel vectors: 1 table load per item, then ONE 64-bit shared
//      load and ONE fma per pixel, the fused consumer, one coalesced streaming vector store.
template <int MODE, int V, bool PLAIN>
__global__ void __launch_bounds__(kThreads)
upsample_kernel(UpParams p) {
  extern __shared__ __align__(16) float smem[];
  const int G = p.grid;
  float2* xtab = reinterpret_cast<float2*>(smem);               // [out_w rounded up to V]
  const int wpad = (p.out_w + V - 1) / V * V;
  float2* rowbuf = reinterpret_cast<float2*>(smem + 2 * wpad);  // [kBand][kRowStride] (v[a], v[a+1]-v[a])
  float* gbuf = smem + 2 * wpad + 2 * kBand * kRowStride;       // [rows needed <= G][G]
  const int map = blockIdx.y;
  const int y_first = blockIdx.x * kBand;
  const int rows_here = min(kBand, p.out_h - y_first);
  const float* g = p.scores + (long long)map * p.map_stride;

  band_setup<MODE == RZ_UP_SIGMOID>(p, g, G, wpad, y_first, rows_here, xtab, rowbuf, gbuf);
  // RZ_UP_MASK_BITS: the band's mask as a bitmap in shared memory ([kBand][words per row]); threads OR their
  // V bits in, the band then leaves as whole 32-bit words (32 pixels per word instead of 32 bytes)
  const int wpw = (p.out_w + 31) >> 5;
  unsigned int* bitrows = reinterpret_cast<unsigned int*>(gbuf + G * G);
  if constexpr (MODE == RZ_UP_MASK_BITS) {
    for (int i = threadIdx.x; i < kBand * wpw; i += kThreads) bitrows[i] = 0u;
    __syncthreads();
  }

  // work item = V adjacent canvas columns x kRowGroup consecutive rows: the x-table entry is
  // loaded once per item and the row loop is unrolled with immediate shared-memory offsets
  const int vec_per_row = wpad / V;
  const int n_groups = (rows_here + kRowGroup - 1) / kRowGroup;
  const int items = vec_per_row * n_groups;
  unsigned long long best = 0ull;
  int grp = threadIdx.x / vec_per_row;
  int xv = threadIdx.x - grp * vec_per_row;
  for (int it = threadIdx.x; it < items; it += kThreads) {
    float2 t[V];
    if constexpr (V == 4) {
      const float4 t0 = *reinterpret_cast<const float4*>(xtab + xv * 4);
      const float4 t1 = *reinterpret_cast<const float4*>(xtab + xv * 4 + 2);
      t[0] = make_float2(t0.x, t0.y); t[1] = make_float2(t0.z, t0.w);
      t[2] = make_float2(t1.x, t1.y); t[3] = make_float2(t1.z, t1.w);
    } else if constexpr (V == 2) {
      const float4 t0 = *reinterpret_cast<const float4*>(xtab + xv * 2);
      t[0] = make_float2(t0.x, t0.y); t[1] = make_float2(t0.z, t0.w);
    } else {
      t[0] = xtab[xv];
    }
    const float2* src[V];
    float w1[V];
    bool inside[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int a = __float_as_int(t[k].x);
      inside[k] = PLAIN || a >= 0;
      src[k] = rowbuf + (grp * kRowGroup) * kRowStride + (a & 0xffff);
      w1[k] = t[k].y;
    }
    const int y0 = y_first + grp * kRowGroup;
    long long o = ((long long)map * p.out_h + y0) * p.out_w + (long long)xv * V;
#pragma unroll
    for (int rr = 0; rr < kRowGroup; ++rr, o += p.out_w) {
      const int y = y0 + rr;
      if (y >= p.out_h) break;
      bool row_in = true;
      if constexpr (!PLAIN) { const int iy = y - p.off_y; row_in = (iy >= 0 && iy < p.interp_h); }
      float val[V];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float2 pd = src[k][rr * kRowStride];
        const float v = fmaf(w1[k], pd.y, pd.x);
        val[k] = (PLAIN || (inside[k] && row_in)) ? v : (MODE == RZ_UP_SIGMOID ? 0.5f * p.fill : p.fill);
      }
      if constexpr (MODE == RZ_UP_RAW || MODE == RZ_UP_SIGMOID) {
        if constexpr (MODE == RZ_UP_SIGMOID) {
#pragma unroll
          for (int k = 0; k < V; ++k) val[k] = sigmoid_from_half(val[k]);
        }
        float* out = static_cast<float*>(p.out) + o;
        if constexpr (V == 4) {
          __stcs(reinterpret_cast<float4*>(out), make_float4(val[0], val[1], val[2], val[3]));
        } else if constexpr (V == 2) {
          __stcs(reinterpret_cast<float2*>(out), make_float2(val[0], val[1]));
        } else {
          __stcs(out, val[0]);
        }
      } else if constexpr (MODE == RZ_UP_MASK) {
        // sigmoid(v) > t  <=>  v > logit(t) (monotone); compare in the score domain
        unsigned char* out = static_cast<unsigned char*>(p.out) + o;
        if constexpr (V == 4) {
          unsigned int w = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) w |= (val[k] > p.thr_logit ? 1u : 0u) << (8 * k);
          __stcs(reinterpret_cast<unsigned int*>(out), w);
        } else if constexpr (V == 2) {
          const unsigned short w = (unsigned short)((val[0] > p.thr_logit ? 1u : 0u) |
                                                    ((val[1] > p.thr_logit ? 1u : 0u) << 8));
          __stcs(reinterpret_cast<unsigned short*>(out), w);
        } else {
          out[0] = val[0] > p.thr_logit ? 1 : 0;
        }
      } else if constexpr (MODE == RZ_UP_MASK_BITS) {
        unsigned int nib = 0;
#pragma unroll
        for (int k = 0; k < V; ++k) nib |= (xv * V + k < p.out_w && val[k] > p.thr_logit ? 1u : 0u) << k;
        const int x0 = xv * V;                               // a multiple of V <= 4: the bits stay in one word
        if (nib) atomicOr(bitrows + (grp * kRowGroup + rr) * wpw + (x0 >> 5), nib << (x0 & 31));
      } else {  // ARGMAX
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const unsigned int flat = (unsigned int)(y * p.out_w + xv * V + k);
          const unsigned long long key = argmax_key(val[k], flat);
          best = key > best ? key : best;
        }
      }
    }
    xv += kThreads;
    while (xv >= vec_per_row) { xv -= vec_per_row; ++grp; }
  }
  if constexpr (MODE == RZ_UP_MASK_BITS) {
    __syncthreads();
    unsigned int* out = static_cast<unsigned int*>(p.out) + ((long long)map * p.out_h + y_first) * wpw;
    for (int i = threadIdx.x; i < rows_here * wpw; i += kThreads) __stcs(out + i, bitrows[i]);
  }
  if constexpr (MODE == RZ_UP_ARGMAX) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    __shared__ unsigned long long wbest[kThreads / 32];
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b = wbest[0];
      for (int w = 1; w < kThreads / 32; ++w) b = wbest[w] > b ? wbest[w] : b;
      atomicMax(p.keys + 2 * map, b);  // key of map m lives in out[2*m] until decoded
    }
  }
}

// keys live in out[2*m] during the reduction; decode in place to (x, y)
__global__ void argmax_decode_kernel(long long* out, int maps, int out_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= maps) return;
  const unsigned long long key = (unsigned long long)out[2 * i];
  const unsigned int flat = 0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull);
  out[2 * i] = (long long)(flat % (unsigned int)out_w);
  out[2 * i + 1] = (long long)(flat / (unsigned int)out_w);
}


// ---- fused consumer: threshold statistics without ever writing the pixel map ------------------
// For every map: hist_all[k] = number of canvas pixels whose interpolated score v satisfies
// (#thresholds below v) == k, hist_gt[k] = the same restricted to ground-truth pixels, and max v.
// A suffix sum over k gives |P_t| and |P_t & G| for all thresholds at once -- the Dice sweep of
// exp/cxr_pt/inference/segmentation_utils.py:255-261 and compute_specificity (:136-158) -- from one
// pass that reads 5.5 KB of scores + H*W bytes of mask per map instead of writing H*W*4 bytes,
// copying them to the host and thresholding them 101 times.
constexpr int kMaxThr = 127;
constexpr int kBins = kMaxThr + 1;

__device__ __forceinline__ unsigned int order_bits(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kThreads)
map_stats_kernel(UpParams p) {
  extern __shared__ __align__(16) float smem[];
  const int G = p.grid;
  float2* xtab = reinterpret_cast<float2*>(smem);
  const int wpad = p.out_w;
  float2* rowbuf = reinterpret_cast<float2*>(smem + 2 * wpad);
  float* gbuf = smem + 2 * wpad + 2 * kBand * kRowStride;
  float* thr = gbuf + G * G;                                          // [kBins]: thresholds, then +inf
  unsigned int* hist = reinterpret_cast<unsigned int*>(thr + kBins);  // [warps][2][kBins]
  const int map = blockIdx.y;
  const int y_first = blockIdx.x * kBand;
  const int rows_here = min(kBand, p.out_h - y_first);
  const float* g = p.scores + (long long)map * p.map_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kBins; i += kThreads) thr[i] = i < p.n_thr ? __ldg(p.thr + i) : INFINITY;
  for (int i = threadIdx.x; i < (kThreads / 32) * 2 * kBins; i += kThreads) hist[i] = 0u;
  band_setup<false>(p, g, G, wpad, y_first, rows_here, xtab, rowbuf, gbuf);   // ends with __syncthreads
  unsigned int* hw = hist + warp * 2 * kBins;
  const unsigned char* gt = p.gt != nullptr ? p.gt + (long long)map * p.out_h * p.out_w : nullptr;
  const int items = rows_here * p.out_w;
  float vmax = -INFINITY;
  for (int it0 = warp * 32; it0 < items; it0 += kThreads) {           // warp-uniform trip count
    const int it = it0 + lane;
    const bool valid = it < items;
    int k = -1;
    bool isgt = false;
    if (valid) {
      const int r = it / p.out_w, x = it - r * p.out_w;
      const int y = y_first + r;
      const float2 t = xtab[x];
      const int a = __float_as_int(t.x);
      const int iy = y - p.off_y;
      float v = p.fill;
      if (a >= 0 && iy >= 0 && iy < p.interp_h) {
        const float2 pd = rowbuf[r * kRowStride + (a & 0xffff)];
        v = fmaf(t.y, pd.y, pd.x);
      }
      vmax = fmaxf(vmax, v);
      // k = number of thresholds strictly below v (thr ascending, padded with +inf): 7 probes
      int lo = 0;
#pragma unroll
      for (int step = kBins / 2; step > 0; step >>= 1)
        if (thr[lo + step - 1] < v) lo += step;
      k = lo;
      if (gt != nullptr) isgt = gt[(long long)y * p.out_w + x] != 0;
    }
    // warp-aggregated histogram update: neighbouring pixels mostly share a bin
    const unsigned int peers = __match_any_sync(0xffffffffu, k);
    const unsigned int gtb = __ballot_sync(0xffffffffu, isgt);
    if (k >= 0 && lane == __ffs(peers) - 1) {
      hw[k] += __popc(peers);
      hw[kBins + k] += __popc(peers & gtb);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kBins; i += kThreads) {
    unsigned int s = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += hist[w * 2 * kBins + i];
    const int k = i % kBins;
    if (s != 0u && k <= p.n_thr)
      atomicAdd((i < kBins ? p.hist_all : p.hist_gt) + (long long)map * (p.n_thr + 1) + k, s);
  }
  vmax = rz::warp_max(vmax);
  if (lane == 0 && vmax > -INFINITY) atomicMax(p.vmax_bits + map, order_bits(vmax));
}

__global__ void decode_max_kernel(unsigned int* bits, int maps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= maps) return;
  const unsigned int u = bits[i];
  bits[i] = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;       // back to the float's bit pattern
}

template <int MODE, int V>
int launch_v(const UpParams& p, int maps, cudaStream_t s) {
  dim3 grid((p.out_h + kBand - 1) / kBand, maps), block(kThreads);
  const int wpad = (p.out_w + V - 1) / V * V;
  const size_t smem = (size_t)(2 * wpad + 2 * kBand * kRowStride + p.grid * p.grid +
                               (MODE == RZ_UP_MASK_BITS ? kBand * ((p.out_w + 31) / 32) : 0)) * sizeof(float);
  // PLAIN: the resized grid covers the canvas exactly (BlipImageProcessor branch) -- no fill
  const bool plain = p.off_x == 0 && p.off_y == 0 && p.interp_h == p.out_h && p.interp_w == p.out_w;
  if (plain) {
    RZ_CUDA_OK(cudaFuncSetAttribute(upsample_kernel<MODE, V, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upsample_kernel<MODE, V, true><<<grid, block, smem, s>>>(p);
  } else {
    RZ_CUDA_OK(cudaFuncSetAttribute(upsample_kernel<MODE, V, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upsample_kernel<MODE, V, false><<<grid, block, smem, s>>>(p);
  }
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}

template <int MODE>
int launch_mode(const UpParams& p, int maps, cudaStream_t s) {
  if (MODE == RZ_UP_ARGMAX) return launch_v<MODE, 1>(p, maps, s);
  // vector stores need every canvas row to start on a vector boundary
  const uintptr_t base = reinterpret_cast<uintptr_t>(p.out);
  const int esz = (MODE == RZ_UP_MASK) ? 1 : 4;
  if (MODE == RZ_UP_MASK_BITS) {      // no per-pixel stores: only the column padding of the item grid matters
    if (base % 4) return RZ_ERR_ALIGNMENT;
    return launch_v<MODE, 4>(p, maps, s);
  }
  if (p.out_w % 4 == 0 && base % (4 * esz) == 0) return launch_v<MODE, 4>(p, maps, s);
  if (p.out_w % 2 == 0 && base % (2 * esz) == 0) return launch_v<MODE, 2>(p, maps, s);
  return launch_v<MODE, 1>(p, maps, s);
}

}  // namespace

extern "C" int rz_upsample_maps(const float* scores, long long map_stride, int maps, int grid,
                                int out_h, int out_w, int interp_h, int interp_w, int off_y,
                                int off_x, float fill, int mode, float threshold, void* out,
                                void* stream) {
  if (scores == nullptr || out == nullptr) return RZ_ERR_INVALID;
  if (maps < 0 || grid <= 0 || grid > kMaxGrid || out_h <= 0 || out_w <= 0 || interp_h <= 0 ||
      interp_w <= 0)
    return RZ_ERR_INVALID;
  if (out_w > kMaxOutW) return RZ_ERR_UNSUPPORTED;
  if (maps == 0) return RZ_OK;
  if (maps > 65535) return RZ_ERR_UNSUPPORTED;  // gridDim.y; callers chunk larger batches
  if ((long long)out_h * out_w >= (1ll << 32)) return RZ_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  UpParams p;
  p.scores = scores; p.map_stride = map_stride; p.grid = grid;
  p.out_h = out_h; p.out_w = out_w; p.interp_h = interp_h; p.interp_w = interp_w;
  p.off_y = off_y; p.off_x = off_x; p.fill = fill;
  p.scale_h = (float)grid / (float)interp_h;
  p.scale_w = (float)grid / (float)interp_w;
  // logit(t); t<=0 -> everything passes, t>=1 -> nothing does
  if (threshold <= 0.f) p.thr_logit = -INFINITY;
  else if (threshold >= 1.f) p.thr_logit = INFINITY;
  else p.thr_logit = logf(threshold / (1.0f - threshold));
  p.out = out;
  p.keys = nullptr;
  switch (mode) {
    case RZ_UP_RAW: return launch_mode<RZ_UP_RAW>(p, maps, s);
    case RZ_UP_SIGMOID: return launch_mode<RZ_UP_SIGMOID>(p, maps, s);
    case RZ_UP_MASK: return launch_mode<RZ_UP_MASK>(p, maps, s);
    case RZ_UP_MASK_BITS: return launch_mode<RZ_UP_MASK_BITS>(p, maps, s);
    case RZ_UP_ARGMAX: {
      unsigned long long* keys = static_cast<unsigned long long*>(out);
      RZ_CUDA_OK(cudaMemsetAsync(keys, 0, 2 * sizeof(unsigned long long) * (size_t)maps, s));
      p.keys = keys;
      int rc = launch_mode<RZ_UP_ARGMAX>(p, maps, s);
      if (rc != RZ_OK) return rc;
      argmax_decode_kernel<<<(maps + 255) / 256, 256, 0, s>>>(static_cast<long long*>(out), maps, out_w);
      RZ_LAUNCH_OK();
      rz_count_launch();
      return RZ_OK;
    }
    default: return RZ_ERR_INVALID;
  }
}

extern "C" int rz_map_threshold_stats(const float* scores, long long map_stride, int maps, int grid,
                                      int out_h, int out_w, int interp_h, int interp_w, int off_y,
                                      int off_x, float fill, const unsigned char* gt_masks,
                                      const float* thresholds_logit, int n_thresholds,
                                      unsigned int* hist_all, unsigned int* hist_gt, float* max_score,
                                      void* stream) {
  if (scores == nullptr || thresholds_logit == nullptr || hist_all == nullptr || hist_gt == nullptr ||
      max_score == nullptr)
    return RZ_ERR_INVALID;
  if (maps < 0 || grid <= 0 || grid > kMaxGrid || out_h <= 0 || out_w <= 0 || interp_h <= 0 ||
      interp_w <= 0 || n_thresholds <= 0)
    return RZ_ERR_INVALID;
  if (n_thresholds > kMaxThr || out_w > kMaxOutW || maps > 65535) return RZ_ERR_UNSUPPORTED;
  if (maps == 0) return RZ_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  UpParams p;
  p.scores = scores; p.map_stride = map_stride; p.grid = grid;
  p.out_h = out_h; p.out_w = out_w; p.interp_h = interp_h; p.interp_w = interp_w;
  p.off_y = off_y; p.off_x = off_x; p.fill = fill;
  p.scale_h = (float)grid / (float)interp_h;
  p.scale_w = (float)grid / (float)interp_w;
  p.thr_logit = 0.f; p.out = nullptr; p.keys = nullptr;
  p.gt = gt_masks; p.thr = thresholds_logit; p.n_thr = n_thresholds;
  p.hist_all = hist_all; p.hist_gt = hist_gt; p.vmax_bits = reinterpret_cast<unsigned int*>(max_score);
  const size_t bins = (size_t)maps * (n_thresholds + 1) * sizeof(unsigned int);
  RZ_CUDA_OK(cudaMemsetAsync(hist_all, 0, bins, s));
  RZ_CUDA_OK(cudaMemsetAsync(hist_gt, 0, bins, s));
  RZ_CUDA_OK(cudaMemsetAsync(max_score, 0, (size_t)maps * sizeof(float), s));
  const size_t smem = (size_t)(2 * out_w + 2 * kBand * kRowStride + grid * grid + kBins) * sizeof(float) +
                      (size_t)(kThreads / 32) * 2 * kBins * sizeof(unsigned int);
  RZ_CUDA_OK(cudaFuncSetAttribute(map_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid_dim((out_h + kBand - 1) / kBand, maps);
  map_stats_kernel<<<grid_dim, kThreads, smem, s>>>(p);
  RZ_LAUNCH_OK();
  decode_max_kernel<<<(maps + 255) / 256, 256, 0, s>>>(reinterpret_cast<unsigned int*>(max_score), maps);
  RZ_LAUNCH_OK();
  rz_count_launch(2);
  return RZ_OK;
}
