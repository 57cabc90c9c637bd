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
    if (MODE == RZ_UP_SIGMOID) { v0 *= 0.5f; v1 *= 0.5f; }     // the consumer wants v / 2
    rowbuf[r * kRowStride + c] = make_float2(v0, v1 - v0);
  }
  __syncthreads();

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

template <int MODE, int V>
int launch_v(const UpParams& p, int maps, cudaStream_t s) {
  dim3 grid((p.out_h + kBand - 1) / kBand, maps), block(kThreads);
  const int wpad = (p.out_w + V - 1) / V * V;
  const size_t smem = (size_t)(2 * wpad + 2 * kBand * kRowStride + p.grid * p.grid) * sizeof(float);
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
