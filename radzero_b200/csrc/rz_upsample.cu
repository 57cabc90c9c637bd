// K8+K9: bilinear upsample (align_corners=False) of patch-grid similarity maps to pixel
// resolution, fused with sigmoid / threshold / global argmax.  Pure write bandwidth:
// the grid rows a band needs are staged in shared memory, interpolated vertically once per
// output row, and every thread emits V horizontally adjacent pixels as one vector store.
// Replaces F.interpolate + consumers in exp/cxr_pt/inference/segmentation_utils.py:36-122,
// :225, :258 and exp/cxr_pt/inference/grounding_utils.py:166-261.
#include "rz_common.cuh"

namespace {

constexpr int kBand = 8;       // canvas rows per CTA
constexpr int kThreads = 256;
constexpr int kMaxGrid = 64;

struct UpParams {
  const float* scores;
  long long map_stride;
  int grid, out_h, out_w, interp_h, interp_w, off_y, off_x;
  float fill, scale_h, scale_w, thr_logit;
  void* out;
  unsigned long long* keys;
};

__device__ __forceinline__ void src_index(float scale, int dst, int n_in, int& i0, int& i1, float& w1) {
  // ATen upsample_bilinear2d, align_corners=False: src = max(0, scale*(dst+0.5)-0.5)
  float src = fmaxf(fmaf(scale, (float)dst + 0.5f, -0.5f), 0.0f);
  i0 = min((int)src, n_in - 1);
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  w1 = src - (float)i0;
}

__device__ __forceinline__ float fast_sigmoid(float v) { return 1.0f / (1.0f + __expf(-v)); }

__device__ __forceinline__ unsigned long long argmax_key(float v, unsigned int flat) {
  unsigned int u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
}

template <int MODE, int V>
__global__ void __launch_bounds__(kThreads)
upsample_kernel(UpParams p) {
  extern __shared__ float smem[];
  float* gbuf = smem;                                   // [rows_needed][grid]
  float* rowbuf = smem + kMaxGrid * kMaxGrid;           // [kBand][grid]
  const int map = blockIdx.y;
  const int y_first = blockIdx.x * kBand;
  const int rows_here = min(kBand, p.out_h - y_first);
  const float* g = p.scores + (long long)map * p.map_stride;
  const int G = p.grid;

  // grid rows this band touches
  int ylo = G, yhi = -1;
  for (int r = 0; r < rows_here; ++r) {
    const int iy = y_first + r - p.off_y;
    if (iy < 0 || iy >= p.interp_h) continue;
    int a, b; float w;
    src_index(p.scale_h, iy, G, a, b, w);
    ylo = min(ylo, a); yhi = max(yhi, b);
  }
  if (yhi >= ylo) {
    const int n = (yhi - ylo + 1) * G;
    for (int i = threadIdx.x; i < n; i += kThreads) gbuf[i] = __ldg(g + ylo * G + i);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows_here * G; i += kThreads) {
    const int r = i / G, c = i - r * G;
    const int iy = y_first + r - p.off_y;
    float v = 0.f;
    if (iy >= 0 && iy < p.interp_h) {
      int a, b; float w;
      src_index(p.scale_h, iy, G, a, b, w);
      v = (1.0f - w) * gbuf[(a - ylo) * G + c] + w * gbuf[(b - ylo) * G + c];
    }
    rowbuf[r * G + c] = v;
  }
  __syncthreads();

  const int vec_per_row = p.out_w / V;
  unsigned long long best = 0ull;
  for (int i = threadIdx.x; i < rows_here * vec_per_row; i += kThreads) {
    const int r = i / vec_per_row;
    const int xv = i - r * vec_per_row;
    const int y = y_first + r;
    const int iy = y - p.off_y;
    const bool row_in = (iy >= 0 && iy < p.interp_h);
    float val[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int ix = xv * V + k - p.off_x;
      float v = p.fill;
      if (row_in && ix >= 0 && ix < p.interp_w) {
        int a, b; float w;
        src_index(p.scale_w, ix, G, a, b, w);
        v = (1.0f - w) * rowbuf[r * G + a] + w * rowbuf[r * G + b];
      }
      val[k] = v;
    }
    const long long o = ((long long)map * p.out_h + y) * p.out_w + (long long)xv * V;
    if constexpr (MODE == RZ_UP_RAW || MODE == RZ_UP_SIGMOID) {
      if constexpr (MODE == RZ_UP_SIGMOID) {
#pragma unroll
        for (int k = 0; k < V; ++k) val[k] = fast_sigmoid(val[k]);
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
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) out[k] = val[k] > p.thr_logit ? 1 : 0;
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

template <int MODE>
int launch_mode(const UpParams& p, int maps, cudaStream_t s) {
  dim3 grid((p.out_h + kBand - 1) / kBand, maps), block(kThreads);
  const size_t smem = (size_t)(kMaxGrid * kMaxGrid + kBand * kMaxGrid) * sizeof(float);
  const bool al16 = (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  int v = 1;
  if (MODE == RZ_UP_ARGMAX) v = (p.out_w % 4 == 0) ? 4 : (p.out_w % 2 == 0 ? 2 : 1);
  else if (al16 && p.out_w % 4 == 0) v = 4;
  else if (al16 && p.out_w % 2 == 0 && MODE != RZ_UP_MASK) v = 2;
  if (v == 4) {
    cudaFuncSetAttribute(upsample_kernel<MODE, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    upsample_kernel<MODE, 4><<<grid, block, smem, s>>>(p);
  } else if (v == 2) {
    cudaFuncSetAttribute(upsample_kernel<MODE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    upsample_kernel<MODE, 2><<<grid, block, smem, s>>>(p);
  } else {
    cudaFuncSetAttribute(upsample_kernel<MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    upsample_kernel<MODE, 1><<<grid, block, smem, s>>>(p);
  }
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
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
