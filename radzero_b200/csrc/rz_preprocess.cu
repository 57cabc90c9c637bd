// K11: image preprocessing of the zero-shot evaluators on the GPU.
//
// Replaces, for a batch of same-sized raw images already in device memory,
//   collate_fn  exp/cxr_pt/inference/dataset.py:31-51  cv2.normalize(.., 0, 255, NORM_MINMAX, CV_8U)
//   image_processor(...)  exp/cxr_pt/model/processing.py:85-101  BlipImageProcessor at 518: convert to RGB,
//               PIL bicubic resize of the uint8 image, * 1/255, (x - mean) / std, channels first
// with results BIT-IDENTICAL to OpenCV + Pillow + transformers (oracle/preprocess.py restates and pins the
// arithmetic): the min-max stretch is one float FMA rounded half-to-even, the resize is Pillow's 8-bit
// fixed-point separable filter (22 fractional bits, uint8 intermediate, horizontal pass first), the
// normalisation is a 256-entry table per channel.
//
//   pp_coeff_kernel     filter taps of both passes, in double, no FMA contraction   (2 tiny CTAs)
//   pp_minmax_kernel    per-image min / max partials                               read raw once
//   pp_horizontal_kernel  raw row -> uint8 (FMA) in shared memory -> W_out outputs  read raw, write tmp
//   pp_vertical_kernel  tmp rows -> uint8 -> table -> 3 x H_out x W_out floats      write pixel_values
// HBM-bound on the output (3 x 518 x 518 x 4 B per image); the uint8 intermediate stays in L2.
#include <cfloat>

#include "rz_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPrecisionBits = 32 - 8 - 2;    // Pillow Resample.c
constexpr int kMinMaxSplit = 16;              // partial (min, max) pairs per image

struct Taps {        // one pass: for output index o, source window [xmin[o], xmin[o] + cnt[o]) and its weights
  int* xmin; int* cnt; int* w; int ksize;
  // ksize <= 12: the weights again as byte planes for dp4a, 9 words per output: planes 0 / 1 (bits 0-7 / 8-15,
  // unsigned) and 2 (bits 16-23, signed) of taps [4g, 4g + 4), g = 0..2:  w = p0 + 256 p1 + 65536 p2 exactly
  unsigned* planes;
};
constexpr int kPlaneTaps = 12;

__host__ __device__ inline int pil_ksize(int in_size, int out_size) {
  double scale = (double)in_size / (double)out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  double support = 2.0 * filterscale;
  return (int)ceil(support) * 2 + 1;
}

// Pillow's bicubic_filter with a = -0.5, every operation rounded separately (as x86-64 C without FMA)
__device__ double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) {
    double t = __dsub_rn(__dmul_rn(__dadd_rn(a, 2.0), x), __dadd_rn(a, 3.0));
    return __dadd_rn(__dmul_rn(__dmul_rn(t, x), x), 1.0);
  }
  if (x < 2.0) {
    double t = __dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0);
    t = __dsub_rn(__dmul_rn(t, x), 4.0);
    return __dmul_rn(t, a);
  }
  return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc (Resample.c) for the full-image box.  blockIdx.x = pass.
__global__ void pp_coeff_kernel(Taps th, int w_in, int w_out, Taps tv, int h_in, int h_out) {
  const Taps t = blockIdx.x == 0 ? th : tv;
  const int in_size = blockIdx.x == 0 ? w_in : h_in;
  const int out_size = blockIdx.x == 0 ? w_out : h_out;
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = 1.0 / filterscale;
  for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) {
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int lo = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (hi > in_size) hi = in_size;
    const int n = hi - lo;
    int* wk = t.w + (long long)xx * t.ksize;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss);
      ww = __dadd_rn(ww, bicubic_filter(arg));
    }
    for (int x = 0; x < t.ksize; ++x) {
      int q = 0;
      if (x < n) {
        const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss);
        double v = bicubic_filter(arg);
        if (ww != 0.0) v = __ddiv_rn(v, ww);
        q = (int)__dadd_rn(__dmul_rn(v, (double)(1 << kPrecisionBits)), v < 0.0 ? -0.5 : 0.5);
      }
      wk[x] = q;
    }
    t.xmin[xx] = lo;
    t.cnt[xx] = n;
    if (t.ksize <= kPlaneTaps) {
      for (int pl = 0; pl < 3; ++pl)
        for (int g = 0; g < 3; ++g) {
          unsigned word = 0;
          for (int i = 0; i < 4; ++i) {
            const int k = 4 * g + i;
            const int wv = k < t.ksize ? wk[k] : 0;
            word |= (unsigned)((wv >> (8 * pl)) & 0xff) << (8 * i);
          }
          t.planes[(long long)xx * 9 + pl * 3 + g] = word;
        }
    }
  }
}

template <typename T> __device__ __forceinline__ float as_float(T v) { return (float)v; }

// per-image min / max over all channels (cv::minMaxIdx), exact for every type: the reduction runs in the
// element's own domain (packed SIMD min / max for 8- and 16-bit pixels: four or two per instruction) and is
// widened to double only at the end.  16-byte loads over the aligned body of this CTA's slice.
template <typename T> struct MinMax;
template <> struct MinMax<uint8_t> {
  uint32_t lo = 0xffffffffu, hi = 0u;
  __device__ __forceinline__ void word(uint32_t w) { lo = __vminu4(lo, w); hi = __vmaxu4(hi, w); }
  __device__ __forceinline__ void one(uint8_t v) { word(v * 0x01010101u); }
  __device__ __forceinline__ void fin(double& a, double& b) const {
    uint32_t l = min(min(lo & 255u, (lo >> 8) & 255u), min((lo >> 16) & 255u, lo >> 24));
    uint32_t h = max(max(hi & 255u, (hi >> 8) & 255u), max((hi >> 16) & 255u, hi >> 24));
    a = (double)l; b = (double)h;
  }
};
template <> struct MinMax<uint16_t> {
  uint32_t lo = 0xffffffffu, hi = 0u;
  __device__ __forceinline__ void word(uint32_t w) { lo = __vminu2(lo, w); hi = __vmaxu2(hi, w); }
  __device__ __forceinline__ void one(uint16_t v) { word(v * 0x00010001u); }
  __device__ __forceinline__ void fin(double& a, double& b) const {
    a = (double)min(lo & 0xffffu, lo >> 16); b = (double)max(hi & 0xffffu, hi >> 16);
  }
};
template <> struct MinMax<int16_t> {
  uint32_t lo = 0x7fff7fffu, hi = 0x80008000u;
  __device__ __forceinline__ void word(uint32_t w) { lo = __vmins2(lo, w); hi = __vmaxs2(hi, w); }
  __device__ __forceinline__ void one(int16_t v) { word((uint32_t)(uint16_t)v * 0x00010001u); }
  __device__ __forceinline__ void fin(double& a, double& b) const {
    a = (double)min((int)(int16_t)(lo & 0xffffu), (int)(int16_t)(lo >> 16));
    b = (double)max((int)(int16_t)(hi & 0xffffu), (int)(int16_t)(hi >> 16));
  }
};
template <> struct MinMax<int32_t> {
  int lo = 0x7fffffff, hi = (int)0x80000000;
  __device__ __forceinline__ void word(uint32_t w) { lo = min(lo, (int)w); hi = max(hi, (int)w); }
  __device__ __forceinline__ void one(int32_t v) { word((uint32_t)v); }
  __device__ __forceinline__ void fin(double& a, double& b) const { a = (double)lo; b = (double)hi; }
};
template <> struct MinMax<float> {
  float lo = FLT_MAX, hi = -FLT_MAX;
  __device__ __forceinline__ void word(uint32_t w) { const float f = __uint_as_float(w); lo = fminf(lo, f); hi = fmaxf(hi, f); }
  __device__ __forceinline__ void one(float v) { lo = fminf(lo, v); hi = fmaxf(hi, v); }
  __device__ __forceinline__ void fin(double& a, double& b) const { a = (double)lo; b = (double)hi; }
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
pp_minmax_kernel(const T* __restrict__ raw, long long per_image, double* __restrict__ part) {
  constexpr int kVec = 16 / (int)sizeof(T);
  const int img = blockIdx.y, split = blockIdx.x;
  const T* src = raw + (long long)img * per_image;
  const long long per = (per_image + kMinMaxSplit - 1) / kMinMaxSplit;
  const long long i0 = (long long)split * per, i1 = min(per_image, i0 + per);
  MinMax<T> mm;
  if (i0 < i1) {
    // first element of the slice whose address is 16-byte aligned
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(src + i0);
    long long v0 = i0 + (long long)(((16 - (a0 & 15)) & 15) / sizeof(T));
    if (v0 > i1) v0 = i1;
    const long long nvec = (i1 - v0) / kVec;
    for (long long i = i0 + threadIdx.x; i < v0; i += kThreads) mm.one(src[i]);
    const uint4* vp = reinterpret_cast<const uint4*>(src + v0);
    for (long long j = threadIdx.x; j < nvec; j += 2 * kThreads) {
      const uint4 q = rz::ldg_stream_u4(vp + j);
      const bool two = j + kThreads < nvec;
      const uint4 r = two ? rz::ldg_stream_u4(vp + j + kThreads) : q;
      mm.word(q.x); mm.word(q.y); mm.word(q.z); mm.word(q.w);
      mm.word(r.x); mm.word(r.y); mm.word(r.z); mm.word(r.w);
    }
    for (long long i = v0 + nvec * kVec + threadIdx.x; i < i1; i += kThreads) mm.one(src[i]);
  }
  double lo, hi;
  mm.fin(lo, hi);            // a thread that saw no element holds the type's (max, min): neutral for the reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ double slo[kThreads / 32], shi[kThreads / 32];
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kThreads / 32; ++k) { lo = fmin(lo, slo[k]); hi = fmax(hi, shi[k]); }
    part[((long long)img * kMinMaxSplit + split) * 2 + 0] = lo;
    part[((long long)img * kMinMaxSplit + split) * 2 + 1] = hi;
  }
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ uint8_t stretch8(float v, float a, float b) {
  int q = __float2int_rn(__fmaf_rn(v, a, b));              // one float FMA, cvRound: round half to even
  return (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

// Horizontal pass.  A CTA stages the tap table (window start + weights of every output column) in shared
// memory once and then walks `rows_per_warp` rows per warp.  One WARP owns a row at a time (no block-level
// synchronisation in the row loop): it stretches the raw row to uint8 (saturate(rint(fma(src, a, b)))) into
// its private shared-memory row -- 16-byte loads where the row is aligned -- and each lane then accumulates
// output columns lane, lane + 32, ... with a branch-free tap loop (weights past the window are zero, the
// row is padded by ksize bytes so the taps stay in bounds).
template <typename T>
__device__ __forceinline__ void stretch_row(const T* __restrict__ src, int n_el, int W, int Wp, int C, float a,
                                            float b, uint8_t* srow, int lane) {   // Wp = channel stride in srow
  constexpr int kVec = 16 / (int)sizeof(T);
  const bool vec = C == 1 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (n_el % kVec) == 0;
  if (vec) {
    const uint4* vp = reinterpret_cast<const uint4*>(src);
    for (int j = lane; j < n_el / kVec; j += 32) {
      const uint4 q = rz::ldg_stream_u4(vp + j);
      const T* e = reinterpret_cast<const T*>(&q);
      uint8_t o[kVec];
#pragma unroll
      for (int u = 0; u < kVec; ++u) o[u] = stretch8(as_float(e[u]), a, b);
      if (kVec == 16) *reinterpret_cast<uint4*>(srow + 16 * j) = *reinterpret_cast<const uint4*>(o);
      else if (kVec == 8) *reinterpret_cast<uint2*>(srow + 8 * j) = *reinterpret_cast<const uint2*>(o);
      else *reinterpret_cast<uint32_t*>(srow + 4 * j) = *reinterpret_cast<const uint32_t*>(o);
    }
  } else if (C == 1) {
    for (int i = lane; i < n_el; i += 32) srow[i] = stretch8(as_float(src[i]), a, b);
  } else {
    for (int i = lane; i < n_el; i += 32) {
      const int x = i / C, c = i - x * C;
      srow[c * Wp + x] = stretch8(as_float(src[i]), a, b);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
pp_horizontal_kernel(const T* __restrict__ raw, int H, int W, int C, const double* __restrict__ part,
                     Taps t, int w_out, int pitch, uint8_t* __restrict__ tmp, int rows_per_warp, int taps_in_smem) {
  extern __shared__ __align__(16) uint8_t hs[];      // [tap table][warps][C][Wp]
  const int img = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = t.ksize;
  const int Wp = (W + ks + 15) & ~15;                // row + tap padding
  int* s_xmin = reinterpret_cast<int*>(hs);
  int* s_w = s_xmin + w_out;
  const size_t tab_bytes = taps_in_smem ? (((size_t)w_out * (1 + ks) * 4 + 15) & ~(size_t)15) : 0;
  if (taps_in_smem) {
    for (int i = threadIdx.x; i < w_out; i += kThreads) s_xmin[i] = __ldg(t.xmin + i);
    for (int i = threadIdx.x; i < w_out * ks; i += kThreads) s_w[i] = __ldg(t.w + i);
  }
  uint8_t* srow = hs + tab_bytes + (size_t)warp * C * Wp;
  for (int i = lane; i < C * Wp; i += 32) srow[i] = 0;                   // (the padding must be finite)
  __syncthreads();
  const int* xm = taps_in_smem ? s_xmin : t.xmin;
  const int* wt = taps_in_smem ? s_w : t.w;
  float a, b;
  {
    double lo = DBL_MAX, hi = -DBL_MAX;
    if (lane < kMinMaxSplit) {
      lo = part[((long long)img * kMinMaxSplit + lane) * 2 + 0];
      hi = part[((long long)img * kMinMaxSplit + lane) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    // cv::normalize, NORM_MINMAX: scale = (255 - 0) * (1 / (smax - smin)), shift = 0 - smin * scale
    const double scale = __dmul_rn(255.0, (hi - lo > DBL_EPSILON) ? __ddiv_rn(1.0, __dsub_rn(hi, lo)) : 0.0);
    const double shift = __dsub_rn(0.0, __dmul_rn(lo, scale));
    a = (float)scale;
    b = (float)shift;
  }
  const int y0 = (blockIdx.x * (kThreads / 32) + warp) * rows_per_warp;
  const int n_el = W * C;
  for (int y = y0; y < min(H, y0 + rows_per_warp); ++y) {
    const T* src = raw + ((long long)img * H + y) * n_el;
    __syncwarp();
    stretch_row<T>(src, n_el, W, Wp, C, a, b, srow, lane);
    __syncwarp();
    for (int c = 0; c < C; ++c) {
      const uint8_t* sc = srow + c * Wp;
      uint8_t* dst = tmp + (((long long)img * C + c) * H + y) * pitch;
      for (int xo = lane; xo < w_out; xo += 32) {
        const uint8_t* s = sc + xm[xo];
        const int* wk = wt + xo * ks;
        int acc = 1 << (kPrecisionBits - 1);
#pragma unroll 4
        for (int k = 0; k < ks; ++k) acc += (int)s[k] * wk[k];
        dst[xo] = clip8(acc);
      }
    }
  }
}

__device__ __forceinline__ int dp4a_uu(unsigned a, unsigned b, int c) {
  int d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp4a_us(unsigned a, unsigned b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// one output pixel from a 16-byte aligned window: realign by `sh` bits, nine dp4a (three byte planes of the
// weights), recombine modulo 2^32, round, shift, clip -- Pillow's 32-bit accumulator bit for bit
__device__ __forceinline__ uint32_t tap12(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, unsigned sh,
                                          const unsigned (&w0)[3], const unsigned (&w1)[3], const unsigned (&w2)[3]) {
  const uint32_t s0 = __funnelshift_r(q0, q1, sh), s1 = __funnelshift_r(q1, q2, sh), s2 = __funnelshift_r(q2, q3, sh);
  const int a0 = dp4a_uu(s2, w0[2], dp4a_uu(s1, w0[1], dp4a_uu(s0, w0[0], 1 << (kPrecisionBits - 1))));
  const int a1 = dp4a_uu(s2, w1[2], dp4a_uu(s1, w1[1], dp4a_uu(s0, w1[0], 0)));
  const int a2 = dp4a_us(s2, w2[2], dp4a_us(s1, w2[1], dp4a_us(s0, w2[0], 0)));
  return (uint32_t)clip8((int)((unsigned)a0 + ((unsigned)a1 << 8) + ((unsigned)a2 << 16)));
}

// Horizontal pass for windows of at most 12 taps (down-sampling by up to 2.75: the chest-X-ray case).  A CTA
// stretches a band of kBandH rows to uint8 in shared memory; thread t then owns output columns t, t + 256, ...
// for the whole band: its window start and its nine weight words (byte planes, see Taps) stay in registers,
// a pixel is four aligned 32-bit loads, three funnel shifts and nine dp4a -- 4 + 12 instead of 18 + 9 load and
// multiply instructions, with the same 32-bit accumulator bit for bit (the planes recombine modulo 2^32).
constexpr int kBandH = 16;
template <typename T>
__global__ void __launch_bounds__(kThreads)
pp_horizontal_dp4a_kernel(const T* __restrict__ raw, int H, int W, int C, const double* __restrict__ part,
                          Taps t, int w_out, int pitch, uint8_t* __restrict__ tmp, int Hp) {
  // Hp > 0: the intermediate is written TRANSPOSED, [image][channel][output column][Hp source rows] -- a
  // thread's 16 band rows of one column are one 16-byte store, and the vertical pass finds a column's taps in
  // consecutive bytes (pp_vertical_dp4a_kernel).  Hp = 0: row-major [image][channel][row][pitch].
  extern __shared__ __align__(16) uint8_t hs[];      // [C][kBandH][Wp]
  const int img = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = (W + 16 + 15) & ~15;                // a 16-byte window may start at the last pixel
  const int y0 = blockIdx.x * kBandH;
  const int rows = min(kBandH, H - y0);
  for (int i = threadIdx.x * 16; i < C * kBandH * Wp; i += kThreads * 16)
    *reinterpret_cast<uint4*>(hs + i) = make_uint4(0u, 0u, 0u, 0u);
  float a, b;
  {
    double lo = DBL_MAX, hi = -DBL_MAX;
    if (lane < kMinMaxSplit) {
      lo = part[((long long)img * kMinMaxSplit + lane) * 2 + 0];
      hi = part[((long long)img * kMinMaxSplit + lane) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    // cv::normalize, NORM_MINMAX: scale = (255 - 0) * (1 / (smax - smin)), shift = 0 - smin * scale
    const double scale = __dmul_rn(255.0, (hi - lo > DBL_EPSILON) ? __ddiv_rn(1.0, __dsub_rn(hi, lo)) : 0.0);
    const double shift = __dsub_rn(0.0, __dmul_rn(lo, scale));
    a = (float)scale;
    b = (float)shift;
  }
  // 8-bit sources: the stretch has 256 possible inputs -- one table per CTA (the same stretch8, bit for bit)
  // replaces a convert / FMA / round / clamp chain per pixel by a byte lookup
  __shared__ uint8_t s8[256];
  constexpr bool kByteSrc = sizeof(T) == 1;
  if (kByteSrc) s8[threadIdx.x & 255] = stretch8((float)(threadIdx.x & 255), a, b);
  __syncthreads();
  const int n_el = W * C;
  for (int r = warp; r < rows; r += kThreads / 32) {
    const T* src = raw + ((long long)img * H + y0 + r) * n_el;
    if (kByteSrc && C == 1 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (n_el & 15) == 0) {
      const uint4* vp = reinterpret_cast<const uint4*>(src);
      uint8_t* srow = hs + r * Wp;
      for (int j = lane; j < n_el / 16; j += 32) {
        const uint4 q = rz::ldg_stream_u4(vp + j);
        const uint32_t in[4] = {q.x, q.y, q.z, q.w};
        uint32_t o[4];
#pragma unroll
        for (int w = 0; w < 4; ++w)
          o[w] = (uint32_t)s8[in[w] & 255u] | ((uint32_t)s8[(in[w] >> 8) & 255u] << 8) |
                 ((uint32_t)s8[(in[w] >> 16) & 255u] << 16) | ((uint32_t)s8[in[w] >> 24] << 24);
        *reinterpret_cast<uint4*>(srow + 16 * j) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    } else {
      stretch_row<T>(src, n_el, W, kBandH * Wp, C, a, b, hs + r * Wp, lane);
    }
  }
  __syncthreads();
  for (int xo = threadIdx.x; xo < w_out; xo += kThreads) {
    const int xm = __ldg(t.xmin + xo);
    const unsigned* pw = t.planes + (long long)xo * 9;
    unsigned w0[3], w1[3], w2[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) { w0[g] = __ldg(pw + g); w1[g] = __ldg(pw + 3 + g); w2[g] = __ldg(pw + 6 + g); }
    const int al = xm & ~3;
    const unsigned sh = 8u * (unsigned)(xm & 3);
    for (int c = 0; c < C; ++c) {
      const uint8_t* sc = hs + (size_t)c * kBandH * Wp + al;
      if (Hp > 0) {
        uint32_t packed[kBandH / 4];
#pragma unroll
        for (int r = 0; r < kBandH; ++r) {               // rows past the image read the zeroed band: result 0
          const uint32_t* q = reinterpret_cast<const uint32_t*>(sc + r * Wp);
          const uint32_t q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
          const uint32_t v = tap12(q0, q1, q2, q3, sh, w0, w1, w2);
          packed[r >> 2] = (r & 3) == 0 ? v : (packed[r >> 2] | (v << (8 * (r & 3))));
        }
        uint8_t* dst = tmp + (((long long)img * C + c) * w_out + xo) * Hp + y0;
        *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      } else {
        uint8_t* dst = tmp + (((long long)img * C + c) * H + y0) * pitch + xo;
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
          const uint32_t* q = reinterpret_cast<const uint32_t*>(sc + r * Wp);
          dst[(long long)r * pitch] = (uint8_t)tap12(q[0], q[1], q[2], q[3], sh, w0, w1, w2);
        }
      }
    }
  }
}

struct NormParams { float mean[3], std[3]; double rescale; };

template <typename TOut> __device__ __forceinline__ TOut to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Vertical pass + rescale + normalise.  One CTA = `rows` consecutive output rows of one image.  The source
// rows those outputs tap (a contiguous band of the uint8 intermediate) are staged in shared memory once;
// thread x then walks the taps of its column and maps the uint8 result through the per-channel table.
template <typename TOut>
__global__ void __launch_bounds__(kThreads)
pp_vertical_kernel(const uint8_t* __restrict__ tmp, int H, int C, int pitch, Taps t, int h_out, int w_out,
                   NormParams np, TOut* __restrict__ out, int rows, int band_cap) {
  extern __shared__ uint8_t band[];          // [C][band_cap][pitch]
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += kThreads) {
    const int c = i >> 8, v = i & 255;
    // image_transforms.rescale: float64 product cast to float32; normalize: (x - mean) / std in float32
    const float r = (float)__dmul_rn((double)v, np.rescale);
    lut[c][v] = __fdiv_rn(__fsub_rn(r, np.mean[c]), np.std[c]);
  }
  const int img = blockIdx.y;
  const int y0 = blockIdx.x * rows, y1 = min(h_out, y0 + rows);
  const int b0 = __ldg(t.xmin + y0);
  const int b1 = __ldg(t.xmin + y1 - 1) + __ldg(t.cnt + y1 - 1);       // xmin is non-decreasing
  const int nb = b1 - b0;                                              // <= band_cap by construction
  const int vec = pitch / 16;
  for (int c = 0; c < C; ++c) {
    const uint4* src = reinterpret_cast<const uint4*>(tmp + (((long long)img * C + c) * H + b0) * pitch);
    uint4* dst = reinterpret_cast<uint4*>(band + (size_t)c * band_cap * pitch);
    for (int i = threadIdx.x; i < nb * vec; i += kThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  for (int yo = y0; yo < y1; ++yo) {
    const int lo = __ldg(t.xmin + yo) - b0, n = __ldg(t.cnt + yo);
    const int* wk = t.w + (long long)yo * t.ksize;
    for (int x = threadIdx.x; x < w_out; x += kThreads) {
      uint8_t q[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c < C) {
          const uint8_t* s = band + ((size_t)c * band_cap + lo) * pitch + x;
          int acc = 1 << (kPrecisionBits - 1);
          for (int k = 0; k < n; ++k) acc += (int)s[k * pitch] * __ldg(wk + k);
          q[c] = clip8(acc);
        } else {
          q[c] = q[0];                                     // convert_to_rgb replicates a grey plane
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        out[(((long long)img * 3 + c) * h_out + yo) * w_out + x] = to_out<TOut>(lut[c][q[c]]);
    }
  }
}

// Vertical pass over the TRANSPOSED intermediate (both windows <= 12 taps): a column's taps are consecutive
// bytes, so a pixel is the dp4a tap of the horizontal pass; then the per-channel normalisation table.
// (kernel below) CTA = (image, kRowBandV output rows, ALL columns): whole output rows are written by one CTA
// within microseconds, so L2 sees complete lines (column stripes per CTA left half-written sectors behind).
#ifndef RZ_PP_ROWLANES
#define RZ_PP_ROWLANES 4
#endif
constexpr int kRowLanesV = RZ_PP_ROWLANES;
constexpr int kRowBandV = 16;
__host__ __device__ inline int transposed_pitch(int H) { return (H + 64 + 15) & ~15; }   // kWinV bytes of slack
constexpr int kWinV = 64;                            // staged bytes per column: the band's source window, 16-byte aligned
#ifndef RZ_PP_VWARPS
#define RZ_PP_VWARPS 16
#endif
template <typename TOut, int CH>
__global__ void __launch_bounds__(32 * RZ_PP_VWARPS)
pp_vertical_dp4a_kernel(const uint8_t* __restrict__ tmpT, int Hp, Taps t, int h_out, int w_out,
                        NormParams np, TOut* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t vs[];      // [kRowBandV][12 words] tap table, then [CH][w_out][kWinS] windows
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += (int)blockDim.x) {
    const int c = i >> 8, v = i & 255;
    const float r = (float)__dmul_rn((double)v, np.rescale);
    lut[c][v] = __fdiv_rn(__fsub_rn(r, np.mean[c]), np.std[c]);
  }
  constexpr int kWinS = kWinV + 4;                   // column pitch in shared memory: 17 words (odd) -> no bank conflicts
  const int img = blockIdx.y, y0 = blockIdx.x * kRowBandV, y1 = min(h_out, y0 + kRowBandV);
  uint32_t* tab = reinterpret_cast<uint32_t*>(vs);   // per band row: ymin (relative to the window), 9 plane words, pad
  uint8_t* cols = vs + kRowBandV * 48;
  const int b0 = __ldg(t.xmin + y0) & ~15;           // the window starts on a 16-byte boundary of the column
  for (int i = threadIdx.x; i < CH * w_out * (kWinV / 16); i += (int)blockDim.x) {
    const int cc = i >> 2, v = i & 3;                // cc = channel * w_out + column
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(tmpT + ((long long)img * CH * w_out + cc) * Hp + b0) + v);
    uint32_t* d = reinterpret_cast<uint32_t*>(cols + (size_t)cc * kWinS + 16 * v);
    d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
  }
  for (int i = threadIdx.x; i < (y1 - y0) * 12; i += (int)blockDim.x) {
    const int r = i / 12, k = i - r * 12, yo = y0 + r;
    tab[i] = k == 0 ? (uint32_t)(__ldg(t.xmin + yo) - b0) : (k < 10 ? __ldg(t.planes + (long long)yo * 9 + k - 1) : 0u);
  }
  __syncthreads();
  // warp = (block of 128 columns, row lane); lane l owns columns 128 cb + 32 k + l, k = 0..3: neighbouring lanes
  // read neighbouring columns (pitch 17 words: conflict-free) and every store instruction writes 128
  // consecutive bytes of one output row
  const int nblk = (w_out + 127) >> 7;
  const int lane = threadIdx.x & 31;
  for (int wi = threadIdx.x >> 5; wi < nblk * kRowLanesV; wi += (int)blockDim.x >> 5) {
    const int rl = wi / nblk;                         // row lane: rows y0 + rl, y0 + rl + kRowLanesV, ...
    const int xb = (wi - rl * nblk) * 128 + lane;
    const uint8_t* cbase = cols + (size_t)xb * kWinS;
    for (int yo = y0 + rl; yo < y1; yo += kRowLanesV) {
      const uint32_t* trow = tab + (yo - y0) * 12;
      const uint4 ta = *reinterpret_cast<const uint4*>(trow);
      const uint4 tb = *reinterpret_cast<const uint4*>(trow + 4);
      const uint2 tc = *reinterpret_cast<const uint2*>(trow + 8);
      const unsigned w0[3] = {ta.y, ta.z, ta.w}, w1[3] = {tb.x, tb.y, tb.z}, w2[3] = {tb.w, tc.x, tc.y};
      const int ym = (int)ta.x;
      const uint8_t* win = cbase + (ym & ~3);
      const unsigned sh = 8u * (unsigned)(ym & 3);
      uint32_t px[CH][4];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          px[c][k] = 0u;
          if (xb + 32 * k < w_out) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(win + ((size_t)c * w_out + 32 * k) * kWinS);
            px[c][k] = tap12(q[0], q[1], q[2], q[3], sh, w0, w1, w2);
          }
        }
      }
      TOut* o = out + ((long long)img * 3 * h_out + yo) * w_out + xb;
#pragma unroll
      for (int c = 0; c < 3; ++c, o += (long long)h_out * w_out) {
        const uint32_t* p4 = px[CH == 3 ? c : 0];    // convert_to_rgb replicates a grey plane
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (xb + 32 * k < w_out) o[32 * k] = to_out<TOut>(lut[c][p4[k]]);
      }
    }
  }
}

struct Plan {
  int ks_h, ks_v, pitch;
  size_t off_part, off_taps, off_tmp, total;
};

Plan make_plan(int images, int H, int W, int C, int h_out, int w_out) {
  Plan p;
  p.ks_h = pil_ksize(W, w_out);
  p.ks_v = pil_ksize(H, h_out);
  p.pitch = (w_out + 15) / 16 * 16;
  size_t o = 0;
  p.off_part = o; o += (size_t)images * kMinMaxSplit * 2 * sizeof(double);
  p.off_taps = o; o += ((size_t)w_out * (2 + p.ks_h + 9) + (size_t)h_out * (2 + p.ks_v + 9)) * sizeof(int);
  o = (o + 255) / 256 * 256;
  const size_t row_major = (size_t)images * C * H * p.pitch;
  const size_t transposed = (size_t)images * C * w_out * (size_t)((H + 64 + 15) & ~15);
  p.off_tmp = o; o += row_major > transposed ? row_major : transposed;
  p.total = o;
  return p;
}

template <typename T>
int run(const T* raw, int images, int H, int W, int C, int h_out, int w_out, const NormParams& np,
        void* out, int out_dtype, uint8_t* ws, cudaStream_t s) {
  const Plan pl = make_plan(images, H, W, C, h_out, w_out);
  double* part = reinterpret_cast<double*>(ws + pl.off_part);
  int* ti = reinterpret_cast<int*>(ws + pl.off_taps);
  Taps th, tv;
  th.xmin = ti; ti += w_out; th.cnt = ti; ti += w_out; th.w = ti; ti += (size_t)w_out * pl.ks_h; th.ksize = pl.ks_h;
  th.planes = reinterpret_cast<unsigned*>(ti); ti += (size_t)w_out * 9;
  tv.xmin = ti; ti += h_out; tv.cnt = ti; ti += h_out; tv.w = ti; ti += (size_t)h_out * pl.ks_v; tv.ksize = pl.ks_v;
  tv.planes = reinterpret_cast<unsigned*>(ti);
  uint8_t* tmp = ws + pl.off_tmp;
  pp_coeff_kernel<<<2, kThreads, 0, s>>>(th, W, w_out, tv, H, h_out);
  RZ_LAUNCH_OK();
  pp_minmax_kernel<T><<<dim3(kMinMaxSplit, images), kThreads, 0, s>>>(raw, (long long)H * W * C, part);
  RZ_LAUNCH_OK();
  // horizontal: one warp per row at a time, `rows_h` rows per warp; shared memory = tap table + one stretched
  // (padded) row per warp
  const size_t smem_f = (size_t)C * kBandH * ((W + 16 + 15) & ~15);
  const int Hp = transposed_pitch(H);
  const size_t smem_vt = (size_t)C * w_out * (kWinV + 4) + (size_t)kRowBandV * 48;
  // the source window of kRowBandV output rows (+ 15 bytes of alignment slack + a 16-byte tap read) fits kWinV
  const bool win_ok = (long long)(kRowBandV - 1) * H / h_out + 2 + 15 + 16 <= kWinV;
  const bool fast_h = pl.ks_h <= kPlaneTaps && smem_f <= 200 * 1024 && !getenv("RZ_PP_GENERIC_H");
  // both windows short: transposed intermediate + dp4a vertical pass
  const bool fast_v = fast_h && pl.ks_v <= kPlaneTaps && win_ok && smem_vt <= 200 * 1024 && !getenv("RZ_PP_GENERIC_V");
  if (fast_h) {
    if (smem_f > 48 * 1024)
      RZ_CUDA_OK(cudaFuncSetAttribute(pp_horizontal_dp4a_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    pp_horizontal_dp4a_kernel<T><<<dim3((H + kBandH - 1) / kBandH, images), kThreads, smem_f, s>>>(
        raw, H, W, C, part, th, w_out, pl.pitch, tmp, fast_v ? Hp : 0);
    RZ_LAUNCH_OK();
  } else {
  const int rows_h = 8;
  const int warps = kThreads / 32;
  const size_t tab = (((size_t)w_out * (1 + pl.ks_h) * 4 + 15) & ~(size_t)15);
  const size_t rows_b = (size_t)warps * C * ((W + pl.ks_h + 15) & ~15);
  const int taps_in_smem = tab + rows_b <= 160 * 1024 ? 1 : 0;
  const size_t smem = rows_b + (taps_in_smem ? tab : 0);
  if (smem > 200 * 1024) return RZ_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    RZ_CUDA_OK(cudaFuncSetAttribute(pp_horizontal_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pp_horizontal_kernel<T><<<dim3((H + rows_h * warps - 1) / (rows_h * warps), images), kThreads, smem, s>>>(
      raw, H, W, C, part, th, w_out, pl.pitch, tmp, rows_h, taps_in_smem);
  RZ_LAUNCH_OK();
  }
  if (fast_v) {
    const dim3 gt((h_out + kRowBandV - 1) / kRowBandV, images);
    const int vwarps = kRowLanesV * ((w_out + 127) / 128);   // one warp per (128-column block, row lane)
    const int vthreads = 32 * (vwarps > RZ_PP_VWARPS ? RZ_PP_VWARPS : vwarps);
#define RZ_PP_LAUNCH_V(TO, CHN)                                                                                   \
  do {                                                                                                            \
    RZ_CUDA_OK(cudaFuncSetAttribute(pp_vertical_dp4a_kernel<TO, CHN>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem_vt));                                                               \
    pp_vertical_dp4a_kernel<TO, CHN><<<gt, vthreads, smem_vt, s>>>(tmp, Hp, tv, h_out, w_out, np, (TO*)out);      \
  } while (0)
    if (out_dtype == RZ_F32) { if (C == 1) RZ_PP_LAUNCH_V(float, 1); else RZ_PP_LAUNCH_V(float, 3); }
    else if (out_dtype == RZ_F16) { if (C == 1) RZ_PP_LAUNCH_V(__half, 1); else RZ_PP_LAUNCH_V(__half, 3); }
    else { if (C == 1) RZ_PP_LAUNCH_V(__nv_bfloat16, 1); else RZ_PP_LAUNCH_V(__nv_bfloat16, 3); }
#undef RZ_PP_LAUNCH_V
    RZ_LAUNCH_OK();
    rz_count_launch(4);
    return RZ_OK;
  }
  // vertical: `rows_v` output rows per CTA share a band of at most ceil(rows_v * H / h_out) + ksize source rows
  const int rows_v = 8;
  const int band_cap = (int)(((long long)rows_v * H + h_out - 1) / h_out) + pl.ks_v + 2;
  const size_t smem_v = (size_t)C * band_cap * pl.pitch;
  if (smem_v > 200 * 1024) return RZ_ERR_UNSUPPORTED;
  const dim3 gv((h_out + rows_v - 1) / rows_v, images);
  if (out_dtype == RZ_F32) {
    if (smem_v > 48 * 1024)
      RZ_CUDA_OK(cudaFuncSetAttribute(pp_vertical_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v));
    pp_vertical_kernel<float><<<gv, kThreads, smem_v, s>>>(tmp, H, C, pl.pitch, tv, h_out, w_out, np, (float*)out, rows_v, band_cap);
  } else if (out_dtype == RZ_F16) {
    if (smem_v > 48 * 1024)
      RZ_CUDA_OK(cudaFuncSetAttribute(pp_vertical_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v));
    pp_vertical_kernel<__half><<<gv, kThreads, smem_v, s>>>(tmp, H, C, pl.pitch, tv, h_out, w_out, np, (__half*)out, rows_v, band_cap);
  } else {
    if (smem_v > 48 * 1024)
      RZ_CUDA_OK(cudaFuncSetAttribute(pp_vertical_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v));
    pp_vertical_kernel<__nv_bfloat16><<<gv, kThreads, smem_v, s>>>(tmp, H, C, pl.pitch, tv, h_out, w_out, np,
                                                                   (__nv_bfloat16*)out, rows_v, band_cap);
  }
  RZ_LAUNCH_OK();
  rz_count_launch(4);
  return RZ_OK;
}

}  // namespace

extern "C" size_t rz_preprocess_workspace_bytes(int images, int height, int width, int channels, int out_h,
                                                int out_w) {
  if (images <= 0 || height <= 0 || width <= 0 || (channels != 1 && channels != 3) || out_h <= 0 || out_w <= 0)
    return 0;
  return make_plan(images, height, width, channels, out_h, out_w).total;
}

extern "C" int rz_preprocess_images(const void* raw, int raw_dtype, int images, int height, int width,
                                    int channels, int out_h, int out_w, const float* mean_host,
                                    const float* std_host, double rescale_factor, void* pixel_values,
                                    int out_dtype, void* workspace, size_t workspace_bytes, void* stream) {
  if (!raw || !pixel_values || !workspace || !mean_host || !std_host) return RZ_ERR_INVALID;
  if (images <= 0 || height <= 0 || width <= 0 || out_h <= 0 || out_w <= 0) return RZ_ERR_INVALID;
  if (channels != 1 && channels != 3) return RZ_ERR_INVALID;
  if (images > 65535) return RZ_ERR_INVALID;
  if (out_dtype != RZ_F32 && out_dtype != RZ_F16 && out_dtype != RZ_BF16) return RZ_ERR_INVALID;
  if (workspace_bytes < rz_preprocess_workspace_bytes(images, height, width, channels, out_h, out_w))
    return RZ_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return RZ_ERR_ALIGNMENT;
  NormParams np;
  for (int c = 0; c < 3; ++c) { np.mean[c] = mean_host[c]; np.std[c] = std_host[c]; }
  np.rescale = rescale_factor;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  switch (raw_dtype) {
    case RZ_IMG_U8: return run<uint8_t>((const uint8_t*)raw, images, height, width, channels, out_h, out_w, np, pixel_values, out_dtype, ws, s);
    case RZ_IMG_U16: return run<uint16_t>((const uint16_t*)raw, images, height, width, channels, out_h, out_w, np, pixel_values, out_dtype, ws, s);
    case RZ_IMG_I16: return run<int16_t>((const int16_t*)raw, images, height, width, channels, out_h, out_w, np, pixel_values, out_dtype, ws, s);
    case RZ_IMG_I32: return run<int32_t>((const int32_t*)raw, images, height, width, channels, out_h, out_w, np, pixel_values, out_dtype, ws, s);
    case RZ_IMG_F32: return run<float>((const float*)raw, images, height, width, channels, out_h, out_w, np, pixel_values, out_dtype, ws, s);
    default: return RZ_ERR_INVALID;
  }
}
