// Diagnostics: a single-CTA tcgen05.mma probe.  The caller builds the exact shared-memory
// byte images of the A and B operands plus the 64-bit matrix descriptors (start address
// relative to the image) and the instruction descriptor; the kernel issues the MMAs and
// dumps TMEM.  tests/test_umma_probe.py uses it to pin, on real hardware, the layouts the
// production kernels assume (K-major / MN-major SWIZZLE_128B operands, M=64 and M=128
// accumulator lane mapping).
#include "rz_common.cuh"
#include "rz_umma.cuh"

namespace {

using namespace rz::umma;

__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}

__global__ void __launch_bounds__(128)
umma_probe_kernel(const uint4* __restrict__ a_image, int a_bytes, const uint4* __restrict__ b_image,
                  int b_bytes, unsigned long long a_desc, unsigned long long b_desc,
                  int a_step_bytes, int b_step_bytes, int k_steps, unsigned int idesc,
                  unsigned int d_tmem_offset, int ncols, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  // 1024-byte aligned operand images
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = base;
  uint8_t* b_smem = base + ((a_bytes + 1023) & ~1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(a_smem)[i] = a_image[i];
  for (int i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(b_smem)[i] = b_image[i];
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // sentinel fill so that untouched lanes/columns are recognisable in the dump
  {
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = __float_as_uint(-12345.0f);
    for (int c = 0; c < ncols; c += 8) tmem_st_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + c, s);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint64_t ad = a_desc + (uint64_t)(smem_u32(a_smem) >> 4);
    const uint64_t bd = b_desc + (uint64_t)(smem_u32(b_smem) >> 4);
    for (int k = 0; k < k_steps; ++k)
      mma_f16_ss(tmem_base + d_tmem_offset, desc_advance(ad, (uint32_t)(k * a_step_bytes)),
                 desc_advance(bd, (uint32_t)(k * b_step_bytes)), idesc, k > 0 ? 1u : 0u);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < ncols; c += 8) {
    uint32_t r[8];
    tmem_ld_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) out[(long long)tid * ncols + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" int rz_umma_probe(const void* a_image, int a_bytes, const void* b_image, int b_bytes,
                             unsigned long long a_desc, unsigned long long b_desc,
                             int a_step_bytes, int b_step_bytes, int k_steps, unsigned int idesc,
                             unsigned int d_tmem_offset, int ncols, float* out, void* stream) {
  if (!a_image || !b_image || !out) return RZ_ERR_INVALID;
  if (a_bytes <= 0 || b_bytes <= 0 || (a_bytes & 15) || (b_bytes & 15)) return RZ_ERR_INVALID;
  if (ncols <= 0 || (ncols & 7) || ncols > 512 || k_steps <= 0) return RZ_ERR_INVALID;
  const size_t smem = (size_t)((a_bytes + 1023) & ~1023) + ((b_bytes + 1023) & ~1023) + 1024;
  if (smem > 227 * 1024) return RZ_ERR_UNSUPPORTED;
  RZ_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a_image), a_bytes, static_cast<const uint4*>(b_image), b_bytes,
      a_desc, b_desc, a_step_bytes, b_step_bytes, k_steps, idesc, d_tmem_offset, ncols, out);
  RZ_LAUNCH_OK();
  rz_count_launch();
  return RZ_OK;
}
