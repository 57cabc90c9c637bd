// MUFU.EX2 throughput per SM on this GPU: N independent ex2 chains per thread, all SMs busy.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(float* out, int iters, float seed) {
  float x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.f) out[0] = s;
}
int main() {
  float* d; cudaMalloc(&d, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int warps : {4, 8, 16, 32}) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    k<8><<<sms, warps * 32>>>(d, 100, -1.0f);
    cudaEventRecord(a);
    k<8><<<sms, warps * 32>>>(d, iters, -1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)sms * warps * 32 * 8 * iters;
    printf("warps/SM %2d: %.1f G ex2/s total, %.2f ex2/clk/SM at %d MHz nominal\n", warps, ops / ms / 1e6,
           ops / ms / 1e3 / sms / (clk / 1e3) , clk / 1000);
  }
  return 0;
}
