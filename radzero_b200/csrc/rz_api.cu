// Library-level entry points of librz_b200.so: version, error text, launch counter.
#include <atomic>
#include <cstring>
#include <mutex>

#include "rz_common.cuh"

namespace {
std::atomic<long long> g_launches{0};
std::mutex g_err_mu;
char g_last_err[256] = "";
}  // namespace

void rz_note_cuda_error(cudaError_t e) {
  std::lock_guard<std::mutex> lock(g_err_mu);
  std::strncpy(g_last_err, cudaGetErrorString(e), sizeof(g_last_err) - 1);
}

void rz_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int rz_sm_count() {
  // cached PER DEVICE: a process may drive several GPUs
  static std::atomic<int> cached[64];
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64) {
    const int c = cached[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
  return n;
}

extern "C" int rz_version(void) { return 1000 * 0 + 1; }

extern "C" const char* rz_strerror(int code) {
  switch (code) {
    case RZ_OK: return "ok";
    case RZ_ERR_INVALID: return "invalid argument (shape, null pointer or size)";
    case RZ_ERR_CUDA: return "CUDA runtime error (see rz_last_cuda_error)";
    case RZ_ERR_UNSUPPORTED: return "request outside what this build implements";
    case RZ_ERR_ALIGNMENT: return "pointer or stride misaligned";
    default: return "unknown error";
  }
}

extern "C" const char* rz_last_cuda_error(void) { return g_last_err; }
extern "C" long long rz_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int rz_device_sm_count(void) { return rz_sm_count(); }
