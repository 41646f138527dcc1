// Status plumbing, launch accounting and device queries for libpgt_b200.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace pgt {
static thread_local char g_last_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
}  // namespace pgt

extern "C" const char* pgt_strerror(int status) {
  switch (status) {
    case PGT_OK: return "ok";
    case PGT_ERR_INVALID: return "invalid argument (shape / alignment / null pointer)";
    case PGT_ERR_CUDA: return "CUDA runtime error (see pgt_last_cuda_error)";
    case PGT_ERR_UNSUPPORTED: return "configuration not covered by the sm_100a kernels";
    case PGT_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    default: return "unknown status";
  }
}
extern "C" const char* pgt_last_cuda_error(void) { return pgt::g_last_error; }
extern "C" int pgt_version(void) { return 100; }
extern "C" int64_t pgt_launch_count(void) { return pgt::g_launches.load(); }
extern "C" void pgt_reset_launch_count(void) { pgt::g_launches.store(0); }
