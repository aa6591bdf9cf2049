// Library-wide state of libsgb200: thread-local error text, ABI version, launch counter.
#include "common.cuh"

namespace sgb {
static thread_local std::string t_last_error;
std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { t_last_error = msg; }

int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64) {
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;   // B200
  if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}
}  // namespace sgb

extern "C" const char* sgb_last_error(void) { return sgb::t_last_error.c_str(); }
extern "C" int sgb_abi_version(void) { return 1; }
extern "C" int64_t sgb_launch_count(void) { return (int64_t)sgb::g_launches.load(); }
