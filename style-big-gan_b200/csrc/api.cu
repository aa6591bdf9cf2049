// Library-wide state of libsgb200: thread-local error text, ABI version, launch counter.
#include "common.cuh"

namespace sgb {
static thread_local std::string t_last_error;
std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { t_last_error = msg; }
}  // namespace sgb

extern "C" const char* sgb_last_error(void) { return sgb::t_last_error.c_str(); }
extern "C" int sgb_abi_version(void) { return 1; }
extern "C" int64_t sgb_launch_count(void) { return (int64_t)sgb::g_launches.load(); }
