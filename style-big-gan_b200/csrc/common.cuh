// Shared helpers for libsgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include "../../include/sgb200.h"

namespace sgb {

// SM count of the current device (148 on a B200), queried once per device ordinal
int num_sms();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel instantiation, device),
// from whichever thread launches first (forward runs on the caller's thread, backward on autograd's worker threads).
// Use inside the templated launcher of `kern` so that the static table is per instantiation.
#define SGB_SET_MAX_SMEM(kern, bytes)                                                                    \
  do {                                                                                                   \
    static std::atomic<int> done_[64];                                                                   \
    int dev_ = 0;                                                                                        \
    cudaGetDevice(&dev_);                                                                                \
    const int want_ = (int)(bytes);                                                                      \
    if (dev_ < 0 || dev_ >= 64 || done_[dev_].load(std::memory_order_acquire) < want_) {                 \
      cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, want_);   \
      if (e_ != cudaSuccess) (void)cudaGetLastError();      /* do not leave the error for the next launch check */ \
      SGB_REQUIRE(e_ == cudaSuccess, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e_));    \
      if (dev_ >= 0 && dev_ < 64) done_[dev_].store(want_, std::memory_order_release);                   \
    }                                                                                                    \
  } while (0)

// ---- error reporting (thread-local text, non-zero return codes; never abort) ----
void set_error(const std::string& msg);
extern std::atomic<long long> g_launches;

#define SGB_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      ::sgb::set_error(std::string(__func__) + ": " + (msg));    \
      return 1;                                                  \
    }                                                            \
  } while (0)

// call after every launch: counts it and turns launch-configuration errors into a return code
#define SGB_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    ::sgb::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
    cudaError_t e_ = cudaPeekAtLastError();                                                  \
    if (e_ != cudaSuccess) {                                                                 \
      ::sgb::set_error(std::string(__func__) + ": launch failed: " + cudaGetErrorString(e_)); \
      return 2;                                                                              \
    }                                                                                        \
  } while (0)

// ---- scalar type traits: storage type T, compute type (fp32, or fp64 for double) ----
template <class T> struct Acc { typedef float type; };
template <> struct Acc<double> { typedef double type; };

template <class T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v);
template <> __device__ __forceinline__ float to_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ double to_acc<double>(double v) { return v; }

template <class T> __device__ __forceinline__ T from_acc(typename Acc<T>::type v);
template <> __device__ __forceinline__ float from_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ double from_acc<double>(double v) { return v; }

// 16-byte vector of T
template <class T> struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union { uint4 raw; T v[16 / sizeof(T)]; };
};

// streaming (read-once) 128-bit global load / store: keep L1 for data that is reused
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// host-side dtype dispatch
#define SGB_DISPATCH_DTYPE(dtype, ...)                                               \
  switch (dtype) {                                                                   \
    case SGB_F32:  { typedef float T;          __VA_ARGS__; break; }                  \
    case SGB_F16:  { typedef __half T;         __VA_ARGS__; break; }                  \
    case SGB_BF16: { typedef __nv_bfloat16 T;  __VA_ARGS__; break; }                  \
    case SGB_F64:  { typedef double T;         __VA_ARGS__; break; }                  \
    default: ::sgb::set_error("unsupported dtype"); return 1;                        \
  }

}  // namespace sgb
