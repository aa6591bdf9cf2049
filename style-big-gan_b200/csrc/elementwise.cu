// Activation-sized passes of modulated_conv2d that are not convolutions (reference:
// train_parts/generators.py:80-87 `x * styles`, `fma(x, dcoefs, noise)`; fma.py:15-58 and its
// `_unbroadcast` reductions).  All HBM-bound; s / t / reduction outputs are accumulator-typed (fp32, or
// fp64 for SGB_F64 tensors).
//
//   sgb_scale_nc    y[n,c,h,w] = x[n,c,h,w] * s[n,c] (+ t[n,h,w])
//   sgb_mul_sum_hw  out[n,c]   = sum_{h,w} a * b           (gradient wrt styles / dcoefs)
//   sgb_sum_c       out[n,h,w] = sum_c a                   (gradient wrt noise)
#include "common.cuh"

namespace sgb {

struct EwParams {
  const void* a; const void* b; const void* s; const void* t; void* y;
  int n, c, h, w;
  int64_t as[4], bs[4], ys[4];
};

// ---- scale ---------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) scale_nc_generic_kernel(EwParams p, int c_fast) {
  typedef typename Acc<T>::type A;
  const int64_t total = (int64_t)p.n * p.c * p.h * p.w;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const A* sp = (const A*)p.s; const A* tp = (const A*)p.t;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int n, c, y, x; int64_t r = idx;
    if (c_fast) { c = (int)(r % p.c); r /= p.c; x = (int)(r % p.w); r /= p.w; y = (int)(r % p.h); n = (int)(r / p.h); }
    else        { x = (int)(r % p.w); r /= p.w; y = (int)(r % p.h); r /= p.h; c = (int)(r % p.c); n = (int)(r / p.c); }
    A v = to_acc<T>(((const T*)p.a)[n * p.as[0] + c * p.as[1] + y * p.as[2] + x * p.as[3]]) * sp[(int64_t)n * p.c + c];
    if (tp) v += tp[((int64_t)n * p.h + y) * p.w + x];
    ((T*)p.y)[n * p.ys[0] + c * p.ys[1] + y * p.ys[2] + x * p.ys[3]] = from_acc<T>(v);
  }
}

// dense NCHW planes, hw % VEC == 0, 16B-aligned: one (n,c) plane per blockIdx.y, 128-bit accesses
template <class T>
__global__ void __launch_bounds__(256) scale_nc_plane_kernel(EwParams p) {
  typedef typename Acc<T>::type A;
  constexpr int VEC = Vec16<T>::N;
  const int64_t plane = blockIdx.y + (int64_t)blockIdx.z * gridDim.y;
  if (plane >= (int64_t)p.n * p.c) return;
  const int64_t hw = (int64_t)p.h * p.w;
  const int64_t nv = hw / VEC;
  const A sc = ((const A*)p.s)[plane];
  const A* tp = p.t ? (const A*)p.t + (plane / p.c) * hw : nullptr;
  const uint4* src = (const uint4*)((const T*)p.a + plane * hw);
  uint4* dst = (uint4*)((T*)p.y + plane * hw);
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    Vec16<T> in, out; in.raw = ld_stream(src + v);
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      A r = to_acc<T>(in.v[j]) * sc;
      if (tp) r += tp[v * VEC + j];
      out.v[j] = from_acc<T>(r);
    }
    st_stream(dst + v, out.raw);
  }
}

// dense channels_last, c % VEC == 0: a thread owns VEC channels of one pixel
template <class T>
__global__ void __launch_bounds__(256) scale_nc_cl_kernel(EwParams p) {
  typedef typename Acc<T>::type A;
  constexpr int VEC = Vec16<T>::N;
  const int64_t cv = p.c / VEC;
  const int64_t hw = (int64_t)p.h * p.w;
  const int64_t nv = (int64_t)p.n * hw * cv;
  const A* sp = (const A*)p.s; const A* tp = (const A*)p.t;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = v / cv; const int c0 = (int)(v - pix * cv) * VEC; const int64_t n = pix / hw;
    Vec16<T> in, out; in.raw = ld_stream((const uint4*)p.a + v);
    const A add = tp ? tp[pix] : A(0);
#pragma unroll
    for (int j = 0; j < VEC; j++) out.v[j] = from_acc<T>(to_acc<T>(in.v[j]) * sp[n * p.c + c0 + j] + add);
    st_stream((uint4*)p.y + v, out.raw);
  }
}

// ---- sum over h,w of a*b -> [n,c] --------------------------------------------------------------------
// w-contiguous layouts: one warp per (n,c) plane
template <class T>
__global__ void __launch_bounds__(256) mul_sum_hw_plane_kernel(EwParams p) {
  typedef typename Acc<T>::type A;
  const int64_t plane = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (plane >= (int64_t)p.n * p.c) return;
  const int n = (int)(plane / p.c), c = (int)(plane % p.c);
  const T* ap = (const T*)p.a + n * p.as[0] + c * p.as[1];
  const T* bp = (const T*)p.b + n * p.bs[0] + c * p.bs[1];
  const int hw = p.h * p.w;
  A acc = A(0);
  for (int i = threadIdx.x & 31; i < hw; i += 32) {
    const int y = i / p.w, x = i - y * p.w;
    acc += to_acc<T>(ap[y * p.as[2] + x * p.as[3]]) * to_acc<T>(bp[y * p.bs[2] + x * p.bs[3]]);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) ((A*)p.y)[plane] = acc;
}

// c-contiguous layouts: grid (c tiles, row splits, n); block (32, 8); atomics into a zeroed output
template <class T>
__global__ void __launch_bounds__(256) mul_sum_hw_cl_kernel(EwParams p) {
  typedef typename Acc<T>::type A;
  __shared__ A part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int n = blockIdx.z;
  const int hw = p.h * p.w;
  A acc = A(0);
  if (c < p.c) {
    const T* ap = (const T*)p.a + n * p.as[0] + c * p.as[1];
    const T* bp = (const T*)p.b + n * p.bs[0] + c * p.bs[1];
    for (int i = blockIdx.y * 8 + threadIdx.y; i < hw; i += gridDim.y * 8) {
      const int y = i / p.w, x = i - y * p.w;
      acc += to_acc<T>(ap[y * p.as[2] + x * p.as[3]]) * to_acc<T>(bp[y * p.bs[2] + x * p.bs[3]]);
    }
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < p.c) {
    A t = A(0);
    for (int i = 0; i < 8; i++) t += part[i][threadIdx.x];
    atomicAdd((A*)p.y + (int64_t)n * p.c + c, t);
  }
}

// ---- sum over c -> [n,h,w] ---------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) sum_c_plane_kernel(EwParams p) {   // thread per pixel, loop over c
  typedef typename Acc<T>::type A;
  const int64_t total = (int64_t)p.n * p.h * p.w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx; const int x = (int)(r % p.w); r /= p.w; const int y = (int)(r % p.h); const int n = (int)(r / p.h);
    const T* ap = (const T*)p.a + n * p.as[0] + y * p.as[2] + x * p.as[3];
    A acc = A(0);
    for (int c = 0; c < p.c; c++) acc += to_acc<T>(ap[c * p.as[1]]);
    ((A*)p.y)[idx] = acc;
  }
}

template <class T>
__global__ void __launch_bounds__(256) sum_c_cl_kernel(EwParams p) {      // warp per pixel, lanes over c
  typedef typename Acc<T>::type A;
  const int64_t total = (int64_t)p.n * p.h * p.w;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t idx = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); idx < total; idx += wstride) {
    int64_t r = idx; const int x = (int)(r % p.w); r /= p.w; const int y = (int)(r % p.h); const int n = (int)(r / p.h);
    const T* ap = (const T*)p.a + n * p.as[0] + y * p.as[2] + x * p.as[3];
    A acc = A(0);
    for (int c = threadIdx.x & 31; c < p.c; c += 32) acc += to_acc<T>(ap[c * p.as[1]]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) ((A*)p.y)[idx] = acc;
  }
}

static bool dense_nchw(const int64_t s[4], int c, int h, int w) {
  return s[3] == 1 && s[2] == w && s[1] == (int64_t)h * w && s[0] == (int64_t)c * h * w;
}
static bool dense_nhwc(const int64_t s[4], int c, int h, int w) {
  return s[1] == 1 && s[3] == c && s[2] == (int64_t)w * c && s[0] == (int64_t)h * w * c;
}

}  // namespace sgb

using namespace sgb;

extern "C" int sgb_scale_nc(const void* x, const void* s, const void* t, void* y, int dtype,
                            int n, int c, int h, int w, const int64_t x_strides[4], const int64_t y_strides[4],
                            void* stream) {
  SGB_REQUIRE(n >= 0 && c >= 0 && h >= 0 && w >= 0, "negative size");
  const int64_t total = (int64_t)n * c * h * w;
  if (total == 0) return 0;
  SGB_REQUIRE(x && s && y, "x, s and y must not be NULL");
  EwParams p; p.a = x; p.b = nullptr; p.s = s; p.t = t; p.y = y; p.n = n; p.c = c; p.h = h; p.w = w;
  for (int i = 0; i < 4; i++) { p.as[i] = x_strides[i]; p.bs[i] = 0; p.ys[i] = y_strides[i]; }
  cudaStream_t st = (cudaStream_t)stream;
  SGB_DISPATCH_DTYPE(dtype, {
    constexpr int VEC = Vec16<T>::N;
    const bool al = aligned16(x) && aligned16(y);
    const int64_t hw = (int64_t)h * w;
    if (al && dense_nchw(p.as, c, h, w) && dense_nchw(p.ys, c, h, w) && hw % VEC == 0 && hw >= VEC) {
      int64_t planes = (int64_t)n * c, gy = planes, gz = 1;
      if (gy > 65535) { gz = ceil_div(gy, 65535); gy = 65535; }
      int64_t gx = ceil_div(hw / VEC, 256); if (gx > 64) gx = 64;
      scale_nc_plane_kernel<T><<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), 256, 0, st>>>(p);
    } else if (al && dense_nhwc(p.as, c, h, w) && dense_nhwc(p.ys, c, h, w) && c % VEC == 0) {
      int64_t blocks = ceil_div(total / VEC, 256); if (blocks > num_sms() * 16) blocks = num_sms() * 16;
      scale_nc_cl_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(p);
    } else {
      int64_t blocks = ceil_div(total, 256); if (blocks > num_sms() * 16) blocks = num_sms() * 16;
      scale_nc_generic_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(p, (y_strides[1] == 1 && c > 1) ? 1 : 0);
    }
    SGB_LAUNCH_CHECK();
  });
  return 0;
}

extern "C" int sgb_mul_sum_hw(const void* a, const void* b, void* out, int dtype, int n, int c, int h, int w,
                              const int64_t a_strides[4], const int64_t b_strides[4], void* stream) {
  SGB_REQUIRE(n >= 0 && c >= 0 && h >= 0 && w >= 0, "negative size");
  if ((int64_t)n * c == 0) return 0;
  SGB_REQUIRE(out, "out must not be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = dtype == SGB_F64 ? 8 : 4;
  if ((int64_t)h * w == 0) { cudaMemsetAsync(out, 0, esz * n * c, st); return 0; }
  SGB_REQUIRE(a && b, "a and b must not be NULL");
  EwParams p; p.a = a; p.b = b; p.s = nullptr; p.t = nullptr; p.y = out; p.n = n; p.c = c; p.h = h; p.w = w;
  for (int i = 0; i < 4; i++) { p.as[i] = a_strides[i]; p.bs[i] = b_strides[i]; p.ys[i] = 0; }
  SGB_DISPATCH_DTYPE(dtype, {
    if (a_strides[1] == 1 && c > 1) {
      cudaError_t e = cudaMemsetAsync(out, 0, esz * n * c, st);
      SGB_REQUIRE(e == cudaSuccess, "memset failed");
      SGB_REQUIRE(n <= 65535, "batch too large for the channels_last reduction");
      int gy = (int)ceil_div((int64_t)h * w, 8 * 32); if (gy < 1) gy = 1; if (gy > 256) gy = 256;
      mul_sum_hw_cl_kernel<T><<<dim3((unsigned)ceil_div(c, 32), (unsigned)gy, (unsigned)n), dim3(32, 8), 0, st>>>(p);
    } else {
      int64_t planes = (int64_t)n * c;
      mul_sum_hw_plane_kernel<T><<<(unsigned)ceil_div(planes, 8), 256, 0, st>>>(p);
    }
    SGB_LAUNCH_CHECK();
  });
  return 0;
}

extern "C" int sgb_sum_c(const void* a, void* out, int dtype, int n, int c, int h, int w, const int64_t a_strides[4],
                         void* stream) {
  SGB_REQUIRE(n >= 0 && c >= 0 && h >= 0 && w >= 0, "negative size");
  const int64_t total = (int64_t)n * h * w;
  if (total == 0) return 0;
  SGB_REQUIRE(a && out, "a and out must not be NULL");
  EwParams p; p.a = a; p.b = nullptr; p.s = nullptr; p.t = nullptr; p.y = out; p.n = n; p.c = c; p.h = h; p.w = w;
  for (int i = 0; i < 4; i++) { p.as[i] = a_strides[i]; p.bs[i] = 0; p.ys[i] = 0; }
  cudaStream_t st = (cudaStream_t)stream;
  SGB_DISPATCH_DTYPE(dtype, {
    if (a_strides[1] == 1 && c > 1) {
      int64_t blocks = ceil_div(total, 8); if (blocks > num_sms() * 16) blocks = num_sms() * 16;
      sum_c_cl_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(p);
    } else {
      int64_t blocks = ceil_div(total, 256); if (blocks > num_sms() * 16) blocks = num_sms() * 16;
      sum_c_plane_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(p);
    }
    SGB_LAUNCH_CHECK();
  });
  return 0;
}

// ---- nan_to_num over a list of tensors (the per-parameter gradient clean-up before the optimizer step) -------------
namespace sgb {
constexpr int NTN_MAX = 96;                       // tensors per launch (pointer table travels as a kernel parameter)
struct NtnParams { float* p[NTN_MAX]; int64_t n[NTN_MAX]; float nan, pinf, ninf; };

__global__ void __launch_bounds__(256) nan_to_num_multi_kernel(NtnParams q) {
  float* p = q.p[blockIdx.y];
  const int64_t n = q.n[blockIdx.y];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = p[i];
    const float r = (v != v) ? q.nan : (v == INFINITY ? q.pinf : (v == -INFINITY ? q.ninf : v));
    if (!(r == v)) p[i] = r;                      // gradients are almost always finite: read-only pass
  }
}
}  // namespace sgb

extern "C" int sgb_nan_to_num_multi(void* const* ptrs, const int64_t* numels, int count, float nan, float posinf, float neginf,
                                    void* stream) {
  SGB_REQUIRE(count >= 0, "negative count");
  SGB_REQUIRE(count == 0 || (ptrs && numels), "ptrs and numels must not be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < count; base += sgb::NTN_MAX) {
    sgb::NtnParams q; q.nan = nan; q.pinf = posinf; q.ninf = neginf;
    int m = 0;
    int64_t largest = 0;
    for (int i = base; i < count && m < sgb::NTN_MAX; i++) {
      SGB_REQUIRE(numels[i] >= 0, "negative numel");
      if (numels[i] == 0) continue;
      SGB_REQUIRE(ptrs[i] != nullptr, "NULL tensor with numel > 0");
      q.p[m] = (float*)ptrs[i]; q.n[m] = numels[i]; m++;
      if (numels[i] > largest) largest = numels[i];
    }
    if (m == 0) continue;
    int64_t bx = sgb::ceil_div(largest, (int64_t)256 * 8); if (bx < 1) bx = 1; if (bx > 64) bx = 64;
    sgb::nan_to_num_multi_kernel<<<dim3((unsigned)bx, (unsigned)m), 256, 0, st>>>(q);
    SGB_LAUNCH_CHECK();
  }
  return 0;
}

