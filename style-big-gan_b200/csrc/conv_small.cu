// 1x1 convolution with very few output channels (toRGB: C -> 3; the data gradient of fromRGB): a bandwidth kernel.
//
//   y[n, o, h, w] = sum_c x[n, c, h, w] * s[n, c] * W[o, c]            (conv2d;  W[c, o] for conv_transpose2d)
//
// The tensor-core kernels spend a 128-row MMA tile and a full pipeline on 3 useful output columns (measured: 0.8 % tensor
// activity, ~1.1 TB/s); here a thread owns one pixel, streams its channels_last row of x with 16-byte loads, and keeps
// the <= 8 accumulators in registers.  The (style-scaled) weights of the block's sample sit in shared memory and are
// read as broadcasts.  HBM traffic = x once + y once.
#include <cstdlib>
#include "common.cuh"

namespace sgb {

constexpr int SMALL_THREADS = 256;

struct SmallParams {
  sgb_conv_desc d;
  const void* x; const void* w; void* y;
  int hw;
};

// G threads share a pixel: thread g of the group loads the 16-byte chunks g, g + G, ... of the pixel's channel row, so the G
// lanes of a group read G * 16 contiguous bytes per step (G = 8: one 128-byte line; with one thread per pixel every load
// instruction of a warp touched 32 different rows, 16 bytes of each 32-byte sector: 0.39 of the copy rate in the step), and the
// CO partial sums are combined with group shuffles.  G = 1 is the original one-thread-per-pixel form (few channels).
template <class T, int CO, int G>
__global__ void __launch_bounds__(SMALL_THREADS) conv1x1_small_kernel(SmallParams p) {
  constexpr int TC = 16 / sizeof(T);
  constexpr int PB = SMALL_THREADS / G;                  // pixels per block
  extern __shared__ float wsm[];                         // [CO][ci], scaled by the sample's style
  const sgb_conv_desc& d = p.d;
  const int n = blockIdx.y;
  const T* w = (const T*)p.w;
  const float* sc = d.in_scale ? (const float*)d.in_scale + (int64_t)n * d.ci : nullptr;
  for (int i = threadIdx.x; i < CO * d.ci; i += SMALL_THREADS) {
    const int o = i / d.ci, c = i - o * d.ci;
    float v = 0.f;
    if (o < d.co) v = to_acc<T>(d.transposed ? w[(int64_t)c * d.co + o] : w[(int64_t)o * d.ci + c]);
    wsm[i] = sc ? v * sc[c] : v;
  }
  __syncthreads();
  const int g = threadIdx.x % G;
  const int pix = blockIdx.x * PB + threadIdx.x / G;
  if (pix >= p.hw) return;                               // the whole group leaves together (group-mask shuffles below)
  const int h = pix / d.in_w, wq = pix - h * d.in_w;
  const T* xp = (const T*)p.x + (int64_t)n * d.x_strides[0] + (int64_t)h * d.x_strides[2] + (int64_t)wq * d.x_strides[3];
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; o++) acc[o] = 0.f;
#pragma unroll 4
  for (int c = g * TC; c < d.ci; c += G * TC) {
    Vec16<T> v;
    v.raw = __ldg((const uint4*)(xp + c));
#pragma unroll
    for (int o = 0; o < CO; o++) {
#pragma unroll
      for (int q = 0; q < TC; q += 4) {
        const float4 w4 = *(const float4*)(wsm + o * d.ci + c + q);
        acc[o] += to_acc<T>(v.v[q]) * w4.x + to_acc<T>(v.v[q + 1]) * w4.y + to_acc<T>(v.v[q + 2]) * w4.z + to_acc<T>(v.v[q + 3]) * w4.w;
      }
    }
  }
  if (G > 1) {
    const unsigned gmask = (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
#pragma unroll
    for (int off = G >> 1; off > 0; off >>= 1)
#pragma unroll
      for (int o = 0; o < CO; o++) acc[o] += __shfl_xor_sync(gmask, acc[o], off);
    if (g != 0) return;
  }
  T* yp = (T*)p.y + (int64_t)n * d.y_strides[0] + (int64_t)h * d.y_strides[2] + (int64_t)wq * d.y_strides[3];
#pragma unroll
  for (int o = 0; o < CO; o++)
    if (o < d.co) yp[(int64_t)o * d.y_strides[1]] = from_acc<T>(acc[o]);
}

bool conv_small_eligible(const sgb_conv_desc* d) {
  if (d->dtype != SGB_F32 && d->dtype != SGB_F16 && d->dtype != SGB_BF16) return false;
  if (d->force_simt) return false;                       // A-B switches of the other kernels keep their meaning
  if (d->kh != 1 || d->kw != 1 || d->stride != 1 || d->pad_y != 0 || d->pad_x != 0 || d->groups != 1) return false;
  if (d->co > 8 || d->ci > 1024) return false;
  if (d->out_h != d->in_h || d->out_w != d->in_w) return false;
  if (d->out_scale || d->noise || d->bias || d->act != 0) return false;
  const int tc = d->dtype == SGB_F32 ? 4 : 8;
  if (d->ci % tc != 0 || d->x_strides[1] != 1) return false;
  if (d->x_strides[0] % tc || d->x_strides[2] % tc || d->x_strides[3] % tc) return false;
  if ((int64_t)d->in_h * d->in_w >= (int64_t)1 << 31 || d->n > 65535) return false;
  return true;
}

template <class T>
static int launch_small(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  SmallParams p; p.d = *d; p.x = x; p.w = w; p.y = y; p.hw = d->in_h * d->in_w;
  SGB_REQUIRE(aligned16(x), "x must be 16-byte aligned");
  constexpr int TC = 16 / sizeof(T);
  // SGB_SMALL_G=8: eight lanes per pixel.  Measured level with one thread per pixel (r2_run37.sh: 1.48 vs 1.38 ms per step for the
  // family, step 68.75 vs 68.68 ms) -- the half-used sectors of the one-thread form are served by L1 on the next load -- so the
  // simpler form stays the default
  static const int env_g = [] { const char* e = getenv("SGB_SMALL_G"); return e ? atoi(e) : 1; }();
  const bool grouped = env_g >= 8 && d->ci >= 8 * TC;            // at least one chunk per lane of a group of 8
  const int pb = grouped ? SMALL_THREADS / 8 : SMALL_THREADS;
  const dim3 grid((unsigned)ceil_div(p.hw, pb), (unsigned)d->n);
  if (grouped) {
    if (d->co <= 4) conv1x1_small_kernel<T, 4, 8><<<grid, SMALL_THREADS, sizeof(float) * 4 * d->ci, s>>>(p);
    else conv1x1_small_kernel<T, 8, 8><<<grid, SMALL_THREADS, sizeof(float) * 8 * d->ci, s>>>(p);
  } else {
    if (d->co <= 4) conv1x1_small_kernel<T, 4, 1><<<grid, SMALL_THREADS, sizeof(float) * 4 * d->ci, s>>>(p);
    else conv1x1_small_kernel<T, 8, 1><<<grid, SMALL_THREADS, sizeof(float) * 8 * d->ci, s>>>(p);
  }
  SGB_LAUNCH_CHECK();
  return 0;
}

int conv_forward_small(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  if (d->dtype == SGB_F16) return launch_small<__half>(d, x, w, y, s);
  if (d->dtype == SGB_BF16) return launch_small<__nv_bfloat16>(d, x, w, y, s);
  return launch_small<float>(d, x, w, y, s);
}

}  // namespace sgb
