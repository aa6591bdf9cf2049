// upfirdn2d for sm_100a: zero-insert (up), pad/crop, FIR, decimate (down).
// Replaces upfirdn2d.cu:29-200 / upfirdn2d.cpp:16-94 of the reference.  The backward pass is the same
// operator with up<->down swapped and the filter flipped (upfirdn2d.py:246-264), so this one entry point
// covers every gradient order.
//
// HBM-bound stencil: algorithmic bytes = (N*C*Hin*Win + N*C*Hout*Wout) * sizeof(T).
//
// Kernels:
//   upfirdn2d_generic_kernel  any filter size / up / down / padding / layout (strides), one output per thread.
//   upfirdn2d_tile_kernel<..> NCHW planes, compile-time (up, down, taps): a CTA stages the input patch of a
//                             32x(8*ROWS) output tile in shared memory with coalesced loads, the filter lives in
//                             registers, the polyphase structure of up=2 is resolved at compile time (only
//                             the (taps/up)^2 live taps are visited), and each thread produces a 1x4 strip.
//   upfirdn2d_nhwc_kernel<..> channels_last: a thread owns 16 bytes of channels of one output pixel and walks
//                             the live taps; neighbouring pixels are served by L1/L2.
#include "common.cuh"

namespace sgb {

struct UpfirdnParams {
  const void* x; const float* f; void* y;
  int n, c, in_h, in_w, out_h, out_w;
  int64_t xs[4], ys[4];
  int fh, fw; int64_t f_sy, f_sx;
  int upx, upy, downx, downy, padx0, pady0, flip;
  float gain;
  int c_fast;   // 1: iterate channels fastest (channels_last output)
};

__device__ __forceinline__ int floor_div_i(int a, int b) {   // b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(UpfirdnParams p) {
  typedef typename Acc<T>::type A;
  const int64_t total = (int64_t)p.n * p.c * p.out_h * p.out_w;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int n, c, oy, ox;
    int64_t t = idx;
    if (p.c_fast) {
      c = (int)(t % p.c); t /= p.c; ox = (int)(t % p.out_w); t /= p.out_w; oy = (int)(t % p.out_h); n = (int)(t / p.out_h);
    } else {
      ox = (int)(t % p.out_w); t /= p.out_w; oy = (int)(t % p.out_h); t /= p.out_h; c = (int)(t % p.c); n = (int)(t / p.c);
    }
    // position of tap (0,0) of this output in the zero-inserted, un-padded image
    const int base_y = oy * p.downy - p.pady0;
    const int base_x = ox * p.downx - p.padx0;
    // first tap whose position is a multiple of `up` (only those carry data)
    int ty0 = ((-base_y) % p.upy + p.upy) % p.upy;
    int tx0 = ((-base_x) % p.upx + p.upx) % p.upx;
    const T* xp = (const T*)p.x + n * p.xs[0] + c * p.xs[1];
    A acc = A(0);
    for (int ty = ty0; ty < p.fh; ty += p.upy) {
      const int iy = (base_y + ty) / p.upy;        // exact division
      if (iy < 0 || iy >= p.in_h) continue;
      const int fy = p.flip ? ty : p.fh - 1 - ty;
      for (int tx = tx0; tx < p.fw; tx += p.upx) {
        const int ix = (base_x + tx) / p.upx;
        if (ix < 0 || ix >= p.in_w) continue;
        const int fx = p.flip ? tx : p.fw - 1 - tx;
        acc += to_acc<T>(xp[iy * p.xs[2] + ix * p.xs[3]]) * A(p.f[fy * p.f_sy + fx * p.f_sx]);
      }
    }
    acc *= A(p.gain);
    ((T*)p.y)[n * p.ys[0] + c * p.ys[1] + oy * p.ys[2] + ox * p.ys[3]] = from_acc<T>(acc);
  }
}

// ------------------------------------------------------------------------------------------------------
// NCHW tile kernel.  UP, DOWN in {1,2} (same in x and y), FT = taps per axis (<= 8), output tile
// TW x TH = 64 x 16 per CTA of 256 threads; each thread computes a 1 x 4 strip (4 consecutive ox).
// The (fixed) filter is copied into shared memory already flipped and pre-multiplied by gain.
template <class T, int UP, int DOWN, int FT>
__global__ void __launch_bounds__(256) upfirdn2d_tile_kernel(UpfirdnParams p) {
  typedef typename Acc<T>::type A;
  constexpr int TW = 64, TH = 16;
  // input extent needed by a TW x TH output tile: taps span FT positions in the upsampled domain
  constexpr int IN_W = ((TW - 1) * DOWN + FT - 1) / UP + 2;
  constexpr int IN_H = ((TH - 1) * DOWN + FT - 1) / UP + 2;
  constexpr int PITCH = IN_W | 1;   // odd pitch: no bank conflicts for the column walks
  __shared__ A sx[IN_H * PITCH];
  __shared__ A sf[FT * FT];

  const int tiles_x = (p.out_w + TW - 1) / TW;
  const int tile_x = (blockIdx.x % tiles_x) * TW;
  const int tile_y = (blockIdx.x / tiles_x) * TH;
  const int64_t plane = blockIdx.y + (int64_t)blockIdx.z * gridDim.y;   // n*C + c
  if (plane >= (int64_t)p.n * p.c) return;
  const int n = (int)(plane / p.c), c = (int)(plane % p.c);

  for (int i = threadIdx.x; i < FT * FT; i += 256) {
    int ty = i / FT, tx = i % FT;
    A v = A(0);
    if (ty < p.fh && tx < p.fw) {
      int fy = p.flip ? ty : p.fh - 1 - ty, fx = p.flip ? tx : p.fw - 1 - tx;
      v = A(p.f[fy * p.f_sy + fx * p.f_sx]) * A(p.gain);
    }
    sf[i] = v;
  }

  // first input sample touched by the tile (floor division: padding makes it negative)
  const int in_x0 = floor_div_i(tile_x * DOWN - p.padx0 + UP - 1, UP);   // ceil((tile_x*DOWN - padx0)/UP)
  const int in_y0 = floor_div_i(tile_y * DOWN - p.pady0 + UP - 1, UP);
  const T* xp = (const T*)p.x + n * p.xs[0] + c * p.xs[1];
  for (int i = threadIdx.x; i < IN_H * IN_W; i += 256) {
    int ry = i / IN_W, rx = i - ry * IN_W;
    int iy = in_y0 + ry, ix = in_x0 + rx;
    A v = A(0);
    if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v = to_acc<T>(xp[iy * p.xs[2] + ix * p.xs[3]]);
    sx[ry * PITCH + rx] = v;
  }
  __syncthreads();

  // thread -> 2 x 2 outputs: lanes walk x (coalesced stores), warps walk y
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int jy = 0; jy < TH / 8; jy++) {
    const int oy = tile_y + warp + 8 * jy;
    if (oy >= p.out_h) break;
    T* yp = (T*)p.y + n * p.ys[0] + c * p.ys[1] + oy * p.ys[2];
    const int base_y = oy * DOWN - p.pady0;
    const int ty0 = ((-base_y) % UP + UP) % UP;
#pragma unroll
    for (int jx = 0; jx < TW / 32; jx++) {
      const int ox = tile_x + lane + 32 * jx;
      if (ox >= p.out_w) break;
      const int base_x = ox * DOWN - p.padx0;
      const int tx0 = ((-base_x) % UP + UP) % UP;
      A acc = A(0);
#pragma unroll
      for (int a = 0; a < (FT + UP - 1) / UP; a++) {
        const int ty = ty0 + a * UP;
        if (ty >= FT) break;
        const int ry = (base_y + ty) / UP - in_y0;
#pragma unroll
        for (int b = 0; b < (FT + UP - 1) / UP; b++) {
          const int tx = tx0 + b * UP;
          if (tx >= FT) break;
          const int rx = (base_x + tx) / UP - in_x0;
          acc += sx[ry * PITCH + rx] * sf[ty * FT + tx];
        }
      }
      yp[ox] = from_acc<T>(acc);
    }
  }
}

}  // namespace sgb

using namespace sgb;

template <class T>
static int launch_upfirdn(const UpfirdnParams& p, cudaStream_t s) {
  const int64_t total = (int64_t)p.n * p.c * p.out_h * p.out_w;
  // tile kernel: NCHW-like output (w contiguous), square integer factors in {1,2}, small filters
  const bool sq = (p.upx == p.upy) && (p.downx == p.downy) && p.upx <= 2 && p.downx <= 2 && !(p.upx == 2 && p.downx == 2);
  const bool small = p.fh <= 4 && p.fw <= 4;
  const int64_t planes = (int64_t)p.n * p.c;
  if (sq && small && p.ys[3] == 1 && p.xs[3] == 1 && p.out_w >= 16 && p.out_h >= 4) {
    const int tiles = ((p.out_w + 63) / 64) * ((p.out_h + 15) / 16);
    int64_t gy = planes, gz = 1;
    if (gy > 65535) { gz = ceil_div(gy, 65535); gy = 65535; }
    dim3 grid((unsigned)tiles, (unsigned)gy, (unsigned)gz);
    if (p.upx == 1 && p.downx == 1) upfirdn2d_tile_kernel<T, 1, 1, 4><<<grid, 256, 0, s>>>(p);
    else if (p.upx == 2)            upfirdn2d_tile_kernel<T, 2, 1, 4><<<grid, 256, 0, s>>>(p);
    else                            upfirdn2d_tile_kernel<T, 1, 2, 4><<<grid, 256, 0, s>>>(p);
    SGB_LAUNCH_CHECK();
    return 0;
  }
  int64_t blocks = ceil_div(total, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  upfirdn2d_generic_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int sgb_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                             int n, int c, int in_h, int in_w, const int64_t x_strides[4],
                             int out_h, int out_w, const int64_t y_strides[4],
                             int fh, int fw, int64_t f_sy, int64_t f_sx,
                             int upx, int upy, int downx, int downy, int padx0, int pady0,
                             int flip, float gain, void* stream) {
  SGB_REQUIRE(x && f && y, "x, f and y must not be NULL");
  SGB_REQUIRE(n >= 0 && c >= 0 && in_h >= 1 && in_w >= 1, "bad input size");
  SGB_REQUIRE(out_h >= 1 && out_w >= 1, "output must be at least 1x1");
  SGB_REQUIRE(fh >= 1 && fw >= 1, "f must be at least 1x1");
  SGB_REQUIRE(upx >= 1 && upy >= 1, "upsampling factor must be at least 1");
  SGB_REQUIRE(downx >= 1 && downy >= 1, "downsampling factor must be at least 1");
  if ((int64_t)n * c == 0) return 0;
  UpfirdnParams p;
  p.x = x; p.f = f; p.y = y; p.n = n; p.c = c; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
  for (int i = 0; i < 4; i++) { p.xs[i] = x_strides[i]; p.ys[i] = y_strides[i]; }
  p.fh = fh; p.fw = fw; p.f_sy = f_sy; p.f_sx = f_sx;
  p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0;
  p.flip = flip ? 1 : 0; p.gain = gain;
  p.c_fast = (y_strides[1] == 1 && c > 1) ? 1 : 0;
  cudaStream_t s = (cudaStream_t)stream;
  SGB_DISPATCH_DTYPE(dtype, return launch_upfirdn<T>(p, s));
  return 0;
}
