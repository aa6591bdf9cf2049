// upfirdn2d for sm_100a: zero-insert (up), pad/crop, FIR, decimate (down).
// Replaces upfirdn2d.cu:29-200 / upfirdn2d.cpp:16-94 of the reference.  The backward pass is the same
// operator with up<->down swapped and the filter flipped (upfirdn2d.py:246-264), so this one entry point
// covers every gradient order.
//
// HBM-bound stencil: algorithmic bytes = (N*C*Hin*Win + N*C*Hout*Wout) * sizeof(T).
//
// Kernels:
//   upfirdn2d_generic_kernel  any filter size / up / down / padding / layout (strides), one output per thread.
//   upfirdn2d_tile_kernel<..> NCHW planes, compile-time (up, down) in {(1,1),(2,1),(1,2)}, filters up to 4x4: a CTA
//                             stages the input patch of a 64x16 output tile in shared memory with coalesced loads;
//                             only the live taps of the polyphase structure are visited; lanes walk x so stores coalesce.
//   upfirdn2d_cl_kernel<..>   channels_last, same (up, down) set: a thread owns 16 bytes of channels (128-bit loads and
//                             stores) and PX = 4 consecutive output pixels of one row; per live filter row it loads the
//                             few input pixels those 4 outputs share into registers once (7 / 4 / 10 vectors for
//                             (1,1) / (2,1) / (1,2)) and applies the taps with compile-time indexing.  CTAs cover
//                             compact 4-row tiles so the vertical reuse is served by L1.
#include <cstdlib>
#include <type_traits>
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

// ------------------------------------------------------------------------------------------------------
// channels_last kernel.  blockDim = (CVB, 256 / CVB): x = channel vector, y = (row in tile, x group).
// PADPAR = padx0 & 1 (only matters for UP == 2: fixes the polyphase pattern at compile time).
template <class T, int UP, int DOWN, int PADPAR>
__global__ void __launch_bounds__(256) upfirdn2d_cl_kernel(UpfirdnParams p, int cv_total, int cvb, int txg) {
  typedef typename Acc<T>::type A;
  constexpr int VEC = Vec16<T>::N;
  constexpr int FT = 4, PX = 4, TY = 4;
  constexpr int NIX = ((PX - 1) * DOWN + FT - 1) / UP + 1;
  __shared__ float sf[FT * FT];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < FT * FT) {
    const int ty = tid / FT, tx = tid % FT;
    float v = 0.f;
    if (ty < p.fh && tx < p.fw) {
      const int fy = p.flip ? ty : p.fh - 1 - ty, fx = p.flip ? tx : p.fw - 1 - tx;
      v = p.f[fy * p.f_sy + fx * p.f_sx] * p.gain;
    }
    sf[tid] = v;
  }
  __syncthreads();

  const int cchunks = (cv_total + cvb - 1) / cvb;
  const int n = blockIdx.z / cchunks;
  const int cv = (blockIdx.z - n * cchunks) * cvb + threadIdx.x;
  const int ry = threadIdx.y / txg, xg = threadIdx.y - ry * txg;
  const int oy = blockIdx.y * TY + ry;
  const int ox0 = (blockIdx.x * txg + xg) * PX;
  if (cv >= cv_total || oy >= p.out_h || ox0 >= p.out_w) return;
  const int c0 = cv * VEC;

  A acc[PX][VEC];
#pragma unroll
  for (int j = 0; j < PX; j++)
#pragma unroll
    for (int e = 0; e < VEC; e++) acc[j][e] = A(0);

  const T* xn = (const T*)p.x + (int64_t)n * p.xs[0] + c0;
  const int base_y = oy * DOWN - p.pady0;
  const int ty0 = ((-base_y) % UP + UP) % UP;
  // first input column any of the PX outputs touches: ceil((ox0*DOWN - padx0) / UP)
  const int bx = ox0 * DOWN - p.padx0;
  const int ix_first = (UP == 1) ? bx : ((bx + PADPAR) >> 1);      // UP == 2: bx has parity PADPAR (ox0 is even)
  const bool interior = ix_first >= 0 && ix_first + NIX <= p.in_w;
  for (int ty = ty0; ty < FT; ty += UP) {
    const int iy = (base_y + ty) / UP;
    if (iy < 0 || iy >= p.in_h) continue;
    const float4 fr = *(const float4*)(sf + ty * FT);
    const float frow[4] = {fr.x, fr.y, fr.z, fr.w};
    const T* xr = xn + (int64_t)iy * p.xs[2];
    Vec16<T> in[NIX];
    if (interior) {                                   // all NIX columns inside the image: no per-load bounds checks
      const T* xc = xr + (int64_t)ix_first * p.xs[3];
#pragma unroll
      for (int i = 0; i < NIX; i++) in[i].raw = __ldg((const uint4*)(xc + (int64_t)i * p.xs[3]));
    } else {
#pragma unroll
      for (int i = 0; i < NIX; i++) {
        const int ix = ix_first + i;
        in[i].raw = (ix >= 0 && ix < p.in_w) ? __ldg((const uint4*)(xr + (int64_t)ix * p.xs[3])) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int j = 0; j < PX; j++) {
#pragma unroll
      for (int tx = 0; tx < FT; tx++) {
        // position of this tap in the zero-inserted image relative to UP*ix_first
        const int rel = j * DOWN + tx - ((UP == 2) ? PADPAR : 0);
        if (UP == 2 && (rel & 1)) continue;          // zero-inserted sample
        const int i = rel / UP;
        if (i < 0 || i >= NIX) continue;
        const A fv = A(frow[tx]);
#pragma unroll
        for (int e = 0; e < VEC; e++) acc[j][e] += to_acc<T>(in[i].v[e]) * fv;
      }
    }
  }
  T* yp = (T*)p.y + (int64_t)n * p.ys[0] + (int64_t)oy * p.ys[2] + c0;
#pragma unroll
  for (int j = 0; j < PX; j++) {
    if (ox0 + j >= p.out_w) break;
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VEC; e++) o.v[e] = from_acc<T>(acc[j][e]);
    *(uint4*)(yp + (int64_t)(ox0 + j) * p.ys[3]) = o.raw;
  }
}


template <class T, int UP, int DOWN, int PADPAR>
__global__ void __launch_bounds__(256) upfirdn2d_cl2_kernel(UpfirdnParams p, int cv_total, int cvb, int txg) {
  typedef typename Acc<T>::type A;
  constexpr int VEC = Vec16<T>::N;
  constexpr int FT = 4, PX = 4, TY = 4;
  constexpr int NIX = ((PX - 1) * DOWN + FT - 1) / UP + 1;
  __shared__ float sf[FT * FT];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < FT * FT) {
    const int ty = tid / FT, tx = tid % FT;
    float v = 0.f;
    if (ty < p.fh && tx < p.fw) {
      const int fy = p.flip ? ty : p.fh - 1 - ty, fx = p.flip ? tx : p.fw - 1 - tx;
      v = p.f[fy * p.f_sy + fx * p.f_sx] * p.gain;
    }
    sf[tid] = v;
  }
  __syncthreads();

  const int cchunks = (cv_total + cvb - 1) / cvb;
  const int n = blockIdx.z / cchunks;
  const int cv = (blockIdx.z - n * cchunks) * cvb + threadIdx.x;
  const int ry = threadIdx.y / txg, xg = threadIdx.y - ry * txg;
  const int oy = blockIdx.y * TY + ry;
  const int ox0 = (blockIdx.x * txg + xg) * PX;
  if (cv >= cv_total || oy >= p.out_h || ox0 >= p.out_w) return;
  const int c0 = cv * VEC;

  A acc[PX][VEC];
#pragma unroll
  for (int j = 0; j < PX; j++)
#pragma unroll
    for (int e = 0; e < VEC; e++) acc[j][e] = A(0);

  const T* xn = (const T*)p.x + (int64_t)n * p.xs[0] + c0;
  const int base_y = oy * DOWN - p.pady0;
  const int ty0 = ((-base_y) % UP + UP) % UP;
  // first input column any of the PX outputs touches: ceil((ox0*DOWN - padx0) / UP)
  const int bx = ox0 * DOWN - p.padx0;
  const int ix_first = (UP == 1) ? bx : ((bx + PADPAR) >> 1);      // UP == 2: bx has parity PADPAR (ox0 is even)
  const bool interior = ix_first >= 0 && ix_first + NIX <= p.in_w;
  // all live filter rows are loaded before the first FMA: NR x NIX 128-bit loads in flight per thread (the one-row-at-a-time
  // form of upfirdn2d_cl_kernel stalls on every row: 45 % of its samples are FFMAs waiting for the row's loads)
  constexpr int NR = (FT + UP - 1) / UP;
  Vec16<T> in[NR][NIX];
  bool live[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const int ty = ty0 + r * UP;
    const int iy = (base_y + ty) / UP;
    live[r] = ty < FT && iy >= 0 && iy < p.in_h;
    const T* xr = xn + (int64_t)(live[r] ? iy : 0) * p.xs[2];
    if (interior) {
      const T* xc = xr + (int64_t)ix_first * p.xs[3];
#pragma unroll
      for (int i = 0; i < NIX; i++) in[r][i].raw = live[r] ? __ldg((const uint4*)(xc + (int64_t)i * p.xs[3])) : make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
      for (int i = 0; i < NIX; i++) {
        const int ix = ix_first + i;
        in[r][i].raw = (live[r] && ix >= 0 && ix < p.in_w) ? __ldg((const uint4*)(xr + (int64_t)ix * p.xs[3])) : make_uint4(0, 0, 0, 0);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const int ty = ty0 + r * UP;
    if (ty >= FT) continue;
    const float4 fr = *(const float4*)(sf + ty * FT);
    const float frow[4] = {fr.x, fr.y, fr.z, fr.w};
#pragma unroll
    for (int j = 0; j < PX; j++) {
#pragma unroll
      for (int tx = 0; tx < FT; tx++) {
        const int rel = j * DOWN + tx - ((UP == 2) ? PADPAR : 0);
        if (UP == 2 && (rel & 1)) continue;
        const int i = rel / UP;
        if (i < 0 || i >= NIX) continue;
        const A fv = A(frow[tx]);
#pragma unroll
        for (int e = 0; e < VEC; e++) acc[j][e] += to_acc<T>(in[r][i].v[e]) * fv;
      }
    }
  }
  T* yp = (T*)p.y + (int64_t)n * p.ys[0] + (int64_t)oy * p.ys[2] + c0;
#pragma unroll
  for (int j = 0; j < PX; j++) {
    if (ox0 + j >= p.out_w) break;
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VEC; e++) o.v[e] = from_acc<T>(acc[j][e]);
    *(uint4*)(yp + (int64_t)(ox0 + j) * p.ys[3]) = o.raw;
  }
}



// ------------------------------------------------------------------------------------------------------
// channels_last STRIP kernel (the default for channels_last tensors, filters up to 4x4, (up, down) in {(1,1),(2,1),(1,2)}).
// A thread owns 16 bytes of channels x PX consecutive output columns and walks DOWN a strip of `rows` output rows.  Every
// input row of the strip is loaded ONCE (NIX 128-bit loads, the next row prefetched while the current one is used) and
// scattered into a ring of SLOTS row accumulators; an output row is stored as soon as its last input row has been added.
// Compared with upfirdn2d_cl_kernel (a thread = 4 outputs of ONE row, 4 filter rows x NIX loads each, tiny CTAs) this cuts
// the load instructions per output 4x and amortises CTA start-up over rows x PX outputs per thread.
// The ring indices are compile-time: the input-row loop is unrolled by UNR and slot(k, ty) only depends on k mod UNR.
// Rank-1 filters (setup_filter's outer product, e.g. [1,3,3,1] x [1,3,3,1]) are detected in the kernel and applied
// separably: FT/UP horizontal + 1 vertical FMA per (input row, tap row) instead of FT/UP x taps (halves the FP32 work, which
// is what bounds 16-bit tensors: two elements per 4 bytes of traffic, fp32 accumulation).
//   PARX = padx0 & 1, PARY = pady0 & 1 (UP == 2 only: fix the polyphase pattern at compile time).
template <class T, int UP, int DOWN, int PARX, int PARY>
__global__ void __launch_bounds__(256) upfirdn2d_strip_kernel(UpfirdnParams p, int cv_total, int cvb, int xgs, int rows) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int FT = 4;
  constexpr int PX = (VEC == 4) ? 4 : 2;
  constexpr int NIX = ((PX - 1) * DOWN + FT - 1) / UP + 1;
  constexpr int SLOTS = (DOWN == 2) ? 2 : 4;
  constexpr int UNR = (UP == 2) ? 2 : 4;
  __shared__ float sf[FT * FT];
  __shared__ float sfx[FT], sfy[FT];
  __shared__ int s_sep;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < FT * FT) {
    const int ty = tid / FT, tx = tid % FT;
    float v = 0.f;
    if (ty < p.fh && tx < p.fw) {
      const int fy = p.flip ? ty : p.fh - 1 - ty, fx = p.flip ? tx : p.fw - 1 - tx;
      v = p.f[fy * p.f_sy + fx * p.f_sx] * p.gain;
    }
    sf[tid] = v;
  }
  __syncthreads();
  if (tid == 0) {      // rank-1 test: f[ty][tx] == fy[ty] * fx[tx] with fx = the row of the largest tap
    int best = 0;
    for (int i = 1; i < FT * FT; i++) if (fabsf(sf[i]) > fabsf(sf[best])) best = i;
    const int r0 = best / FT, c0 = best % FT;
    const float piv = sf[best];
    int sep = piv != 0.f;
    float fx_[FT], fy_[FT];
    for (int i = 0; i < FT; i++) { fx_[i] = sf[r0 * FT + i]; fy_[i] = sep ? sf[i * FT + c0] / piv : 0.f; }
    for (int i = 0; i < FT * FT && sep; i++)
      if (fabsf(sf[i] - fy_[i / FT] * fx_[i % FT]) > 1e-6f * fabsf(piv)) sep = 0;
    for (int i = 0; i < FT; i++) { sfx[i] = fx_[i]; sfy[i] = fy_[i]; }
    s_sep = sep;
  }
  __syncthreads();
  const bool sep = s_sep != 0;

  const int cchunks = (cv_total + cvb - 1) / cvb;
  const int n = blockIdx.z / cchunks;
  const int cv = (blockIdx.z - n * cchunks) * cvb + threadIdx.x;
  const int oy0 = blockIdx.y * rows;
  const int ox0 = (blockIdx.x * xgs + threadIdx.y) * PX;
  if (cv >= cv_total || ox0 >= p.out_w) return;
  const int c0 = cv * VEC;
  const int R = min(rows, p.out_h - oy0);

  const int base_y = oy0 * DOWN - p.pady0;
  const int iy0 = (UP == 1) ? base_y : ((base_y + PARY) >> 1);          // UP == 2: base_y has parity PARY (oy0 is even)
  const int K = (UP == 2) ? ((R + 2 - PARY) / 2 + 1) : ((DOWN == 2) ? 2 * R + 2 : R + 3);
  const int bx = ox0 * DOWN - p.padx0;
  const int ix_first = (UP == 1) ? bx : ((bx + PARX) >> 1);             // UP == 2: bx has parity PARX (ox0 is even)
  const bool interior = ix_first >= 0 && ix_first + NIX <= p.in_w;
  const T* xn = (const T*)p.x + (int64_t)n * p.xs[0] + c0;
  T* yn = (T*)p.y + (int64_t)n * p.ys[0] + c0;

  float fr[FT * FT], fxr[FT], fyr[FT];
#pragma unroll
  for (int i = 0; i < FT * FT; i++) fr[i] = sf[i];
#pragma unroll
  for (int i = 0; i < FT; i++) { fxr[i] = sfx[i]; fyr[i] = sfy[i]; }

  float acc[SLOTS][PX][VEC];
#pragma unroll
  for (int s = 0; s < SLOTS; s++)
#pragma unroll
    for (int j = 0; j < PX; j++)
#pragma unroll
      for (int e = 0; e < VEC; e++) acc[s][j][e] = 0.f;

  auto load_row = [&](int k, Vec16<T> (&in)[NIX]) -> bool {
    const int iy = iy0 + k;
    if (k >= K || iy < 0 || iy >= p.in_h) return false;
    const T* xr = xn + (int64_t)iy * p.xs[2];
    if (interior) {
      const T* xc = xr + (int64_t)ix_first * p.xs[3];
#pragma unroll
      for (int i = 0; i < NIX; i++) in[i].raw = __ldg((const uint4*)(xc + (int64_t)i * p.xs[3]));
    } else {
#pragma unroll
      for (int i = 0; i < NIX; i++) {
        const int ix = ix_first + i;
        in[i].raw = (ix >= 0 && ix < p.in_w) ? __ldg((const uint4*)(xr + (int64_t)ix * p.xs[3])) : make_uint4(0, 0, 0, 0);
      }
    }
    return true;
  };
  auto store_row = [&](int jc, float (&a)[PX][VEC]) {
    if (jc >= 0 && jc < R) {
      T* yp = yn + (int64_t)(oy0 + jc) * p.ys[2];
#pragma unroll
      for (int j = 0; j < PX; j++) {
        if (ox0 + j < p.out_w) {
          Vec16<T> o;
#pragma unroll
          for (int e = 0; e < VEC; e++) o.v[e] = from_acc<T>(a[j][e]);
          *(uint4*)(yp + (int64_t)(ox0 + j) * p.ys[3]) = o.raw;
        }
      }
#pragma unroll
      for (int j = 0; j < PX; j++)
#pragma unroll
        for (int e = 0; e < VEC; e++) a[j][e] = 0.f;
    }
  };

  Vec16<T> nxt[NIX];
  bool nxt_ok = load_row(0, nxt);
  for (int k0 = 0; k0 < K; k0 += UNR) {
#pragma unroll
    for (int kk = 0; kk < UNR; kk++) {
      const int k = k0 + kk;
      if (k < K) {
        Vec16<T> cur[NIX];
#pragma unroll
        for (int i = 0; i < NIX; i++) cur[i].raw = nxt[i].raw;
        const bool cur_ok = nxt_ok;
        nxt_ok = load_row(k + 1, nxt);
        if (cur_ok) {
          float h[PX][VEC];
          if (sep) {                  // horizontal pass once per input row
#pragma unroll
            for (int j = 0; j < PX; j++) {
#pragma unroll
              for (int e = 0; e < VEC; e++) h[j][e] = 0.f;
#pragma unroll
              for (int tx = 0; tx < FT; tx++) {
                const int rel = j * DOWN + tx - ((UP == 2) ? PARX : 0);
                if (UP == 2 && (rel & 1)) continue;
                const int i = rel / UP;
                if (rel < 0 || i >= NIX) continue;
#pragma unroll
                for (int e = 0; e < VEC; e++) h[j][e] += to_acc<T>(cur[i].v[e]) * fxr[tx];
              }
            }
          }
#pragma unroll
          for (int ty = 0; ty < FT; ty++) {
            // output row (relative to oy0) this (input row, tap row) pair feeds, and its ring slot (compile-time)
            int jrel; int slot;
            if (UP == 1 && DOWN == 1) { jrel = k - ty; slot = (kk - ty) & 3; }
            else if (DOWN == 2) { if ((kk - ty) & 1) continue; jrel = (k - ty) >> 1; slot = ((kk - ty + 8) >> 1) & 1; }
            else { jrel = 2 * k + PARY - ty; slot = (2 * kk + PARY - ty) & 3; }
            if (jrel < 0 || jrel >= R) continue;
            if (sep) {
              const float fv = fyr[ty];
#pragma unroll
              for (int j = 0; j < PX; j++)
#pragma unroll
                for (int e = 0; e < VEC; e++) acc[slot][j][e] += h[j][e] * fv;
            } else {
#pragma unroll
              for (int j = 0; j < PX; j++) {
#pragma unroll
                for (int tx = 0; tx < FT; tx++) {
                  const int rel = j * DOWN + tx - ((UP == 2) ? PARX : 0);
                  if (UP == 2 && (rel & 1)) continue;
                  const int i = rel / UP;
                  if (rel < 0 || i >= NIX) continue;
                  const float fv = fr[ty * FT + tx];
#pragma unroll
                  for (int e = 0; e < VEC; e++) acc[slot][j][e] += to_acc<T>(cur[i].v[e]) * fv;
                }
              }
            }
          }
        }
        // output rows whose last input row was k
        if (UP == 1 && DOWN == 1) store_row(k - 3, acc[(kk - 3) & 3]);
        else if (DOWN == 2) { if (kk & 1) store_row((k - 3) >> 1, acc[((kk - 3 + 8) >> 1) & 1]); }
        else { store_row(2 * k + PARY - 3, acc[(2 * kk + PARY - 3) & 3]); store_row(2 * k + PARY - 2, acc[(2 * kk + PARY - 2) & 3]); }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------
// Separable FIR, no resampling (the two big forms of the hot path: after the transposed convolution of an up-sampling layer
// and in front of the strided convolution of a down-sampling layer), channels_last.  The host has factored the 4 x 4 filter
// into fy (x) fx (setup_filter builds it as an outer product; sgb200/ops/upfirdn2d.py checks and caches that per filter tensor).
// A thread owns 16 bytes of channels x PX output columns and walks down a strip of output rows: every input row is loaded
// and converted ONCE, filtered horizontally (4 FMAs per element) into a ring of the last four rows, and an output row is the
// vertical 4-tap combination of the ring (4 FMAs per element): 8 FMAs and one conversion per output element instead of 16 + 7
// in upfirdn2d_cl_kernel.  That matters for 16-bit tensors, where the stencil is bound by instruction issue, not by HBM
// (ncu, profiles/README.md: fp16 [4,32,1025,1025] at 25 % of DRAM throughput with 58 % of the issue slots busy).
struct FirSepParams {
  const void* x; void* y;
  int n, c, in_h, in_w, out_h, out_w;
  int64_t xs[4], ys[4];
  int padx0, pady0;
  float fx[4], fy[4];          // taps in application order: y[oy][ox] = sum fy[ty] fx[tx] x[oy - pady0 + ty][ox - padx0 + tx]
  int cv_total, cvb, xgs, rows;
};

template <class T>
__global__ void __launch_bounds__(256, 2) fir_sep_strip_kernel(FirSepParams p) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int PX = (VEC == 4) ? 4 : 2;
  constexpr int NIX = PX + 3;
  const int cchunks = (p.cv_total + p.cvb - 1) / p.cvb;
  const int n = blockIdx.z / cchunks;
  const int cv = (blockIdx.z - n * cchunks) * p.cvb + threadIdx.x;
  const int oy0 = blockIdx.y * p.rows;
  const int ox0 = (blockIdx.x * p.xgs + threadIdx.y) * PX;
  if (cv >= p.cv_total || ox0 >= p.out_w) return;
  const int R = min(p.rows, p.out_h - oy0);
  const int K = R + 3;
  const int iy0 = oy0 - p.pady0;
  const int ix_first = ox0 - p.padx0;
  const bool interior = ix_first >= 0 && ix_first + NIX <= p.in_w;
  const T* xn = (const T*)p.x + (int64_t)n * p.xs[0] + cv * VEC;
  T* yn = (T*)p.y + (int64_t)n * p.ys[0] + cv * VEC;
  const float fx0 = p.fx[0], fx1 = p.fx[1], fx2 = p.fx[2], fx3 = p.fx[3];
  const float fy0 = p.fy[0], fy1 = p.fy[1], fy2 = p.fy[2], fy3 = p.fy[3];

  float h[4][PX][VEC];          // ring: horizontally filtered input rows k - 3 ... k
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int j = 0; j < PX; j++)
#pragma unroll
      for (int e = 0; e < VEC; e++) h[r][j][e] = 0.f;

  for (int k0 = 0; k0 < K; k0 += 4) {
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      const int k = k0 + kk;
      if (k < K) {
        const int iy = iy0 + k;
        float v[NIX][VEC];
        if (iy >= 0 && iy < p.in_h) {
          const T* xr = xn + (int64_t)iy * p.xs[2];
          Vec16<T> in[NIX];
          if (interior) {
            const T* xc = xr + (int64_t)ix_first * p.xs[3];
#pragma unroll
            for (int i = 0; i < NIX; i++) in[i].raw = __ldg((const uint4*)(xc + (int64_t)i * p.xs[3]));
          } else {
#pragma unroll
            for (int i = 0; i < NIX; i++) {
              const int ix = ix_first + i;
              in[i].raw = (ix >= 0 && ix < p.in_w) ? __ldg((const uint4*)(xr + (int64_t)ix * p.xs[3])) : make_uint4(0, 0, 0, 0);
            }
          }
#pragma unroll
          for (int i = 0; i < NIX; i++)
#pragma unroll
            for (int e = 0; e < VEC; e++) v[i][e] = to_acc<T>(in[i].v[e]);
#pragma unroll
          for (int j = 0; j < PX; j++)
#pragma unroll
            for (int e = 0; e < VEC; e++)
              h[kk][j][e] = v[j][e] * fx0 + v[j + 1][e] * fx1 + v[j + 2][e] * fx2 + v[j + 3][e] * fx3;
        } else {
#pragma unroll
          for (int j = 0; j < PX; j++)
#pragma unroll
            for (int e = 0; e < VEC; e++) h[kk][j][e] = 0.f;
        }
        const int jr = k - 3;                 // output row completed by input row k: rows k-3 .. k sit in slots (kk+1..kk+4) & 3
        if (jr >= 0 && jr < R) {
          T* yp = yn + (int64_t)(oy0 + jr) * p.ys[2];
#pragma unroll
          for (int j = 0; j < PX; j++) {
            if (ox0 + j < p.out_w) {
              Vec16<T> o;
#pragma unroll
              for (int e = 0; e < VEC; e++)
                o.v[e] = from_acc<T>(h[(kk + 1) & 3][j][e] * fy0 + h[(kk + 2) & 3][j][e] * fy1 + h[(kk + 3) & 3][j][e] * fy2 + h[kk][j][e] * fy3);
              *(uint4*)(yp + (int64_t)(ox0 + j) * p.ys[3]) = o.raw;
            }
          }
        }
      }
    }
  }
}

// 16-bit tensors: the same strip walk with the input rows PREFETCHED through shared memory.  The register version above keeps
// 5 x 16 bytes per thread in flight at 2 CTAs / SM (120 registers): ncu shows fp16 [4,32,1025,1025] at 40 % of DRAM throughput
// with 45 % of the issue slots busy and 22 % of the warp slots occupied -- latency-bound.  Here every thread owns NIX private
// 16-byte slots in each of four stages and issues the cp.async of row k + 3 before it filters row k (zero fill outside the
// image, so there are no edge branches); the slots are thread-private, so cp.async.wait_group is the only synchronisation.
template <class T, int PX>
__global__ void __launch_bounds__(256, 2) fir_sep_strip_pf_kernel(FirSepParams p) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int NIX = PX + 3;
  constexpr int ST = 4;                     // stages = the unroll of the row loop, so the stage index is static
  extern __shared__ __align__(16) uint4 fir_sm[];          // [ST][NIX][256]
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int cchunks = (p.cv_total + p.cvb - 1) / p.cvb;
  const int n = blockIdx.z / cchunks;
  const int cv = (blockIdx.z - n * cchunks) * p.cvb + threadIdx.x;
  const int oy0 = blockIdx.y * p.rows;
  const int ox0 = (blockIdx.x * p.xgs + threadIdx.y) * PX;
  if (cv >= p.cv_total || ox0 >= p.out_w) return;
  const int R = min(p.rows, p.out_h - oy0);
  const int K = R + 3;
  const int iy0 = oy0 - p.pady0;
  const int ix_first = ox0 - p.padx0;
  const T* xn = (const T*)p.x + (int64_t)n * p.xs[0] + cv * VEC;
  T* yn = (T*)p.y + (int64_t)n * p.ys[0] + cv * VEC;
  const float fx0 = p.fx[0], fx1 = p.fx[1], fx2 = p.fx[2], fx3 = p.fx[3];
  const float fy0 = p.fy[0], fy1 = p.fy[1], fy2 = p.fy[2], fy3 = p.fy[3];
  const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(fir_sm) + tid * 16;
  bool col_ok[NIX];
#pragma unroll
  for (int i = 0; i < NIX; i++) col_ok[i] = ix_first + i >= 0 && ix_first + i < p.in_w;

  auto issue = [&](int k, int stage) {
    const int iy = iy0 + k;
    const bool row_ok = k < K && iy >= 0 && iy < p.in_h;
    const T* xr = xn + (int64_t)iy * p.xs[2] + (int64_t)ix_first * p.xs[3];
#pragma unroll
    for (int i = 0; i < NIX; i++) {
      const bool ok = row_ok && col_ok[i];
      const void* src = ok ? (const void*)(xr + (int64_t)i * p.xs[3]) : (const void*)p.x;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(sm0 + (stage * NIX + i) * (256 * 16)), "l"(src), "r"(ok ? 16u : 0u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float h[4][PX][VEC];          // ring: horizontally filtered input rows k - 3 ... k
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int j = 0; j < PX; j++)
#pragma unroll
      for (int e = 0; e < VEC; e++) h[r][j][e] = 0.f;

  issue(0, 0); issue(1, 1); issue(2, 2);
  for (int k0 = 0; k0 < K; k0 += 4) {
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      const int k = k0 + kk;
      issue(k + 3, (kk + 3) & 3);                       // always commits a group (empty beyond the strip): the wait count stays 3
      asm volatile("cp.async.wait_group 3;" ::: "memory");
      if (k < K) {
        float v[NIX][VEC];
#pragma unroll
        for (int i = 0; i < NIX; i++) {
          Vec16<T> in;
          in.raw = fir_sm[(kk * NIX + i) * 256 + tid];
#pragma unroll
          for (int e = 0; e < VEC; e++) v[i][e] = to_acc<T>(in.v[e]);
        }
#pragma unroll
        for (int j = 0; j < PX; j++)
#pragma unroll
          for (int e = 0; e < VEC; e++)
            h[kk][j][e] = v[j][e] * fx0 + v[j + 1][e] * fx1 + v[j + 2][e] * fx2 + v[j + 3][e] * fx3;
        const int jr = k - 3;                 // output row completed by input row k
        if (jr >= 0 && jr < R) {
          T* yp = yn + (int64_t)(oy0 + jr) * p.ys[2];
#pragma unroll
          for (int j = 0; j < PX; j++) {
            if (ox0 + j < p.out_w) {
              Vec16<T> o;
#pragma unroll
              for (int e = 0; e < VEC; e++)
                o.v[e] = from_acc<T>(h[(kk + 1) & 3][j][e] * fy0 + h[(kk + 2) & 3][j][e] * fy1 + h[(kk + 3) & 3][j][e] * fy2 + h[kk][j][e] * fy3);
              *(uint4*)(yp + (int64_t)(ox0 + j) * p.ys[3]) = o.raw;
            }
          }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace sgb

using namespace sgb;

template <class T>
static int launch_strip(const UpfirdnParams& p, cudaStream_t s) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int PX = (VEC == 4) ? 4 : 2;
  const int cv_total = p.c / VEC;
  int cvb = 1; while (cvb < cv_total && cvb < 32) cvb <<= 1;
  const int xgs = 256 / cvb;                            // x groups (of PX outputs) per CTA
  const int64_t gx = ceil_div(p.out_w, xgs * PX), gzz = (int64_t)p.n * ceil_div(cv_total, cvb);
  int rows = p.out_h >= 128 ? 32 : (p.out_h >= 32 ? 16 : 8);
  while (rows > 8 && gx * ceil_div(p.out_h, rows) * gzz < 2 * (int64_t)num_sms()) rows >>= 1;
  const int64_t gy = ceil_div(p.out_h, rows);
  if (gy > 65535 || gzz > 65535) return 1;
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gzz), block(cvb, xgs);
  const int px = p.padx0 & 1, py = p.pady0 & 1;
  if (p.upx == 1 && p.downx == 1)           upfirdn2d_strip_kernel<T, 1, 1, 0, 0><<<grid, block, 0, s>>>(p, cv_total, cvb, xgs, rows);
  else if (p.downx == 2)                    upfirdn2d_strip_kernel<T, 1, 2, 0, 0><<<grid, block, 0, s>>>(p, cv_total, cvb, xgs, rows);
  else if (px == 0 && py == 0)              upfirdn2d_strip_kernel<T, 2, 1, 0, 0><<<grid, block, 0, s>>>(p, cv_total, cvb, xgs, rows);
  else if (px == 1 && py == 1)              upfirdn2d_strip_kernel<T, 2, 1, 1, 1><<<grid, block, 0, s>>>(p, cv_total, cvb, xgs, rows);
  else if (px == 0)                         upfirdn2d_strip_kernel<T, 2, 1, 0, 1><<<grid, block, 0, s>>>(p, cv_total, cvb, xgs, rows);
  else                                      upfirdn2d_strip_kernel<T, 2, 1, 1, 0><<<grid, block, 0, s>>>(p, cv_total, cvb, xgs, rows);
  SGB_LAUNCH_CHECK();
  return 0;
}
template <> int launch_strip<double>(const UpfirdnParams&, cudaStream_t) { return 1; }

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
  // channels_last kernel: channel-contiguous tensors whose channel count allows 16-byte vectors
  constexpr int VEC = Vec16<T>::N;
  const bool cl_ok = sq && small && p.xs[1] == 1 && p.ys[1] == 1 && p.c % VEC == 0 && aligned16(p.x) && aligned16(p.y) &&
                     p.xs[0] % VEC == 0 && p.xs[2] % VEC == 0 && p.xs[3] % VEC == 0 &&
                     p.ys[0] % VEC == 0 && p.ys[2] % VEC == 0 && p.ys[3] % VEC == 0 && !std::is_same<T, double>::value;
  // the strip kernel is opt-in (SGB_FIR_STRIP=1): measured on a B200 it does not beat upfirdn2d_cl_kernel (profiles/README.md:
  // 0.223 vs 0.248 ms on fp16 [4,32,1025,1025] but 0.292 vs 0.277 ms on fp32 [32,64,257,257], slower for up / down 2)
  static const int fir_strip = [] { const char* e = getenv("SGB_FIR_STRIP"); return e ? atoi(e) : 0; }();
  if (cl_ok && fir_strip) {
    if (int r = launch_strip<T>(p, s)) { if (r < 0) return 1; } else return 0;      // > 0: not taken, fall through
  }
  if (cl_ok) {
    const int cv_total = p.c / VEC;
    int cvb = 1; while (cvb < cv_total && cvb < 32) cvb <<= 1;
    const int txg = 256 / cvb / 4;                       // x groups per CTA (4 rows per CTA)
    const int64_t gx = ceil_div(p.out_w, txg * 4), gy = ceil_div(p.out_h, 4), gz = (int64_t)p.n * ceil_div(cv_total, cvb);
    if (gy <= 65535 && gz <= 65535) {
      dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(cvb, 256 / cvb);
      const int pp = p.padx0 & 1;
      static const int v2 = [] { const char* e = getenv("SGB_FIR_V2"); return e ? atoi(e) : 1; }();
      // measured (profiles/README.md): hoisting all loads wins for up = 2 (2 live filter rows, 48-64 registers: fp32
      // [32,64,128,128] 0.148 -> 0.124 ms = 5.4 TB/s) and loses for the 4-row forms (132-255 registers, occupancy)
      if (v2 && p.upx == 2) {
        if (pp == 0) upfirdn2d_cl2_kernel<T, 2, 1, 0><<<grid, block, 0, s>>>(p, cv_total, cvb, txg);
        else         upfirdn2d_cl2_kernel<T, 2, 1, 1><<<grid, block, 0, s>>>(p, cv_total, cvb, txg);
        SGB_LAUNCH_CHECK();
        return 0;
      }
      if (p.upx == 1 && p.downx == 1)      upfirdn2d_cl_kernel<T, 1, 1, 0><<<grid, block, 0, s>>>(p, cv_total, cvb, txg);
      else if (p.upx == 2 && pp == 0)      upfirdn2d_cl_kernel<T, 2, 1, 0><<<grid, block, 0, s>>>(p, cv_total, cvb, txg);
      else if (p.upx == 2)                 upfirdn2d_cl_kernel<T, 2, 1, 1><<<grid, block, 0, s>>>(p, cv_total, cvb, txg);
      else                                 upfirdn2d_cl_kernel<T, 1, 2, 0><<<grid, block, 0, s>>>(p, cv_total, cvb, txg);
      SGB_LAUNCH_CHECK();
      return 0;
    }
  }
  int64_t blocks = ceil_div(total, 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
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

/* upfirdn2d without resampling for a 4 x 4 (or smaller) filter the caller has factored into fy (x) fx -- see fir_sep_strip_kernel.
 * fx / fy: HOST arrays of 4 taps in application order (flip and gain already applied; unused taps zero). */
extern "C" int sgb_upfirdn2d_sep(const void* x, const float* fx, const float* fy, void* y, int dtype,
                                 int n, int c, int in_h, int in_w, const int64_t x_strides[4],
                                 int out_h, int out_w, const int64_t y_strides[4], int padx0, int pady0, void* stream) {
  SGB_REQUIRE(x && fx && fy && y, "x, fx, fy and y must not be NULL");
  SGB_REQUIRE(dtype == SGB_F32 || dtype == SGB_F16 || dtype == SGB_BF16, "unsupported dtype");
  SGB_REQUIRE(n >= 0 && c >= 1 && in_h >= 1 && in_w >= 1 && out_h >= 1 && out_w >= 1, "bad size");
  if (n == 0) return 0;
  const int vec = dtype == SGB_F32 ? 4 : 8;
  SGB_REQUIRE(x_strides[1] == 1 && y_strides[1] == 1 && c % vec == 0 && aligned16(x) && aligned16(y), "needs channels_last tensors with 16-byte channel vectors");
  for (int i : {0, 2, 3}) SGB_REQUIRE(x_strides[i] % vec == 0 && y_strides[i] % vec == 0, "strides must keep 16-byte alignment");
  FirSepParams p;
  p.x = x; p.y = y; p.n = n; p.c = c; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
  for (int i = 0; i < 4; i++) { p.xs[i] = x_strides[i]; p.ys[i] = y_strides[i]; p.fx[i] = fx[i]; p.fy[i] = fy[i]; }
  p.padx0 = padx0; p.pady0 = pady0;
  const int PX = vec == 4 ? 4 : 2;
  p.cv_total = c / vec;
  int cvb = 1; while (cvb < p.cv_total && cvb < 32) cvb <<= 1;
  p.cvb = cvb; p.xgs = 256 / cvb;
  const int64_t gx = ceil_div(out_w, p.xgs * PX), gz = (int64_t)n * ceil_div(p.cv_total, cvb);
  int rows = out_h >= 128 ? 32 : (out_h >= 32 ? 16 : 8);
  while (rows > 8 && gx * ceil_div(out_h, rows) * gz < 4 * (int64_t)num_sms()) rows >>= 1;
  p.rows = rows;
  const int64_t gy = ceil_div(out_h, rows);
  SGB_REQUIRE(gy <= 65535 && gz <= 65535, "grid too large");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(cvb, p.xgs);
  cudaStream_t s = (cudaStream_t)stream;
  // SGB_FIR_PF: 0 = register kernel for every type (A/B), 1 = prefetching kernel for 16-bit tensors only, 2 (default) = for
  // fp32 tensors as well (PX = 4: 7 slots x 4 stages = 112 KB per CTA, two CTAs still fit an SM).  Measured (r2_run28.sh,
  // r2_run30.sh): fp16 [4,32,1025,1025] 3.30 -> 4.91 TB/s, fp32 [8,64,257,257] 5.13 -> 5.97 TB/s (0.91 of the copy rate),
  // ffhq256 step 69.5 -> 68.2 ms
  static const int use_pf = [] { const char* e = getenv("SGB_FIR_PF"); return e ? atoi(e) : 2; }();
  constexpr int PF_SMEM = 4 * 5 * 256 * 16, PF_SMEM32 = 4 * 7 * 256 * 16;
  if (dtype == SGB_F32 && use_pf >= 2) {
    SGB_SET_MAX_SMEM((fir_sep_strip_pf_kernel<float, 4>), PF_SMEM32);
    fir_sep_strip_pf_kernel<float, 4><<<grid, block, PF_SMEM32, s>>>(p);
  } else if (dtype == SGB_F32) fir_sep_strip_kernel<float><<<grid, block, 0, s>>>(p);
  else if (!use_pf) {
    if (dtype == SGB_F16) fir_sep_strip_kernel<__half><<<grid, block, 0, s>>>(p);
    else fir_sep_strip_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(p);
  } else if (dtype == SGB_F16) {
    SGB_SET_MAX_SMEM((fir_sep_strip_pf_kernel<__half, 2>), PF_SMEM);
    fir_sep_strip_pf_kernel<__half, 2><<<grid, block, PF_SMEM, s>>>(p);
  } else {
    SGB_SET_MAX_SMEM((fir_sep_strip_pf_kernel<__nv_bfloat16, 2>), PF_SMEM);
    fir_sep_strip_pf_kernel<__nv_bfloat16, 2><<<grid, block, PF_SMEM, s>>>(p);
  }
  SGB_LAUNCH_CHECK();
  return 0;
}
