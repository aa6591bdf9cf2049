// Generic SIMT implicit-GEMM convolution for sm_100a: every geometry the hot path can ask for
// (conv2d / conv_transpose2d, any stride / padding / groups / kernel size / layout, f16 / bf16 / f32 / f64),
// with fp32 (fp64) accumulation.  This is the shape-complete path; descriptors that qualify are routed to
// the tcgen05 kernel in conv_umma.cu instead (see conv_dispatch.cu).  Replaces the aten/cuDNN calls of
// conv2d_gradfix.py:112-114 and :143-145.
//
// GEMM view per group:  D[m, o] = sum_{tap, c} A[m, (tap, c)] * W[(tap, c), o]
//   m = (n, oy, ox) output pixel;  A gathers x[n, c, iy, ix]:
//     conv2d:            iy = oy*stride + ky - pad_y
//     conv_transpose2d:  iy = (oy + pad_y - ky) / stride   when divisible
//   Tile 64 x 64 x 16, 256 threads, 4 x 4 outputs per thread, operands staged in shared memory as
//   accumulator type.  Fused prologue (in_scale) and epilogue (out_scale, noise, bias_act) as described in
//   include/sgb200.h.
#include "common.cuh"
#include "act.cuh"

namespace sgb {

constexpr int BM = 64, BN = 64, BK = 16;

struct ConvParams {
  sgb_conv_desc d;
  const void* x; const void* w; void* y;     // forward
  const void* dy; void* dw;                  // wgrad
  int ci_g, co_g;                            // channels per group
  int x_cl, y_cl;                            // channel-contiguous layouts
  int splits;                                // wgrad split-K
};

template <class T>
__device__ __forceinline__ typename Acc<T>::type load_scaled(const T* p, const typename Acc<T>::type* scale, int64_t sidx) {
  typename Acc<T>::type v = to_acc<T>(*p);
  if (scale) v *= scale[sidx];
  return v;
}

// ---------------------------------------------------------------------------------------------------
// forward (conv2d or conv_transpose2d)
template <class T, bool CL>
__global__ void __launch_bounds__(256) conv_fwd_simt_kernel(ConvParams p) {
  typedef typename Acc<T>::type A;
  const sgb_conv_desc& d = p.d;
  __shared__ A As[BK][BM + 4];
  __shared__ A Bs[BK][BN + 4];

  const int g = blockIdx.z;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int o0 = blockIdx.y * BN;               // within group
  const int64_t M = (int64_t)d.n * d.out_h * d.out_w;
  const int ohw = d.out_h * d.out_w;
  const int tid = threadIdx.x;
  const A* in_scale = (const A*)d.in_scale;

  // A-load assignment: 4 elements per thread
  //   !CL: m = tid % 64 (fixed), k = tid / 64 + 4 i       (coalesced over pixels)
  //    CL: k = tid % 16 (fixed), m = tid / 16 + 16 i      (coalesced over channels)
  int am[4], ak[4];
  int an[4], aoy[4], aox[4];
  for (int i = 0; i < 4; i++) {
    am[i] = CL ? (tid / 16 + 16 * i) : (tid % 64);
    ak[i] = CL ? (tid % 16) : (tid / 64 + 4 * i);
    int64_t m = m0 + am[i];
    if (m < M) { an[i] = (int)(m / ohw); int r = (int)(m - (int64_t)an[i] * ohw); aoy[i] = r / d.out_w; aox[i] = r - aoy[i] * d.out_w; }
    else { an[i] = -1; aoy[i] = 0; aox[i] = 0; }
  }
  // B-load assignment: o = tid % 64, k = tid / 64 + 4 i
  const int bo = tid % 64;

  const int ty = tid / 16, tx = tid % 16;       // micro-tile: rows ty*4.., cols tx*4..
  A acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = A(0);

  const int taps = d.kh * d.kw;
  for (int tap = 0; tap < taps; tap++) {
    const int ky = tap / d.kw, kx = tap - ky * d.kw;
    const int wy = d.flip ? d.kh - 1 - ky : ky, wx = d.flip ? d.kw - 1 - kx : kx;
    for (int c0 = 0; c0 < p.ci_g; c0 += BK) {
      // ---- stage A
#pragma unroll
      for (int i = 0; i < 4; i++) {
        A v = A(0);
        const int c = c0 + ak[i];
        if (an[i] >= 0 && c < p.ci_g) {
          int iy, ix; bool ok;
          if (!d.transposed) {
            iy = aoy[i] * d.stride + ky - d.pad_y; ix = aox[i] * d.stride + kx - d.pad_x;
            ok = iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w;
          } else {
            const int ty_ = aoy[i] + d.pad_y - ky, tx_ = aox[i] + d.pad_x - kx;
            ok = ty_ >= 0 && tx_ >= 0 && (ty_ % d.stride == 0) && (tx_ % d.stride == 0);
            iy = ty_ / d.stride; ix = tx_ / d.stride;
            ok = ok && iy < d.in_h && ix < d.in_w;
          }
          if (ok) {
            const int cg = g * p.ci_g + c;
            v = load_scaled<T>((const T*)p.x + an[i] * d.x_strides[0] + cg * d.x_strides[1] + iy * d.x_strides[2] + ix * d.x_strides[3],
                               in_scale, (int64_t)an[i] * d.ci + cg);
          }
        }
        As[ak[i]][am[i]] = v;
      }
      // ---- stage B (weights)
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int k = tid / 64 + 4 * i;
        const int c = c0 + k, o = o0 + bo;
        A v = A(0);
        if (c < p.ci_g && o < p.co_g) {
          int64_t widx;
          if (!d.transposed) widx = (((int64_t)(g * p.co_g + o) * p.ci_g + c) * d.kh + wy) * d.kw + wx;     // [co, ci_g, kh, kw]
          else               widx = (((int64_t)(g * p.ci_g + c) * p.co_g + o) * d.kh + wy) * d.kw + wx;     // [ci, co_g, kh, kw]
          v = to_acc<T>(((const T*)p.w)[widx]);
        }
        Bs[k][bo] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; k++) {
        A a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
      }
      __syncthreads();
    }
  }

  // ---- epilogue
  const A* out_scale = (const A*)d.out_scale;
  const A* noise = (const A*)d.noise;
  const A alpha = A(d.alpha), gain = A(d.gain), clamp = A(d.clamp);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int n = (int)(m / ohw); const int r = (int)(m - (int64_t)n * ohw); const int oy = r / d.out_w, ox = r - oy * d.out_w;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int o = o0 + tx * 4 + j;
      if (o >= p.co_g) continue;
      const int og = g * p.co_g + o;
      A v = acc[i][j];
      if (out_scale) v *= out_scale[(int64_t)n * d.co + og];
      if (noise) v += noise[((int64_t)n * d.out_h + oy) * d.out_w + ox];
      if (d.act) {
        if (d.bias) v += to_acc<T>(((const T*)d.bias)[og]);
        v = act_forward<A>(d.act, v, alpha, gain, clamp);
      }
      ((T*)p.y)[n * d.y_strides[0] + og * d.y_strides[1] + oy * d.y_strides[2] + ox * d.y_strides[3]] = from_acc<T>(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// weight gradient of a (non-transposed) conv:  dw[o, c, ky, kx] = sum_{n,oy,ox} dy[n,o,oy,ox] * x[n,c,iy,ix]
// grid: (o tiles, c tiles * taps, groups * splits); accumulates with atomics into a zeroed dw.
template <class T, bool CL>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(ConvParams p) {
  typedef typename Acc<T>::type A;
  const sgb_conv_desc& d = p.d;
  __shared__ A As[BK][BM + 4];   // [pixel][o]
  __shared__ A Bs[BK][BN + 4];   // [pixel][c]

  const int ctiles = (p.ci_g + BN - 1) / BN;
  const int tap = blockIdx.y / ctiles;
  const int c0 = (blockIdx.y - tap * ctiles) * BN;
  const int o0 = blockIdx.x * BM;
  const int g = blockIdx.z / p.splits;
  const int split = blockIdx.z - g * p.splits;
  const int ky = tap / d.kw, kx = tap - ky * d.kw;
  const int ohw = d.out_h * d.out_w;
  const int64_t P = (int64_t)d.n * ohw;
  const int64_t chunk = ((P + p.splits - 1) / p.splits + BK - 1) / BK * BK;
  const int64_t p_begin = split * chunk;
  const int64_t p_end = (p_begin + chunk < P) ? p_begin + chunk : P;
  const int tid = threadIdx.x;
  const A* in_scale = (const A*)d.in_scale;

  // load assignment: !CL: pixel = tid % 16 (fast), ch = tid / 16 + 16 i ;  CL: ch = tid % 64 (fast), pixel = tid / 64 + 4 i
  const int ty = tid / 16, tx = tid % 16;
  A acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = A(0);

  for (int64_t pb = p_begin; pb < p_end; pb += BK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int pk = CL ? (tid / 64 + 4 * i) : (tid % 16);
      const int ch = CL ? (tid % 64) : (tid / 16 + 16 * i);
      const int64_t pix = pb + pk;
      A va = A(0), vb = A(0);
      if (pix < p_end) {
        const int n = (int)(pix / ohw); const int r = (int)(pix - (int64_t)n * ohw); const int oy = r / d.out_w, ox = r - oy * d.out_w;
        const int o = o0 + ch;
        if (o < p.co_g) {
          const int og = g * p.co_g + o;
          va = to_acc<T>(((const T*)p.dy)[n * d.y_strides[0] + og * d.y_strides[1] + oy * d.y_strides[2] + ox * d.y_strides[3]]);
        }
        const int c = c0 + ch;
        const int iy = oy * d.stride + ky - d.pad_y, ix = ox * d.stride + kx - d.pad_x;
        if (c < p.ci_g && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w) {
          const int cg = g * p.ci_g + c;
          vb = load_scaled<T>((const T*)p.x + n * d.x_strides[0] + cg * d.x_strides[1] + iy * d.x_strides[2] + ix * d.x_strides[3],
                              in_scale, (int64_t)n * d.ci + cg);
        }
      }
      As[pk][ch] = va;
      Bs[pk][ch] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      A a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }

  const int wy = d.flip ? d.kh - 1 - ky : ky, wx = d.flip ? d.kw - 1 - kx : kx;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int o = o0 + ty * 4 + i;
    if (o >= p.co_g) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int c = c0 + tx * 4 + j;
      if (c >= p.ci_g) continue;
      const int64_t widx = (((int64_t)(g * p.co_g + o) * p.ci_g + c) * d.kh + wy) * d.kw + wx;
      atomicAdd((A*)p.dw + widx, acc[i][j]);
    }
  }
}

int conv_forward_simt(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  ConvParams p; p.d = *d; p.x = x; p.w = w; p.y = y; p.dy = nullptr; p.dw = nullptr;
  p.ci_g = d->ci / d->groups; p.co_g = d->co / d->groups; p.splits = 1;
  p.x_cl = (d->x_strides[1] == 1 && d->ci > 1); p.y_cl = (d->y_strides[1] == 1 && d->co > 1);
  const int64_t M = (int64_t)d->n * d->out_h * d->out_w;
  const int64_t gx = ceil_div(M, BM), gy = ceil_div(p.co_g, BN);
  SGB_REQUIRE(gx <= 0x7fffffff && gy <= 65535 && d->groups <= 65535, "problem too large for the SIMT conv grid");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)d->groups);
  SGB_DISPATCH_DTYPE(d->dtype, {
    if (p.x_cl) conv_fwd_simt_kernel<T, true><<<grid, 256, 0, s>>>(p);
    else        conv_fwd_simt_kernel<T, false><<<grid, 256, 0, s>>>(p);
    SGB_LAUNCH_CHECK();
  });
  return 0;
}

int conv_wgrad_simt(const sgb_conv_desc* d, const void* x, const void* dy, void* dw, cudaStream_t s) {
  ConvParams p; p.d = *d; p.x = x; p.w = nullptr; p.y = nullptr; p.dy = dy; p.dw = dw;
  p.ci_g = d->ci / d->groups; p.co_g = d->co / d->groups;
  p.x_cl = (d->x_strides[1] == 1 && d->ci > 1); p.y_cl = (d->y_strides[1] == 1 && d->co > 1);
  const size_t esz = d->dtype == SGB_F64 ? 8 : 4;
  const int64_t wnum = (int64_t)d->co * p.ci_g * d->kh * d->kw;
  cudaError_t e = cudaMemsetAsync(dw, 0, esz * wnum, s);
  SGB_REQUIRE(e == cudaSuccess, "memset failed");
  const int64_t P = (int64_t)d->n * d->out_h * d->out_w;
  if (P == 0) return 0;
  const int taps = d->kh * d->kw;
  const int64_t gx = ceil_div(p.co_g, BM), gy = ceil_div(p.ci_g, BN) * taps;
  int64_t splits = ceil_div((int64_t)num_sms() * 4, gx * gy * d->groups);
  const int64_t max_splits = ceil_div(P, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  SGB_REQUIRE(gy <= 65535 && d->groups * splits <= 65535, "problem too large for the SIMT wgrad grid");
  p.splits = (int)splits;
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)(d->groups * splits));
  SGB_DISPATCH_DTYPE(d->dtype, {
    if (p.x_cl && p.y_cl) conv_wgrad_simt_kernel<T, true><<<grid, 256, 0, s>>>(p);
    else                  conv_wgrad_simt_kernel<T, false><<<grid, 256, 0, s>>>(p);
    SGB_LAUNCH_CHECK();
  });
  return 0;
}

}  // namespace sgb
