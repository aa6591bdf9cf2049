// bias_act for sm_100a: y = clamp(act(x + b) * gain), its first derivative (grad = 1) and second
// derivative (grad = 2).  Replaces bias_act.cu:23-147 / bias_act.cpp:32-90 of the reference.
//
// HBM-bound elementwise kernel (algorithmic bytes: 2*numel*sizeof(T) forward, 3*numel*sizeof(T)
// backward).  Design: 128-bit streaming loads/stores (L1::no_allocate), UNROLL independent vectors in
// flight per thread, persistent grid-stride over a grid of 148*k CTAs, and a division-free bias index:
// the (channel, offset-in-channel) pair of each thread is advanced incrementally instead of doing the
// reference's per-element (i / stepB) % sizeB.
#include "common.cuh"
#include "act.cuh"

namespace sgb {

struct BiasActParams {
  const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
  float alpha, gain, clamp;
  int64_t size_x, size_b, step_b;
  int bmode;   // 0 none, 1 plane (step_b % VEC == 0), 2 inner (step_b == 1, size_b % VEC == 0), 3 generic
};

template <class T, int ACT, int G, int UNROLL>
__global__ void __launch_bounds__(256) bias_act_vec_kernel(BiasActParams p) {
  typedef typename Acc<T>::type A;
  constexpr int VEC = Vec16<T>::N;
  // derivative kernels mask on the STORED y: a saturated output was rounded to T, so compare against the
  // clamp value rounded the same way (fp16(181.02) = 181.0 would otherwise never look saturated)
  const A alpha = A(p.alpha), gain = A(p.gain);
  const A clamp = (G > 0 && p.clamp >= 0.f) ? A(to_acc<T>(from_acc<T>(A(p.clamp)))) : A(p.clamp);
  const int64_t nvec = p.size_x / VEC;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t vi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;

  // incremental bias index state
  const T* bp = (const T*)p.b;
  int64_t chan = 0, rem = 0, d_chan = 0, d_rem = 0, period = 1;
  if (p.bmode == 1) {            // plane: all VEC elements of a vector share one bias value
    period = p.step_b / VEC;
    int64_t q = vi / period; rem = vi - q * period; chan = q % p.size_b;
    int64_t dq = stride / period; d_rem = stride - dq * period; d_chan = dq % p.size_b;
  } else if (p.bmode == 2) {     // inner: bias index == element index mod size_b, vector-aligned
    period = p.size_b / VEC;
    chan = vi % period; d_chan = stride % period;
  }

  for (; vi < nvec; vi += stride * UNROLL) {
    Vec16<T> vx[UNROLL], vxr[UNROLL], vyr[UNROLL], vdy[UNROLL];
    int64_t ch[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int64_t v = vi + (int64_t)u * stride;
      ch[u] = chan;
      if (p.bmode == 1) {
        rem += d_rem; chan += d_chan;
        if (rem >= period) { rem -= period; chan += 1; }
        if (chan >= p.size_b) chan -= p.size_b;
      } else if (p.bmode == 2) {
        chan += d_chan; if (chan >= period) chan -= period;
      }
      if (v < nvec) {
        vx[u].raw = ld_stream((const uint4*)p.x + v);
        if (G > 0 && p.xref) vxr[u].raw = ld_stream((const uint4*)p.xref + v);
        if (G > 0 && p.yref) vyr[u].raw = ld_stream((const uint4*)p.yref + v);
        if (G == 2 && p.dy)  vdy[u].raw = ld_stream((const uint4*)p.dy + v);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int64_t v = vi + (int64_t)u * stride;
      if (v >= nvec) break;
      A bias[VEC];
      if (p.bmode == 0) {
#pragma unroll
        for (int j = 0; j < VEC; j++) bias[j] = A(0);
      } else if (p.bmode == 1) {
        A bv = to_acc<T>(bp[ch[u]]);
#pragma unroll
        for (int j = 0; j < VEC; j++) bias[j] = bv;
      } else if (p.bmode == 2) {
        Vec16<T> vb; vb.raw = *((const uint4*)bp + ch[u]);
#pragma unroll
        for (int j = 0; j < VEC; j++) bias[j] = to_acc<T>(vb.v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; j++) bias[j] = to_acc<T>(bp[((v * VEC + j) / p.step_b) % p.size_b]);
      }
      Vec16<T> out;
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        A x = to_acc<T>(vx[u].v[j]);
        A xr = (G > 0 && p.xref) ? to_acc<T>(vxr[u].v[j]) : A(0);
        A yr = (G > 0 && p.yref) ? to_acc<T>(vyr[u].v[j]) : A(0);
        A dy = (G == 2 && p.dy) ? to_acc<T>(vdy[u].v[j]) : A(1);
        if (G == 0) x += bias[j]; else xr += bias[j];
        out.v[j] = from_acc<T>(act_eval<A, ACT, G>(x, xr, yr, dy, alpha, gain, clamp));
      }
      st_stream((uint4*)p.y + v, out.raw);
    }
  }
}

// scalar kernel: tails, unaligned pointers, odd sizes
template <class T, int ACT, int G>
__global__ void __launch_bounds__(256) bias_act_scalar_kernel(BiasActParams p, int64_t begin) {
  typedef typename Acc<T>::type A;
  // derivative kernels mask on the STORED y: a saturated output was rounded to T, so compare against the
  // clamp value rounded the same way (fp16(181.02) = 181.0 would otherwise never look saturated)
  const A alpha = A(p.alpha), gain = A(p.gain);
  const A clamp = (G > 0 && p.clamp >= 0.f) ? A(to_acc<T>(from_acc<T>(A(p.clamp)))) : A(p.clamp);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.size_x; i += stride) {
    A x = to_acc<T>(((const T*)p.x)[i]);
    A b = p.b ? to_acc<T>(((const T*)p.b)[(i / p.step_b) % p.size_b]) : A(0);
    A xr = (G > 0 && p.xref) ? to_acc<T>(((const T*)p.xref)[i]) : A(0);
    A yr = (G > 0 && p.yref) ? to_acc<T>(((const T*)p.yref)[i]) : A(0);
    A dy = (G == 2 && p.dy) ? to_acc<T>(((const T*)p.dy)[i]) : A(1);
    if (G == 0) x += b; else xr += b;
    ((T*)p.y)[i] = from_acc<T>(act_eval<A, ACT, G>(x, xr, yr, dy, alpha, gain, clamp));
  }
}

template <class T, int ACT, int G>
static int launch_bias_act(BiasActParams p, cudaStream_t stream) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int UNROLL = 4;
  bool vec_ok = aligned16(p.x) && aligned16(p.y) && (!p.xref || aligned16(p.xref)) &&
                (!p.yref || aligned16(p.yref)) && (!p.dy || aligned16(p.dy));
  p.bmode = 0;
  if (p.b) {
    if (p.step_b % VEC == 0) p.bmode = 1;
    else if (p.step_b == 1 && p.size_b % VEC == 0 && aligned16(p.b)) p.bmode = 2;
    else p.bmode = 3;
  }
  int64_t done = 0;
  if (vec_ok && p.size_x >= VEC) {
    int64_t nvec = p.size_x / VEC;
    int64_t blocks = ceil_div(nvec, 256 * UNROLL);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    bias_act_vec_kernel<T, ACT, G, UNROLL><<<(unsigned)blocks, 256, 0, stream>>>(p);
    SGB_LAUNCH_CHECK();
    done = nvec * VEC;
  }
  if (done < p.size_x) {
    int64_t rest = p.size_x - done;
    int64_t blocks = ceil_div(rest, 256);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    bias_act_scalar_kernel<T, ACT, G><<<(unsigned)blocks, 256, 0, stream>>>(p, done);
    SGB_LAUNCH_CHECK();
  }
  return 0;
}

template <class T, int ACT>
static int dispatch_grad(const BiasActParams& p, int grad, cudaStream_t s) {
  if (grad == 0) return launch_bias_act<T, ACT, 0>(p, s);
  if (grad == 1) return launch_bias_act<T, ACT, 1>(p, s);
  if (ACT >= SGB_ACT_TANH) return launch_bias_act<T, ACT, 2>(p, s);
  set_error("sgb_bias_act: activation has no second derivative kernel (it is identically zero)");
  return 1;
}

template <class T>
static int dispatch_act(const BiasActParams& p, int grad, int act, cudaStream_t s) {
  switch (act) {
    case SGB_ACT_LINEAR:   return dispatch_grad<T, SGB_ACT_LINEAR>(p, grad, s);
    case SGB_ACT_RELU:     return dispatch_grad<T, SGB_ACT_RELU>(p, grad, s);
    case SGB_ACT_LRELU:    return dispatch_grad<T, SGB_ACT_LRELU>(p, grad, s);
    case SGB_ACT_TANH:     return dispatch_grad<T, SGB_ACT_TANH>(p, grad, s);
    case SGB_ACT_SIGMOID:  return dispatch_grad<T, SGB_ACT_SIGMOID>(p, grad, s);
    case SGB_ACT_ELU:      return dispatch_grad<T, SGB_ACT_ELU>(p, grad, s);
    case SGB_ACT_SELU:     return dispatch_grad<T, SGB_ACT_SELU>(p, grad, s);
    case SGB_ACT_SOFTPLUS: return dispatch_grad<T, SGB_ACT_SOFTPLUS>(p, grad, s);
    case SGB_ACT_SWISH:    return dispatch_grad<T, SGB_ACT_SWISH>(p, grad, s);
  }
  set_error("sgb_bias_act: unknown activation id");
  return 1;
}

// ---- per-channel sum: x[outer, C, inner] -> out[C] (fp32) -------------------------------------------
// grid (C-tile, splits); each CTA reduces a slab and does one atomicAdd per channel.
template <class T>
__global__ void __launch_bounds__(256) sum_to_channel_inner_kernel(const T* __restrict__ x,
                                                                   typename Acc<T>::type* __restrict__ out,
                                                                   int64_t outer, int64_t C, int64_t inner) {
  // inner > 1: one CTA per (channel, split of outer); threads sweep the contiguous inner run
  const int64_t c = blockIdx.x;
  typedef typename Acc<T>::type A;
  A acc = A(0);
  for (int64_t o = blockIdx.y; o < outer; o += gridDim.y) {
    const T* row = x + (o * C + c) * inner;
    for (int64_t i = threadIdx.x; i < inner; i += blockDim.x) acc += to_acc<T>(row[i]);
  }
  __shared__ double part[8];
  double v = warp_sum((double)acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; for (int i = 0; i < 8; i++) t += part[i];
    atomicAdd(out + c, (A)t);
  }
}

template <class T>
__global__ void __launch_bounds__(256) sum_to_channel_last_kernel(const T* __restrict__ x,
                                                                  typename Acc<T>::type* __restrict__ out,
                                                                  int64_t rows, int64_t C) {
  // inner == 1: column sums of a [rows, C] matrix; thread.x walks channels (coalesced), thread.y rows
  typedef typename Acc<T>::type A;
  __shared__ A part[8][33];
  const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
  A acc = A(0);
  if (c < C)
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8)
      acc += to_acc<T>(x[r * C + c]);
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    A t = 0; for (int i = 0; i < 8; i++) t += part[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

}  // namespace sgb

using namespace sgb;

extern "C" int sgb_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                            int dtype, int grad, int act, float alpha, float gain, float clamp,
                            int64_t size_x, int64_t size_b, int64_t step_b, void* stream) {
  SGB_REQUIRE(size_x >= 0, "negative size");
  if (size_x == 0) return 0;
  SGB_REQUIRE(x && y, "x and y must not be NULL");
  SGB_REQUIRE(grad >= 0 && grad <= 2, "grad must be 0, 1 or 2");
  SGB_REQUIRE(!b || (size_b >= 1 && step_b >= 1), "bias needs size_b >= 1 and step_b >= 1");
  BiasActParams p;
  p.x = x; p.b = b; p.xref = xref; p.yref = yref; p.dy = dy; p.y = y;
  p.alpha = alpha; p.gain = gain; p.clamp = clamp;
  p.size_x = size_x; p.size_b = b ? size_b : 1; p.step_b = b ? step_b : 1; p.bmode = 0;
  cudaStream_t s = (cudaStream_t)stream;
  SGB_DISPATCH_DTYPE(dtype, return dispatch_act<T>(p, grad, act, s));
  return 0;
}

extern "C" int sgb_sum_to_channel(const void* x, void* out, int dtype, int64_t outer, int64_t size_c, int64_t inner,
                                  void* stream) {
  SGB_REQUIRE(out && size_c >= 1, "out must not be NULL");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, (dtype == SGB_F64 ? sizeof(double) : sizeof(float)) * size_c, s);
  SGB_REQUIRE(e == cudaSuccess, "memset failed");
  if (outer <= 0 || inner <= 0) return 0;
  SGB_REQUIRE(x, "x must not be NULL");
  SGB_DISPATCH_DTYPE(dtype, {
    if (inner == 1) {
      int64_t gx = ceil_div(size_c, 32);
      int64_t gy = ceil_div(outer, 8 * 16); if (gy < 1) gy = 1;
      int64_t cap = ceil_div((int64_t)num_sms() * 8, gx); if (gy > cap) gy = cap; if (gy > 65535) gy = 65535;
      SGB_REQUIRE(gx <= 0x7fffffff, "too many channels");
      sum_to_channel_last_kernel<T><<<dim3((unsigned)gx, (unsigned)gy), dim3(32, 8), 0, s>>>((const T*)x, (Acc<T>::type*)out, outer, size_c);
    } else {
      SGB_REQUIRE(size_c <= 0x7fffffff, "too many channels");
      int64_t gy = ceil_div((int64_t)num_sms() * 4, size_c); if (gy > outer) gy = outer; if (gy < 1) gy = 1;
      if (gy > 65535) gy = 65535;
      sum_to_channel_inner_kernel<T><<<dim3((unsigned)size_c, (unsigned)gy), 256, 0, s>>>((const T*)x, (Acc<T>::type*)out, outer, size_c, inner);
    }
    SGB_LAUNCH_CHECK();
  });
  return 0;
}
