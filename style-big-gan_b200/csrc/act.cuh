// Activation table of bias_act (reference bias_act.py:23-33, kernel math bias_act.cu:52-146), shared by
// the standalone bias_act kernels and the fused convolution epilogues.
#pragma once
#include "common.cuh"

namespace sgb {

// One element.  G: 0 forward, 1 first derivative, 2 second derivative.
// x: forward input (+bias already added) for G == 0, otherwise the incoming gradient.
// xr: xref + bias, yr: yref, dy: extra multiplier (1 unless G == 2).
template <class A, int ACT, int G>
__device__ __forceinline__ A act_eval(A x, A xr, A yr, A dy, A alpha, A gain, A clamp) {
  const A one = A(1), two = A(2);
  const A kExpRange = A(80), kHalfExpRange = A(40);
  const A kSeluScale = A(1.0507009873554804934193349852946);
  const A kSeluAlpha = A(1.6732632423543772848170429916717);
  A yy = (gain != A(0)) ? yr / gain : A(0);   // activation output before gain
  A y = A(0);
  if (ACT == SGB_ACT_LINEAR) {
    if (G <= 1) y = x;
  } else if (ACT == SGB_ACT_RELU) {
    if (G == 0) y = x > A(0) ? x : A(0);
    if (G == 1) y = yy > A(0) ? x : A(0);
  } else if (ACT == SGB_ACT_LRELU) {
    if (G == 0) y = x > A(0) ? x : x * alpha;
    if (G == 1) y = yy > A(0) ? x : x * alpha;
  } else if (ACT == SGB_ACT_TANH) {
    if (G == 0) { A c = exp(x); A d = one / c; y = x < -kExpRange ? -one : (x > kExpRange ? one : (c - d) / (c + d)); }
    if (G == 1) y = x * (one - yy * yy);
    if (G == 2) y = x * (one - yy * yy) * (-two * yy);
  } else if (ACT == SGB_ACT_SIGMOID) {
    if (G == 0) y = x < -kExpRange ? A(0) : one / (exp(-x) + one);
    if (G == 1) y = x * yy * (one - yy);
    if (G == 2) y = x * yy * (one - yy) * (one - two * yy);
  } else if (ACT == SGB_ACT_ELU) {
    if (G == 0) y = x >= A(0) ? x : exp(x) - one;
    if (G == 1) y = yy >= A(0) ? x : x * (yy + one);
    if (G == 2) y = yy >= A(0) ? A(0) : x * (yy + one);
  } else if (ACT == SGB_ACT_SELU) {
    if (G == 0) y = x >= A(0) ? kSeluScale * x : (kSeluScale * kSeluAlpha) * (exp(x) - one);
    if (G == 1) y = yy >= A(0) ? x * kSeluScale : x * (yy + kSeluScale * kSeluAlpha);
    if (G == 2) y = yy >= A(0) ? A(0) : x * (yy + kSeluScale * kSeluAlpha);
  } else if (ACT == SGB_ACT_SOFTPLUS) {
    if (G == 0) y = x > kExpRange ? x : log(exp(x) + one);
    if (G == 1) y = x * (one - exp(-yy));
    if (G == 2) { A c = exp(-yy); y = x * c * (one - c); }
  } else if (ACT == SGB_ACT_SWISH) {
    if (G == 0) {
      y = x < -kExpRange ? A(0) : x / (exp(-x) + one);
    } else {
      A c = exp(xr), d = c + one;
      if (G == 1) y = xr > kHalfExpRange ? x : x * c * (xr + d) / (d * d);
      else        y = xr > kHalfExpRange ? A(0) : x * c * (xr * (two - d) + two * d) / (d * d * d);
      yr = xr < -kExpRange ? A(0) : xr / (exp(-xr) + one) * gain;   // swish keeps x, not y: rebuild y for the clamp mask
    }
  }
  y *= gain * dy;
  if (clamp >= A(0)) {
    if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= A(0) ? clamp : -clamp);
    else        y = (yr > -clamp && yr < clamp) ? y : A(0);
  }
  return y;
}


// runtime-act forward evaluation (fused epilogues): bias already added to x
template <class A>
__device__ __forceinline__ A act_forward(int act, A x, A alpha, A gain, A clamp) {
  switch (act) {
    case SGB_ACT_LINEAR:   return act_eval<A, SGB_ACT_LINEAR, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
    case SGB_ACT_RELU:     return act_eval<A, SGB_ACT_RELU, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
    case SGB_ACT_LRELU:    return act_eval<A, SGB_ACT_LRELU, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
    case SGB_ACT_TANH:     return act_eval<A, SGB_ACT_TANH, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
    case SGB_ACT_SIGMOID:  return act_eval<A, SGB_ACT_SIGMOID, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
    case SGB_ACT_ELU:      return act_eval<A, SGB_ACT_ELU, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
    case SGB_ACT_SELU:     return act_eval<A, SGB_ACT_SELU, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
    case SGB_ACT_SOFTPLUS: return act_eval<A, SGB_ACT_SOFTPLUS, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
    case SGB_ACT_SWISH:    return act_eval<A, SGB_ACT_SWISH, 0>(x, A(0), A(0), A(1), alpha, gain, clamp);
  }
  return x;
}

}  // namespace sgb
