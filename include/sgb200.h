/*
 * sgb200.h -- C ABI of libsgb200.so: hand-written sm_100a kernels for the StyleGAN2-ADA op hot
 * path (bias_act, upfirdn2d, conv2d / conv_transpose2d / weight-gradient, modulation helpers).
 *
 * This is the drop-in boundary.  Each entry point names the reference interface it replaces
 * (paths relative to /root/reference).  Rules, identical for every call:
 *   - plain pointers and sizes only; the CALLER owns every buffer (PyTorch allocates them),
 *     the library never allocates, frees or retains device memory beyond the call;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); no host sync;
 *   - the caller has made the right device current; the library is re-entrant;
 *   - return 0 on success; non-zero => sgb_last_error() (thread-local text).  Never aborts,
 *     never falls back to another implementation.
 * Strides are in ELEMENTS, logical order (n, c, h, w), so NCHW-contiguous and channels_last
 * tensors are both accepted exactly like the reference plugins accept them
 * (upfirdn2d.cpp:49-55, bias_act.cpp:47-51).
 */
#ifndef SGB200_H_
#define SGB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types (reference: AT_DISPATCH_FLOATING_TYPES_AND_HALF, bias_act.cpp:79; bf16 is new) */
enum { SGB_F32 = 0, SGB_F16 = 1, SGB_BF16 = 2, SGB_F64 = 3 };

/* activation ids = `cuda_idx` of bias_act.py:23-33 */
enum { SGB_ACT_LINEAR = 1, SGB_ACT_RELU = 2, SGB_ACT_LRELU = 3, SGB_ACT_TANH = 4, SGB_ACT_SIGMOID = 5,
       SGB_ACT_ELU = 6, SGB_ACT_SELU = 7, SGB_ACT_SOFTPLUS = 8, SGB_ACT_SWISH = 9 };

const char* sgb_last_error(void);
int sgb_abi_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t sgb_launch_count(void);

/* ---- bias_act ------------------------------------------------------------------------------
 * Replaces bias_act_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)
 * (bias_act.cpp:32-90, kernel bias_act.cu:23-147).  NULL pointer == the reference's empty tensor.
 *   grad 0: y = clamp(act(x + b) * gain)
 *   grad 1: x holds dy;  y = x * gain * act'(.)   masked to 0 where |yref| >= clamp
 *   grad 2: second derivative (tanh ... swish), `dy` is the extra multiplier
 * x / xref / yref / dy / y share one dense layout of `size_x` elements; the bias index of element i
 * is (i / step_b) % size_b  (step_b = x.stride(dim)), exactly bias_act.cu:44. clamp < 0 disables. */
int sgb_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                 int dtype, int grad, int act, float alpha, float gain, float clamp,
                 int64_t size_x, int64_t size_b, int64_t step_b, void* stream);

/* Reduction outputs below are ACCUMULATOR-typed: fp32, or fp64 when dtype == SGB_F64.
 * Sum of x viewed as [outer, size_c, inner] (dense) over outer and inner -> out[size_c] (overwritten).
 * Replaces the `dx.sum([...])` that produces db in bias_act.py:172-173. */
int sgb_sum_to_channel(const void* x, void* out, int dtype, int64_t outer, int64_t size_c, int64_t inner,
                       void* stream);

/* ---- upfirdn2d -----------------------------------------------------------------------------
 * Replaces upfirdn2d_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)
 * (upfirdn2d.cpp:16-94, kernels upfirdn2d.cu:29-200).  The caller computes
 *   out_w = (in_w*upx + padx0 + padx1 - fw + downx) / downx   (upfirdn2d.cpp:31-32) and allocates y.
 * f is fp32 [fh, fw] with element strides (f_sy, f_sx). */
int sgb_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                  int n, int c, int in_h, int in_w, const int64_t x_strides[4],
                  int out_h, int out_w, const int64_t y_strides[4],
                  int fh, int fw, int64_t f_sy, int64_t f_sx,
                  int upx, int upy, int downx, int downy, int padx0, int pady0,
                  int flip, float gain, void* stream);

/* The same operator for the case the hot path is dominated by: no resampling (up = down = 1), a filter of at most 4 x 4
 * taps that is an outer product fy (x) fx (upfirdn2d.setup_filter builds it that way, upfirdn2d.py:105-106), channels_last
 * tensors with 16-byte channel vectors.  fx / fy are HOST arrays of 4 taps in application order (filter flip and gain applied
 * by the caller, unused taps zero):  y[oy][ox] = sum_{ty,tx} fy[ty] fx[tx] x[oy - pady0 + ty][ox - padx0 + tx]. */
int sgb_upfirdn2d_sep(const void* x, const float* fx, const float* fy, void* y, int dtype,
                      int n, int c, int in_h, int in_w, const int64_t x_strides[4],
                      int out_h, int out_w, const int64_t y_strides[4], int padx0, int pady0, void* stream);

/* ---- convolution ---------------------------------------------------------------------------
 * Replaces the aten/cuDNN calls of conv2d_gradfix.py:112-114 (conv2d / conv_transpose2d forward, which
 * is also the data gradient of the other one, :125-128) and :143-145 (weight gradient), plus the
 * weight flip of conv2d_resample.py:35-36 (flip_weight) so no flipped copy is materialised. */
typedef struct {
  int32_t dtype;          /* SGB_* of x, w and y */
  int32_t transposed;     /* 0: conv2d (correlation), 1: conv_transpose2d */
  int32_t n, ci, co;      /* batch, input channels of x, output channels of y (totals over groups) */
  int32_t in_h, in_w, out_h, out_w;
  int32_t kh, kw, stride, pad_y, pad_x, groups;
  int32_t flip;           /* 1: use w[..., kh-1-ky, kw-1-kx] (true convolution) */
  int64_t x_strides[4];   /* elements, (n, c, h, w) */
  int64_t y_strides[4];
  /* weight is dense: conv2d [co, ci/groups, kh, kw]; conv_transpose2d [ci, co/groups, kh, kw] */
  /* optional fused prologue / epilogue (all may be NULL / 0): */
  /* in_scale / out_scale / noise are accumulator-typed (fp32; fp64 when dtype == SGB_F64) */
  const void* in_scale;   /* [n, ci]  x is multiplied by in_scale[n, c] while it is loaded (style modulation) */
  const void* out_scale;  /* [n, co]  accumulator * out_scale[n, o]            (demodulation)  */
  const void* noise;      /* [n, out_h, out_w] added after out_scale                         */
  const void*  bias;      /* [co] dtype of x; with act/gain/clamp below = a fused bias_act (grad 0) */
  int32_t act;            /* 0 = no fused bias_act, else SGB_ACT_* */
  float alpha, gain, clamp;
  /* math mode and scratch */
  int32_t strict_fp32;    /* SGB_F32 only: 1 = fp32 FFMA arithmetic (1e-4 class); 0 = TF32 tensor cores allowed
                             (what torch.backends.cudnn.allow_tf32 means for the reference's cuDNN convs) */
  int32_t force_simt;     /* A/B testing: 1 = never use the tensor-core kernels; 2 = tensor cores but not the halo-tile kernels */
  int32_t halo_gt;        /* tuning / test knob: tiles per super-tile of the halo-tile kernel (1, 2, 4); 0 = automatic */
  void*   workspace;      /* caller-owned scratch of >= sgb_conv2d_workspace_bytes(d) bytes, 16-byte aligned, or
                             NULL (then only the SIMT kernels are used).  Holds the re-packed weights. */
  int64_t workspace_bytes;
} sgb_conv_desc;

/* scratch bytes the tensor-core path needs for this descriptor (0 = it cannot be used) */
int64_t sgb_conv2d_workspace_bytes(const sgb_conv_desc* d);

int sgb_conv2d_forward(const sgb_conv_desc* d, const void* x, const void* w, void* y, void* stream);

/* dw (accumulator-typed, dense, same shape as the weight of a NON-transposed conv described by d, overwritten) =
 * sum over n, oy, ox of dy[n, o, oy, ox] * x[n, c, oy*stride + ky - pad, ...]; in_scale (if set) scales x.
 * d->x_strides describe x, d->y_strides describe dy. `flip` stores the result flipped. */
int sgb_conv2d_wgrad(const sgb_conv_desc* d, const void* x, const void* dy, void* dw, void* stream);

/* Which kernel sgb_conv2d_forward will run for this descriptor, for reporting only: 1 = the tcgen05 (tensor-core)
 * implicit-GEMM kernels, 3 = the tcgen05 kernel whose activation patches are staged by TMA (cp.async.bulk.tensor),
 * 2 = the small-output-channel 1x1 bandwidth kernel (toRGB), 0 = the generic SIMT kernel. */
int sgb_conv2d_uses_tensor_cores(const sgb_conv_desc* d);
/* same question for sgb_conv2d_wgrad: 0 = SIMT, non-zero = tensor cores; 2 = the halo-tile kernels, the only ones that
 * accept d->out_scale in a weight-gradient descriptor (there it is a per-sample scale [N, co] on dy: the style modulation
 * of a transposed convolution, whose input plays the role of dy). */
int sgb_conv2d_wgrad_uses_tensor_cores(const sgb_conv_desc* d);

/* ---- modulation helpers (activation-sized passes of modulated_conv2d, generators.py:80-87; fma.py) ----
 * y[n,c,h,w] = x[n,c,h,w] * s[n,c] (+ t[n,h,w] if t != NULL).  s, t accumulator-typed, dense. */
int sgb_scale_nc(const void* x, const void* s, const void* t, void* y, int dtype,
                 int n, int c, int h, int w, const int64_t x_strides[4], const int64_t y_strides[4], void* stream);
/* out[n,c] = sum_hw a*b   (accumulator-typed, overwritten)  -- d(styles), d(dcoefs) */
int sgb_mul_sum_hw(const void* a, const void* b, void* out, int dtype, int n, int c, int h, int w,
                   const int64_t a_strides[4], const int64_t b_strides[4], void* stream);
/* out[n,h,w] = sum_c a  (accumulator-typed, overwritten)  -- d(noise) */
int sgb_sum_c(const void* a, void* out, int dtype, int n, int c, int h, int w, const int64_t a_strides[4],
              void* stream);

/* In-place nan_to_num over a list of dense fp32 device tensors, 96 tensors per launch: the per-parameter
 * `param.grad.nan_to_num(nan=0, posinf=1e5, neginf=-1e5)` in front of the optimizer step (train_parts/trainers.py:745-748)
 * as 1-2 launches per module instead of one per parameter.  ptrs / numels are HOST arrays of length count. */
int sgb_nan_to_num_multi(void* const* ptrs, const int64_t* numels, int count, float nan, float posinf, float neginf,
                         void* stream);

/* ---- backward of the convolution's fused epilogue (sgb_conv_desc: out_scale / noise / bias / act / gain / clamp) ----
 * One pass over channels_last (dy, y) replacing bias_act backward + its bias sum (bias_act.py:161-175) and the fma
 * backward passes (fma.py:37-58):   dz = dy * gain * act'(y), masked where |y| >= clamp;
 *   dconv = dz * out_scale[n,c]  (dtype of y);   dbias[c] = sum dz;   dnoise[n,hw] = sum_c dz;
 *   dscale[n,c] = sum_hw dz * conv, conv re-derived from y.   bias / out_scale / noise / dbias / dnoise / dscale may be NULL;
 * y may be NULL for a linear epilogue without clamp and without dscale (nothing depends on it: the reference does not
 * keep y for linear either, bias_act.py:152-155, which is what lets callers add to the output in place);
 * the three reduction outputs are fp32 and are overwritten.  act: SGB_ACT_LINEAR or SGB_ACT_LRELU. */
int sgb_fused_epilogue_bwd(const void* dy, const void* y, void* dconv, const void* bias, const void* out_scale,
                           const void* noise, void* dbias, void* dnoise, void* dscale, int dtype, int n, int c, int hw,
                           int act, float alpha, float gain, float clamp, void* stream);

/* forward of that epilogue as one stand-alone channels_last pass (generators.py:83 fma + :328 bias_act):
 *   y = clamp(act(x * out_scale[n,c] + noise[n,hw] + bias[c]) * gain);  bias / out_scale / noise may be NULL */
int sgb_scale_bias_act(const void* x, const void* bias, const void* out_scale, const void* noise, void* y, int dtype,
                       int n, int c, int hw, int act, float alpha, float gain, float clamp, void* stream);

/* backward of the fused style modulation (sgb_conv_desc.in_scale) in one channels_last pass: gx = g * s[n,c] (dtype of g),
 * gs[n,c] = sum_hw g * x (fp32, overwritten).  gx or gs may be NULL.  Replaces the two passes over g of the reference's
 * `x * styles` backward (generators.py:80). */
int sgb_mod_bwd(const void* g, const void* x, const void* s, void* gx, void* gs, int dtype, int n, int c, int hw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGB200_H_ */
