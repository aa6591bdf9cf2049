// C-ABI entry points for convolution: validation + routing between the tcgen05 implicit-GEMM kernels (conv_tma.cuh: TMA-staged
// stride 1; conv_halo.cuh: every other geometry up to 3 x 3; conv_umma.cu: per-tap fallback; conv_wgrad_*.cu: weight gradient),
// the small-output 1 x 1 bandwidth kernel (conv_small.cu) and the shape-complete SIMT kernels (conv_simt.cu).
#include "common.cuh"

namespace sgb {
int conv_forward_simt(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);
int conv_wgrad_simt(const sgb_conv_desc* d, const void* x, const void* dy, void* dw, cudaStream_t s);
bool conv_umma_eligible(const sgb_conv_desc* d);
int conv_forward_umma(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);
bool conv_wgrad_umma_eligible(const sgb_conv_desc* d);
bool conv_halo_eligible(const sgb_conv_desc* d);
int conv_forward_halo(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);
int conv_wgrad_umma(const sgb_conv_desc* d, const void* x, const void* dy, void* dw, cudaStream_t s);
bool conv_wgrad_halo_eligible(const sgb_conv_desc* d);
int conv_wgrad_halo(const sgb_conv_desc* d, const void* x, const void* dy, void* dw, cudaStream_t s);
bool conv_tma_eligible(const sgb_conv_desc* d);
int conv_forward_tma(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);
bool conv_small_eligible(const sgb_conv_desc* d);
int conv_forward_small(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);

static int validate(const sgb_conv_desc* d) {
  SGB_REQUIRE(d != nullptr, "descriptor is NULL");
  SGB_REQUIRE(d->dtype >= SGB_F32 && d->dtype <= SGB_F64, "unsupported dtype");
  SGB_REQUIRE(d->n >= 0 && d->ci >= 1 && d->co >= 1, "bad channel / batch counts");
  SGB_REQUIRE(d->groups >= 1 && d->ci % d->groups == 0 && d->co % d->groups == 0, "channels must divide into groups");
  SGB_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->stride >= 1, "bad kernel size / stride");
  SGB_REQUIRE(d->pad_y >= 0 && d->pad_x >= 0, "padding must be non-negative");
  SGB_REQUIRE(d->in_h >= 1 && d->in_w >= 1 && d->out_h >= 1 && d->out_w >= 1, "bad spatial size");
  if (!d->transposed) {
    SGB_REQUIRE((d->out_h - 1) * d->stride + d->kh - 2 * d->pad_y <= d->in_h &&
                (d->out_w - 1) * d->stride + d->kw - 2 * d->pad_x <= d->in_w, "output larger than the input allows");
  } else {
    SGB_REQUIRE(d->out_h >= (d->in_h - 1) * d->stride + d->kh - 2 * d->pad_y &&
                d->out_w >= (d->in_w - 1) * d->stride + d->kw - 2 * d->pad_x, "transposed output smaller than the input needs");
  }
  return 0;
}
}  // namespace sgb

using namespace sgb;

// 2 = halo-tile tensor-core kernels (16-bit: conv_wgrad_halo_kernel; fp32: conv_wgrad_tf32_kernel on kind::tf32 with the
// SWIZZLE_128B_BASE32B operand layout, or the 3 x bf16 split with SGB_WGRAD_FP32=split), 1 = per-tap tensor-core kernel
// (16-bit types, geometries the halo kernels do not take), 0 = SIMT
static int wgrad_route(const sgb_conv_desc* d) {
  if (!conv_wgrad_umma_eligible(d)) return 0;
  if (conv_wgrad_halo_eligible(d)) return 2;
  return d->dtype == SGB_F32 ? 0 : 1;
}

// 0 = SIMT; non-zero = tensor cores, 2 = the halo-tile kernels (the ones that also take out_scale = a scale on dy)
extern "C" int sgb_conv2d_wgrad_uses_tensor_cores(const sgb_conv_desc* d) {
  return (d && !d->transposed) ? wgrad_route(d) : 0;
}

// 1 = tcgen05 kernels (cp.async-staged patches), 3 = tcgen05 kernel with TMA-staged patches (conv_tma.cuh), 2 = the small-co
// 1x1 bandwidth kernel (conv_small.cu), 0 = generic SIMT kernel
extern "C" int sgb_conv2d_uses_tensor_cores(const sgb_conv_desc* d) {
  if (!d) return 0;
  if (conv_small_eligible(d)) return 2;
  if (!conv_umma_eligible(d)) return 0;
  return (d->force_simt != 2 && conv_tma_eligible(d)) ? 3 : 1;
}

extern "C" int sgb_conv2d_forward(const sgb_conv_desc* d, const void* x, const void* w, void* y, void* stream) {
  if (int r = validate(d)) return r;
  if ((int64_t)d->n * d->out_h * d->out_w == 0) return 0;
  SGB_REQUIRE(x && w && y, "x, w and y must not be NULL");
  cudaStream_t s = (cudaStream_t)stream;
  if (conv_small_eligible(d)) return conv_forward_small(d, x, w, y, s);
  if (conv_umma_eligible(d)) {
    if (d->force_simt != 2 && conv_tma_eligible(d)) return conv_forward_tma(d, x, w, y, s);
    if (d->force_simt != 2 && conv_halo_eligible(d)) return conv_forward_halo(d, x, w, y, s);
    return conv_forward_umma(d, x, w, y, s);
  }
  return conv_forward_simt(d, x, w, y, s);
}

extern "C" int sgb_conv2d_wgrad(const sgb_conv_desc* d, const void* x, const void* dy, void* dw, void* stream) {
  if (int r = validate(d)) return r;
  SGB_REQUIRE(!d->transposed, "wgrad takes the non-transposed description (swap x and dy for conv_transpose2d)");
  SGB_REQUIRE(dw, "dw must not be NULL");
  SGB_REQUIRE((x && dy) || (int64_t)d->n == 0, "x and dy must not be NULL");
  SGB_REQUIRE(!d->out_scale || wgrad_route(d) == 2, "wgrad: out_scale (a scale on dy) is only taken by the halo-tile kernels");
  switch (wgrad_route(d)) {
    case 2: return conv_wgrad_halo(d, x, dy, dw, (cudaStream_t)stream);
    case 1: return conv_wgrad_umma(d, x, dy, dw, (cudaStream_t)stream);
    default: break;
  }
  return conv_wgrad_simt(d, x, dy, dw, (cudaStream_t)stream);
}
