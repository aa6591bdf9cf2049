// Host side of the TMA-staged convolution kernel (conv_tma.cuh): eligibility, (BN, KB, GT) choice, the driver entry point
// for tensor-map encoding, and the fp16 / fp32 instantiations.
#include <cstdlib>
#include "conv_tma.cuh"

namespace sgb {

PFN_encodeTiled tma_encode_fn() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (PFN_encodeTiled)p;
  }();
  return fn;
}

// which convolutions take the TMA kernel: stride 1 (conv2d and conv_transpose2d), kernels up to 3x3, fp16 / fp32-as-TF32,
// images of at least 16 rows (smaller ones: conv_halo_kernel's tiles straddle images, per-image tiles would be mostly padding),
// a channel count whose bytes are a multiple of 16 (tensor-map strides); fused epilogue: linear / lrelu.
// SGB_TMA: 1 (default) = on, 0 = off (A/B)
static bool tma_geometry_ok(const sgb_conv_desc* d) {
  static const int on = [] { const char* e = getenv("SGB_TMA"); return e ? atoi(e) : 1; }();
  if (!on) return false;
  if (d->dtype != SGB_F16 && d->dtype != SGB_F32) return false;
  if ((d->stride != 1 && d->stride != 2) || d->kh > 3 || d->kw > 3 || d->groups != 1) return false;
  // stride 2 (the 3 x 3 convolution of the down path / the data gradient of the up path): built, exact, and OFF by default
  // (SGB_TMA_S2=1 switches it on) -- measured level with conv_halo_kernel's MODE 1 on [8,64,257,257] -> 128 (90.7 vs 90.1 us) and
  // 5-20 % behind it on the small 512-channel layers; ffhq256 step 70.6 vs 70.1 ms, config-f 53.4 vs 52.1 (r2_run29.sh)
  static const int s2_on = [] { const char* e = getenv("SGB_TMA_S2"); return e ? atoi(e) : 0; }();
  if (d->stride == 2 && (!s2_on || d->transposed || d->kh != 3 || d->kw != 3)) return false;
  if (d->act != 0 && d->act != SGB_ACT_LINEAR && d->act != SGB_ACT_LRELU) return false;
  if (d->bias && d->act == 0) return false;
  static const int force_g = [] { const char* e = getenv("SGB_TMA_FORCE"); return e ? atoi(e) : 0; }();
  // images of at least 16 x 16: one 16-row tile per image column strip without padding rows.  Measured on the 512-channel
  // layers (profiles/README.md): 16^2 and 32^2 images run 12-25 % faster here than on conv_halo_kernel at both 4 and 32 images,
  // 8^2 images (half of every tile is padding) 25 % slower
  if (d->stride == 2) {
    if ((d->out_h < 8 || d->out_w < 8) && !force_g) return false;
    if (d->out_h != (d->in_h + 2 * d->pad_y - d->kh) / 2 + 1 || d->out_w != (d->in_w + 2 * d->pad_x - d->kw) / 2 + 1) return false;
  } else
  if ((d->out_h < 16 || d->out_w < 16) && !force_g) return false;
  if (d->stride == 2) {}
  else if (!d->transposed) { if (d->out_h != d->in_h + 2 * d->pad_y - d->kh + 1 || d->out_w != d->in_w + 2 * d->pad_x - d->kw + 1) return false; }
  else { if (d->out_h != d->in_h - 2 * d->pad_y + d->kh - 1 || d->out_w != d->in_w - 2 * d->pad_x + d->kw - 1) return false;
         if (d->pad_y > d->kh - 1 || d->pad_x > d->kw - 1) return false; }
  const int es = d->dtype == SGB_F16 ? 2 : 4;
  if ((d->ci * es) % 16 != 0) return false;
  // tensor-map limits: strides < 2^40 bytes, multiples of 16 bytes; dims < 2^32
  for (int i : {0, 2, 3}) if ((d->x_strides[i] * es) % 16 != 0 || d->x_strides[i] * es >= ((int64_t)1 << 40)) return false;
  if (d->x_strides[1] != 1 || d->y_strides[1] != 1) return false;
  return true;
}

// (BN, KB, GT) for this descriptor, or false when the kernel has no instantiation for it
static bool tma_pick(const sgb_conv_desc* d, int& bn, int& kb, int& gt) {
  bn = conv_bn(d);
  const int es = d->dtype == SGB_F16 ? 2 : 4;
  static const int force = [] { const char* e = getenv("SGB_TMA_FORCE"); return e ? atoi(e) : 0; }();
  if (d->stride == 2) {
    // 64-byte K blocks (a patch is 33 rows x 2 planes x 9 columns: 38 KB per stage), one tile per stage, <= 128 output channels
    // per tile (the packed weights are laid out for the tile width chosen here)
    if (bn > 128) bn = 128;
    if (bn < 32) return false;
    kb = 64; gt = 1;
    const int64_t tiles2 = (int64_t)d->n * ((d->out_h + 15) / 16) * ((d->out_w + 7) / 8) * ((d->co + bn - 1) / bn);
    return tiles2 * 4 >= num_sms() || force;
  }
  // measured (benchmarks/experiments/tma_check.py): ahead of conv_halo_kernel up to 128 output channels, behind it at 256
  // (GT = 2 against the halo kernel's wider super-tiles: the L2 weight stream per MMA doubles)
  if (bn < 32 || bn > 128) return false;
  // K block = one swizzle row per pixel: 64 bytes (SWIZZLE_64B) when that already holds all channels, else 128 bytes; the
  // super-tile width keeps a patch stage near 40 KB either way (18 x 34 x 64 B, 18 x 18 x 128 B)
  kb = (d->ci * es <= 64) ? 64 : 128;
  gt = (kb == 64) ? 4 : 2;
  if (kb == 64 && bn > 64) return false;
  // enough tiles (SGB_TMA_FORCE=1: take small problems too -- tests)
  const int64_t tiles = (int64_t)d->n * ((d->out_h + 15) / 16) * ((d->out_w + 8 * gt - 1) / (8 * gt)) * ((d->co + bn - 1) / bn);
  // a quarter of the SMs must get a tile (it was "every SM": the small-batch 512-channel layers, 64-128 tiles, are latency-bound
  // on either kernel and this one has the shorter per-K-block critical path)
  if (tiles * 4 < num_sms() && !force) return false;
  return true;
}

bool conv_tma_eligible(const sgb_conv_desc* d) {
  int bn, kb, gt;
  return tma_geometry_ok(d) && tma_pick(d, bn, kb, gt) && tma_encode_fn() != nullptr;
}

int conv_forward_tma(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  int bn, kb, gt;
  SGB_REQUIRE(tma_geometry_ok(d) && tma_pick(d, bn, kb, gt), "geometry not supported by the TMA kernel");
  if (d->dtype == SGB_F16) return dispatch_tma<__half, 0>(bn, kb, gt, d, x, w, y, s);
  return dispatch_tma<float, 2>(bn, kb, gt, d, x, w, y, s);
}

}  // namespace sgb
