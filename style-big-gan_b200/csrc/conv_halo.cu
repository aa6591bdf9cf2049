// Host-side dispatch of the halo-tile convolution kernel (conv_halo.cuh): geometry -> MODE, output-channel tile BN,
// tiles per super-tile GT.  The kernels themselves are instantiated per dtype in conv_halo_{f16,bf16,f32}.cu.
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

namespace sgb {

int conv_halo_f16(int bn, int mode, int gt, const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);
int conv_halo_bf16(int bn, int mode, int gt, const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);
int conv_halo_f32(int bn, int mode, int gt, const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s);

// geometry the halo kernel handles; 0 / 1 / 2 = MODE, -1 = not eligible
int conv_halo_mode(const sgb_conv_desc* d) {
  if (d->kh > 3 || d->kw > 3) return -1;
  if ((int64_t)d->n * d->x_strides[0] >= (int64_t)1 << 31) return -1;        // int32 offsets in the producer
  if (d->act != 0 && d->act != SGB_ACT_LINEAR && d->act != SGB_ACT_LRELU) return -1;
  if (d->stride == 1) {
    // geometry must be the "every tap stays inside the padded image" kind (always true for valid descriptors)
    if (!d->transposed) { if (d->out_h != d->in_h + 2 * d->pad_y - d->kh + 1 || d->out_w != d->in_w + 2 * d->pad_x - d->kw + 1) return -1; }
    else { if (d->out_h != d->in_h - 2 * d->pad_y + d->kh - 1 || d->out_w != d->in_w - 2 * d->pad_x + d->kw - 1) return -1;
           if (d->pad_y > d->kh - 1 || d->pad_x > d->kw - 1) return -1; }
    return 0;
  }
  if (d->stride == 2) {
    if (d->kh != 3 || d->kw != 3) return -1;
    return d->transposed ? 2 : 1;
  }
  return -1;
}
bool conv_halo_eligible(const sgb_conv_desc* d) { return conv_halo_mode(d) >= 0; }

// output-channel tile of the tensor-core kernels for this descriptor (also fixes the packed-weight layout)
int conv_bn(const sgb_conv_desc* d) {
  int bn = pick_bn(d->co);
  static const int m2bn = [] { const char* e = getenv("SGB_HALO_M2BN"); return e ? atoi(e) : 128; }();
  const int mode = (d->force_simt != 2) ? conv_halo_mode(d) : -1;
  if (mode == 2 && bn > m2bn) bn = m2bn;   // 4 accumulators must fit 512 TMEM columns
  // small images at small batch (the 4x4 ... 32x32 layers of a 4-image batch): a 256-channel output tile leaves a handful of
  // CTAs, each streaming megabytes of weights through one SM (measured: [4,512,9,9] stride 2 at 0.11 ms = 85 GB/s of weight
  // traffic).  Narrower output tiles spread the weight stream over more SMs; the patch re-reads this costs are tiny here.
  static const int par_bn = [] { const char* e = getenv("SGB_HALO_PARBN"); return e ? atoi(e) : 1; }();
  if (mode >= 0 && par_bn) {
    int64_t tiles;
    if (mode == 0) tiles = ceil_div((int64_t)d->n * (d->out_h + d->kh - 1), 16) * ceil_div(d->out_w, 8);
    else if (mode == 1) tiles = (int64_t)d->n * ceil_div(d->out_h, 16) * ceil_div(d->out_w, 8);
    else tiles = ceil_div((int64_t)d->n * (((d->out_h - 1 + d->pad_y) >> 1) + 2), 16) * ceil_div(((d->out_w - 1 + d->pad_x) >> 1) + 1, 8);
    while (bn > 32 && tiles * ceil_div(d->co, bn) < num_sms()) bn >>= 1;
  }
  return bn;
}

static bool gt_supported(int bn, int mode, int gt) {
  if (gt == 1) return true;
  if (mode == 1) return false;
  if (mode == 0) return gt * bn <= 512;
  return bn >= 32 && 4 * gt * bn <= 512;
}

// tiles per super-tile: more tiles = more MMAs per weight tile pulled from L2 (the bound at GT = 1), as long as TMEM
// and shared memory hold the accumulators / the wider patch and all SMs still get work
static int pick_gt(const sgb_conv_desc* d, int mode, int bn) {
  static const int env_gt = [] { const char* e = getenv("SGB_HALO_GT"); return e ? atoi(e) : 0; }();
  const int forced = d->halo_gt ? d->halo_gt : env_gt;
  const int nph = (mode == 2) ? 4 : 1;
  const int cols = (mode == 2) ? ((d->out_w - 1 + d->pad_x) >> 1) + 1 : d->out_w;      // width of the tile space
  const int rows = (mode == 0) ? d->out_h + d->kh - 1 : ((mode == 2) ? ((d->out_h - 1 + d->pad_y) >> 1) + 1 : d->out_h);
  const int64_t ntiles = (d->co + bn - 1) / bn;
  int gt = 4;
  if (forced == 1 || forced == 2 || forced == 4) gt = forced;
  while (gt > 1 && !gt_supported(bn, mode, gt)) gt >>= 1;
  if (forced) return gt;
  // keep TMEM double buffering (epilogue overlapped with the next tile's MMAs) unless K is long enough to amortise it
  static const int m2gt = [] { const char* e = getenv("SGB_HALO_M2GT"); return e ? atoi(e) : 0; }();
  if (mode == 2 && m2gt) { gt = m2gt; while (gt > 1 && !gt_supported(bn, mode, gt)) gt >>= 1; }
  else
  while (gt > 1 && nph * gt * bn * 2 > 512 && d->ci < 256) gt >>= 1;
  // do not pad narrow images, keep every SM busy
  const bool keep_cols = (mode == 2 && m2gt);      // experiment: accept the padded last column tile
  while (gt > 1 && ((!keep_cols && cols % (8 * gt) != 0) ||
                    ceil_div((int64_t)d->n * rows, 16) * ceil_div(cols, 8 * gt) * ntiles < num_sms())) gt >>= 1;
  return gt;
}

int conv_forward_halo(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  const int mode = conv_halo_mode(d);
  SGB_REQUIRE(mode >= 0, "geometry not supported");
  const int bn = conv_bn(d);
  const int gt = pick_gt(d, mode, bn);
  if (d->dtype == SGB_F16) return conv_halo_f16(bn, mode, gt, d, x, w, y, s);
  if (d->dtype == SGB_BF16) return conv_halo_bf16(bn, mode, gt, d, x, w, y, s);
  return conv_halo_f32(bn, mode, gt, d, x, w, y, s);
}

}  // namespace sgb
