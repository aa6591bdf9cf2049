// Instantiates the halo-tile convolution kernels (conv_halo.cuh) for one element type.
#include "conv_halo.cuh"

namespace sgb {
int conv_halo_f32(int bn, int mode, int gt, const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  return dispatch_halo<float, 2>(bn, mode, gt, d, x, w, y, s);
}
}  // namespace sgb
