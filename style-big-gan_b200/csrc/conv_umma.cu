// tcgen05 implicit-GEMM convolution (placeholder until the kernel lands: nothing is eligible yet).
#include "common.cuh"
namespace sgb {
bool conv_umma_eligible(const sgb_conv_desc*) { return false; }
int conv_forward_umma(const sgb_conv_desc*, const void*, const void*, void*, cudaStream_t) {
  set_error("conv_forward_umma: not built");
  return 1;
}
}  // namespace sgb
