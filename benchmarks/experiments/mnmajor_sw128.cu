// Experiment (not part of the library): does tcgen05.mma take MN-major bf16 operands in the 128-byte-swizzled layout, and
// how do row-shifted start addresses (the halo trick: one pixel = one 128-byte K row) interact with the swizzle?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mnmajor_sw128 mnmajor_sw128.cu && ./mnmajor_sw128 <shift> <base_off_mode> <sbo_rows>
// A[m=128][k=K] and B[k][n=64] are both MN-major: element (mn, k) of a 64-wide MN block lives at
//   blk*LBO + (k/8)*SBO + (k%8)*128 + ((mn%64)/8 ^ (k%8))*16 + (mn%8)*2      (absolute-address swizzle, base 1024-aligned)
// B is read with its start address moved by `shift` rows; base_off_mode 1 sets the descriptor's base_offset to (start>>7)&7.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* Ag, const __nv_bfloat16* Bg, float* D, int K, int shift, int bmode, int sbo_rows, int KB_rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int SBO = sbo_rows * 128;                 // bytes between 8-row K blocks
  const int a_blk = (K / 8) * SBO;                // bytes per 64-channel block of A
  uint8_t* sa = smem; uint8_t* sb = smem + 2 * a_blk + 1024;
  sb = (uint8_t*)(((uintptr_t)sb + 1023) & ~(uintptr_t)1023);
  // A: logical rows k = 0..K-1 packed in 8-row blocks at SBO
  for (int i = tid; i < 128 * K; i += 128) {
    const int kk = i / 128, m = i % 128;
    const int addr = (m / 64) * a_blk + (kk / 8) * SBO + (kk % 8) * 128 + ((((m % 64) / 8) ^ (kk % 8)) * 16) + (m % 8) * 2;
    *(__nv_bfloat16*)(sa + addr) = Ag[kk * 128 + m];
  }
  // B: physical rows r = 0..KB_rows-1 stored CONTIGUOUSLY (128 B per row, swizzle by absolute row); logical k reads row k + shift
  for (int i = tid; i < 64 * KB_rows; i += 128) {
    const int r = i / 64, n = i % 64;
    const int addr = r * 128 + (((n / 8) ^ (r % 8)) * 16) + (n % 8) * 2;
    *(__nv_bfloat16*)(sb + addr) = Bg[r * 64 + n];
  }
  if (warp == 0) {
    if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tslot;
  if (tid == 0) {
    // idesc: fp32 accum (1<<4), a/b = bf16 (1<<7, 1<<10), a_major/b_major = MN (1<<15, 1<<16), N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int kk = 0; kk < K / 16; kk++) {
      const uint32_t a_start = smem_u32(sa) + kk * 2 * SBO;
      const uint32_t b_start = smem_u32(sb) + (shift + kk * 16) * 128;      // B rows are contiguous: SBO_b = 1024
      auto desc = [&](uint32_t start, uint32_t lbo, uint32_t sbo) {
        uint64_t d = 0;
        d |= (uint64_t)((start >> 4) & 0x3FFF);
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        if (bmode == 1) d |= (uint64_t)((start >> 7) & 7) << 49;
        d |= (uint64_t)2 << 61;                                             // SWIZZLE_128B
        return d;
      };
      const uint64_t da = desc(a_start, a_blk, SBO), db = desc(b_start, 0, 1024);
      const uint32_t acc = kk > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  // wait
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DN;\n\tbra W;\n\tDN:\n\t}" :: "r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 16; e++) D[tid * 64 + c + e] = __uint_as_float(v[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem));
}

int main(int argc, char** argv) {
  const int shift = argc > 1 ? atoi(argv[1]) : 0, bmode = argc > 2 ? atoi(argv[2]) : 0, sbo_rows = argc > 3 ? atoi(argv[3]) : 8;
  const int K = 32, KB_rows = K + 16;
  std::vector<__nv_bfloat16> A(K * 128), B(KB_rows * 64);
  std::vector<float> Af(K * 128), Bf(KB_rows * 64);
  srand(1);
  for (size_t i = 0; i < A.size(); i++) { float v = (rand() % 17 - 8) / 8.f; A[i] = __float2bfloat16(v); Af[i] = v; }
  for (size_t i = 0; i < B.size(); i++) { float v = (rand() % 17 - 8) / 8.f; B[i] = __float2bfloat16(v); Bf[i] = v; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<1, 128, 100 * 1024>>>(dA, dB, dD, K, shift, bmode, sbo_rows, KB_rows);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> D(128 * 64);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; m++) for (int n = 0; n < 64; n++) {
    double r = 0; for (int kk = 0; kk < K; kk++) r += (double)Af[kk * 128 + m] * Bf[(kk + shift) * 64 + n];
    maxerr = fmax(maxerr, fabs(r - D[m * 64 + n])); maxref = fmax(maxref, fabs(r));
  }
  printf("shift=%d base_off_mode=%d sbo_rows=%d: max err %.4g (max ref %.4g) %s\n", shift, bmode, sbo_rows, maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
  return 0;
}
