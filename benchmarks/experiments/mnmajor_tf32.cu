// Experiment (not part of the library): kind::tf32 with MN-major fp32 operands.  The no-swizzle layout returns zeros; the
// layout type SWIZZLE_128B_BASE32B (descriptor layout type 1: 128-byte rows swizzled in 32-byte units over 4 K rows) is the
// one meant for 32-bit MN-major operands.  Checks it, and row-shifted start addresses (one pixel = one 128-byte K row).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mnmajor_tf32 mnmajor_tf32.cu && ./mnmajor_tf32 <shift> <base_off_mode> <variant>
// A[m=128][k=K] and B[k][n=64] are both MN-major: element (mn, k) of a 32-wide MN block lives at
//   blk*LBO + k*128 + (((mn%32)/8 ^ (k%4)) * 32) + (mn%8)*4      (absolute-address swizzle, base 1024-aligned; SBO = 512)
// B is read with its start address moved by `shift` rows.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) k(const float* Ag, const float* Bg, float* D, int K, int shift, int bmode, int variant, int KB_rows, int m64) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a_blk = K * 128;                      // bytes per 32-channel block of A (rows contiguous)
  const int b_blk = KB_rows * 128;
  uint8_t* sa = smem; uint8_t* sb = smem + 4 * a_blk;
  sb = (uint8_t*)(((uintptr_t)sb + 1023) & ~(uintptr_t)1023);
  auto swz = [&](int mn32, int r) {               // byte offset inside the 128-byte row r of a 32-channel block
    int q = mn32 / 8;
    if (variant == 0) q ^= (r & 3);
    if (variant == 2) q ^= ((r >> 1) & 3);
    return q * 32 + (mn32 % 8) * 4;
  };
  for (int i = tid; i < 128 * K; i += 128) {
    const int kk = i / 128, m = i % 128;
    *(float*)(sa + (m / 32) * a_blk + kk * 128 + swz(m % 32, kk)) = Ag[kk * 128 + m];
  }
  for (int i = tid; i < 64 * KB_rows; i += 128) {
    const int r = i / 64, n = i % 64;
    *(float*)(sb + (n / 32) * b_blk + r * 128 + swz(n % 32, r)) = Bg[r * 64 + n];
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
    // idesc: fp32 accum (1<<4), a/b = tf32 (2<<7, 2<<10), a_major/b_major = MN (1<<15, 1<<16), N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | (((m64 ? 64u : 128u) >> 4) << 24);
    for (int kk = 0; kk < K / 8; kk++) {
      const uint32_t a_start = smem_u32(sa) + kk * 8 * 128;
      const uint32_t b_start = smem_u32(sb) + (shift + kk * 8) * 128;
      auto desc = [&](uint32_t start, uint32_t lbo, uint32_t sbo) {
        uint64_t d = 0;
        d |= (uint64_t)((start >> 4) & 0x3FFF);
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        if (bmode == 1) d |= (uint64_t)((start >> 7) & 7) << 49;
        d |= (uint64_t)(variant == 3 ? 0 : 1) << 61;                        // SWIZZLE_128B_BASE32B
        return d;
      };
      const uint64_t da = desc(a_start, a_blk, 512), db = desc(b_start, b_blk, 512);
      const uint32_t acc = kk > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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
  const int shift = argc > 1 ? atoi(argv[1]) : 0, bmode = argc > 2 ? atoi(argv[2]) : 0, variant = argc > 3 ? atoi(argv[3]) : 0;
  const int K = 32, KB_rows = K + 16;
  std::vector<float> A(K * 128), B(KB_rows * 64);
  std::vector<float> Af(K * 128), Bf(KB_rows * 64);
  srand(1);
  for (size_t i = 0; i < A.size(); i++) { float v = (rand() % 17 - 8) / 8.f; A[i] = v; Af[i] = v; }
  for (size_t i = 0; i < B.size(); i++) { float v = (rand() % 17 - 8) / 8.f; B[i] = v; Bf[i] = v; }
  float *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int m64 = argc > 4 ? atoi(argv[4]) : 0;
  cudaMemset(dD, 0xff, 128 * 64 * 4);
  k<<<1, 128, 100 * 1024>>>(dA, dB, dD, K, shift, bmode, variant, KB_rows, m64);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> D(128 * 64);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  if (m64) {      // M = 64: which TMEM lane holds which row of D?
    for (int m = 0; m < 64; m++) {
      int found = -1;
      for (int l = 0; l < 128 && found < 0; l++) {
        bool ok = true;
        for (int n = 0; n < 64 && ok; n++) {
          double r = 0; for (int kk = 0; kk < K; kk++) r += (double)Af[kk * 128 + m] * Bf[(kk + shift) * 64 + n];
          ok = fabs(r - D[l * 64 + n]) < 1e-3;
        }
        if (ok) found = l;
      }
      printf("row %d -> lane %d\n", m, found);
    }
    return 0;
  }
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; m++) for (int n = 0; n < 64; n++) {
    double r = 0; for (int kk = 0; kk < K; kk++) r += (double)Af[kk * 128 + m] * Bf[(kk + shift) * 64 + n];
    maxerr = fmax(maxerr, fabs(r - D[m * 64 + n])); maxref = fmax(maxref, fabs(r));
  }
  printf("shift=%d base_off_mode=%d variant=%d: max err %.4g (max ref %.4g) %s\n", shift, bmode, variant, maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
  return 0;
}
