// tcgen05 implicit-GEMM convolution for sm_100a (channels_last activations).
//
//   D[m, o] = sum_{tap, c} A[m, (tap, c)] * W[o, (tap, c)]      m = output pixel (n, oy, ox), 128 per CTA
//
// Warp roles in a 192-thread CTA (one CTA per 128-pixel x BN-channel output tile):
//   warps 0-3  producers: thread r owns row r of the A tile.  For every K block (one filter tap x BK channels,
//              128 bytes) it gathers its pixel's channel vector from global memory with 8 LDG.128 (zero when the
//              tap falls into padding), optionally multiplies by the per-sample style in_scale[n, c] and rounds
//              to the MMA operand type, and writes it to shared memory in the canonical K-major UMMA layout
//              (8-row x 16-byte core matrices, no swizzle: chunk j of row r sits at j*2048 + r*16).  Afterwards the
//              same four warps run the epilogue: tcgen05.ld of their 32 TMEM lanes, demodulation scale, noise,
//              bias + activation + clamp, 128-bit stores.
//   warp 4     MMA issuer: one elected lane issues tcgen05.mma.cta_group::1 (M=128, N=BN, K=32 bytes) from
//              shared-memory descriptors into a TMEM accumulator, tcgen05.commit releases the stage.
//   warp 5     weight loader: the weights were re-packed (pack kernel below: flip, transpose, type conversion,
//              zero padding) into the exact shared-memory image of each B tile, so one cp.async.bulk (TMA bulk
//              copy, complete_tx on the stage's mbarrier) per K block brings BN x 128 bytes.
// Synchronisation: full[s] (128 producer arrivals + 1 expect_tx arrival), empty[s] (tcgen05.commit),
// accum (tcgen05.commit after the last K block).
//
// Supported: f16 / bf16 (kind::f16) and f32 storage with TF32 math (kind::tf32), fp32 accumulation; conv2d and
// conv_transpose2d, any kernel size / stride / padding, groups == 1.
#include "common.cuh"
#include "act.cuh"
#include "umma.cuh"

namespace sgb {

constexpr int KB_BYTES = 128;    // bytes of K per block per row (8 chunks of 16 B)
constexpr int A_STAGE_BYTES = UM * KB_BYTES;   // 16 KB
constexpr int NUM_PRODUCERS = 128;

struct UmmaParams {
  sgb_conv_desc d;
  const void* x; const void* wpack; void* y;
  int64_t M;            // n * out_h * out_w
  int ohw, taps, cblocks, ntiles;
  int tc;               // elements per 16-byte chunk
  int vec_store;        // co allows 16-byte stores
};

// ---- weight re-pack ------------------------------------------------------------------------------------
// out image per (ntile, tap, cblock): [chunk j (8)][row r (BN)][16 bytes]; zero outside (co, ci).
// One thread = one (ntile, cblock, chunk j, row r) and ALL taps: the taps of a (output, input) channel pair are adjacent in the
// parameter ([O, I, kh, kw], or [I, O, kh, kw] for conv_transpose2d), so a thread reads tc short contiguous runs and writes one
// 16-byte chunk per tap; consecutive threads (r) write consecutive chunks.  (The first version handled one ELEMENT per thread
// with six 64-bit divisions each and stride-9 gathers: 22 us for a 512 x 512 x 3 x 3 fp32 weight, a quarter of the time of
// the small layers' convolutions, profiles/README.md.)
template <class T, int KIND, int TC>
__global__ void __launch_bounds__(256) pack_weights_kernel(const T* __restrict__ w, void* __restrict__ out, sgb_conv_desc d,
                                                           int bn, int ntiles, int cblocks) {
  const int taps = d.kh * d.kw;
  const int total = ntiles * cblocks * 8 * bn;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx % bn;
    int t = idx / bn;
    const int j = t & 7; t >>= 3;
    const int cb = t % cblocks, nt = t / cblocks;
    const int o = nt * bn + r, c0 = (cb * 8 + j) * TC;
    const bool o_ok = o < d.co;
    // element (o, c, tap) of the parameter: base + c * cstep + tap
    const int64_t base = d.transposed ? (int64_t)o * taps : (int64_t)o * d.ci * taps;
    const int64_t cstep = d.transposed ? (int64_t)d.co * taps : (int64_t)taps;
    uint4* dst = (uint4*)out + (((int64_t)nt * taps * cblocks + cb) * 8 + j) * bn + r;
    const int64_t tap_stride = (int64_t)cblocks * 8 * bn;            // 16-byte chunks between taps
    for (int tap = 0; tap < taps; tap++) {
      const int st = d.flip ? taps - 1 - tap : tap;                    // flipping both axes reverses the linear tap index
      Vec16<T> v;
#pragma unroll
      for (int e = 0; e < TC; e++) {
        const int c = c0 + e;
        const float f = (o_ok && c < d.ci) ? (float)to_acc<T>(w[base + (int64_t)c * cstep + st]) : 0.f;
        if (KIND == 2) ((uint32_t*)&v.raw)[e] = f32_to_tf32(f);
        else v.v[e] = from_acc<T>(f);
      }
      dst[tap * tap_stride] = v.raw;
    }
  }
}

// ---- main kernel ---------------------------------------------------------------------------------------
template <class T, int KIND, int BN, int STAGES>
__global__ void __launch_bounds__(192, 1) conv_umma_kernel(UmmaParams p) {
  constexpr int TC = 16 / sizeof(T);                 // elements per 16-byte chunk
  constexpr int BK = 8 * TC;                         // channels per K block
  constexpr int B_STAGE_BYTES = BN * KB_BYTES;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = make_idesc(KIND, BN);

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const sgb_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * UM;
  const int ntile = blockIdx.y;
  const int num_kblocks = p.taps * p.cblocks;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; s++) { mbar_init(smem_u32(&full_bar[s]), NUM_PRODUCERS + 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
      mbar_init(smem_u32(&accum_bar), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 4) {
    // =========================== producers ===========================
    const int r = threadIdx.x;                       // row of the tile
    const int64_t m = m0 + r;
    const bool row_ok = m < p.M;
    int n = 0, oy = 0, ox = 0;
    if (row_ok) { n = (int)(m / p.ohw); const int rem = (int)(m - (int64_t)n * p.ohw); oy = rem / d.out_w; ox = rem - oy * d.out_w; }
    const T* xn = (const T*)p.x + (int64_t)n * d.x_strides[0];
    const float* sc = d.in_scale ? (const float*)d.in_scale + (int64_t)n * d.ci : nullptr;
    uint8_t* a_row = smem + r * 16;
    int kb = 0;
    for (int tap = 0; tap < p.taps; tap++) {
      const int ky = tap / d.kw, kx = tap - ky * d.kw;
      int iy, ix; bool ok = row_ok;
      if (!d.transposed) {
        iy = oy * d.stride + ky - d.pad_y; ix = ox * d.stride + kx - d.pad_x;
        ok = ok && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w;
      } else {
        const int ty = oy + d.pad_y - ky, tx = ox + d.pad_x - kx;
        ok = ok && ty >= 0 && tx >= 0 && (ty % d.stride == 0) && (tx % d.stride == 0);
        iy = ty / d.stride; ix = tx / d.stride;
        ok = ok && iy < d.in_h && ix < d.in_w;
      }
      const T* src = xn + (int64_t)iy * d.x_strides[2] + (int64_t)ix * d.x_strides[3];
      for (int cb = 0; cb < p.cblocks; cb++, kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        const int c0 = cb * BK;
        // issue the global loads before waiting for the stage: latency overlaps the wait
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int c = c0 + j * TC;
          v[j] = (ok && c < d.ci) ? __ldg((const uint4*)(src + c)) : make_uint4(0, 0, 0, 0);
        }
        if (sc) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const int c = c0 + j * TC;
            if (ok && c < d.ci) {
              if (KIND == 2) {
                const float4 s4 = __ldg((const float4*)(sc + c));
                float* f = (float*)&v[j];
                f[0] *= s4.x; f[1] *= s4.y; f[2] *= s4.z; f[3] *= s4.w;
              } else {
                const float4 sa = __ldg((const float4*)(sc + c)), sb = __ldg((const float4*)(sc + c + 4));
                const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                T* h = (T*)&v[j];
#pragma unroll
                for (int e = 0; e < 8; e++) h[e] = from_acc<T>(to_acc<T>(h[e]) * sv[e]);
              }
            }
          }
        }
        if (KIND == 2) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            float* f = (float*)&v[j]; uint32_t* u = (uint32_t*)&v[j];
            u[0] = f32_to_tf32(f[0]); u[1] = f32_to_tf32(f[1]); u[2] = f32_to_tf32(f[2]); u[3] = f32_to_tf32(f[3]);
          }
        }
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        uint8_t* dst = a_row + s * STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < 8; j++) *(uint4*)(dst + j * (UM * 16)) = v[j];
        fence_proxy_async();
        mbar_arrive(smem_u32(&full_bar[s]));
      }
    }

    // =========================== epilogue ===========================
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float* out_scale = d.out_scale ? (const float*)d.out_scale + (int64_t)n * d.co : nullptr;
    const float nz = (d.noise && row_ok) ? ((const float*)d.noise)[((int64_t)n * d.out_h + oy) * d.out_w + ox] : 0.f;
    const float alpha = d.alpha, gain = d.gain, clamp = d.clamp;
    T* yrow = (T*)p.y + (int64_t)n * d.y_strides[0] + (int64_t)oy * d.y_strides[2] + (int64_t)ox * d.y_strides[3];
    const int o_base = ntile * BN;
#pragma unroll 1
    for (int cc = 0; cc < BN; cc += 16) {
      uint32_t acc[16];
      tmem_ld16(lane_addr + cc, acc);          // warp-collective: every lane participates, stores are predicated
      if (!row_ok) continue;
      float val[16];
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const int o = o_base + cc + e;
        float a = __uint_as_float(acc[e]);
        if (o < d.co) {
          if (out_scale) a *= out_scale[o];
          a += nz;
          if (d.act) {
            if (d.bias) a += to_acc<T>(((const T*)d.bias)[o]);
            a = act_forward<float>(d.act, a, alpha, gain, clamp);
          }
        }
        val[e] = a;
      }
      if (p.vec_store) {
#pragma unroll
        for (int g = 0; g < 16 / TC; g++) {
          const int o = o_base + cc + g * TC;
          if (o < d.co) {
            Vec16<T> pk;
#pragma unroll
            for (int e = 0; e < TC; e++) pk.v[e] = from_acc<T>(val[g * TC + e]);
            *(uint4*)(yrow + o) = pk.raw;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const int o = o_base + cc + e;
          if (o < d.co) yrow[(int64_t)o * d.y_strides[1]] = from_acc<T>(val[e]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      for (int kb = 0; kb < num_kblocks; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {          // 4 x (2 chunks of 16 bytes) = 128 bytes of K
          const uint64_t adesc = make_smem_desc(a_addr + kk * 2 * (UM * 16), UM * 16, 128);
          const uint64_t bdesc = make_smem_desc(b_addr + kk * 2 * (BN * 16), BN * 16, 128);
          umma<KIND>(tmem_base, adesc, bdesc, IDESC, (kb > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));     // frees the stage once these MMAs have read it
      }
      umma_commit(smem_u32(&accum_bar));          // accumulator complete
    }
    __syncwarp();
  } else {
    // =========================== weight loader ===========================
    if (lane == 0) {
      const uint8_t* wsrc = (const uint8_t*)p.wpack + (int64_t)ntile * num_kblocks * B_STAGE_BYTES;
      for (int kb = 0; kb < num_kblocks; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        const uint32_t bar = smem_u32(&full_bar[s]);
        mbar_arrive_expect_tx(bar, B_STAGE_BYTES);
        bulk_copy_g2s(smem_u32(smem + s * STAGE_BYTES + A_STAGE_BYTES), wsrc + (int64_t)kb * B_STAGE_BYTES, B_STAGE_BYTES, bar);
      }
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ---- weight gradient ---------------------------------------------------------------------------------------
//   dW[o, c, tap] = sum_p dy[p, o] * xs[p_in(p, tap), c]        p = output pixel (n, oy, ox)
// One CTA owns (128 output channels) x (BNC input channels) x (one filter tap) x (one slice of the pixels = split-K).
// Both operands are [pixel][channel] in memory with the channel contiguous, i.e. MN-major for a GEMM whose K is the
// pixel index: the 128 producer threads copy 16-byte channel chunks of PPB pixels per K block into the canonical
// MN-major UMMA layout ([chunk][pixel][16 B]: SBO = PPB*16 between channel chunks, LBO = 128 between groups of 8 pixels),
// the x operand gathered at the tap's offset (zero in the padding) and multiplied by in_scale when set.
// The fp32 accumulator tile is added to dW with red.global.add.f32 (dW zeroed by the launcher).
struct WgradParams {
  sgb_conv_desc d;
  const void* x; const void* dy; float* dw;
  int64_t P;            // n * out_h * out_w
  int ohw, taps, ctiles, otiles, splits;
  int64_t chunk;        // pixels per split (multiple of PPB)
};

template <class T, int KIND, int BNC, int STAGES>
__global__ void __launch_bounds__(160, 1) conv_wgrad_umma_kernel(WgradParams p) {
  constexpr int TC = 16 / sizeof(T);
  constexpr int PPB = 8 * TC;                        // pixels per K block: 64 (16-bit) / 32 (tf32) -> 4 MMAs
  constexpr int SPLIT = 128 / PPB;                   // producer thread groups per pixel
  constexpr int A_CHUNKS = UM / TC;                  // 16-byte chunks per pixel of the dy tile
  constexpr int B_CHUNKS = BNC / TC;
  constexpr int A_PER_T = A_CHUNKS / SPLIT, B_PER_T = B_CHUNKS / SPLIT;
  constexpr int A_BYTES = UM * PPB * (int)sizeof(T); // 16 KB
  constexpr int B_BYTES = BNC * PPB * (int)sizeof(T);
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BNC < 32 ? 32 : BNC;
  constexpr uint32_t IDESC = make_idesc(KIND, BNC, 1);
  constexpr int K_PER_MMA = 32 / (int)sizeof(T);     // 16 / 8 pixels

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const sgb_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int otile = blockIdx.x;
  const int tap = blockIdx.y / p.ctiles, ctile = blockIdx.y - tap * p.ctiles;
  const int split = blockIdx.z;
  const int64_t p_begin = (int64_t)split * p.chunk;
  const int64_t p_end = (p_begin + p.chunk < p.P) ? p_begin + p.chunk : p.P;
  const int num_kblocks = p_end > p_begin ? (int)((p_end - p_begin + PPB - 1) / PPB) : 0;
  const int ky = tap / d.kw, kx = tap - ky * d.kw;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; s++) { mbar_init(smem_u32(&full_bar[s]), NUM_PRODUCERS); mbar_init(smem_u32(&empty_bar[s]), 1); }
      mbar_init(smem_u32(&accum_bar), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 4) {
    const int t = threadIdx.x;
    const int pl = t % PPB, grp = t / PPB;            // pixel within the block, channel group
    const int o_first = otile * UM + grp * (A_PER_T * TC);
    const int c_first = ctile * BNC + grp * (B_PER_T * TC);
    for (int kb = 0; kb < num_kblocks; kb++) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      const int64_t pix = p_begin + (int64_t)kb * PPB + pl;
      const bool pok = pix < p_end;
      int n = 0, oy = 0, ox = 0;
      if (pok) { n = (int)(pix / p.ohw); const int rem = (int)(pix - (int64_t)n * p.ohw); oy = rem / d.out_w; ox = rem - oy * d.out_w; }
      const int iy = oy * d.stride + ky - d.pad_y, ix = ox * d.stride + kx - d.pad_x;
      const bool xok = pok && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w;
      const T* dyp = (const T*)p.dy + (int64_t)n * d.y_strides[0] + (int64_t)oy * d.y_strides[2] + (int64_t)ox * d.y_strides[3];
      const T* xp = (const T*)p.x + (int64_t)n * d.x_strides[0] + (int64_t)iy * d.x_strides[2] + (int64_t)ix * d.x_strides[3];
      const float* sc = d.in_scale ? (const float*)d.in_scale + (int64_t)n * d.ci : nullptr;
      uint4 va[A_PER_T], vb[B_PER_T];
#pragma unroll
      for (int j = 0; j < A_PER_T; j++) {
        const int o = o_first + j * TC;
        va[j] = (pok && o < d.co) ? __ldg((const uint4*)(dyp + o)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < B_PER_T; j++) {
        const int c = c_first + j * TC;
        vb[j] = (xok && c < d.ci) ? __ldg((const uint4*)(xp + c)) : make_uint4(0, 0, 0, 0);
      }
      if (sc) {
#pragma unroll
        for (int j = 0; j < B_PER_T; j++) {
          const int c = c_first + j * TC;
          if (xok && c < d.ci) {
            if (KIND == 2) {
              const float4 s4 = __ldg((const float4*)(sc + c));
              float* f = (float*)&vb[j];
              f[0] *= s4.x; f[1] *= s4.y; f[2] *= s4.z; f[3] *= s4.w;
            } else {
              const float4 sa = __ldg((const float4*)(sc + c)), sb = __ldg((const float4*)(sc + c + 4));
              const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              T* h = (T*)&vb[j];
#pragma unroll
              for (int e = 0; e < 8; e++) h[e] = from_acc<T>(to_acc<T>(h[e]) * sv[e]);
            }
          }
        }
      }
      if (KIND == 2) {
#pragma unroll
        for (int j = 0; j < A_PER_T; j++) {
          float* f = (float*)&va[j]; uint32_t* u = (uint32_t*)&va[j];
          u[0] = f32_to_tf32(f[0]); u[1] = f32_to_tf32(f[1]); u[2] = f32_to_tf32(f[2]); u[3] = f32_to_tf32(f[3]);
        }
#pragma unroll
        for (int j = 0; j < B_PER_T; j++) {
          float* f = (float*)&vb[j]; uint32_t* u = (uint32_t*)&vb[j];
          u[0] = f32_to_tf32(f[0]); u[1] = f32_to_tf32(f[1]); u[2] = f32_to_tf32(f[2]); u[3] = f32_to_tf32(f[3]);
        }
      }
      mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
      uint8_t* a_dst = smem + s * STAGE_BYTES + pl * 16;
      uint8_t* b_dst = a_dst + A_BYTES;
#pragma unroll
      for (int j = 0; j < A_PER_T; j++) *(uint4*)(a_dst + (grp * A_PER_T + j) * (PPB * 16)) = va[j];
#pragma unroll
      for (int j = 0; j < B_PER_T; j++) *(uint4*)(b_dst + (grp * B_PER_T + j) * (PPB * 16)) = vb[j];
      fence_proxy_async();
      mbar_arrive(smem_u32(&full_bar[s]));
    }

    // epilogue: lane = output channel, columns = input channels of this tile
    if (num_kblocks > 0) {
      mbar_wait(smem_u32(&accum_bar), 0);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int o = otile * UM + threadIdx.x;
      const int wy = d.flip ? d.kh - 1 - ky : ky, wx = d.flip ? d.kw - 1 - kx : kx;
#pragma unroll 1
      for (int cc = 0; cc < BNC; cc += 16) {
        uint32_t acc[16];
        tmem_ld16(lane_addr + cc, acc);
        if (o >= d.co) continue;
#pragma unroll
        for (int e = 0; e < 16; e++) {
          const int c = ctile * BNC + cc + e;
          if (c < d.ci) atomicAdd(p.dw + (((int64_t)o * d.ci + c) * d.kh + wy) * d.kw + wx, __uint_as_float(acc[e]));
        }
      }
      tc_fence_before();
    }
  } else {
    if (lane == 0) {
      for (int kb = 0; kb < num_kblocks; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < PPB / K_PER_MMA; kk++) {
          const uint64_t adesc = make_smem_desc(a_addr + kk * K_PER_MMA * 16, 128, PPB * 16);
          const uint64_t bdesc = make_smem_desc(b_addr + kk * K_PER_MMA * 16, 128, PPB * 16);
          umma<KIND>(tmem_base, adesc, bdesc, IDESC, (kb > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));
      }
      if (num_kblocks > 0) umma_commit(smem_u32(&accum_bar));
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side -------------------------------------------------------------------------------------------
int pick_bn(int co) {
  if (co <= 16) return 16;
  if (co <= 32) return 32;
  if (co <= 64) return 64;
  if (co <= 128) return 128;
  return 256;
}
static int stages_for(int bn) { return bn == 256 ? 4 : (bn == 128 ? 3 : 4); }

int elem_size(int dtype) { return dtype == SGB_F32 ? 4 : 2; }

// re-pack w into d->workspace as B-tile images for N tiles of `bn` channels (order [ntile][tap][cblock])
int pack_weights_umma(const sgb_conv_desc* d, const void* w, int bn, cudaStream_t s) {
  const int tc = 16 / elem_size(d->dtype);
  const int ntiles = (d->co + bn - 1) / bn, cblocks = (d->ci + 8 * tc - 1) / (8 * tc);
  const int64_t total = (int64_t)ntiles * cblocks * 8 * bn;           // one thread per 16-byte chunk position, all taps
  SGB_REQUIRE(total * d->kh * d->kw * tc < ((int64_t)1 << 31), "weight tensor too large for the pack kernel");
  int64_t blocks = ceil_div(total, 256); if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (d->dtype == SGB_F16)       pack_weights_kernel<__half, 0, 8><<<(unsigned)blocks, 256, 0, s>>>((const __half*)w, d->workspace, *d, bn, ntiles, cblocks);
  else if (d->dtype == SGB_BF16) pack_weights_kernel<__nv_bfloat16, 1, 8><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)w, d->workspace, *d, bn, ntiles, cblocks);
  else                           pack_weights_kernel<float, 2, 4><<<(unsigned)blocks, 256, 0, s>>>((const float*)w, d->workspace, *d, bn, ntiles, cblocks);
  SGB_LAUNCH_CHECK();
  return 0;
}

bool conv_umma_eligible(const sgb_conv_desc* d) {
  if (d->dtype != SGB_F32 && d->dtype != SGB_F16 && d->dtype != SGB_BF16) return false;
  if (d->dtype == SGB_F32 && d->strict_fp32) return false;
  if (d->force_simt == 1) return false;
  if (d->groups != 1) return false;
  const int tc = 16 / elem_size(d->dtype);
  if (d->ci % tc != 0) return false;
  if (d->x_strides[1] != 1 || d->y_strides[1] != 1) return false;              // channels_last only
  if (d->x_strides[0] % tc || d->x_strides[2] % tc || d->x_strides[3] % tc) return false;
  if (d->kh * d->kw > 49) return false;
  if (!d->workspace || d->workspace_bytes < sgb_conv2d_workspace_bytes(d)) return false;
  return true;
}

template <class T, int KIND, int BN>
static int launch_umma(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 3 : 4);
  constexpr int TC = 16 / sizeof(T);
  UmmaParams p; p.d = *d; p.x = x; p.y = y; p.wpack = d->workspace;
  p.M = (int64_t)d->n * d->out_h * d->out_w; p.ohw = d->out_h * d->out_w; p.taps = d->kh * d->kw;
  p.cblocks = (d->ci + 8 * TC - 1) / (8 * TC); p.ntiles = (d->co + BN - 1) / BN; p.tc = TC;
  const bool y_al = aligned16(y) && d->y_strides[0] % TC == 0 && d->y_strides[2] % TC == 0 && d->y_strides[3] % TC == 0;
  p.vec_store = (d->co % TC == 0 && y_al) ? 1 : 0;
  SGB_REQUIRE(aligned16(x) && aligned16(d->workspace), "x and workspace must be 16-byte aligned");
  // 1) re-pack the weights into B-tile images
  if (int r = pack_weights_umma(d, w, BN, s)) return r;
  // 2) implicit GEMM
  const size_t smem = (size_t)STAGES * (A_STAGE_BYTES + BN * KB_BYTES) + 1024;
  auto kern = conv_umma_kernel<T, KIND, BN, STAGES>;
  SGB_SET_MAX_SMEM(kern, (int)smem);
  const int64_t gx = ceil_div(p.M, UM);
  SGB_REQUIRE(gx <= 0x7fffffff && p.ntiles <= 65535, "problem too large for the UMMA conv grid");
  kern<<<dim3((unsigned)gx, (unsigned)p.ntiles), 192, smem, s>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

template <class T, int KIND>
static int dispatch_bn(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  switch (conv_bn(d)) {
    case 16:  return launch_umma<T, KIND, 16>(d, x, w, y, s);
    case 32:  return launch_umma<T, KIND, 32>(d, x, w, y, s);
    case 64:  return launch_umma<T, KIND, 64>(d, x, w, y, s);
    case 128: return launch_umma<T, KIND, 128>(d, x, w, y, s);
    default:  return launch_umma<T, KIND, 256>(d, x, w, y, s);
  }
}

int conv_forward_umma(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  if (d->dtype == SGB_F16) return dispatch_bn<__half, 0>(d, x, w, y, s);
  if (d->dtype == SGB_BF16) return dispatch_bn<__nv_bfloat16, 1>(d, x, w, y, s);
  return dispatch_bn<float, 2>(d, x, w, y, s);
}


bool conv_wgrad_umma_eligible(const sgb_conv_desc* d) {
  if (d->dtype != SGB_F32 && d->dtype != SGB_F16 && d->dtype != SGB_BF16) return false;
  if (d->dtype == SGB_F32 && d->strict_fp32) return false;
  if (d->force_simt == 1) return false;
  if (d->groups != 1 || d->transposed) return false;
  const int tc = 16 / elem_size(d->dtype);
  if (d->ci % tc != 0 || d->co % tc != 0) return false;
  if (d->x_strides[1] != 1 || d->y_strides[1] != 1) return false;
  if (d->x_strides[0] % tc || d->x_strides[2] % tc || d->x_strides[3] % tc) return false;
  if (d->y_strides[0] % tc || d->y_strides[2] % tc || d->y_strides[3] % tc) return false;
  if (d->kh * d->kw > 49) return false;
  return true;
}

template <class T, int KIND, int BNC>
static int launch_wgrad_umma(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t s) {
  constexpr int STAGES = BNC == 256 ? 4 : (BNC == 128 ? 3 : 4);
  constexpr int TC = 16 / sizeof(T);
  constexpr int PPB = 8 * TC;
  WgradParams p; p.d = *d; p.x = x; p.dy = dy; p.dw = dw;
  p.P = (int64_t)d->n * d->out_h * d->out_w; p.ohw = d->out_h * d->out_w; p.taps = d->kh * d->kw;
  p.ctiles = (d->ci + BNC - 1) / BNC; p.otiles = (d->co + UM - 1) / UM;
  SGB_REQUIRE(aligned16(x) && aligned16(dy), "x and dy must be 16-byte aligned");
  const int64_t tiles = (int64_t)p.otiles * p.ctiles * p.taps;
  int64_t splits = ceil_div((int64_t)num_sms() * 2, tiles);
  const int64_t max_splits = ceil_div(p.P, (int64_t)PPB * 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.chunk = ceil_div(ceil_div(p.P, splits), PPB) * PPB;
  splits = ceil_div(p.P, p.chunk);
  p.splits = (int)splits;
  SGB_REQUIRE((int64_t)p.ctiles * p.taps <= 65535 && splits <= 65535, "problem too large for the UMMA wgrad grid");
  const size_t smem = (size_t)STAGES * (UM * PPB * sizeof(T) + BNC * PPB * sizeof(T)) + 1024;
  auto kern = conv_wgrad_umma_kernel<T, KIND, BNC, STAGES>;
  SGB_SET_MAX_SMEM(kern, (int)smem);
  kern<<<dim3((unsigned)p.otiles, (unsigned)(p.ctiles * p.taps), (unsigned)splits), 160, smem, s>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

template <class T, int KIND>
static int dispatch_wgrad_bn(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t s) {
  const int ci = d->ci;
  if (ci <= 32) return launch_wgrad_umma<T, KIND, 32>(d, x, dy, dw, s);
  if (ci <= 64) return launch_wgrad_umma<T, KIND, 64>(d, x, dy, dw, s);
  if (ci <= 128) return launch_wgrad_umma<T, KIND, 128>(d, x, dy, dw, s);
  return launch_wgrad_umma<T, KIND, 256>(d, x, dy, dw, s);
}

int conv_wgrad_umma(const sgb_conv_desc* d, const void* x, const void* dy, void* dw, cudaStream_t s) {
  const int64_t wnum = (int64_t)d->co * d->ci * d->kh * d->kw;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * wnum, s);
  SGB_REQUIRE(e == cudaSuccess, "memset failed");
  if ((int64_t)d->n * d->out_h * d->out_w == 0) return 0;
  if (d->dtype == SGB_F16) return dispatch_wgrad_bn<__half, 0>(d, x, dy, (float*)dw, s);
  if (d->dtype == SGB_BF16) return dispatch_wgrad_bn<__nv_bfloat16, 1>(d, x, dy, (float*)dw, s);
  return dispatch_wgrad_bn<float, 2>(d, x, dy, (float*)dw, s);
}

}  // namespace sgb

extern "C" int64_t sgb_conv2d_workspace_bytes(const sgb_conv_desc* d) {
  if (!d || d->groups != 1) return 0;
  if (d->dtype != SGB_F32 && d->dtype != SGB_F16 && d->dtype != SGB_BF16) return 0;
  const int es = sgb::elem_size(d->dtype);
  const int tc = 16 / es;
  const int bn = sgb::conv_bn(d);
  const int64_t ntiles = (d->co + bn - 1) / bn, cblocks = (d->ci + 8 * tc - 1) / (8 * tc);
  return ntiles * d->kh * d->kw * cblocks * bn * sgb::KB_BYTES;
}
