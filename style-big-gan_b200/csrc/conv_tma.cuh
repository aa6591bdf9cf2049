// tcgen05 implicit-GEMM convolution with TMA-staged activation patches (stride 1, kernels up to 3x3, channels_last):
// forward of conv2d AND of conv_transpose2d stride 1 (= the data gradient of the other one).  S2 = 1: the 3 x 3 stride-2
// convolution through a tensor map with element stride 2 along W (two boxes per K block: the even- and the odd-column plane of the
// patch; a tap moves the descriptor's start inside its plane, consecutive output rows are two patch rows apart) -- exact, opt-in
// (conv_tma.cu: measured no faster than conv_halo_kernel's MODE 1).
//
// conv_halo_kernel stages the input patch of a super-tile with 8 producer warps (cp.async, 16 bytes per instruction, the
// source address of every chunk computed in registers).  The ncu captures of round 2 (profiles/README.md) show those
// warps and the LSU queue as the limiter on the layers with few channels, which are the HBM-bound ones.  Here ONE thread
// issues ONE `cp.async.bulk.tensor.4d` per K block: the patch is a box {KB bytes of channels, HC columns, HR rows, 1 image}
// of the NHWC tensor map; rows / columns outside the image (the convolution's zero padding) and channels beyond Ci are
// filled with zeros by the TMA unit, and the landed bytes are counted on an mbarrier.
//
// Shared-memory layout of a patch = what TMA writes: pixel-major rows of KB bytes (KB = 128: SWIZZLE_128B, KB = 64:
// SWIZZLE_64B), row p at stage + p * KB, its 16-byte chunks XOR-ed with the row phase ((address >> 7) & 7, resp. & 3).
// That is the canonical K-major swizzled UMMA operand: a core-matrix group = 8 consecutive pixels of one patch row,
// SBO = one patch row (HC * KB bytes) = the next row of the 16 x 8 output tile.  The halo trick of conv_halo_kernel
// carries over: a filter tap (ky, kx) and the tile index g within the super-tile only move the descriptor's start address
// by whole pixels.  Established by experiment on a B200 (benchmarks/experiments/tma_check.py, SGB_TMA_BO=0|1): the UMMA
// swizzle is a function of the ABSOLUTE shared-memory address, so a start address shifted by whole rows needs base_offset = 0;
// putting the start's swizzle phase into the descriptor's base-offset field gives wrong results for every shifted tap.
//
// Weights: the pre-packed no-swizzle B tiles of conv_umma.cu (pack_weights_umma), streamed with cp.async.bulk as before.
// Tiles are per image (a super-tile = 16 x 8*GT output pixels of one image), so one box covers a patch.
//
// Warps: 0-3 epilogue (TMEM -> registers -> shared staging -> 128-byte global stores), 4 MMA issuer + TMEM owner,
// 5 weight loader, 6 patch loader (TMA), 7-10 operand pass (only when the patch needs touching after it lands: style
// modulation `x * s[n, c]` and / or round-to-nearest TF32 conversion for fp32 tensors), which hands the stage to the MMA
// warp through a second mbarrier.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

namespace sgb {

constexpr int TMA_THREADS = 11 * 32;
constexpr int TMA_MAX_SA = 6, TMA_MAX_SB = 8;
constexpr int TMA_TILE_H = 16, TMA_TILE_W = 8;

struct TmaConvParams {
  sgb_conv_desc d;
  const void* wpack; void* y;
  int HR, HC;               // patch rows / columns (stride 2: columns of ONE column-parity plane)
  int plane_bytes;          // stride 2: bytes between the even-column and the odd-column plane of a stage (multiple of 1024)
  int sbo_bytes;            // bytes between consecutive output rows of the tile in the patch (one patch row, two for stride 2)
  int top, left;            // patch origin relative to the tile origin (input coordinates)
  int row_tiles, col_tiles, ntiles;
  int64_t total_tiles;
  int taps, cblocks;        // K blocks of KB bytes
  int a_stage_bytes;
  int sa, sb, tps;
  int stg_off;
  int vec_store;
  int pass;                 // 1: the operand pass runs (in_scale and / or TF32 rounding)
  int wres;                 // 1: the whole packed weight tensor is loaded into shared memory once and stays resident
  int w_bytes;              //    its size
  int bo_mode;              // descriptor base offset: 0 = always zero (correct), 1 = swizzle phase of the start address (the
                            // experiment that showed it is wrong; kept for benchmarks/experiments/tma_check.py)
  int tap_aoff[9];          // per tap: patch offset in pixels
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               :: "r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_decode_tile(const TmaConvParams& p, int gt, int64_t t, int& ntile, int& n, int& oy0, int& ox0) {
  ntile = (int)(t % p.ntiles);
  int64_t mt = t / p.ntiles;
  ox0 = (int)(mt % p.col_tiles) * (TMA_TILE_W * gt);
  mt /= p.col_tiles;
  oy0 = (int)(mt % p.row_tiles) * TMA_TILE_H;
  n = (int)(mt / p.row_tiles);
}

template <class T, int KIND, int BN, int KB, int GT, int S2>
__global__ void __launch_bounds__(TMA_THREADS, 1) conv_tma_kernel(const __grid_constant__ CUtensorMap tmx, TmaConvParams p) {
  constexpr int NPL = S2 ? 2 : 1;                      // column-parity planes per stage
  constexpr int TC = 16 / sizeof(T);
  constexpr int CH = KB / 16;                          // 16-byte chunks of K per stage
  constexpr int KSTEPS = KB / 32;                      // MMAs per tap per tile (32 bytes of K each)
  constexpr int B_TAP_BYTES = BN * KB;
  constexpr int NACC = GT;
  static_assert(NACC * BN <= 512, "accumulators exceed TMEM");
  constexpr int NBUF = (NACC * BN * 2 <= 512) ? 2 : 1;
  constexpr uint32_t NEED_COLS = NACC * BN * NBUF;
  constexpr uint32_t TMEM_COLS = NEED_COLS <= 32 ? 32 : (NEED_COLS <= 64 ? 64 : (NEED_COLS <= 128 ? 128 : (NEED_COLS <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = make_idesc(KIND, BN);
  constexpr uint32_t A_LAYOUT = (KB == 128) ? 2u : 4u;  // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  constexpr int PHASE_MASK = (KB == 128) ? 7 : 3;
  constexpr int SLAB = (BN * (int)sizeof(T) >= 128) ? 128 / (int)sizeof(T) : BN;
  constexpr int SLABB = SLAB * (int)sizeof(T);
  constexpr int PITCH = SLABB + 16;
  constexpr int LPP = SLABB / 16;
  constexpr int PPI = 32 / LPP;
  constexpr int NQ = 32 / PPI;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t a_full[TMA_MAX_SA], a_ready[TMA_MAX_SA], a_empty[TMA_MAX_SA], b_full[TMA_MAX_SB], b_empty[TMA_MAX_SB], acc_full[2], acc_empty[2];
  __shared__ uint64_t w_full;
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // swizzled stages: 1024-byte aligned

  const sgb_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SA = p.sa, SB = p.sb;
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + SA * p.a_stage_bytes;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < TMA_MAX_SA; s++) {
        mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_ready[s]), 4); mbar_init(smem_u32(&a_empty[s]), 1);
      }
      for (int s = 0; s < TMA_MAX_SB; s++) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
      for (int s = 0; s < 2; s++) { mbar_init(smem_u32(&acc_full[s]), 1); mbar_init(smem_u32(&acc_empty[s]), 4); }
      mbar_init(smem_u32(&w_full), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 6) {
    // =========================== patch loader: one TMA box per K block ===========================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" :: "l"(&tmx) : "memory");
      int sa = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (uint32_t)(NPL * p.HR * p.HC * KB);
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int ntile, n, oy0, ox0;
        tma_decode_tile(p, GT, tile, ntile, n, oy0, ox0);
        for (int cb = 0; cb < p.cblocks; cb++) {
          mbar_wait(smem_u32(&a_empty[sa]), ph ^ 1);
          const uint32_t bar = smem_u32(&a_full[sa]);
          mbar_arrive_expect_tx(bar, bytes);
          if (S2) {
            // stride 2: the tensor map walks the columns with element stride 2, so one box is the even-column (q = 0) or
            // the odd-column (q = 1) plane of the patch: 8 consecutive outputs of a row are again consecutive pixels
            const uint32_t dst = smem_u32(a_base + sa * p.a_stage_bytes);
#pragma unroll
            for (int q = 0; q < 2; q++)
              tma_load_4d(dst + q * p.plane_bytes, &tmx, cb * (KB / (int)sizeof(T)), 2 * ox0 - p.left + q, 2 * oy0 - p.top, n, bar);
          } else {
            tma_load_4d(smem_u32(a_base + sa * p.a_stage_bytes), &tmx, cb * (KB / (int)sizeof(T)), ox0 - p.left, oy0 - p.top, n, bar);
          }
          if (++sa == SA) { sa = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 7) {
    // =========================== operand pass (style scale / TF32 rounding), in place ===========================
    if (p.pass) {
      const int t = threadIdx.x - 7 * 32;
      const float* scb = (const float*)d.in_scale;
      const int nchunks = p.HR * p.HC * CH;
      int sa = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int ntile, n, oy0, ox0;
        tma_decode_tile(p, GT, tile, ntile, n, oy0, ox0);
        for (int cb = 0; cb < p.cblocks; cb++) {
          mbar_wait(smem_u32(&a_full[sa]), ph);
          for (int pl = 0; pl < NPL; pl++) {
          uint8_t* st = a_base + sa * p.a_stage_bytes + pl * p.plane_bytes;
          const uint32_t st_addr = smem_u32(st);
          for (int q = t; q < nchunks; q += 128) {
            const int pix = q / CH, pc = q - pix * CH;                 // physical chunk position within the row
            const int phase = (int)(((st_addr + (uint32_t)pix * KB) >> 7) & PHASE_MASK);
            const int lc = pc ^ phase;                                 // logical chunk = channel group
            const int ch = cb * (KB / (int)sizeof(T)) + lc * TC;
            if (ch < d.ci) {
              uint4* ptr = (uint4*)(st + (size_t)pix * KB + pc * 16);
              uint4 v = *ptr;
              if (KIND == 2) {
                float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y), f2 = __uint_as_float(v.z), f3 = __uint_as_float(v.w);
                if (scb) {
                  const float4 s4 = __ldg((const float4*)(scb + (int64_t)n * d.ci + ch));
                  f0 *= s4.x; f1 *= s4.y; f2 *= s4.z; f3 *= s4.w;
                }
                v.x = f32_to_tf32(f0); v.y = f32_to_tf32(f1); v.z = f32_to_tf32(f2); v.w = f32_to_tf32(f3);
              } else {
                const float* sp = scb + (int64_t)n * d.ci + ch;
                const float4 sa4 = __ldg((const float4*)sp), sb4 = __ldg((const float4*)(sp + 4));
                const float sv[8] = {sa4.x, sa4.y, sa4.z, sa4.w, sb4.x, sb4.y, sb4.z, sb4.w};
                T* h = (T*)&v;
#pragma unroll
                for (int e = 0; e < 8; e++) h[e] = from_acc<T>(to_acc<T>(h[e]) * sv[e]);
              }
              *ptr = v;
            }
          }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&a_ready[sa]));
          if (++sa == SA) { sa = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    const uint32_t sbo = (uint32_t)p.sbo_bytes;
    const uint32_t a_hi0 = smem_desc_hi(sbo) | (A_LAYOUT << 29), b_hi = smem_desc_hi(128);
    const uint32_t b_lo_base = smem_desc_lo(smem_u32(b_base), BN * 16);
    const uint32_t b_stage_u = (uint32_t)(p.tps * B_TAP_BYTES) >> 4;
    const uint32_t a_base_addr = smem_u32(a_base);
    int sa = 0, sb = 0, li = 0;
    uint32_t pha = 0, phb = 0;
    constexpr int HALVES_M = 8 / CH;
    const int cb128_m = (p.cblocks + HALVES_M - 1) / HALVES_M;
    const uint32_t leader = elect_one();            // the whole warp runs the loop; this lane issues
    if (p.wres) { mbar_wait(smem_u32(&w_full), 0); tc_fence_after(); }
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, li++) {
      const int buf = (NBUF == 2) ? (li & 1) : 0;
      const uint32_t eph = (NBUF == 2) ? ((li >> 1) & 1) : (li & 1);
      mbar_wait(smem_u32(&acc_empty[buf]), eph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * (NACC * BN);
      const int ntile_m = (int)(tile % p.ntiles);
      for (int cb = 0; cb < p.cblocks; cb++) {
        mbar_wait(smem_u32(p.pass ? &a_ready[sa] : &a_full[sa]), pha);
        tc_fence_after();
        const uint32_t a_stage = a_base_addr + sa * (uint32_t)p.a_stage_bytes;
        if (p.wres) {
          // resident weights: tile (ntile, tap, cb) sits where the packed image has it
          for (int tap = 0; tap < p.taps; tap++) {
            const uint32_t woff = (uint32_t)(((ntile_m * p.taps + tap) * cb128_m + cb / HALVES_M) * (BN * 128) + (cb % HALVES_M) * B_TAP_BYTES);
            const uint32_t b_lo = b_lo_base + (woff >> 4);
#pragma unroll
            for (int g = 0; g < GT; g++) {
              const uint32_t start = a_stage + (uint32_t)(p.tap_aoff[tap] + g * TMA_TILE_W) * KB;
              const uint32_t a_lo = ((start >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
              for (int kk = 0; kk < KSTEPS; kk++)
                umma_lh_pred<KIND>(leader, tmem_d + g * BN, a_lo + kk * 2, a_hi0, b_lo + kk * (2 * BN), b_hi, IDESC, (uint32_t)((cb | tap | kk) != 0));
            }
          }
        } else
        for (int tap0 = 0; tap0 < p.taps; tap0 += p.tps) {
          mbar_wait(smem_u32(&b_full[sb]), phb);
          tc_fence_after();
          const uint32_t b_lo0 = b_lo_base + sb * b_stage_u;
          for (int t = 0; t < p.tps; t++) {
            const int tap = tap0 + t;
            const uint32_t b_lo = b_lo0 + t * (B_TAP_BYTES >> 4);
#pragma unroll
            for (int g = 0; g < GT; g++) {
              const uint32_t start = a_stage + (uint32_t)(p.tap_aoff[tap] + g * TMA_TILE_W) * KB;
              const uint32_t a_hi = a_hi0 | (p.bo_mode ? (((start >> 7) & 7u) << 17) : 0u);
              const uint32_t a_lo = ((start >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
              for (int kk = 0; kk < KSTEPS; kk++)
                umma_lh_pred<KIND>(leader, tmem_d + g * BN, a_lo + kk * 2, a_hi, b_lo + kk * (2 * BN), b_hi, IDESC, (uint32_t)((cb | tap | kk) != 0));
            }
          }
          umma_commit_pred(leader, smem_u32(&b_empty[sb]));
          if (++sb == SB) { sb = 0; phb ^= 1; }
        }
        umma_commit_pred(leader, smem_u32(&a_empty[sa]));
        if (++sa == SA) { sa = 0; pha ^= 1; }
      }
      umma_commit_pred(leader, smem_u32(&acc_full[buf]));
    }
  } else if (warp == 5) {
    // =========================== weight loader ===========================
    if (lane == 0 && p.wres) {
      const uint32_t bar = smem_u32(&w_full);
      mbar_arrive_expect_tx(bar, (uint32_t)p.w_bytes);
      for (int off = 0; off < p.w_bytes; off += 16384) {
        const int nb = (p.w_bytes - off < 16384) ? p.w_bytes - off : 16384;
        bulk_copy_g2s(smem_u32(b_base) + off, (const uint8_t*)p.wpack + off, (uint32_t)nb, bar);
      }
    } else if (lane == 0) {
      int sb = 0;
      uint32_t phb = 0;
      // packed image: [ntile][tap][128-byte channel block][chunk (8)][row (BN)][16 B]; a stage takes CH chunks of it
      constexpr int HALVES = 8 / CH;
      const int cb128 = (p.cblocks + HALVES - 1) / HALVES;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int ntile = (int)(tile % p.ntiles);
        const uint8_t* wsrc = (const uint8_t*)p.wpack + (int64_t)ntile * p.taps * cb128 * (BN * 128);
        for (int cb = 0; cb < p.cblocks; cb++) {
          const uint8_t* wcb = wsrc + (int64_t)(cb / HALVES) * (BN * 128) + (cb % HALVES) * B_TAP_BYTES;
          for (int tap0 = 0; tap0 < p.taps; tap0 += p.tps) {
            mbar_wait(smem_u32(&b_empty[sb]), phb ^ 1);
            const uint32_t bar = smem_u32(&b_full[sb]);
            mbar_arrive_expect_tx(bar, p.tps * B_TAP_BYTES);
            const uint32_t dst = smem_u32(b_base + sb * (p.tps * B_TAP_BYTES));
            for (int t = 0; t < p.tps; t++)
              bulk_copy_g2s(dst + t * B_TAP_BYTES, wcb + (int64_t)(tap0 + t) * cb128 * (BN * 128), B_TAP_BYTES, bar);
            if (++sb == SB) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 0-3) ===========================
    const int m = threadIdx.x;                         // TMEM lane = tile row
    uint8_t* stg = smem + p.stg_off + warp * (32 * PITCH);
    // fused tail (sgb_conv_desc.out_scale / noise / bias / act): y = clamp(act(acc * out_scale[n,o] + noise[n,oy,ox] + bias[o]) * gain).
    // The per-channel vectors of the tile are staged once per tile in a per-warp slice of shared memory and read back as
    // broadcast float4 loads (conv_halo_kernel fetched them with one __ldg per element, which made its epilogue the bottleneck).
    float* s_sc = (float*)(smem + p.stg_off + 4 * 32 * PITCH) + warp * (2 * BN);
    float* s_bs = s_sc + BN;
    const bool fused = d.out_scale != nullptr || d.noise != nullptr || d.act != 0;
    const bool has_act = d.act != 0;
    const float alpha = d.alpha, gain = d.gain, clampv = d.clamp;
    const int q_chunk = lane % LPP;
    const int q_pix = lane / LPP;
    int li = 0;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, li++) {
      int ntile, n, oy0, ox0;
      tma_decode_tile(p, GT, tile, ntile, n, oy0, ox0);
      const int buf = (NBUF == 2) ? (li & 1) : 0;
      const uint32_t fph = (NBUF == 2) ? ((li >> 1) & 1) : (li & 1);
      const int o_base = ntile * BN;
      // this thread's row of the tile and the rows it stores in the store phase (same for every accumulator up to the x shift)
      const int oy = oy0 + (m >> 3), oxr = ox0 + (m & 7);
      const int64_t ybase = (int64_t)n * d.y_strides[0];
      int64_t yoff[NQ];
      int qox[NQ];
#pragma unroll
      for (int k = 0; k < NQ; k++) {
        const int mm = warp * 32 + q_pix + PPI * k;
        const int oy2 = oy0 + (mm >> 3);
        qox[k] = ox0 + (mm & 7);
        yoff[k] = (oy2 < d.out_h) ? (ybase + (int64_t)oy2 * d.y_strides[2]) : -1;
      }
      if (fused) {
        for (int i = lane; i < BN; i += 32) {
          const int o = o_base + i;
          s_sc[i] = (d.out_scale && o < d.co) ? ((const float*)d.out_scale)[(int64_t)n * d.co + o] : 1.f;
          s_bs[i] = (d.bias && o < d.co) ? to_acc<T>(((const T*)d.bias)[o]) : 0.f;
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&acc_full[buf]), fph);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < GT; g++) {
        const int ox = oxr + g * TMA_TILE_W;
        const bool row_ok = oy < d.out_h && ox < d.out_w;
        const float nz = (d.noise && row_ok) ? ((const float*)d.noise)[((int64_t)n * d.out_h + oy) * d.out_w + ox] : 0.f;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (NACC * BN) + g * BN;
        T* yrow = (T*)p.y + ybase + (int64_t)oy * d.y_strides[2] + (int64_t)ox * d.y_strides[3];
#pragma unroll 1
        for (int cc0 = 0; cc0 < BN; cc0 += SLAB) {
#pragma unroll
          for (int cs = 0; cs < SLAB; cs += 16) {
            const int cc = cc0 + cs;
            uint32_t acc[16];
            tmem_ld16(lane_addr + cc, acc);
            float val[16];
#pragma unroll
            for (int e = 0; e < 16; e++) val[e] = __uint_as_float(acc[e]);
            if (fused) {
#pragma unroll
              for (int q4 = 0; q4 < 4; q4++) {
                const float4 sc4 = *(const float4*)(s_sc + cc + q4 * 4), bs4 = *(const float4*)(s_bs + cc + q4 * 4);
                const float scv[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, bsv[4] = {bs4.x, bs4.y, bs4.z, bs4.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  float v = val[q4 * 4 + e] * scv[e] + nz + bsv[e];
                  if (has_act) {
                    v = (v > 0.f ? v : v * alpha) * gain;                        // explicit compares: NaN propagates like torch
                    if (clampv >= 0.f) v = v < -clampv ? -clampv : (v > clampv ? clampv : v);
                  }
                  val[q4 * 4 + e] = v;
                }
              }
            }
            if (p.vec_store) {
#pragma unroll
              for (int q = 0; q < 16 / TC; q++) {
                Vec16<T> pk;
#pragma unroll
                for (int e = 0; e < TC; e++) pk.v[e] = from_acc<T>(val[q * TC + e]);
                *(uint4*)(stg + lane * PITCH + cs * (int)sizeof(T) + q * 16) = pk.raw;
              }
            } else if (row_ok) {
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const int o = o_base + cc + e;
                if (o < d.co) yrow[(int64_t)o * d.y_strides[1]] = from_acc<T>(val[e]);
              }
            }
          }
          if (p.vec_store) {
            __syncwarp();
            const int o = o_base + cc0 + q_chunk * TC;
            if (o < d.co) {
#pragma unroll
              for (int k = 0; k < NQ; k++) {
                const int ox2 = qox[k] + g * TMA_TILE_W;
                if (yoff[k] >= 0 && ox2 < d.out_w) {
                  const uint4 v = *(const uint4*)(stg + (q_pix + PPI * k) * PITCH + q_chunk * 16);
                  *(uint4*)((T*)p.y + yoff[k] + (int64_t)ox2 * d.y_strides[3] + o) = v;
                }
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled tma_encode_fn();      // conv_tma.cu: cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no -lcuda)

template <class T, int KIND, int BN, int KB, int GT, int S2>
int launch_tma(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  constexpr int TC = 16 / sizeof(T);
  constexpr int SLAB = (BN * (int)sizeof(T) >= 128) ? 128 / (int)sizeof(T) : BN;
  constexpr int PITCH = SLAB * (int)sizeof(T) + 16;
  TmaConvParams p; p.d = *d; p.y = y; p.wpack = d->workspace;
  const int TW = TMA_TILE_W * GT;
  if (S2) {
    p.HR = 2 * (TMA_TILE_H - 1) + d->kh; p.HC = TW + (d->kw - 1 + 1) / 2;      // rows: all of them; columns: one parity plane
    p.top = d->pad_y; p.left = d->pad_x;
  } else {
    p.HR = TMA_TILE_H + d->kh - 1; p.HC = TW + d->kw - 1;
    p.top = d->transposed ? (d->kh - 1 - d->pad_y) : d->pad_y;
    p.left = d->transposed ? (d->kw - 1 - d->pad_x) : d->pad_x;
  }
  p.plane_bytes = (p.HR * p.HC * KB + 1023) / 1024 * 1024;
  p.sbo_bytes = (S2 ? 2 : 1) * p.HC * KB;
  p.row_tiles = (d->out_h + TMA_TILE_H - 1) / TMA_TILE_H;
  p.col_tiles = (d->out_w + TW - 1) / TW;
  p.ntiles = (d->co + BN - 1) / BN;
  p.total_tiles = (int64_t)d->n * p.row_tiles * p.col_tiles * p.ntiles;
  p.taps = d->kh * d->kw;
  p.cblocks = (d->ci * (int)sizeof(T) + KB - 1) / KB;
  for (int tap = 0; tap < p.taps; tap++) {
    const int ky = tap / d->kw, kx = tap - ky * d->kw;
    const int pr = d->transposed ? (d->kh - 1 - ky) : ky, pc = d->transposed ? (d->kw - 1 - kx) : kx;
    p.tap_aoff[tap] = S2 ? ((kx & 1) * (p.plane_bytes / KB) + ky * p.HC + (kx >> 1)) : (pr * p.HC + pc);
  }
  p.a_stage_bytes = (S2 ? 2 : 1) * p.plane_bytes;
  const bool y_al = aligned16(y) && d->y_strides[0] % TC == 0 && d->y_strides[2] % TC == 0 && d->y_strides[3] % TC == 0;
  p.vec_store = (d->co % TC == 0 && y_al) ? 1 : 0;
  static const int no_round = [] { const char* e = getenv("SGB_TMA_NOROUND"); return e ? atoi(e) : 0; }();
  static const int bo_mode = [] { const char* e = getenv("SGB_TMA_BO"); return e ? atoi(e) : 0; }();
  p.pass = (d->in_scale != nullptr || (KIND == 2 && !no_round)) ? 1 : 0;
  p.bo_mode = bo_mode;
  SGB_REQUIRE(aligned16(x) && aligned16(d->workspace), "x and workspace must be 16-byte aligned");
  SGB_REQUIRE(d->act == 0 || d->act == SGB_ACT_LINEAR || d->act == SGB_ACT_LRELU, "fused epilogue supports linear and lrelu only");
  if (d->act == SGB_ACT_LINEAR) p.d.alpha = 1.f;

  // weights first: the pack kernel's launch also binds this thread to the device's primary context, which the driver entry
  // point below needs (an autograd worker thread whose first CUDA call was cuTensorMapEncodeTiled got CUDA_ERROR_INVALID_CONTEXT)
  if (int r = pack_weights_umma(d, w, BN, s)) return r;
  // tensor map of x: dims (innermost first) {C, W, H, N}, byte strides of W, H, N; box {KB bytes of channels, HC, HR, 1}
  CUtensorMap tm;
  const cuuint64_t gdim[4] = {(cuuint64_t)d->ci, (cuuint64_t)d->in_w, (cuuint64_t)d->in_h, (cuuint64_t)d->n};
  const cuuint64_t gstr[3] = {(cuuint64_t)d->x_strides[3] * sizeof(T), (cuuint64_t)d->x_strides[2] * sizeof(T), (cuuint64_t)d->x_strides[0] * sizeof(T)};
  // element stride 2 along W (stride-2 convolution): TMA loads ceil(box / stride) elements, so a plane of HC columns is a box of 2 * HC
  const cuuint32_t box[4] = {(cuuint32_t)(KB / sizeof(T)), (cuuint32_t)((S2 ? 2 : 1) * p.HC), (cuuint32_t)p.HR, 1u};
  const cuuint32_t estr[4] = {1, (cuuint32_t)(S2 ? 2 : 1), 1, 1};
  const CUtensorMapDataType dt = KIND == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (KIND == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  PFN_encodeTiled enc = tma_encode_fn();
  SGB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const CUresult cr = enc(&tm, dt, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          KB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SGB_REQUIRE(cr == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")");

  const int budget = 223 * 1024;                    // 226 KB of dynamic shared memory - 1 KB alignment slack - margin
  const int stg_bytes = 4 * 32 * PITCH + 4 * 2 * BN * (int)sizeof(float);      // store slabs + per-warp scale / bias vectors
  const int b_tap = BN * KB;
  int sa = 0, sb = 0, tps = 1;
  // few channels: the whole packed weight tensor stays in shared memory (no weight stream, no b_full / b_empty hand-shakes)
  {
    const int cb128 = (p.cblocks * KB + 127) / 128;
    const int64_t wb = (int64_t)p.ntiles * p.taps * cb128 * (BN * 128);
    static const int no_res = [] { const char* e = getenv("SGB_TMA_NORES"); return e ? atoi(e) : 0; }();
    p.wres = (!no_res && wb <= 80 * 1024) ? 1 : 0;
    p.w_bytes = (int)wb;
  }
  if (p.wres) {
    sa = (budget - stg_bytes - p.w_bytes) / p.a_stage_bytes;
    if (sa > TMA_MAX_SA) sa = TMA_MAX_SA;
    SGB_REQUIRE(sa >= 2, "shared memory budget exceeded");
    p.sa = sa; p.sb = 2; p.tps = 1;
    p.stg_off = sa * p.a_stage_bytes + (p.w_bytes + 1023) / 1024 * 1024;
  } else {
  for (int cand = p.taps; cand >= 1; cand--) {
    if (p.taps % cand) continue;
    const int b_stage_c = cand * b_tap;
    int sb_c = (64 * 1024 + b_stage_c - 1) / b_stage_c;
    if (sb_c < 2) sb_c = 2;
    if (sb_c > TMA_MAX_SB) sb_c = TMA_MAX_SB;
    const int sa_c = (budget - stg_bytes - sb_c * b_stage_c) / p.a_stage_bytes;
    if (sa_c >= 3 || cand == 1) { tps = cand; sb = sb_c; sa = sa_c > TMA_MAX_SA ? TMA_MAX_SA : sa_c; break; }
  }
  SGB_REQUIRE(sa >= 2 && sb >= 2, "shared memory budget exceeded");
  const int b_stage = tps * b_tap;
  {
    const int extra = (budget - stg_bytes - sa * p.a_stage_bytes - sb * b_stage) / b_stage;
    sb = (sb + extra > TMA_MAX_SB) ? TMA_MAX_SB : sb + extra;
  }
  p.sa = sa; p.sb = sb; p.tps = tps;
  p.stg_off = sa * p.a_stage_bytes + sb * b_stage;
  }
  const size_t smem = (size_t)p.stg_off + stg_bytes + 2048;
  auto kern = conv_tma_kernel<T, KIND, BN, KB, GT, S2>;
  SGB_REQUIRE(smem <= 226 * 1024, "shared memory plan exceeds 226 KB");
  SGB_SET_MAX_SMEM(kern, 226 * 1024);
  const int64_t grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kern<<<(unsigned)grid, TMA_THREADS, smem, s>>>(tm, p);
  SGB_LAUNCH_CHECK();
  return 0;
}

// (BN, KB, GT) -> launcher
template <class T, int KIND>
int dispatch_tma(int bn, int kb, int gt, const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
#define SGB_TMA_CASE(BN_, KB_, GT_, S2_) \
  if (bn == BN_ && kb == KB_ && gt == GT_ && (d->stride == 2) == (S2_ == 1)) return launch_tma<T, KIND, BN_, KB_, GT_, S2_>(d, x, w, y, s);
  SGB_TMA_CASE(32, 64, 4, 0) SGB_TMA_CASE(64, 64, 4, 0)
  SGB_TMA_CASE(32, 128, 2, 0) SGB_TMA_CASE(64, 128, 2, 0) SGB_TMA_CASE(128, 128, 2, 0) SGB_TMA_CASE(256, 128, 2, 0)
  SGB_TMA_CASE(32, 64, 1, 1) SGB_TMA_CASE(64, 64, 1, 1) SGB_TMA_CASE(128, 64, 1, 1)        // stride 2: 64-byte K blocks, one tile per stage
#undef SGB_TMA_CASE
  set_error("conv_tma: no kernel for this (BN, KB, GT)");
  return 1;
}

}  // namespace sgb
