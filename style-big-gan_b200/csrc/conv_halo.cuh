// tcgen05 implicit-GEMM convolution, persistent "halo tile" kernel (channels_last).  Kernel template; the three
// dtype translation units (conv_halo_f16.cu / _bf16.cu / _f32.cu) instantiate it, conv_halo.cu dispatches.
//
// conv_umma.cu re-gathers the A operand once per filter tap (9x the activation traffic through LSU + shared
// memory).  Here the CTA stages the input patch of its output tile ONCE per channel block (64 bytes of K) and the
// filter taps are shared-memory descriptors into the same patch.  Three geometries (template parameter MODE):
//
//   MODE 0  stride 1, conv2d and conv_transpose2d (forward and data gradient of every 3x3 / 1x1 layer)
//           a tile = 16 x 8 output pixels (M = 128, row m = ty*8 + tx); a CTA works on a SUPER-TILE of GT tiles side
//           by side (16 x 8*GT pixels, GT accumulators) so that every weight tile it pulls from L2 feeds GT times
//           the MMAs -- with one tile per weight pass the kernel is bound by the L2 -> shared-memory weight stream.
//           A core matrix = 8 consecutive x of one tile row = 8 consecutive patch pixels (16 B apart);
//           SBO = one patch row = next tile row; the tap / the tile index only move the start address.
//           Rows are "virtual rows": every image contributes out_h + kh - 1 rows (zero padding included), so tiles
//           may straddle images under one addressing scheme; the kh - 1 junk rows per image are dropped.
//   MODE 1  conv2d stride 2 (D down path after the FIR; data gradient of the G up path)
//           tile = 16 x 8 output pixels; patch = (30+kh) x (14+kw) input pixels stored with the columns
//           DE-INTERLEAVED BY PARITY, so the 8 pixels ox..ox+7 of a tap (input columns 2*ox + kx) are again 16 B
//           apart; SBO = two patch rows.
//   MODE 2  conv_transpose2d stride 2 (G up path; data gradient of the D down path)
//           the output splits into 4 parity phases (oy+pad, ox+pad mod 2); a tile = 16 x 8 positions (a, b) of the
//           half-resolution grid, i.e. 4 x 128 output pixels; phase (py, px) only receives the taps with ky = py,
//           kx = px (mod 2), each a stride-1 gather at (a - ky/2, b - kx/2).  4*GT TMEM accumulators per super-tile,
//           no multiplications by the inserted zeros (the per-tap kernel spends 4x the MMAs on them).
//
// Persistent: grid = #SMs, each CTA walks super-tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  Ten warps:
//   warps 0-3  epilogue: tcgen05.ld of an accumulator, demod scale / noise / bias + lrelu|linear + clamp, transpose
//              through a per-warp shared-memory staging slab so that every global store instruction writes whole
//              128-byte lines, release the buffer -- overlaps the next super-tile's MMAs when TMEM has room for two
//   warp 4     MMA issuer (one lane): tcgen05.mma.cta_group::1; owns TMEM
//   warp 5     weight loader: cp.async.bulk of pre-packed B tiles into a ring of SB stages
//   warps 6-9  patch producers: cp.async 16-byte chunks global -> shared (zero-fill for padding), several patches in
//              flight per thread, optional in-place style scaling, fence.proxy.async, mbarrier arrive
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

namespace sgb {

constexpr int TILE_H = 16, TILE_W = 8;
constexpr int HALO_PRODUCERS = 256;       // patch producer threads (8 warps)
constexpr int HALO_THREADS = 192 + HALO_PRODUCERS;
constexpr int HALO_CH = 4;               // 16-byte channel chunks of K per stage (64 bytes)
constexpr int MAX_SA = 6, MAX_SB = 8;

struct HaloParams {
  sgb_conv_desc d;
  const void* x; const void* wpack; void* y;
  int VR;               // rows of the tile-row space per image (MODE 0: out_h + kh - 1; MODE 2: in_h + 1; MODE 1: out_h padded to 16)
  int HR, HC;           // patch rows / column slots per row
  int QP;               // MODE 1: slot offset of the odd-column plane
  int top, left;        // patch origin relative to the tile origin (input coordinates)
  int col_tiles, ntiles;
  int64_t total_tiles;
  int taps, cblocks;
  int lbo;              // bytes between channel chunks of the patch (padded)
  int a_stage_bytes;
  int sa, sb;           // stages
  int tps;              // filter taps per weight stage (divides taps): one mbarrier round trip feeds tps * GT * 2 MMAs
  int stg_off;          // byte offset of the epilogue staging area
  int vec_store;
  int debug;            // SGB_HALO_DEBUG bit mask (experiments only): 1 = no epilogue stores, 2 = no patch loads, 4 = no MMAs, 8 = no weight loads
  int tap_aoff[9];      // per tap: patch offset (pixels) of the A descriptor
  int tap_acc[9];       // per tap: accumulator (output phase) it feeds
};

__device__ __forceinline__ void decode_tile(const HaloParams& p, int gt, int64_t t, int& ntile, int& u0, int& x0) {
  ntile = (int)(t % p.ntiles);
  const int64_t mt = t / p.ntiles;
  x0 = (int)(mt % p.col_tiles) * (TILE_W * gt);
  u0 = (int)(mt / p.col_tiles) * TILE_H;
}

__device__ __forceinline__ void cp_async_wait_dyn(int n) {   // wait until at most n of this thread's groups are pending
  if (n <= 0) cp_async_wait<0>();
  else if (n == 1) cp_async_wait<1>();
  else if (n == 2) cp_async_wait<2>();
  else if (n == 3) cp_async_wait<3>();
  else cp_async_wait<4>();
}

constexpr int halo_max_slots(int mode, int gt) {             // ceil(patch pixels * HALO_CH / HALO_PRODUCERS)
  return (HALO_PRODUCERS == 256) ? (mode == 1 ? 9 : (mode == 2 ? (gt == 1 ? 3 : (gt == 2 ? 5 : 9)) : (gt == 1 ? 3 : (gt == 2 ? 6 : 10))))
                                 : (mode == 1 ? 18 : (mode == 2 ? (gt == 1 ? 5 : (gt == 2 ? 10 : 18)) : (gt == 1 ? 6 : (gt == 2 ? 11 : 20))));
}

template <class T, int KIND, int BN, int MODE, int GT>
__global__ void __launch_bounds__(HALO_THREADS, 1) conv_halo_kernel(HaloParams p) {
  constexpr int TC = 16 / sizeof(T);
  constexpr int CH = HALO_CH;
  constexpr int BK = CH * TC;
  constexpr int B_TAP_BYTES = BN * CH * 16;           // one tap's weight tile for one K stage
  constexpr int NPH = (MODE == 2) ? 4 : 1;            // output phases per tile
  constexpr int NACC = NPH * GT;                      // accumulators per super-tile
  static_assert(NACC * BN <= 512, "accumulators exceed TMEM");
  constexpr int NBUF = (NACC * BN * 2 <= 512) ? 2 : 1;
  constexpr uint32_t NEED_COLS = NACC * BN * NBUF;
  constexpr uint32_t TMEM_COLS = NEED_COLS <= 32 ? 32 : (NEED_COLS <= 64 ? 64 : (NEED_COLS <= 128 ? 128 : (NEED_COLS <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = make_idesc(KIND, BN);
  constexpr int MAX_SLOTS = halo_max_slots(MODE, GT);
  constexpr int PPS = HALO_PRODUCERS / CH;            // patch pixels per slot pass
  // epilogue staging: SLAB columns (128 bytes of output per pixel when the tile is that wide)
  constexpr int SLAB = (BN * (int)sizeof(T) >= 128) ? 128 / (int)sizeof(T) : BN;
  constexpr int SLABB = SLAB * (int)sizeof(T);
  constexpr int PITCH = SLABB + 16;
  constexpr int LPP = SLABB / 16;                     // lanes per pixel in the store phase
  constexpr int PPI = 32 / LPP;                       // pixels per store instruction
  constexpr int NQ = 32 / PPI;                        // store instructions per slab (= LPP)

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t a_full[MAX_SA], a_empty[MAX_SA], b_full[MAX_SB], b_empty[MAX_SB], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const sgb_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SA = p.sa, SB = p.sb;
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + SA * p.a_stage_bytes;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < MAX_SA; s++) { mbar_init(smem_u32(&a_full[s]), HALO_PRODUCERS / 32); mbar_init(smem_u32(&a_empty[s]), 1); }
      for (int s = 0; s < MAX_SB; s++) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
      for (int s = 0; s < 2; s++) { mbar_init(smem_u32(&acc_full[s]), 1); mbar_init(smem_u32(&acc_empty[s]), 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= 6) {
    // =========================== patch producers (cp.async) ===========================
    const int t = threadIdx.x - 192;
    const int j = t & (CH - 1);                        // channel chunk owned by this thread
    const int pl = t / CH;
    const int npix = p.HR * p.HC;
    const int lookahead = SA >= 3 ? SA - 2 : 1;
    const T* xb = (const T*)p.x;
    const float* scb = (const float*)d.in_scale;
    // per slot: patch coordinates (constant over tiles) and destination offset
    int hrc[MAX_SLOTS];                                // hr << 16 | hc, -1 = unused slot
    uint32_t dsl[MAX_SLOTS];
#pragma unroll
    for (int i = 0; i < MAX_SLOTS; i++) {
      const int pix = pl + PPS * i;
      hrc[i] = -1; dsl[i] = 0;
      if (pix < npix) {
        const int hr = pix / p.HC, hc = pix - hr * p.HC;
        hrc[i] = (hr << 16) | hc;
        const int slot = (MODE == 1) ? (hr * p.HC + (hc & 1) * p.QP + (hc >> 1)) : pix;
        dsl[i] = (uint32_t)(j * p.lbo + slot * 16);
      }
    }
    int pa = 0;                                        // patches issued so far
    int pub = 0;                                       // patches published so far
    int sa_i = 0, sa_p = 0;                            // ring positions of the next patch to issue / to publish
    uint32_t ph_i = 0;                                 // phase of the issue position
    int pub_tile_cb = 0;                               // channel block of patch `pub` within its tile
    int64_t pub_tile = blockIdx.x;                     // tile of patch `pub`
    // image / virtual row of that tile's first patch row, refreshed once per tile (the 64-bit divisions of decode_tile were
    // ~15 % of the producers' samples when they ran once per K block: profiles/README.md, small 512-channel layers)
    int pub_n0 = 0, pub_rem0 = 0;
    const bool need_pass = scb || (KIND == 2 && !(p.debug & 64));      // 16-bit tensors without a style scale: publish() touches nothing
    auto pub_decode = [&]() {
      if (need_pass && pub_tile < p.total_tiles) {
        int nt_, u0o, x0_;
        decode_tile(p, GT, pub_tile, nt_, u0o, x0_);
        pub_n0 = u0o / p.VR; pub_rem0 = u0o - pub_n0 * p.VR;
      }
    };
    pub_decode();

    auto publish = [&]() {                             // the oldest unpublished patch has landed: scaling, fence, arrive
      const int sa = sa_p;
      if (++sa_p == SA) sa_p = 0;
      // fp32 tensors (KIND 2): the MMA reads the top 19 bits of every fp32 word, i.e. it TRUNCATES to TF32.  Truncation is
      // biased (every operand shrinks by ~2^-11 on average), and the bias compounds over the layers of a network: the
      // deepest gradients of the golden network came out ~2-3 % short.  The producers therefore round the landed patch to
      // TF32 (cvt.rna, what cuDNN's TF32 convolutions amount to) in the same in-place pass that applies the style scale.
      if (need_pass) {                                    // SGB_HALO_DEBUG=64: A/B switch, leave the truncation to the MMA
        uint8_t* dst = a_base + sa * p.a_stage_bytes;
        const int co = pub_tile_cb * BK + j * TC;
        if (co < d.ci) {
          const int n0 = pub_n0, rem0 = pub_rem0;
#pragma unroll
          for (int i = 0; i < MAX_SLOTS; i++) {
            if (hrc[i] >= 0) {
              int n = n0;
              if (scb) {
                if (MODE == 0) { int r = rem0 + (hrc[i] >> 16); while (r >= p.VR) { r -= p.VR; n++; } }
                if (MODE == 2) { int r = rem0 - p.top + (hrc[i] >> 16); if (r < 0) n--; while (r >= p.VR) { r -= p.VR; n++; } }
                n = n < d.n ? (n < 0 ? 0 : n) : d.n - 1;
              }
              const float* sp = scb + (int64_t)n * d.ci + co;
              uint4* q = (uint4*)(dst + dsl[i]);
              uint4 v = *q;
              if (KIND == 2) {
                float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y), f2 = __uint_as_float(v.z), f3 = __uint_as_float(v.w);
                if (scb) {
                  const float4 s4 = __ldg((const float4*)sp);
                  f0 *= s4.x; f1 *= s4.y; f2 *= s4.z; f3 *= s4.w;
                }
                v.x = f32_to_tf32(f0); v.y = f32_to_tf32(f1); v.z = f32_to_tf32(f2); v.w = f32_to_tf32(f3);
              } else {
                const float4 sa4 = __ldg((const float4*)sp), sb4 = __ldg((const float4*)(sp + 4));
                const float sv[8] = {sa4.x, sa4.y, sa4.z, sa4.w, sb4.x, sb4.y, sb4.z, sb4.w};
                T* h = (T*)&v;
#pragma unroll
                for (int e = 0; e < 8; e++) h[e] = from_acc<T>(to_acc<T>(h[e]) * sv[e]);
              }
              *q = v;
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();                                    // one mbarrier arrival per warp instead of per thread
      if (lane == 0) mbar_arrive(smem_u32(&a_full[sa]));
      pub++;
      if (++pub_tile_cb == p.cblocks) { pub_tile_cb = 0; pub_tile += gridDim.x; pub_decode(); }
    };

    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int ntile, u0, x0;
      decode_tile(p, GT, tile, ntile, u0, x0);
      // per-slot source offsets (elements) of the patch pixels this thread stages; -1 = zero (padding / outside)
      int off[MAX_SLOTS];
      if (p.debug & 32) {
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; i++) off[i] = -1;
      } else {
        const int n0 = u0 / p.VR, r0 = u0 - n0 * p.VR;             // MODE 1: tiles never straddle images
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; i++) {
          off[i] = -1;
          if (hrc[i] >= 0) {
            const int hr = hrc[i] >> 16, hc = hrc[i] & 0xffff;
            int n = n0, iy, ix;
            if (MODE == 0) {
              int r = r0 + hr;
              while (r >= p.VR) { r -= p.VR; n++; }
              iy = r - p.top;
              ix = x0 + hc - p.left;
            } else if (MODE == 1) {
              iy = 2 * r0 - p.top + hr; ix = 2 * x0 - p.left + hc;
            } else {
              int r = r0 - p.top + hr;
              if (r < 0) { r += p.VR; n--; }
              while (r >= p.VR) { r -= p.VR; n++; }
              iy = r; ix = x0 - p.left + hc;
            }
            if (n >= 0 && n < d.n && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w)
              off[i] = (int)(n * d.x_strides[0] + iy * d.x_strides[2] + ix * d.x_strides[3]);
          }
        }
      }
      for (int cb = 0; cb < p.cblocks; cb++, pa++) {
        const int sa = sa_i;
        const int c = cb * BK + j * TC;
        const bool c_ok = c < d.ci;
        mbar_wait(smem_u32(&a_empty[sa]), ph_i ^ 1);
        if (++sa_i == SA) { sa_i = 0; ph_i ^= 1; }
        const uint32_t dst = smem_u32(a_base + sa * p.a_stage_bytes);
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; i++) {
          if (hrc[i] >= 0) {
            const bool ok = c_ok && off[i] >= 0;
            if (!(p.debug & 2)) cp_async16(dst + dsl[i], ok ? (const void*)(xb + off[i] + c) : (const void*)xb, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        if (pa - pub >= lookahead) {                   // keep `lookahead` patches in flight, publish the oldest
          cp_async_wait_dyn(lookahead);
          publish();
        }
      }
    }
    cp_async_wait<0>();
    while (pub < pa) publish();
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the loop (warp-uniform values stay in uniform registers); lane 0 issues.  Descriptors are
    // kept as (lo, hi) halves so that an MMA costs two 32-bit adds; ring positions are counters, not divisions.
    {
      const uint32_t sbo = (uint32_t)((MODE == 1 ? 2 : 1) * p.HC * 16);
      const uint32_t a_hi = smem_desc_hi(sbo), b_hi = smem_desc_hi(128);
      const uint32_t a_lo_base = smem_desc_lo(smem_u32(a_base), (uint32_t)p.lbo);
      const uint32_t b_lo_base = smem_desc_lo(smem_u32(b_base), BN * 16);
      const uint32_t a_stage_u = (uint32_t)p.a_stage_bytes >> 4, kk_u = (uint32_t)(2 * p.lbo) >> 4;
      const uint32_t b_stage_u = (uint32_t)(p.tps * B_TAP_BYTES) >> 4;
      int sa = 0, sb = 0, li = 0;
      uint32_t pha = 0, phb = 0;
      const uint32_t leader = elect_one();
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, li++) {
        const int buf = (NBUF == 2) ? (li & 1) : 0;
        const uint32_t eph = (NBUF == 2) ? ((li >> 1) & 1) : (li & 1);
        mbar_wait(smem_u32(&acc_empty[buf]), eph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * (NACC * BN);
        uint32_t started = 0;                              // bit a set: phase a has been written in this tile
        for (int cb = 0; cb < p.cblocks; cb++) {
          mbar_wait(smem_u32(&a_full[sa]), pha);
          tc_fence_after();
          const uint32_t a_lo0 = a_lo_base + sa * a_stage_u;
          for (int tap0 = 0; tap0 < p.taps; tap0 += p.tps) {
            mbar_wait(smem_u32(&b_full[sb]), phb);
            tc_fence_after();
            const uint32_t b_lo0 = b_lo_base + sb * b_stage_u;
            // warp-uniform issue (umma.cuh: elect_one / umma_lh_pred): every lane runs the loop so that the descriptor
            // arithmetic stays in uniform registers; the elected lane issues
            for (int t = 0; t < p.tps; t++) {
              const int tap = tap0 + t;
              const int acc = (MODE == 2) ? p.tap_acc[tap] : 0;
              const uint32_t a_lo = a_lo0 + (uint32_t)p.tap_aoff[tap];
              const uint32_t b_lo = b_lo0 + t * (B_TAP_BYTES >> 4);
              const uint32_t first = (started >> acc) & 1u;
#pragma unroll
              for (int g = 0; g < GT; g++) {
                if (p.debug & 4) break;
#pragma unroll
                for (int kk = 0; kk < CH / 2; kk++)
                  umma_lh_pred<KIND>(leader, tmem_d + (g * NPH + acc) * BN, a_lo + g * TILE_W + kk * kk_u, a_hi, b_lo + kk * (2 * BN), b_hi, IDESC,
                                     first | (uint32_t)kk);
              }
              started |= 1u << acc;
            }
            umma_commit_pred(leader, smem_u32(&b_empty[sb]));
            if (++sb == SB) { sb = 0; phb ^= 1; }
          }
          umma_commit_pred(leader, smem_u32(&a_empty[sa]));
          if (++sa == SA) { sa = 0; pha ^= 1; }
        }
        umma_commit_pred(leader, smem_u32(&acc_full[buf]));
      }
    }
  } else if (warp == 5) {
    // =========================== weight loader ===========================
    if (lane == 0) {
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
            if (p.debug & 8) {
              mbar_arrive(bar);
            } else {
              mbar_arrive_expect_tx(bar, p.tps * B_TAP_BYTES);
              const uint32_t dst = smem_u32(b_base + sb * (p.tps * B_TAP_BYTES));
              for (int t = 0; t < p.tps; t++)
                bulk_copy_g2s(dst + t * B_TAP_BYTES, wcb + (int64_t)(tap0 + t) * cb128 * (BN * 128), B_TAP_BYTES, bar);
            }
            if (++sb == SB) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 0-3) ===========================
    const int m = threadIdx.x;                         // TMEM lane = tile row
    const float alpha = d.alpha, gain = d.gain, clamp = d.clamp;
    const bool has_act = d.act != 0;                   // only linear (alpha = 1) and lrelu reach this kernel
    uint8_t* stg = smem + p.stg_off + warp * (32 * PITCH);
    const int q_chunk = lane % LPP;                    // store phase: 16-byte chunk of the slab / first pixel of the warp
    const int q_pix = lane / LPP;
    int li = 0;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, li++) {
      int ntile, u0, x0;
      decode_tile(p, GT, tile, ntile, u0, x0);
      const int buf = (NBUF == 2) ? (li & 1) : 0;
      const uint32_t fph = (NBUF == 2) ? ((li >> 1) & 1) : (li & 1);
      const int o_base = ntile * BN;
      const int n0 = u0 / p.VR, r0 = u0 - n0 * p.VR;
      mbar_wait(smem_u32(&acc_full[buf]), fph);
      tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < NACC; a++) {
        if (p.debug & 16) break;
        const int g = a / NPH, ph = a - g * NPH;
        const int py = ph >> 1, px = ph & 1;
        // output pixel of tile row mm (this thread's own row for the maths, other rows in the store phase)
        auto out_pixel = [&](int mm, int& n, int& oy, int& ox) -> bool {
          int r = r0 + (mm >> 3);
          n = n0;
          if (MODE != 1) { while (r >= p.VR) { r -= p.VR; n++; } }
          const int cx = x0 + g * TILE_W + (mm & 7);
          if (MODE == 2) { oy = 2 * r + py - d.pad_y; ox = 2 * cx + px - d.pad_x; }
          else { oy = r; ox = cx; }
          return n < d.n && oy >= 0 && oy < d.out_h && ox >= 0 && ox < d.out_w;
        };
        int n, oy, ox;
        const bool row_ok = out_pixel(m, n, oy, ox);
        const float* out_scale = (d.out_scale && row_ok) ? (const float*)d.out_scale + (int64_t)n * d.co : nullptr;
        const float nz = (d.noise && row_ok) ? ((const float*)d.noise)[((int64_t)n * d.out_h + oy) * d.out_w + ox] : 0.f;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (NACC * BN) + a * BN;
        // store phase addressing: pixel q_pix + PPI * k of this warp, k < NQ
        int64_t yoff[NQ];
#pragma unroll
        for (int k = 0; k < NQ; k++) {
          int n2, oy2, ox2;
          const bool ok2 = out_pixel(warp * 32 + q_pix + PPI * k, n2, oy2, ox2);
          yoff[k] = ok2 ? ((int64_t)n2 * d.y_strides[0] + (int64_t)oy2 * d.y_strides[2] + (int64_t)ox2 * d.y_strides[3]) : -1;
        }
        T* yrow = (T*)p.y + (int64_t)n * d.y_strides[0] + (int64_t)oy * d.y_strides[2] + (int64_t)ox * d.y_strides[3];
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
            if (out_scale || d.noise || has_act) {             // fused modulated-conv / bias_act epilogue
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const int o = o_base + cc + e;
                float v = val[e];
                if (o < d.co) {
                  if (out_scale) v *= __ldg(out_scale + o);
                  v += nz;
                  if (has_act) {
                    if (d.bias) v += to_acc<T>(((const T*)d.bias)[o]);
                    v = (v > 0.f ? v : v * alpha) * gain;
                    if (clamp >= 0.f) v = v < -clamp ? -clamp : (v > clamp ? clamp : v);      // explicit compares: NaN propagates like torch.clamp
                  }
                }
                val[e] = v;
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
                if (yoff[k] >= 0 && !(p.debug & 1)) {
                  const uint4 v = *(const uint4*)(stg + (q_pix + PPI * k) * PITCH + q_chunk * 16);
                  *(uint4*)((T*)p.y + yoff[k] + o) = v;
                }
              }
            }
            __syncwarp();
          }
        }
      }
      // all TMEM reads of this warp are complete (tmem_ld16 waits): hand the buffer back to the MMA warp
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

// ---- launcher (one per instantiation) ------------------------------------------------------------------------
template <class T, int KIND, int BN, int MODE, int GT>
int launch_halo(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  constexpr int TC = 16 / sizeof(T);
  constexpr int CH = HALO_CH;
  constexpr int SLAB = (BN * (int)sizeof(T) >= 128) ? 128 / (int)sizeof(T) : BN;
  constexpr int PITCH = SLAB * (int)sizeof(T) + 16;
  HaloParams p; p.d = *d; p.x = x; p.y = y; p.wpack = d->workspace;
  p.QP = 0;
  const int TW = TILE_W * GT;
  if (MODE == 0) {
    p.VR = d->out_h + d->kh - 1;
    p.HR = TILE_H + d->kh - 1; p.HC = TW + d->kw - 1;
    p.top = d->transposed ? (d->kh - 1 - d->pad_y) : d->pad_y;
    p.left = d->transposed ? (d->kw - 1 - d->pad_x) : d->pad_x;
    p.col_tiles = (d->out_w + TW - 1) / TW;
  } else if (MODE == 1) {
    p.HR = 2 * (TILE_H - 1) + d->kh; p.HC = 2 * (TW - 1) + d->kw;
    p.QP = (p.HC + 1) / 2;
    p.top = d->pad_y; p.left = d->pad_x;
    p.VR = (d->out_h + TILE_H - 1) / TILE_H * TILE_H;
    p.col_tiles = (d->out_w + TW - 1) / TW;
  } else {
    p.top = (d->kh - 1) >> 1; p.left = (d->kw - 1) >> 1;
    p.HR = TILE_H + p.top; p.HC = TW + p.left;
    // virtual rows per image (tiles straddle images as in MODE 0): the last one is an all-zero input row, so that the
    // patch row it shares with the next image's row -1 is zero for both
    const int rows_img = ((d->out_h - 1 + d->pad_y) >> 1) + 1;
    p.VR = rows_img > d->in_h + 1 ? rows_img : d->in_h + 1;
    p.col_tiles = ((((d->out_w - 1 + d->pad_x) >> 1) + 1) + TW - 1) / TW;
  }
  for (int tap = 0; tap < d->kh * d->kw; tap++) {
    const int ky = tap / d->kw, kx = tap - ky * d->kw;
    p.tap_acc[tap] = 0;
    if (MODE == 0) {
      const int pr = d->transposed ? (d->kh - 1 - ky) : ky, pc = d->transposed ? (d->kw - 1 - kx) : kx;
      p.tap_aoff[tap] = pr * p.HC + pc;
    } else if (MODE == 1) {
      p.tap_aoff[tap] = ky * p.HC + (kx & 1) * p.QP + (kx >> 1);
    } else {
      p.tap_aoff[tap] = (p.top - (ky >> 1)) * p.HC + (p.left - (kx >> 1));
      p.tap_acc[tap] = (ky & 1) * 2 + (kx & 1);
    }
  }
  const int64_t row_tiles = (MODE != 1) ? ceil_div((int64_t)d->n * p.VR, TILE_H) : (int64_t)d->n * (p.VR / TILE_H);
  p.ntiles = (d->co + BN - 1) / BN;
  p.total_tiles = row_tiles * p.col_tiles * p.ntiles;
  p.taps = d->kh * d->kw;
  p.cblocks = (d->ci + CH * TC - 1) / (CH * TC);
  int npix = p.HR * p.HC;
  SGB_REQUIRE(npix * CH <= halo_max_slots(MODE, GT) * HALO_PRODUCERS, "patch too large");
  while (npix % 8 != 1) npix++;                       // chunk planes 16 B (mod 128 B) apart: conflict-free 128-bit stores
  p.lbo = npix * 16;
  p.a_stage_bytes = (CH * p.lbo + 127) / 128 * 128;
  const bool y_al = aligned16(y) && d->y_strides[0] % TC == 0 && d->y_strides[2] % TC == 0 && d->y_strides[3] % TC == 0;
  p.vec_store = (d->co % TC == 0 && y_al) ? 1 : 0;
  SGB_REQUIRE(aligned16(x) && aligned16(d->workspace), "x and workspace must be 16-byte aligned");
  SGB_REQUIRE(d->act == 0 || d->act == SGB_ACT_LINEAR || d->act == SGB_ACT_LRELU, "fused epilogue supports linear and lrelu only");
  if (d->act == SGB_ACT_LINEAR) p.d.alpha = 1.f;
  static const int dbg = [] { const char* e = getenv("SGB_HALO_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = dbg;
  if (int r = pack_weights_umma(d, w, BN, s)) return r;
  // shared-memory plan: staging slabs; weight stages of `tps` taps (largest divisor of taps that leaves room: fewer
  // mbarrier round trips per MMA) with enough bytes in flight to cover the L2 latency; the rest for patch stages
  const int budget = 225 * 1024;
  const int stg_bytes = 4 * 32 * PITCH;
  const int b_tap = BN * CH * 16;
  int sa = 0, sb = 0, tps = 1;
  for (int cand = p.taps; cand >= 1; cand--) {
    if (p.taps % cand) continue;
    const int b_stage_c = cand * b_tap;
    int sb_c = (64 * 1024 + b_stage_c - 1) / b_stage_c;              // >= 64 KB of weights in flight ...
    if (sb_c < 2) sb_c = 2;                                          // ... and at least double buffering
    if (sb_c > MAX_SB) sb_c = MAX_SB;
    const int sa_c = (budget - stg_bytes - sb_c * b_stage_c) / p.a_stage_bytes;
    if (sa_c >= 3 || cand == 1) { tps = cand; sb = sb_c; sa = sa_c > MAX_SA ? MAX_SA : sa_c; break; }
  }
  SGB_REQUIRE(sa >= 2 && sb >= 2, "shared memory budget exceeded");
  const int b_stage = tps * b_tap;
  {                                                                  // left-over space: more weight stages
    const int extra = (budget - stg_bytes - sa * p.a_stage_bytes - sb * b_stage) / b_stage;
    sb = (sb + extra > MAX_SB) ? MAX_SB : sb + extra;
  }
  p.sa = sa; p.sb = sb; p.tps = tps;
  p.stg_off = sa * p.a_stage_bytes + sb * b_stage;
  const size_t smem = (size_t)p.stg_off + stg_bytes + 1024;
  auto kern = conv_halo_kernel<T, KIND, BN, MODE, GT>;
  SGB_SET_MAX_SMEM(kern, 226 * 1024);
  const int64_t grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kern<<<(unsigned)grid, HALO_THREADS, smem, s>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

// (BN, MODE, GT) -> launcher; GT combinations limited by TMEM (4*GT*BN <= 512 for MODE 2, GT*BN <= 512 else)
template <class T, int KIND>
int dispatch_halo(int bn, int mode, int gt, const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
#define SGB_HALO_CASE(BN_, MODE_, GT_) \
  if (bn == BN_ && mode == MODE_ && gt == GT_) return launch_halo<T, KIND, BN_, MODE_, GT_>(d, x, w, y, s);
  SGB_HALO_CASE(16, 0, 1) SGB_HALO_CASE(32, 0, 1) SGB_HALO_CASE(64, 0, 1) SGB_HALO_CASE(128, 0, 1) SGB_HALO_CASE(256, 0, 1)
  SGB_HALO_CASE(16, 0, 2) SGB_HALO_CASE(32, 0, 2) SGB_HALO_CASE(64, 0, 2) SGB_HALO_CASE(128, 0, 2) SGB_HALO_CASE(256, 0, 2)
  SGB_HALO_CASE(16, 0, 4) SGB_HALO_CASE(32, 0, 4) SGB_HALO_CASE(64, 0, 4) SGB_HALO_CASE(128, 0, 4)
  SGB_HALO_CASE(16, 1, 1) SGB_HALO_CASE(32, 1, 1) SGB_HALO_CASE(64, 1, 1) SGB_HALO_CASE(128, 1, 1) SGB_HALO_CASE(256, 1, 1)
  SGB_HALO_CASE(16, 2, 1) SGB_HALO_CASE(32, 2, 1) SGB_HALO_CASE(64, 2, 1) SGB_HALO_CASE(128, 2, 1)
  SGB_HALO_CASE(32, 2, 2) SGB_HALO_CASE(64, 2, 2)
  SGB_HALO_CASE(32, 2, 4)
#undef SGB_HALO_CASE
  set_error("conv_halo: no kernel for this (BN, MODE, GT)");
  return 1;
}

}  // namespace sgb
