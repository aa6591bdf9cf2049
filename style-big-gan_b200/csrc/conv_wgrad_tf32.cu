// tcgen05 weight gradient for fp32 tensors: kind::tf32 on MN-major operands in the SWIZZLE_128B_BASE32B layout.
//
//   dW[o, c, ky, kx] = sum_{n, oy, ox} dy[n, oy, ox, o] * xs[n, oy*s + ky - pad, ox*s + kx - pad, c]
//
// Same decomposition as conv_wgrad_halo.cu (a CTA owns 128 output channels x BNC input channels x one filter row x a
// range of TH x 8 pixel tiles; K = output pixel; the kw taps are kw start addresses into ONE staged input patch), but
// the operands go to the tensor core as they are in memory: no conversion pass, one MMA per K step.
//
// kind::tf32 returns zeros for MN-major (channel-contiguous) operands in the no-swizzle layout; the layout it takes is
// descriptor layout type 1, SWIZZLE_128B_BASE32B (established with benchmarks/experiments/mnmajor_tf32.cu):
//   * a 32-channel block of the operand is a stack of 128-byte rows, one row per K index (= pixel), rows contiguous;
//   * inside a row the four 32-byte units are permuted: unit q of row r is stored at unit q ^ (r & 3), where r is taken
//     from the ABSOLUTE shared-memory address (bits 7-8), so a descriptor whose start address is moved by whole rows
//     (the filter tap = a pixel shift) stays consistent with data written once;
//   * LBO = bytes between 32-channel blocks, SBO = 512 (four rows), one MMA consumes K = 8 rows = one tile row.
// Stage = [dy tile: PA blocks x (TH*8 rows)][x patch: BNC/32 blocks x (TH*HC rows)], written by cp.async (zero fill
// outside the image) straight from the channels_last tensors.  Stride 2: the patch columns are stored de-interleaved
// by parity, so that the 8 output pixels of a tile row are consecutive rows for every tap.
// Accumulators (kw x BNC columns) stay in TMEM across all tiles of the CTA; one epilogue adds them to dW (zeroed by
// the launcher) with red.global.add.f32.
// Filter rows stacked in M (kys; stride 1, 3 x 3, <= 64 dy and <= 64 x channels -- the layers whose weight gradient is bound by
// the L2 -> shared-memory stream, because every filter-row CTA re-reads dy and x): the x patch gets kh - 1 extra rows
// (row r <-> input row oy0 - pad + r) and ONE K step per patch row r pairs it with the dy rows oy0 + r - 2 + j, j = 0..3, as
// the four 32-channel M blocks of one MMA: the A descriptor's LBO is one tile row (8 pixels = 1 KB), so M block j is the same
// dy channel block one tile row further down, i.e. filter row ky = 2 - j (j = 3 lands in accumulator lanes that are never
// stored).  The dy stage carries two permanent zero tile rows above and three below the TH loaded rows, so that products
// with rows outside the tile vanish instead of being counted twice.  Per tile row: (dy blocks) x (x blocks) MMAs of
// M = 128, N = 96 (the three kx taps) for all nine taps, against 3 CTAs x (x blocks) MMAs before, with a third of the loads.
// Warps 0-7: cp.async producers (+ in-place style scaling of their own chunks), warps 0-3 then run the epilogue.
// Warp 8: MMA issuer (one lane), owns TMEM.
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

namespace sgb {

constexpr int WT_PRODUCERS = 256;
constexpr int WT_THREADS = WT_PRODUCERS + 32;
constexpr int WT_MAX_STAGES = 8;

struct WgradTf32Params {
  sgb_conv_desc d;
  const float* x; const float* dy; float* dw;
  int TH;                     // tile rows; tile = TH x 8 output pixels
  int row_tiles, col_tiles;   // per image
  int64_t total_tiles;        // n * row_tiles * col_tiles
  int64_t chunk_tiles;        // tiles per split
  int ctiles;
  int HC;                     // patch column slots per row = 7*s + kw
  int QP;                     // slot offset of the odd-column plane (stride 2), 0 for stride 1
  int a_blk, b_blk;           // bytes per 32-channel block (rows * 128, multiples of 512)
  int a_bytes, stage_bytes;
  int stages, lookahead;
  int m64;                    // 1: <= 64 dy channels, M = 64 MMAs (accumulator row r sits in TMEM lane (r % 16) + 32 * (r / 16))
  int stack;                  // 1: the kw taps of a 32-channel block are one MMA (N = 96, LBO = one patch row)
  int kys;                    // 1: the kh filter rows are stacked in M as well (see "filter rows stacked in M" below): one CTA
                              //    computes all nine taps from ONE staging of dy and x
  int pa;                     // kys: 32-channel blocks of dy
  int RB;                     // x patch rows per tile: TH (+ kh - 1 with kys)
};

// byte offset of 16-byte chunk c (0..7) inside row `row` of a block whose base is 512-byte aligned
__device__ __forceinline__ uint32_t swz_row(int row, int c) {
  return (uint32_t)(row * 128 + ((((c >> 1) ^ (row & 3)) << 5) | ((c & 1) << 4)));
}

template <int BNC>
__global__ void __launch_bounds__(WT_THREADS, 1) conv_wgrad_tf32_kernel(WgradTf32Params p) {
  // instruction descriptor: fp32 accumulate, tf32 operands, both MN-major, N >> 3 at [17,23), M >> 4 at [24,29)
  auto idesc_of = [](uint32_t m, uint32_t n) { return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((m >> 4) << 24); };

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the swizzle is a function of the absolute address: blocks must start on 512-byte boundaries (1 KB of slack is allocated)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full_bar[WT_MAX_STAGES], empty_bar[WT_MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const sgb_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int otile = blockIdx.x;
  const int ky = p.kys ? 0 : blockIdx.y / p.ctiles, ctile = blockIdx.y - ky * p.ctiles;
  const int64_t t_begin = (int64_t)blockIdx.z * p.chunk_tiles;
  const int64_t t_end = (t_begin + p.chunk_tiles < p.total_tiles) ? t_begin + p.chunk_tiles : p.total_tiles;
  const int ntiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;
  const int SA = p.stages;
  const int s = d.stride;
  const int o0 = otile * UM, c0 = ctile * BNC;
  const uint32_t need_cols = p.kys ? (uint32_t)(p.pa * (BNC / 32) * 96) : (uint32_t)(d.kw * BNC);
  const uint32_t tmem_cols = need_cols <= 32 ? 32u : (need_cols <= 64 ? 64u : (need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u)));

  // Channels that are never written (beyond co / ci) hold whatever the stage held before: they only reach accumulator
  // rows / columns that the epilogue does not store.  Clear the stages once anyway so that no NaN patterns are fed.
  for (int i = threadIdx.x * 16; i < SA * p.stage_bytes; i += WT_THREADS * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
  if (warp == WT_PRODUCERS / 32) {
    if (lane == 0) {
      for (int i = 0; i < WT_MAX_STAGES; i++) { mbar_init(smem_u32(&full_bar[i]), WT_PRODUCERS / 32); mbar_init(smem_u32(&empty_bar[i]), 1); }
      mbar_init(smem_u32(&accum_bar), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < WT_PRODUCERS / 32) {
    // =========================== producers ===========================
    const int t = threadIdx.x;
    const int cow = (d.co - o0 < UM) ? d.co - o0 : UM;             // valid channels (multiples of 4)
    const int ciw = (d.ci - c0 < BNC) ? d.ci - c0 : BNC;
    const int ca = cow >> 2, cb = ciw >> 2;                        // 16-byte chunks per pixel
    int la = 0; while ((1 << la) < ca) la++;
    int lb = 0; while ((1 << lb) < cb) lb++;
    const int ja = t & ((1 << la) - 1), pa0 = t >> la, ppa = WT_PRODUCERS >> la;   // chunk / first pixel / pixels per pass
    const int jb = t & ((1 << lb) - 1), pb0 = t >> lb, ppb = WT_PRODUCERS >> lb;
    const int npa = p.TH * 8, npb = p.RB * p.HC;
    const int a_row0 = p.kys ? 16 : 0;                              // kys: two zero tile rows precede the loaded dy rows
    const int ppb_div = ppb / p.HC, ppb_mod = ppb - ppb_div * p.HC;
    const int pb0_r = pb0 / p.HC, pb0_c = pb0 - pb0_r * p.HC;
    const uint32_t a_off = (uint32_t)((ja >> 3) * p.a_blk), b_off = (uint32_t)(p.a_bytes + (jb >> 3) * p.b_blk);
    const int a_c = ja & 7, b_c = jb & 7;
    const float* xb = p.x;
    const float* dyb = p.dy;
    const float* scb = (const float*)d.in_scale;                   // per-sample scale of the x channels
    const float* sca = (const float*)d.out_scale;                  // per-sample scale of the dy channels
    const int tiles_per_img = p.row_tiles * p.col_tiles;
    const int lookahead = p.lookahead;

    // tile coordinates advance incrementally (no divisions per tile); one cursor for the loads, one for publishing
    struct Cursor { int n, oy0, ox0; };
    auto cursor_at = [&](int64_t tt) {
      Cursor c;
      c.n = (int)(tt / tiles_per_img);
      const int rem = (int)(tt - (int64_t)c.n * tiles_per_img);
      const int tr = rem / p.col_tiles;
      c.oy0 = tr * p.TH; c.ox0 = (rem - tr * p.col_tiles) * 8;
      return c;
    };
    auto advance = [&](Cursor& c) {
      c.ox0 += 8;
      if (c.ox0 >= p.col_tiles * 8) { c.ox0 = 0; c.oy0 += p.TH; if (c.oy0 >= p.row_tiles * p.TH) { c.oy0 = 0; c.n++; } }
    };
    Cursor cur_i = cursor_at(t_begin), cur_p = cur_i;

    int pub = 0, sa_p = 0;
    auto publish = [&]() {                               // the oldest unpublished tile has landed
      uint8_t* stage = smem + sa_p * p.stage_bytes;
      if (sca && ja < ca) {                              // style scale of a transposed convolution's input (the dy operand)
        const float4 sv = __ldg((const float4*)(sca + (int64_t)cur_p.n * d.co + o0 + ja * 4));
        for (int pp = pa0; pp < npa; pp += ppa) {
          float4* q = (float4*)(stage + a_off + swz_row(pp + a_row0, a_c));
          float4 v = *q;
          v.x *= sv.x; v.y *= sv.y; v.z *= sv.z; v.w *= sv.w;
          *q = v;
        }
      }
      if (scb && jb < cb) {                              // style modulation of x
        const float4 sv = __ldg((const float4*)(scb + (int64_t)cur_p.n * d.ci + c0 + jb * 4));
        int hr = pb0_r, hc = pb0_c;
        for (int hp = pb0; hp < npb; hp += ppb) {
          const int slot = (s == 1) ? hc : ((hc & 1) * p.QP + (hc >> 1));
          float4* q = (float4*)(stage + b_off + swz_row(hr * p.HC + slot, b_c));
          float4 v = *q;
          v.x *= sv.x; v.y *= sv.y; v.z *= sv.z; v.w *= sv.w;
          *q = v;
          hr += ppb_div; hc += ppb_mod;
          if (hc >= p.HC) { hc -= p.HC; hr++; }
        }
      }
      fence_proxy_async();
      __syncwarp();                                      // one arrival per warp
      if (lane == 0) mbar_arrive(smem_u32(&full_bar[sa_p]));
      if (++sa_p == SA) sa_p = 0;
      advance(cur_p);
      pub++;
    };

    int sa_i = 0;
    uint32_t ph_i = 0;
    for (int i = 0; i < ntiles; i++) {
      const int n = cur_i.n, oy0 = cur_i.oy0, ox0 = cur_i.ox0;
      advance(cur_i);
      mbar_wait(smem_u32(&empty_bar[sa_i]), ph_i ^ 1);
      const uint32_t stage = smem_u32(smem + sa_i * p.stage_bytes);
      if (++sa_i == SA) { sa_i = 0; ph_i ^= 1; }
      // dy tile: row pp = ty * 8 + tx
      if (ja < ca) {
        const float* src_n = dyb + (int64_t)n * d.y_strides[0] + o0 + ja * 4;
        const uint32_t dst_j = stage + a_off;
        for (int pp = pa0; pp < npa; pp += ppa) {
          const int oy = oy0 + (pp >> 3), ox = ox0 + (pp & 7);
          const bool ok = oy < d.out_h && ox < d.out_w;
          const float* src = src_n + (int64_t)oy * d.y_strides[2] + (int64_t)ox * d.y_strides[3];
          cp_async16(dst_j + swz_row(pp + a_row0, a_c), ok ? (const void*)src : (const void*)dyb, ok ? 16u : 0u);
        }
      }
      // x patch: row hr <-> input row (oy0 + hr) * s + ky - pad, column hc <-> input column ox0 * s - pad + hc
      if (jb < cb) {
        const float* src_n = xb + (int64_t)n * d.x_strides[0] + c0 + jb * 4;
        const uint32_t dst_j = stage + b_off;
        const int ix0 = ox0 * s - d.pad_x;
        int hr = pb0_r, hc = pb0_c;
        for (int hp = pb0; hp < npb; hp += ppb) {
          const int iy = p.kys ? (oy0 + hr - d.pad_y) : ((oy0 + hr) * s + ky - d.pad_y), ix = ix0 + hc;
          const bool ok = iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w;
          const int slot = (s == 1) ? hc : ((hc & 1) * p.QP + (hc >> 1));
          const float* src = src_n + (int64_t)iy * d.x_strides[2] + (int64_t)ix * d.x_strides[3];
          cp_async16(dst_j + swz_row(hr * p.HC + slot, b_c), ok ? (const void*)src : (const void*)xb, ok ? 16u : 0u);
          hr += ppb_div; hc += ppb_mod;
          if (hc >= p.HC) { hc -= p.HC; hr++; }
        }
      }
      cp_async_commit();
      if (i - pub >= lookahead) {                        // keep `lookahead` tiles in flight, publish the oldest
        cp_async_wait_n(lookahead);
        publish();
      }
    }
    cp_async_wait<0>();
    while (pub < ntiles) publish();

    // =========================== epilogue (warps 0-3: one TMEM lane quadrant each) ===========================
    if (ntiles > 0 && warp < 4 && p.kys) {
      // lane = (shift j = warp, channel within the dy block); columns = [dy block][x block][kx][32 channels]
      mbar_wait(smem_u32(&accum_bar), 0);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int kye = 2 - warp;
      const int wy = d.flip ? d.kh - 1 - kye : kye;
      for (int a = 0; a < p.pa; a++) {
        const int o = o0 + a * 32 + lane;
        for (int b = 0; b < BNC / 32; b++)
          for (int kx = 0; kx < 3; kx++) {
            const int wx = d.flip ? 2 - kx : kx;
#pragma unroll 1
            for (int cc = 0; cc < 32; cc += 16) {
              uint32_t acc[16];
              tmem_ld16(lane_addr + (a * (BNC / 32) + b) * 96 + kx * 32 + cc, acc);
              if (kye < 0 || o >= d.co) continue;
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const int c = c0 + b * 32 + cc + e;
                if (c < d.ci) atomicAdd(p.dw + (((int64_t)o * d.ci + c) * d.kh + wy) * d.kw + wx, __uint_as_float(acc[e]));
              }
            }
          }
      }
      tc_fence_before();
    } else if (ntiles > 0 && warp < 4) {
      mbar_wait(smem_u32(&accum_bar), 0);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int o = p.m64 ? ((lane < 16) ? o0 + warp * 16 + lane : d.co) : o0 + threadIdx.x;
      const int wy = d.flip ? d.kh - 1 - ky : ky;
      for (int kx = 0; kx < d.kw; kx++) {
        const int wx = d.flip ? d.kw - 1 - kx : kx;
#pragma unroll 1
        for (int cc = 0; cc < BNC; cc += 16) {
          uint32_t acc[16];
          // accumulator columns: [tap][channel], or per 32-channel block [tap][32 channels] when the taps are stacked in N
          tmem_ld16(lane_addr + (p.stack ? (cc >> 5) * 96 + kx * 32 + (cc & 31) : kx * BNC + cc), acc);
          if (o >= d.co) continue;
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const int c = c0 + cc + e;
            if (c < d.ci) atomicAdd(p.dw + (((int64_t)o * d.ci + c) * d.kh + wy) * d.kw + wx, __uint_as_float(acc[e]));
          }
        }
      }
      tc_fence_before();
    }
  } else {
    // =========================== MMA issuer ===========================
    // whole warp runs the loop (uniform values), lane 0 issues; descriptors as (lo, hi) halves, ring counters
    {
      const uint32_t leader = elect_one();
      const uint32_t IDESC = idesc_of(p.m64 ? 64u : 128u, (uint32_t)BNC);
      const uint32_t hi = smem_desc_hi(512) | (1u << 29);                     // SBO = 4 rows; layout type 1 (bits 61-63)
      const uint32_t a_lo_base = smem_desc_lo(smem_u32(smem), (uint32_t)p.a_blk);
      const uint32_t b_lo_base = smem_desc_lo(smem_u32(smem) + (uint32_t)p.a_bytes, (uint32_t)p.b_blk);
      const uint32_t stage_u = (uint32_t)p.stage_bytes >> 4;
      const uint32_t b_row_u = (uint32_t)(p.HC * 8);                          // one tile row down the patch, in 16-byte units
      int sa = 0;
      uint32_t pha = 0;
      for (int i = 0; i < ntiles; i++) {
        mbar_wait(smem_u32(&full_bar[sa]), pha);
        tc_fence_after();
        const uint32_t a_lo0 = a_lo_base + sa * stage_u, b_lo0 = b_lo_base + sa * stage_u;
        if (p.kys) {
          // filter rows stacked in M: A blocks = the dy channel block at tile rows r, r + 1, r + 2, r + 3 (LBO = one tile row)
          const uint32_t IDESC96 = idesc_of(128u, 96u);
          const uint32_t ak_lo0 = smem_desc_lo(smem_u32(smem), 1024) + sa * stage_u;
          const uint32_t bs_lo0 = smem_desc_lo(smem_u32(smem) + (uint32_t)p.a_bytes, 128) + sa * stage_u;
          for (int a = 0; a < p.pa; a++)
            for (int b = 0; b < BNC / 32; b++) {
              const uint32_t a_lo = ak_lo0 + a * ((uint32_t)p.a_blk >> 4);
              const uint32_t b_lo = bs_lo0 + b * ((uint32_t)p.b_blk >> 4);
              const uint32_t tm = tmem_base + (a * (BNC / 32) + b) * 96;
#pragma unroll 2
              for (int r = 0; r < p.RB; r++)
                umma_lh_pred<2>(leader, tm, a_lo + r * 64, hi, b_lo + r * b_row_u, hi, IDESC96, (i > 0 || r > 0) ? 1u : 0u);
            }
          umma_commit_pred(leader, smem_u32(&empty_bar[sa]));
        } else if (p.stack) {                         // warp-uniform issue (umma.cuh: elect_one / umma_lh_pred)
          // stride 1, three taps: tap kx of a 32-channel block is the same block one patch row further down, i.e. the
          // next N block of a descriptor whose LBO is one row (128 B).  One N = 96 MMA per block and tile row reads
          // the dy tile once for the three taps instead of three times.
          const uint32_t IDESC96 = idesc_of(p.m64 ? 64u : 128u, 96u);
          const uint32_t bs_lo0 = smem_desc_lo(smem_u32(smem) + (uint32_t)p.a_bytes, 128) + sa * stage_u;
          for (int cbk = 0; cbk < BNC / 32; cbk++) {
            const uint32_t b_lo = bs_lo0 + cbk * ((uint32_t)p.b_blk >> 4);
            const uint32_t tm = tmem_base + cbk * 96;
#pragma unroll 4
            for (int ty = 0; ty < p.TH; ty++)
              umma_lh_pred<2>(leader, tm, a_lo0 + ty * 64, hi, b_lo + ty * b_row_u, hi, IDESC96, (i > 0 || ty > 0) ? 1u : 0u);
          }
          umma_commit_pred(leader, smem_u32(&empty_bar[sa]));
        } else {
          for (int kx = 0; kx < d.kw; kx++) {
            const uint32_t b_lo = b_lo0 + (uint32_t)(((kx % s) * p.QP + kx / s) * 8);
            const uint32_t tm = tmem_base + kx * BNC;
#pragma unroll 4
            for (int ty = 0; ty < p.TH; ty++)
              umma_lh_pred<2>(leader, tm, a_lo0 + ty * 64, hi, b_lo + ty * b_row_u, hi, IDESC, (i > 0 || ty > 0) ? 1u : 0u);
          }
          umma_commit_pred(leader, smem_u32(&empty_bar[sa]));
        }
        if (++sa == SA) { sa = 0; pha ^= 1; }
      }
      if (ntiles > 0) umma_commit_pred(leader, smem_u32(&accum_bar));
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WT_PRODUCERS / 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- host side -------------------------------------------------------------------------------------------
template <int BNC>
static int launch_wgrad_tf32(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
  WgradTf32Params p; p.d = *d; p.x = (const float*)x; p.dy = (const float*)dy; p.dw = dw;
  const int s = d->stride;
  p.HC = 7 * s + d->kw;
  p.QP = (s == 1) ? 0 : (p.HC + 1) / 2;
  p.ctiles = (d->ci + BNC - 1) / BNC;
  static const int env_stack = [] { const char* e = getenv("SGB_WGRAD_STACK"); return e ? atoi(e) : 1; }();
  p.stack = (env_stack && s == 1 && d->kw == 3 && BNC <= 64) ? 1 : 0;
  const int otiles = (d->co + UM - 1) / UM;
  const int budget = 224 * 1024;
  static const int env_th = [] { const char* e = getenv("SGB_WGRAD_TH"); return e ? atoi(e) : 0; }();
  static const int env_la = [] { const char* e = getenv("SGB_WGRAD_LA"); return e ? atoi(e) : 0; }();
  // <= 64 dy channels: two 32-channel blocks and M = 64 MMAs (half the shared-memory reads of the dy operand)
  static const int env_m64 = [] { const char* e = getenv("SGB_WGRAD_M64"); return e ? atoi(e) : 1; }();
  static const int env_kys = [] { const char* e = getenv("SGB_WGRAD_KYS"); return e ? atoi(e) : 1; }();
  p.kys = (env_kys && p.stack && d->kh == 3 && d->co <= 64 && p.ctiles == 1 && otiles == 1) ? 1 : 0;
  p.pa = (d->co + 31) / 32;
  p.m64 = (env_m64 && d->co <= 64 && !p.kys) ? 1 : 0;
  // without M = 64 the M = 128 MMA reads four blocks LBO apart: the last two alias the x patch (inside the stage as long
  // as it is at least as large) and only reach accumulator rows that are never stored
  const int PA = p.kys ? p.pa : ((d->co <= 64 && (p.m64 || BNC >= 64)) ? 2 : 4);
  int TH = (env_th == 2 || env_th == 4 || env_th == 8 || env_th == 16 || env_th == 32 || (p.kys && env_th > 0 && env_th <= 32)) ? env_th : 16, stages = 0;
  for (;;) {
    // kys: TH + 5 tile rows of dy (2 zero rows above, 3 below the loaded ones), TH + 2 patch rows of x
    p.RB = p.kys ? TH + 2 : TH;
    p.a_blk = (p.kys ? TH + 5 : TH) * 8 * 128;
    p.b_blk = (p.RB * p.HC + 3) / 4 * 4 * 128;
    p.a_bytes = PA * p.a_blk;
    p.stage_bytes = (p.a_bytes + (BNC / 32) * p.b_blk + 1023) / 1024 * 1024;
    stages = budget / p.stage_bytes; if (stages > WT_MAX_STAGES) stages = WT_MAX_STAGES;
    if (stages >= 3 || TH == 2) break;
    TH = (p.kys && TH > 4) ? TH - 4 : TH >> 1;        // kys: 16, 12, 8, 4, 2 (the halo rows cost (TH + 2) / TH in MMAs and x loads)
  }
  SGB_REQUIRE(stages >= 2, "wgrad tf32: tile does not fit shared memory");
  p.TH = TH; p.stages = stages;
  p.lookahead = stages - 2 < 1 ? 1 : stages - 2;
  if (env_la > 0 && env_la < stages) p.lookahead = env_la;
  if (p.lookahead > 7) p.lookahead = 7;
  p.row_tiles = (d->out_h + TH - 1) / TH; p.col_tiles = (d->out_w + 7) / 8;
  p.total_tiles = (int64_t)d->n * p.row_tiles * p.col_tiles;
  const int kgroups = p.kys ? 1 : d->kh;
  const int64_t base = (int64_t)otiles * p.ctiles * kgroups;
  int64_t splits = num_sms() / base; if (splits < 1) splits = 1;
  if (splits > p.total_tiles) splits = p.total_tiles;
  p.chunk_tiles = ceil_div(p.total_tiles, splits);
  splits = ceil_div(p.total_tiles, p.chunk_tiles);
  SGB_REQUIRE((int64_t)p.ctiles * kgroups <= 65535 && splits <= 65535, "problem too large for the wgrad tf32 grid");
  SGB_REQUIRE(aligned16(x) && aligned16(dy), "x and dy must be 16-byte aligned");
  const size_t smem = (size_t)stages * p.stage_bytes + 1024;
  auto kern = conv_wgrad_tf32_kernel<BNC>;
  SGB_SET_MAX_SMEM(kern, 226 * 1024);
  kern<<<dim3((unsigned)otiles, (unsigned)(p.ctiles * kgroups), (unsigned)splits), WT_THREADS, smem, st>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

// dw must be zeroed by the caller
int conv_wgrad_tf32(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t s) {
  if (d->ci <= 32) return launch_wgrad_tf32<32>(d, x, dy, dw, s);
  if (d->ci <= 64) return launch_wgrad_tf32<64>(d, x, dy, dw, s);
  return launch_wgrad_tf32<128>(d, x, dy, dw, s);
}

}  // namespace sgb
