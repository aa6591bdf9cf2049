// tcgen05 weight-gradient kernel, "halo tile" form (channels_last, stride 1 or 2, 3x3 and 1x1 filters).
//
//   dW[o, c, ky, kx] = sum_{n, oy, ox} dy[n, oy, ox, o] * xs[n, oy*s + ky - pad, ox*s + kx - pad, c]
//
// A GEMM per filter tap whose K dimension is the output pixel: D_tap[o, c] = sum_p A[o, p] * B_tap[p, c] with
// A = dy and B_tap = x shifted by the tap.  Both operands are "MN-major" for the tensor core (the channel is the
// contiguous dimension in memory), so they are staged in the canonical no-swizzle MN-major layout
//     [16-byte channel chunk j][pixel][16 B]          (SBO = plane stride between chunks, LBO = stride between
//                                                      groups of 8 pixels)
// conv_umma.cu's first wgrad kernel re-gathers x and dy once per tap (9x the traffic through LSU and L2).  Here a CTA
// owns (128 output channels) x (BNC input channels) x (one filter ROW ky = kw taps) x (a range of pixel tiles):
//   * a pixel tile is TH x 8 output pixels of one image (TH = 4 / 8 / 16 chosen by the launcher to fit >= 3 stages);
//   * per tile the CTA stages the dy tile and ONE input patch (TH rows x (7*s + kw) columns) with cp.async (zero
//     fill outside the image); the kw taps of the row are kw B descriptors into the same patch (start address moves
//     by one pixel = 16 B), accumulating into kw TMEM accumulators of BNC columns;
//   * stride 2: the patch columns are stored de-interleaved by parity, so that 8 consecutive output pixels of a row
//     are still 16 B apart for every tap;
//   * accumulators stay in TMEM across ALL tiles of the CTA; a single epilogue adds them to dW with
//     red.global.add.f32 (dW zeroed by the launcher).
// Warps 0-3: cp.async producers (+ optional in-place style scaling of the patch), then the epilogue.
// Warp 4: MMA issuer (one lane), owns TMEM.
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

namespace sgb {

int conv_wgrad_tf32(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t s);   // conv_wgrad_tf32.cu
bool conv_wgrad_kys_eligible(const sgb_conv_desc* d);                                                   // conv_wgrad_kys.cu
int conv_wgrad_kys(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t s);

constexpr int WG_THREADS = 160;
constexpr int WG_LOOKAHEAD = 1;          // tiles in flight per producer thread beyond the one being published
constexpr int NUM_PRODUCERS_WG = 128;
constexpr int WS_PRODUCERS = 256;         // fp32 split kernel: 8 producer / converter warps
constexpr int WS_THREADS = WS_PRODUCERS + 32;

struct WgradHaloParams {
  sgb_conv_desc d;
  const void* x; const void* dy; float* dw;
  int TH;                     // tile rows; tile = TH x 8 output pixels
  int row_tiles, col_tiles;   // per image
  int64_t total_tiles;        // n * row_tiles * col_tiles
  int64_t chunk_tiles;        // tiles per split
  int ctiles;
  int HC;                     // patch column slots per row = 7*s + kw
  int QP;                     // slot offset of the odd-column plane (stride 2), 0 for stride 1
  int a_plane, b_plane;       // bytes between channel chunks
  int a_bytes, stage_bytes;
  int stages;
  int lookahead;              // halo kernel: tiles in flight per producer thread beyond the one being published
  int nstg;                   // fp32 split kernel: staging buffers (tiles in flight from L2 / HBM = nstg - 1)
};

template <class T, int KIND, int BNC>
__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_halo_kernel(WgradHaloParams p) {
  constexpr int TC = 16 / sizeof(T);
  constexpr int KPM = 32 / (int)sizeof(T);           // pixels per MMA (K = 32 bytes)
  constexpr uint32_t IDESC = make_idesc(KIND, BNC, 1);
  constexpr int MAX_STAGES = 8;

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const sgb_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int otile = blockIdx.x;
  const int ky = blockIdx.y / p.ctiles, ctile = blockIdx.y - ky * p.ctiles;
  const int64_t t_begin = (int64_t)blockIdx.z * p.chunk_tiles;
  const int64_t t_end = (t_begin + p.chunk_tiles < p.total_tiles) ? t_begin + p.chunk_tiles : p.total_tiles;
  const int ntiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;
  const int SA = p.stages;
  const int s = d.stride;
  const int o0 = otile * UM, c0 = ctile * BNC;
  const uint32_t tmem_cols = (d.kw * BNC <= 128) ? 128u : ((d.kw * BNC <= 256) ? 256u : 512u);

  // channel chunks that are never written (beyond co / ci) must read as zero: clear the stages once
  for (int i = threadIdx.x * 16; i < SA * p.stage_bytes; i += WG_THREADS * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
  if (warp == 4) {
    if (lane == 0) {
      for (int i = 0; i < MAX_STAGES; i++) { mbar_init(smem_u32(&full_bar[i]), NUM_PRODUCERS_WG / 32); mbar_init(smem_u32(&empty_bar[i]), 1); }
      mbar_init(smem_u32(&accum_bar), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols);
  }
  fence_proxy_async();          // the zero fill above is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 4) {
    // =========================== producers ===========================
    const int t = threadIdx.x;
    const int cow = (d.co - o0 < UM) ? d.co - o0 : UM;             // valid channels of the dy tile (multiple of TC)
    const int ciw = (d.ci - c0 < BNC) ? d.ci - c0 : BNC;           // valid channels of the x slice
    const int cpa = cow / TC, cpb = ciw / TC;
    int la = 0; while ((1 << la) < cpa) la++;                      // chunk lanes padded to a power of two (<= 32)
    int lb = 0; while ((1 << lb) < cpb) lb++;
    const int ja = t & ((1 << la) - 1), pa0 = t >> la, ppa = 128 >> la;   // chunk / first pixel / pixels per pass (dy tile)
    const int jb = t & ((1 << lb) - 1), pb0 = t >> lb, ppb = 128 >> lb;   // same for the x patch
    const int npa = p.TH * 8, npb = p.TH * p.HC;
    const int ppb_div = ppb / p.HC, ppb_mod = ppb - ppb_div * p.HC;
    const int pb0_r = pb0 / p.HC, pb0_c = pb0 - pb0_r * p.HC;
    const T* xb = (const T*)p.x;
    const T* dyb = (const T*)p.dy;
    const float* scb = (const float*)d.in_scale;
    const float* sca = (const float*)d.out_scale;                  // per-sample scale of the dy channels
    const int tiles_per_img = p.row_tiles * p.col_tiles;
    int pub = 0;

    // tile coordinates advance incrementally: one cursor for the loads, one for publishing (the 64-bit division of a
    // per-tile decode was a dependent ~1 K cycle chain in every producer thread, once per pipeline step)
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

    auto publish = [&](int i) {
      const int sa = i % SA;
      const int n = cur_p.n;
      advance(cur_p);
      if (scb && jb < cpb) {                                        // in-place style scaling of the chunks this thread copied
        const float* sp = scb + (int64_t)n * d.ci + c0 + jb * TC;
        float sv[TC];
#pragma unroll
        for (int e = 0; e < TC; e += 4) { const float4 q = __ldg((const float4*)(sp + e)); sv[e] = q.x; sv[e + 1] = q.y; sv[e + 2] = q.z; sv[e + 3] = q.w; }
        uint8_t* bdst = smem + sa * p.stage_bytes + p.a_bytes + jb * p.b_plane;
        int hr = pb0_r, hc = pb0_c;
        for (int hp = pb0; hp < npb; hp += ppb) {
          const int slot = (s == 1) ? hc : ((hc & 1) * p.QP + (hc >> 1));
          uint4* q = (uint4*)(bdst + (hr * p.HC + slot) * 16);
          uint4 v = *q;
          if (KIND == 2) {
            float* f = (float*)&v;
#pragma unroll
            for (int e = 0; e < 4; e++) f[e] *= sv[e % TC];
          } else {
            T* h = (T*)&v;
#pragma unroll
            for (int e = 0; e < TC; e++) h[e] = from_acc<T>(to_acc<T>(h[e]) * sv[e]);
          }
          *q = v;
          hr += ppb_div; hc += ppb_mod;
          if (hc >= p.HC) { hc -= p.HC; hr++; }
        }
      }
      if (sca && ja < cpa) {                                        // same for the dy tile (out_scale)
        const float* sp = sca + (int64_t)n * d.co + o0 + ja * TC;
        float sv[TC];
#pragma unroll
        for (int e = 0; e < TC; e += 4) { const float4 q = __ldg((const float4*)(sp + e)); sv[e] = q.x; sv[e + 1] = q.y; sv[e + 2] = q.z; sv[e + 3] = q.w; }
        uint8_t* adst = smem + sa * p.stage_bytes + ja * p.a_plane;
        for (int pp = pa0; pp < npa; pp += ppa) {
          uint4* q = (uint4*)(adst + pp * 16);
          uint4 v = *q;
          if (KIND == 2) {
            float* f = (float*)&v;
#pragma unroll
            for (int e = 0; e < 4; e++) f[e] *= sv[e % TC];
          } else {
            T* h = (T*)&v;
#pragma unroll
            for (int e = 0; e < TC; e++) h[e] = from_acc<T>(to_acc<T>(h[e]) * sv[e]);
          }
          *q = v;
        }
      }
      fence_proxy_async();
      __syncwarp();                                   // one arrival per warp: 128 serialised arrivals per tile were a fixed cost
      if (lane == 0) mbar_arrive(smem_u32(&full_bar[sa]));
    };

    int sa_i = 0;
    uint32_t ph_i = 0;
    for (int i = 0; i < ntiles; i++) {
      const int sa = sa_i;
      const int n = cur_i.n, oy0 = cur_i.oy0, ox0 = cur_i.ox0;
      advance(cur_i);
      mbar_wait(smem_u32(&empty_bar[sa]), ph_i ^ 1);
      if (++sa_i == SA) { sa_i = 0; ph_i ^= 1; }
      const uint32_t a_dst = smem_u32(smem + sa * p.stage_bytes);
      const uint32_t b_dst = a_dst + p.a_bytes;
      // dy tile: pixel pp = ty * 8 + tx
      if (ja < cpa) {
        const T* src_n = dyb + (int64_t)n * d.y_strides[0] + o0 + ja * TC;
        const uint32_t dst_j = a_dst + ja * p.a_plane;
        for (int pp = pa0; pp < npa; pp += ppa) {
          const int oy = oy0 + (pp >> 3), ox = ox0 + (pp & 7);
          const bool ok = oy < d.out_h && ox < d.out_w;
          const T* src = src_n + (int64_t)oy * d.y_strides[2] + (int64_t)ox * d.y_strides[3];
          cp_async16(dst_j + pp * 16, ok ? (const void*)src : (const void*)dyb, ok ? 16u : 0u);
        }
      }
      // x patch: row hr <-> input row (oy0 + hr) * s + ky - pad, column hc <-> input column ox0 * s - pad + hc
      if (jb < cpb) {
        const T* src_n = xb + (int64_t)n * d.x_strides[0] + c0 + jb * TC;
        const uint32_t dst_j = b_dst + jb * p.b_plane;
        const int ix0 = ox0 * s - d.pad_x;
        int hr = pb0_r, hc = pb0_c;
        for (int hp = pb0; hp < npb; hp += ppb) {
          const int iy = (oy0 + hr) * s + ky - d.pad_y, ix = ix0 + hc;
          const bool ok = iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w;
          const int slot = (s == 1) ? hc : ((hc & 1) * p.QP + (hc >> 1));
          const T* src = src_n + (int64_t)iy * d.x_strides[2] + (int64_t)ix * d.x_strides[3];
          cp_async16(dst_j + (hr * p.HC + slot) * 16, ok ? (const void*)src : (const void*)xb, ok ? 16u : 0u);
          hr += ppb_div; hc += ppb_mod;
          if (hc >= p.HC) { hc -= p.HC; hr++; }
        }
      }
      cp_async_commit();
      if (i - pub >= p.lookahead) {
        cp_async_wait_n(p.lookahead);
        publish(pub++);
      }
    }
    cp_async_wait<0>();
    while (pub < ntiles) publish(pub++);

    // =========================== epilogue: lane = output channel, columns = (tap kx, input channel) ===========
    if (ntiles > 0) {
      mbar_wait(smem_u32(&accum_bar), 0);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int o = o0 + threadIdx.x;
      const int wy = d.flip ? d.kh - 1 - ky : ky;
      for (int kx = 0; kx < d.kw; kx++) {
        const int wx = d.flip ? d.kw - 1 - kx : kx;
#pragma unroll 1
        for (int cc = 0; cc < BNC; cc += 16) {
          uint32_t acc[16];
          tmem_ld16(lane_addr + kx * BNC + cc, acc);
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
      const int mmas = p.TH * 8 / KPM;
      const uint32_t leader = elect_one();
      const uint32_t a_hi = smem_desc_hi((uint32_t)p.a_plane), b_hi = smem_desc_hi((uint32_t)p.b_plane);
      const uint32_t a_lo_base = smem_desc_lo(smem_u32(smem), 128);
      const uint32_t b_lo_base = smem_desc_lo(smem_u32(smem) + (uint32_t)p.a_bytes, (uint32_t)(p.HC * 16));   // LBO: next tile row = next patch row
      const uint32_t stage_u = (uint32_t)p.stage_bytes >> 4;
      const uint32_t b_row_u = (uint32_t)((KPM / 8) * p.HC);     // patch pixels between the first rows of consecutive MMAs
      int sa = 0;
      uint32_t pha = 0;
      for (int i = 0; i < ntiles; i++) {
        mbar_wait(smem_u32(&full_bar[sa]), pha);
        tc_fence_after();
        const uint32_t a_lo0 = a_lo_base + sa * stage_u, b_lo0 = b_lo_base + sa * stage_u;
        for (int kx = 0; kx < d.kw; kx++) {          // warp-uniform issue (umma.cuh: elect_one / umma_lh_pred)
          const uint32_t b_lo = b_lo0 + (uint32_t)((kx % s) * p.QP + kx / s);
          const uint32_t tm = tmem_base + kx * BNC;
#pragma unroll 4
          for (int kk = 0; kk < mmas; kk++)
            umma_lh_pred<KIND>(leader, tm, a_lo0 + kk * KPM, a_hi, b_lo + kk * b_row_u, b_hi, IDESC, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit_pred(leader, smem_u32(&empty_bar[sa]));
        if (++sa == SA) { sa = 0; pha ^= 1; }
      }
      if (ntiles > 0) umma_commit_pred(leader, smem_u32(&accum_bar));
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- fp32 tensors: 3 x bf16 split ----------------------------------------------------------------------------------
// tcgen05 takes MN-major (channel-contiguous) operands for the 16-bit kinds only in the no-swizzle layout used here:
// kind::tf32 with MN-major operands returns zeros (measured; the old per-tap kernel had the same defect).  fp32
// gradients / activations are therefore split into two bf16 terms each, v = hi + lo, and
//     dW  ~=  dy_hi * x_hi  +  dy_hi * x_lo  +  dy_lo * x_hi          (error ~2^-16 relative: better than TF32's 2^-11)
// is accumulated in fp32 by three kind::f16 MMAs per K step.  Data flow per tile:
//   cp.async  fp32 chunks  ->  STAGING buffer (thread-private positions, WG_LOOKAHEAD + 1 buffers, no barriers)
//   the same thread: LDS its two fp32 chunks of an 8-channel group, (x: multiply by the style), split, STS the bf16
//   hi / lo chunks into the MMA stage ([hi planes | lo planes] x [chunk of 8 channels][pixel][16 B]), fence, arrive.
template <int BNC>
__global__ void __launch_bounds__(WS_THREADS, 1) conv_wgrad_split_kernel(WgradHaloParams p) {
  constexpr int KPM = 16;                            // pixels per MMA (bf16: K = 32 bytes)
  constexpr uint32_t IDESC = make_idesc(1, BNC, 1);  // bf16 operands, both MN-major
  constexpr int MAX_STAGES = 4;

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const sgb_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int otile = blockIdx.x;
  const int ky = blockIdx.y / p.ctiles, ctile = blockIdx.y - ky * p.ctiles;
  const int64_t t_begin = (int64_t)blockIdx.z * p.chunk_tiles;
  const int64_t t_end = (t_begin + p.chunk_tiles < p.total_tiles) ? t_begin + p.chunk_tiles : p.total_tiles;
  const int ntiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;
  const int SA = p.stages;
  const int NSTG = p.nstg, lookahead = p.nstg - 1;
  const int s = d.stride;
  const int o0 = otile * UM, c0 = ctile * BNC;
  const uint32_t tmem_cols = (d.kw * BNC <= 128) ? 128u : ((d.kw * BNC <= 256) ? 256u : 512u);
  // staging: fp32 layout [chunk of 4 channels][pixel][16 B] with the same plane strides (PA + BNC/4 planes, PA = 32, or
  // 16 when dy has <= 64 channels); MMA stage: [A hi: PA/2 planes][A lo: PA/2][B hi: BNC/8][B lo: BNC/8],
  // plane = (pixels padded) * 16 B
  uint8_t* stg_base = smem;
  uint8_t* mma_base = smem + NSTG * p.stage_bytes;
  const int a_half = p.a_bytes / 2, b_half = (BNC / 8) * p.b_plane;

  for (int i = threadIdx.x * 16; i < (NSTG + SA) * p.stage_bytes; i += WS_THREADS * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
  if (warp == WS_PRODUCERS / 32) {
    if (lane == 0) {
      for (int i = 0; i < MAX_STAGES; i++) { mbar_init(smem_u32(&full_bar[i]), WS_PRODUCERS / 32); mbar_init(smem_u32(&empty_bar[i]), 1); }
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

  if (warp < WS_PRODUCERS / 32) {
    // =========================== producers / converters (8 warps) ===========================
    const int t = threadIdx.x;
    const int cow = (d.co - o0 < UM) ? d.co - o0 : UM;             // valid channels (multiples of 4)
    const int ciw = (d.ci - c0 < BNC) ? d.ci - c0 : BNC;
    const int ga = (cow + 7) / 8, gb = (ciw + 7) / 8;              // 8-channel groups: one thread item = 2 fp32 chunks
    int la = 0; while ((1 << la) < ga) la++;
    int lb = 0; while ((1 << lb) < gb) lb++;
    const int ja = t & ((1 << la) - 1), pa0 = t >> la, ppa = WS_PRODUCERS >> la;
    const int jb = t & ((1 << lb) - 1), pb0 = t >> lb, ppb = WS_PRODUCERS >> lb;
    const bool a_two = ja * 8 + 4 < cow, b_two = jb * 8 + 4 < ciw;  // second fp32 chunk of the group exists
    const int npa = p.TH * 8, npb = p.TH * p.HC;
    const int ppb_div = ppb / p.HC, ppb_mod = ppb - ppb_div * p.HC;
    const int pb0_r = pb0 / p.HC, pb0_c = pb0 - pb0_r * p.HC;
    const float* xb = (const float*)p.x;
    const float* dyb = (const float*)p.dy;
    const float* scb = (const float*)d.in_scale;
    const float* sca = (const float*)d.out_scale;                  // per-sample scale of the dy channels
    const int tiles_per_img = p.row_tiles * p.col_tiles;
    int pub = 0, sa_p = 0;
    uint32_t ph_p = 0;

    // tile coordinates advance incrementally (no divisions per tile); one cursor for the loads, one for the conversion
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
    // fp32 x8 (two chunks) -> bf16 hi chunk + bf16 lo chunk, with integer / FADD instructions only (cvt runs on the
    // slow conversion pipe: it was 12 % of the kernel's stall samples).  hi = the top 16 bits of v (truncation),
    // lo = the top 16 bits of (v - hi), which is exact in fp32; |v - hi - lo| <= 2^-16 |v|.
    auto split8 = [](const uint4& c0v, const uint4& c1v, const float* sv, uint4& hi, uint4& lo) {
      float f[8] = {__uint_as_float(c0v.x), __uint_as_float(c0v.y), __uint_as_float(c0v.z), __uint_as_float(c0v.w),
                    __uint_as_float(c1v.x), __uint_as_float(c1v.y), __uint_as_float(c1v.z), __uint_as_float(c1v.w)};
      if (sv) {
#pragma unroll
        for (int e = 0; e < 8; e++) f[e] *= sv[e];
      }
      uint32_t h[4], l[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t b0 = __float_as_uint(f[2 * q]), b1 = __float_as_uint(f[2 * q + 1]);
        h[q] = __byte_perm(b0, b1, 0x7632);                               // {hi16(b1), hi16(b0)}
        const float r0 = f[2 * q] - __uint_as_float(b0 & 0xffff0000u);
        const float r1 = f[2 * q + 1] - __uint_as_float(b1 & 0xffff0000u);
        l[q] = __byte_perm(__float_as_uint(r0), __float_as_uint(r1), 0x7632);
      }
      hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
    };

    int stg_p = 0;                                       // staging buffer of the next tile to convert
    auto convert_publish = [&](int i) {                  // tile i has landed in its staging buffer
      const uint8_t* stg = stg_base + stg_p * p.stage_bytes;
      if (++stg_p == NSTG) stg_p = 0;
      mbar_wait(smem_u32(&empty_bar[sa_p]), ph_p ^ 1);
      uint8_t* ms = mma_base + sa_p * p.stage_bytes;
      if (ja < ga) {
        float sv[8];
        if (sca) {
          const float* sp = sca + (int64_t)cur_p.n * d.co + o0 + ja * 8;
#pragma unroll
          for (int e = 0; e < 8; e++) sv[e] = (ja * 8 + e < cow) ? __ldg(sp + e) : 0.f;
        }
        const uint8_t* src = stg + (2 * ja) * p.a_plane;
        uint8_t* dst = ms + ja * p.a_plane;
        for (int pp = pa0; pp < npa; pp += ppa) {
          const uint4 c0v = *(const uint4*)(src + pp * 16), c1v = *(const uint4*)(src + p.a_plane + pp * 16);
          uint4 hi, lo;
          split8(c0v, c1v, sca ? sv : nullptr, hi, lo);
          *(uint4*)(dst + pp * 16) = hi;
          *(uint4*)(dst + a_half + pp * 16) = lo;
        }
      }
      if (jb < gb) {
        float sv[8];
        if (scb) {
          const float* sp = scb + (int64_t)cur_p.n * d.ci + c0 + jb * 8;
#pragma unroll
          for (int e = 0; e < 8; e++) sv[e] = (jb * 8 + e < ciw) ? __ldg(sp + e) : 0.f;
        }
        const uint8_t* src = stg + p.a_bytes + (2 * jb) * p.b_plane;
        uint8_t* dst = ms + 2 * a_half + jb * p.b_plane;
        int hr = pb0_r, hc = pb0_c;
        for (int hp = pb0; hp < npb; hp += ppb) {
          const int slot = (s == 1) ? hc : ((hc & 1) * p.QP + (hc >> 1));
          const int off = (hr * p.HC + slot) * 16;
          const uint4 c0v = *(const uint4*)(src + off), c1v = *(const uint4*)(src + p.b_plane + off);
          uint4 hi, lo;
          split8(c0v, c1v, scb ? sv : nullptr, hi, lo);
          *(uint4*)(dst + off) = hi;
          *(uint4*)(dst + b_half + off) = lo;
          hr += ppb_div; hc += ppb_mod;
          if (hc >= p.HC) { hc -= p.HC; hr++; }
        }
      }
      fence_proxy_async();
      __syncwarp();                                   // one arrival per warp (see above)
      if (lane == 0) mbar_arrive(smem_u32(&full_bar[sa_p]));
      if (++sa_p == SA) { sa_p = 0; ph_p ^= 1; }
      advance(cur_p);
    };

    int stg_i = 0;                                       // staging buffer of the next tile to load
    for (int i = 0; i < ntiles; i++) {
      const int n = cur_i.n, oy0 = cur_i.oy0, ox0 = cur_i.ox0;
      advance(cur_i);
      const uint32_t a_dst = smem_u32(stg_base + stg_i * p.stage_bytes);
      if (++stg_i == NSTG) stg_i = 0;
      const uint32_t b_dst = a_dst + p.a_bytes;
      if (ja < ga) {
        const float* src_n = dyb + (int64_t)n * d.y_strides[0] + o0 + ja * 8;
        const uint32_t dst_j = a_dst + (2 * ja) * p.a_plane;
        for (int pp = pa0; pp < npa; pp += ppa) {
          const int oy = oy0 + (pp >> 3), ox = ox0 + (pp & 7);
          const bool ok = oy < d.out_h && ox < d.out_w;
          const float* src = src_n + (int64_t)oy * d.y_strides[2] + (int64_t)ox * d.y_strides[3];
          cp_async16(dst_j + pp * 16, ok ? (const void*)src : (const void*)dyb, ok ? 16u : 0u);
          if (a_two) cp_async16(dst_j + p.a_plane + pp * 16, ok ? (const void*)(src + 4) : (const void*)dyb, ok ? 16u : 0u);
        }
      }
      if (jb < gb) {
        const float* src_n = xb + (int64_t)n * d.x_strides[0] + c0 + jb * 8;
        const uint32_t dst_j = b_dst + (2 * jb) * p.b_plane;
        const int ix0 = ox0 * s - d.pad_x;
        int hr = pb0_r, hc = pb0_c;
        for (int hp = pb0; hp < npb; hp += ppb) {
          const int iy = (oy0 + hr) * s + ky - d.pad_y, ix = ix0 + hc;
          const bool ok = iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w;
          const int slot = (s == 1) ? hc : ((hc & 1) * p.QP + (hc >> 1));
          const float* src = src_n + (int64_t)iy * d.x_strides[2] + (int64_t)ix * d.x_strides[3];
          const uint32_t dd = dst_j + (hr * p.HC + slot) * 16;
          cp_async16(dd, ok ? (const void*)src : (const void*)xb, ok ? 16u : 0u);
          if (b_two) cp_async16(dd + p.b_plane, ok ? (const void*)(src + 4) : (const void*)xb, ok ? 16u : 0u);
          hr += ppb_div; hc += ppb_mod;
          if (hc >= p.HC) { hc -= p.HC; hr++; }
        }
      }
      cp_async_commit();
      if (i - pub >= lookahead) {                          // keep `lookahead` tiles in flight, convert the oldest
        cp_async_wait_n(lookahead);
        convert_publish(pub++);
      }
    }
    cp_async_wait<0>();
    while (pub < ntiles) convert_publish(pub++);

    // =========================== epilogue (warps 0-3: one TMEM lane quadrant each) ===========================
    if (ntiles > 0 && warp < 4) {
      mbar_wait(smem_u32(&accum_bar), 0);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int o = o0 + threadIdx.x;
      const int wy = d.flip ? d.kh - 1 - ky : ky;
      for (int kx = 0; kx < d.kw; kx++) {
        const int wx = d.flip ? d.kw - 1 - kx : kx;
#pragma unroll 1
        for (int cc = 0; cc < BNC; cc += 16) {
          uint32_t acc[16];
          tmem_ld16(lane_addr + kx * BNC + cc, acc);
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
    {
      const int mmas = p.TH * 8 / KPM;
      const uint32_t a_hi = smem_desc_hi((uint32_t)p.a_plane), b_hi = smem_desc_hi((uint32_t)p.b_plane);
      const uint32_t a_lo_base = smem_desc_lo(smem_u32(mma_base), 128);
      const uint32_t b_lo_base = smem_desc_lo(smem_u32(mma_base) + (uint32_t)(2 * a_half), (uint32_t)(p.HC * 16));
      const uint32_t stage_u = (uint32_t)p.stage_bytes >> 4;
      const uint32_t a_half_u = (uint32_t)a_half >> 4, b_half_u = (uint32_t)b_half >> 4;
      const uint32_t b_row_u = (uint32_t)((KPM / 8) * p.HC);
      int sa = 0;
      uint32_t pha = 0;
      for (int i = 0; i < ntiles; i++) {
        mbar_wait(smem_u32(&full_bar[sa]), pha);
        tc_fence_after();
        const uint32_t a0 = a_lo_base + sa * stage_u, b0 = b_lo_base + sa * stage_u;
        if (lane == 0) {
          for (int kx = 0; kx < d.kw; kx++) {
            const uint32_t bk = b0 + (uint32_t)((kx % s) * p.QP + kx / s);
            const uint32_t tm = tmem_base + kx * BNC;
            for (int kk = 0; kk < mmas; kk++) {
              const uint32_t aa = a0 + kk * KPM, bb = bk + kk * b_row_u;
              umma_lh<1>(tm, aa, a_hi, bb, b_hi, IDESC, (i > 0 || kk > 0) ? 1u : 0u);               // hi * hi
              umma_lh<1>(tm, aa, a_hi, bb + b_half_u, b_hi, IDESC, 1u);                             // hi * lo
              umma_lh<1>(tm, aa + a_half_u, a_hi, bb, b_hi, IDESC, 1u);                             // lo * hi
            }
          }
          umma_commit(smem_u32(&empty_bar[sa]));
        }
        __syncwarp();
        if (++sa == SA) { sa = 0; pha ^= 1; }
      }
      if (ntiles > 0 && lane == 0) umma_commit(smem_u32(&accum_bar));
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WS_PRODUCERS / 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- host side -------------------------------------------------------------------------------------------
bool conv_wgrad_halo_eligible(const sgb_conv_desc* d) {
  if (!((d->kh == 3 && d->kw == 3) || (d->kh == 1 && d->kw == 1))) return false;
  if (d->stride != 1 && d->stride != 2) return false;
  if (d->force_simt == 2) return false;
  return true;      // dtype / layout / alignment conditions are those of conv_wgrad_umma_eligible (checked by the caller)
}

template <class T, int KIND, int BNC>
static int launch_wgrad_halo(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
  constexpr int TC = 16 / sizeof(T);
  WgradHaloParams p; p.d = *d; p.x = x; p.dy = dy; p.dw = dw;
  const int s = d->stride;
  p.HC = 7 * s + d->kw;
  p.QP = (s == 1) ? 0 : (p.HC + 1) / 2;
  p.ctiles = (d->ci + BNC - 1) / BNC;
  const int otiles = (d->co + UM - 1) / UM;
  const int budget = 224 * 1024;
  int TH = 16, stages = 0;
  for (;; TH >>= 1) {
    int npa = TH * 8 + 1;
    int npb = TH * p.HC; while (npb % 8 != 1) npb++;
    p.a_plane = npa * 16; p.b_plane = npb * 16;
    p.a_bytes = (UM / TC) * p.a_plane;
    p.stage_bytes = (p.a_bytes + (BNC / TC) * p.b_plane + 127) / 128 * 128;
    stages = budget / p.stage_bytes; if (stages > 8) stages = 8;
    if (stages >= 3 || TH == 4) break;
  }
  SGB_REQUIRE(stages >= 2, "wgrad halo: tile does not fit shared memory");
  p.TH = TH; p.stages = stages;
  // tiles in flight per producer thread beyond the one being published (SGB_WGRAD_LA overrides); up to 8 stages
  static const int env_la = [] { const char* e = getenv("SGB_WGRAD_LA"); return e ? atoi(e) : 0; }();
  p.lookahead = 1;        // measured (r2_run28.sh, config-f step): 1 tile ahead 52.2 ms, stages - 2 ahead 53.0 ms
  if (env_la > 0 && env_la < stages) p.lookahead = env_la;
  if (p.lookahead > 6) p.lookahead = 6;
  p.row_tiles = (d->out_h + TH - 1) / TH; p.col_tiles = (d->out_w + 7) / 8;
  p.total_tiles = (int64_t)d->n * p.row_tiles * p.col_tiles;
  const int64_t base = (int64_t)otiles * p.ctiles * d->kh;
  int64_t splits = num_sms() / base; if (splits < 1) splits = 1;
  if (splits > p.total_tiles) splits = p.total_tiles;
  p.chunk_tiles = ceil_div(p.total_tiles, splits);
  splits = ceil_div(p.total_tiles, p.chunk_tiles);
  SGB_REQUIRE((int64_t)p.ctiles * d->kh <= 65535 && splits <= 65535, "problem too large for the wgrad halo grid");
  SGB_REQUIRE(aligned16(x) && aligned16(dy), "x and dy must be 16-byte aligned");
  const size_t smem = (size_t)stages * p.stage_bytes + 1024;
  auto kern = conv_wgrad_halo_kernel<T, KIND, BNC>;
  SGB_SET_MAX_SMEM(kern, 226 * 1024);
  kern<<<dim3((unsigned)otiles, (unsigned)(p.ctiles * d->kh), (unsigned)splits), WG_THREADS, smem, st>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

template <int BNC>
static int launch_wgrad_split(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
  WgradHaloParams p; p.d = *d; p.x = x; p.dy = dy; p.dw = dw;
  const int s = d->stride;
  p.HC = 7 * s + d->kw;
  p.QP = (s == 1) ? 0 : (p.HC + 1) / 2;
  p.ctiles = (d->ci + BNC - 1) / BNC;
  const int otiles = (d->co + UM - 1) / UM;
  const int budget = 224 * 1024;
  // Tiles are small (their MMAs take ~1 K cycles) and L2 / HBM latency is several K cycles, so the loads of MANY tiles
  // must be in flight: two MMA stages, every other buffer is fp32 staging; the tile height is the largest that keeps
  // >= 96 KB in flight (pixels per tile stay a multiple of 16 = one bf16 MMA).
  static const int env_sa = [] { const char* e = getenv("SGB_WGRAD_SA"); return e ? atoi(e) : 2; }();
  static const int env_kb = [] { const char* e = getenv("SGB_WGRAD_KB"); return e ? atoi(e) : 32; }();
  int TH = 8, stages = env_sa < 2 ? 2 : (env_sa > 4 ? 4 : env_sa), nstg = 2;
  for (;; TH >>= 1) {
    int npa = TH * 8 + 1;
    int npb = TH * p.HC; while (npb % 8 != 1) npb++;
    p.a_plane = npa * 16; p.b_plane = npb * 16;
    // <= 64 dy channels: half the A planes.  The M = 128 MMA still reads 16 planes per half; what it finds beyond the
    // valid ones (the lo half, the x planes -- all inside the stage) only reaches accumulator rows that are never stored
    p.a_bytes = (d->co <= 64 ? 16 : 32) * p.a_plane;
    p.stage_bytes = (p.a_bytes + (BNC / 4) * p.b_plane + 127) / 128 * 128;     // fp32 staging == bf16 hi + lo planes
    const int total = budget / p.stage_bytes;
    nstg = total - stages; if (nstg > 8) nstg = 8;
    if ((nstg >= 2 && (nstg - 1) * p.stage_bytes >= env_kb * 1024) || TH == 2) break;
  }
  if (nstg < 2) { nstg = 2; }
  SGB_REQUIRE((nstg + stages) * p.stage_bytes <= budget, "wgrad split: tile does not fit shared memory");
  p.nstg = nstg;
  p.TH = TH; p.stages = stages;
  p.row_tiles = (d->out_h + TH - 1) / TH; p.col_tiles = (d->out_w + 7) / 8;
  p.total_tiles = (int64_t)d->n * p.row_tiles * p.col_tiles;
  const int64_t base = (int64_t)otiles * p.ctiles * d->kh;
  int64_t splits = num_sms() / base; if (splits < 1) splits = 1;
  if (splits > p.total_tiles) splits = p.total_tiles;
  p.chunk_tiles = ceil_div(p.total_tiles, splits);
  splits = ceil_div(p.total_tiles, p.chunk_tiles);
  SGB_REQUIRE((int64_t)p.ctiles * d->kh <= 65535 && splits <= 65535, "problem too large for the wgrad split grid");
  SGB_REQUIRE(aligned16(x) && aligned16(dy), "x and dy must be 16-byte aligned");
  const size_t smem = (size_t)(nstg + stages) * p.stage_bytes + 1024;
  auto kern = conv_wgrad_split_kernel<BNC>;
  SGB_SET_MAX_SMEM(kern, 226 * 1024);
  kern<<<dim3((unsigned)otiles, (unsigned)(p.ctiles * d->kh), (unsigned)splits), WS_THREADS, smem, st>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

template <class T, int KIND>
static int dispatch_wgrad_halo(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t s) {
  if (d->ci <= 32) return launch_wgrad_halo<T, KIND, 32>(d, x, dy, dw, s);
  if (d->ci <= 64) return launch_wgrad_halo<T, KIND, 64>(d, x, dy, dw, s);
  return launch_wgrad_halo<T, KIND, 128>(d, x, dy, dw, s);
}

int conv_wgrad_halo(const sgb_conv_desc* d, const void* x, const void* dy, void* dw, cudaStream_t s) {
  const int64_t wnum = (int64_t)d->co * d->ci * d->kh * d->kw;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * wnum, s);
  SGB_REQUIRE(e == cudaSuccess, "memset failed");
  if ((int64_t)d->n * d->out_h * d->out_w == 0) return 0;
  // few-channel 3 x 3 stride-1 layers in 16-bit types: all nine taps from one staging (filter rows stacked in M)
  if (conv_wgrad_kys_eligible(d)) return conv_wgrad_kys(d, x, dy, (float*)dw, s);
  if (d->dtype == SGB_F16) return dispatch_wgrad_halo<__half, 0>(d, x, dy, (float*)dw, s);
  if (d->dtype == SGB_BF16) return dispatch_wgrad_halo<__nv_bfloat16, 1>(d, x, dy, (float*)dw, s);
  // fp32: kind::tf32 takes MN-major operands only in the SWIZZLE_128B_BASE32B layout (conv_wgrad_tf32.cu); the earlier
  // 3 x bf16 split kernel (fp32-exact products, ~1.7x slower) stays selectable with SGB_WGRAD_FP32=split
  static const bool use_split = [] { const char* e = getenv("SGB_WGRAD_FP32"); return e && std::string(e) == "split"; }();
  if (!use_split) return conv_wgrad_tf32(d, x, dy, (float*)dw, s);
  if (d->ci <= 32) return launch_wgrad_split<32>(d, x, dy, (float*)dw, s);
  if (d->ci <= 64) return launch_wgrad_split<64>(d, x, dy, (float*)dw, s);
  return launch_wgrad_split<128>(d, x, dy, (float*)dw, s);
}

}  // namespace sgb
