// tcgen05 weight gradient for the FEW-CHANNEL layers in 16-bit types (fp16 / bf16, channels_last, 3 x 3, stride 1, <= 64 output
// and <= 64 input channels): the 512^2 / 1024^2 layers of config-f, whose weight gradient is a stream over two huge tensors into
// a 9 x Co x Ci result.
//
//   dW[o, c, ky, kx] = sum_{n, oy, ox} dy[n, oy, ox, o] * xs[n, oy + ky - pad, ox + kx - pad, c]
//
// conv_wgrad_halo_kernel gives every filter ROW its own CTAs, so dy and x are read three times from L2, always issues
// M = 128 MMAs (a quarter of the rows is real data when Co = 32) and keeps two tiles in flight per SM: 0.83 ms for
// [4,32,1024,1024] fp16, ten times the HBM time of its operands (profiles/README.md).  Here ONE CTA computes all nine taps
// of a tile from one staging of dy and x: the filter rows are stacked in the M dimension of the MMA.
//
//   * Tile = TH x 8 output pixels of one image.  x patch: (TH + 2) rows x (8 + 2) column slots, row r <-> input row
//     oy0 - pad + r, in the no-swizzle MN-major plane layout of conv_wgrad_halo_kernel ([8-channel chunk][row][slot][16 B]):
//     the tap kx is the B descriptor's start moved by one slot, the second K group of an MMA is the next patch row (LBO).
//   * dy stage: [tile row q][8-channel chunk j][8 pixels][16 B] with q <-> output row oy0 + q - 2.  In this order the M
//     chunks of an MMA (SBO = 128 B apart) run through the chunks of tile row q first and then into tile row q + 1: with Co <= 32
//     (4 chunks per row) the sixteen M chunks of an M = 128 MMA started at row r are the dy rows r, r+1, r+2, r+3, which meet the
//     x patch row r as filter rows ky = 2, 1, 0 (and a fourth shift that lands in accumulator lanes nobody stores).  With
//     Co <= 64 (8 chunks per row) an MMA holds two shifts; a second MMA started at row r + 2 supplies ky = 0.
//     Tile rows q < 2 and q >= TH + 2 are zero for ever (cleared once, never written), so products with rows outside the tile
//     vanish instead of being counted by two tiles.
//   * K steps: one fp16 MMA consumes 16 pixels = two patch rows; a tile takes (TH + 2) / 2 K steps x 3 taps x (1 or 2) MMAs.
//   * Accumulators ([MMA][kx][BNC] columns) stay in TMEM across all tiles of the CTA; one red.global.add epilogue.
// Warps 0-7: cp.async producers with `lookahead` tiles in flight (+ in-place style scaling of their own chunks); warps 0-3 then
// run the epilogue.  Warp 8: MMA issuer (warp-uniform issue), owns TMEM.
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

namespace sgb {

constexpr int WK_PRODUCERS = 256;
constexpr int WK_THREADS = WK_PRODUCERS + 32;
constexpr int WK_MAX_STAGES = 8;

struct WgradKysParams {
  sgb_conv_desc d;
  const void* x; const void* dy; float* dw;
  int TH, RB;                 // tile rows; x patch rows = TH + 2
  int row_tiles, col_tiles;   // per image
  int64_t total_tiles, chunk_tiles;
  int HC;                     // patch column slots per row = 8 + kw - 1
  int cap;                    // dy chunks per pixel as staged: 4 (co <= 32) or 8 (co <= 64)
  int a_row;                  // bytes per staged dy tile row = cap * 128
  int a_bytes, b_plane, stage_bytes;
  int stages, lookahead;
};

template <class T, int KIND, int BNC>
__global__ void __launch_bounds__(WK_THREADS, 1) conv_wgrad_kys_kernel(WgradKysParams p) {
  constexpr int TC = 8;                                // 16-bit elements per 16-byte chunk
  constexpr uint32_t IDESC = make_idesc(KIND, BNC, 1); // M = 128, N = BNC, both operands MN-major

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[WK_MAX_STAGES], empty_bar[WK_MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const sgb_conv_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t_begin = (int64_t)blockIdx.x * p.chunk_tiles;
  const int64_t t_end = (t_begin + p.chunk_tiles < p.total_tiles) ? t_begin + p.chunk_tiles : p.total_tiles;
  const int ntiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;
  const int SA = p.stages;
  const int nm = (p.cap == 4) ? 1 : 2;                 // MMAs per K step and tap
  const uint32_t need_cols = (uint32_t)(nm * 3 * BNC);
  const uint32_t tmem_cols = need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u);

  // zero tile rows of dy, channel chunks beyond co / ci: cleared once, never written afterwards
  for (int i = threadIdx.x * 16; i < SA * p.stage_bytes; i += WK_THREADS * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
  if (warp == WK_PRODUCERS / 32) {
    if (lane == 0) {
      for (int i = 0; i < WK_MAX_STAGES; i++) { mbar_init(smem_u32(&full_bar[i]), WK_PRODUCERS / 32); mbar_init(smem_u32(&empty_bar[i]), 1); }
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

  if (warp < WK_PRODUCERS / 32) {
    // =========================== producers ===========================
    const int t = threadIdx.x;
    const int ca = d.co / TC, cb = d.ci / TC;                      // valid 16-byte chunks per pixel
    int la = 0; while ((1 << la) < ca) la++;
    int lb = 0; while ((1 << lb) < cb) lb++;
    const int ja = t & ((1 << la) - 1), pa0 = t >> la, ppa = WK_PRODUCERS >> la;   // chunk / first pixel / pixels per pass
    const int jb = t & ((1 << lb) - 1), pb0 = t >> lb, ppb = WK_PRODUCERS >> lb;
    const int npa = p.TH * 8, npb = p.RB * p.HC;
    const int ppb_div = ppb / p.HC, ppb_mod = ppb - ppb_div * p.HC;
    const int pb0_r = pb0 / p.HC, pb0_c = pb0 - pb0_r * p.HC;
    const T* xb = (const T*)p.x;
    const T* dyb = (const T*)p.dy;
    const float* scb = (const float*)d.in_scale;                   // per-sample scale of the x channels (style modulation)
    const float* sca = (const float*)d.out_scale;                  // per-sample scale of the dy channels
    const int tiles_per_img = p.row_tiles * p.col_tiles;
    const int lookahead = p.lookahead;
    // byte offset of this thread's dy chunk inside a stage for pixel pp: tile row (pp >> 3) + 2, chunk ja, pixel pp & 7
    auto a_off = [&](int pp) { return (uint32_t)(((pp >> 3) + 2) * p.a_row + ja * 128 + (pp & 7) * 16); };
    const uint32_t b_off = (uint32_t)(p.a_bytes + jb * p.b_plane);

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

    auto scale8 = [&](uint4* q, const float* sv) {
      uint4 v = *q;
      T* h = (T*)&v;
#pragma unroll
      for (int e = 0; e < TC; e++) h[e] = from_acc<T>(to_acc<T>(h[e]) * sv[e]);
      *q = v;
    };
    int pub = 0, sa_p = 0;
    auto publish = [&]() {                               // the oldest unpublished tile has landed
      uint8_t* stage = smem + sa_p * p.stage_bytes;
      if (sca && ja < ca) {
        const float* sp = sca + (int64_t)cur_p.n * d.co + ja * TC;
        const float4 s0 = __ldg((const float4*)sp), s1 = __ldg((const float4*)(sp + 4));
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        for (int pp = pa0; pp < npa; pp += ppa) scale8((uint4*)(stage + a_off(pp)), sv);
      }
      if (scb && jb < cb) {
        const float* sp = scb + (int64_t)cur_p.n * d.ci + jb * TC;
        const float4 s0 = __ldg((const float4*)sp), s1 = __ldg((const float4*)(sp + 4));
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        for (int hp = pb0; hp < npb; hp += ppb) scale8((uint4*)(stage + b_off + hp * 16), sv);
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
      if (ja < ca) {
        const T* src_n = dyb + (int64_t)n * d.y_strides[0] + ja * TC;
        for (int pp = pa0; pp < npa; pp += ppa) {
          const int oy = oy0 + (pp >> 3), ox = ox0 + (pp & 7);
          const bool ok = oy < d.out_h && ox < d.out_w;
          const T* src = src_n + (int64_t)oy * d.y_strides[2] + (int64_t)ox * d.y_strides[3];
          cp_async16(stage + a_off(pp), ok ? (const void*)src : (const void*)dyb, ok ? 16u : 0u);
        }
      }
      if (jb < cb) {
        const T* src_n = xb + (int64_t)n * d.x_strides[0] + jb * TC;
        const uint32_t dst_j = stage + b_off;
        const int iy0 = oy0 - d.pad_y, ix0 = ox0 - d.pad_x;
        int hr = pb0_r, hc = pb0_c;
        for (int hp = pb0; hp < npb; hp += ppb) {
          const int iy = iy0 + hr, ix = ix0 + hc;
          const bool ok = iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w;
          const T* src = src_n + (int64_t)iy * d.x_strides[2] + (int64_t)ix * d.x_strides[3];
          cp_async16(dst_j + hp * 16, ok ? (const void*)src : (const void*)xb, ok ? 16u : 0u);
          hr += ppb_div; hc += ppb_mod;
          if (hc >= p.HC) { hc -= p.HC; hr++; }
        }
      }
      cp_async_commit();
      if (i - pub >= lookahead) {
        cp_async_wait_n(lookahead);
        publish();
      }
    }
    cp_async_wait<0>();
    while (pub < ntiles) publish();

    // =========================== epilogue (warps 0-3: one TMEM lane quadrant each) ===========================
    if (ntiles > 0 && warp < 4) {
      mbar_wait(smem_u32(&accum_bar), 0);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      // accumulator lane = (shift, output channel): 4 shifts x 32 channels (cap 4) or 2 shifts x 64 channels (cap 8)
      const int shift = (p.cap == 4) ? warp : (warp >> 1);
      const int o = (p.cap == 4) ? lane : ((warp & 1) * 32 + lane);
      for (int m = 0; m < nm; m++) {
        const int ky = 2 - 2 * m - shift;            // MMA m starts two dy rows further down
        const int wy = d.flip ? 2 - ky : ky;
        for (int kx = 0; kx < 3; kx++) {
          const int wx = d.flip ? 2 - kx : kx;
#pragma unroll 1
          for (int cc = 0; cc < BNC; cc += 16) {
            uint32_t acc[16];
            tmem_ld16(lane_addr + (m * 3 + kx) * BNC + cc, acc);
            if (ky < 0 || o >= d.co) continue;
#pragma unroll
            for (int e = 0; e < 16; e++) {
              const int c = cc + e;
              if (c < d.ci) atomicAdd(p.dw + (((int64_t)o * d.ci + c) * 3 + wy) * 3 + wx, __uint_as_float(acc[e]));
            }
          }
        }
      }
      tc_fence_before();
    }
  } else {
    // =========================== MMA issuer ===========================
    const uint32_t leader = elect_one();
    const uint32_t a_hi = smem_desc_hi(128), b_hi = smem_desc_hi((uint32_t)p.b_plane);     // SBO: next 8-channel chunk
    const uint32_t a_lo_base = smem_desc_lo(smem_u32(smem), (uint32_t)p.a_row);              // LBO: next tile row (second K group)
    const uint32_t b_lo_base = smem_desc_lo(smem_u32(smem) + (uint32_t)p.a_bytes, (uint32_t)(p.HC * 16));   // LBO: next patch row
    const uint32_t stage_u = (uint32_t)p.stage_bytes >> 4;
    const uint32_t a_row_u = (uint32_t)p.a_row >> 4;
    int sa = 0;
    uint32_t pha = 0;
    for (int i = 0; i < ntiles; i++) {
      mbar_wait(smem_u32(&full_bar[sa]), pha);
      tc_fence_after();
      const uint32_t a_lo0 = a_lo_base + sa * stage_u, b_lo0 = b_lo_base + sa * stage_u;
      for (int m = 0; m < nm; m++)
        for (int kx = 0; kx < 3; kx++) {
          const uint32_t tm = tmem_base + (m * 3 + kx) * BNC;
          const uint32_t a_lo = a_lo0 + 2 * m * a_row_u, b_lo = b_lo0 + kx;
#pragma unroll 3
          for (int r = 0; r < p.RB; r += 2)
            umma_lh_pred<KIND>(leader, tm, a_lo + r * a_row_u, a_hi, b_lo + r * p.HC, b_hi, IDESC, (i > 0 || r > 0) ? 1u : 0u);
        }
      umma_commit_pred(leader, smem_u32(&empty_bar[sa]));
      if (++sa == SA) { sa = 0; pha ^= 1; }
    }
    if (ntiles > 0) umma_commit_pred(leader, smem_u32(&accum_bar));
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WK_PRODUCERS / 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- host side -------------------------------------------------------------------------------------------
// SGB_WGRAD_KYS=0 switches the kernel off (A/B against conv_wgrad_halo_kernel)
bool conv_wgrad_kys_eligible(const sgb_conv_desc* d) {
  static const int on = [] { const char* e = getenv("SGB_WGRAD_KYS"); return e ? atoi(e) : 1; }();
  if (!on || d->force_simt == 2) return false;
  if (d->dtype != SGB_F16 && d->dtype != SGB_BF16) return false;
  if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->groups != 1 || d->transposed) return false;
  if (d->co > 64 || d->ci > 64 || d->co % 8 || d->ci % 8) return false;
  return true;      // layout / alignment conditions: conv_wgrad_umma_eligible (checked by the caller)
}

template <class T, int KIND, int BNC>
static int launch_wgrad_kys(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
  WgradKysParams p; p.d = *d; p.x = x; p.dy = dy; p.dw = dw;
  p.HC = 8 + d->kw - 1;
  p.cap = d->co <= 32 ? 4 : 8;
  p.a_row = p.cap * 128;
  const int budget = 224 * 1024;
  static const int env_th = [] { const char* e = getenv("SGB_WGRAD_TH"); return e ? atoi(e) : 0; }();
  static const int env_la = [] { const char* e = getenv("SGB_WGRAD_LA"); return e ? atoi(e) : 0; }();
  int TH = (env_th >= 2 && env_th <= 32 && env_th % 2 == 0) ? env_th : 16, stages = 0;
  for (;;) {
    p.RB = TH + 2;
    p.a_bytes = (TH + 5) * p.a_row;
    int npb = p.RB * p.HC; while (npb % 8 != 1) npb++;           // odd multiple of 16 bytes between planes: conflict-free cp.async
    p.b_plane = npb * 16;
    p.stage_bytes = (p.a_bytes + (BNC / 8) * p.b_plane + 127) / 128 * 128;
    stages = budget / p.stage_bytes; if (stages > WK_MAX_STAGES) stages = WK_MAX_STAGES;
    if (stages >= 4 || TH == 2) break;
    TH = TH > 4 ? TH - 4 : TH - 2;
  }
  SGB_REQUIRE(stages >= 2, "wgrad kys: tile does not fit shared memory");
  p.TH = TH; p.stages = stages;
  p.lookahead = stages - 2 < 1 ? 1 : stages - 2;
  if (env_la > 0 && env_la < stages) p.lookahead = env_la;
  if (p.lookahead > 7) p.lookahead = 7;
  p.row_tiles = (d->out_h + TH - 1) / TH; p.col_tiles = (d->out_w + 7) / 8;
  p.total_tiles = (int64_t)d->n * p.row_tiles * p.col_tiles;
  int64_t splits = num_sms();
  if (splits > p.total_tiles) splits = p.total_tiles;
  p.chunk_tiles = ceil_div(p.total_tiles, splits);
  splits = ceil_div(p.total_tiles, p.chunk_tiles);
  SGB_REQUIRE(aligned16(x) && aligned16(dy), "x and dy must be 16-byte aligned");
  const size_t smem = (size_t)stages * p.stage_bytes;
  auto kern = conv_wgrad_kys_kernel<T, KIND, BNC>;
  SGB_SET_MAX_SMEM(kern, 226 * 1024);
  kern<<<(unsigned)splits, WK_THREADS, smem, st>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

// dw must be zeroed by the caller
int conv_wgrad_kys(const sgb_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t s) {
  if (d->dtype == SGB_F16) return d->ci <= 32 ? launch_wgrad_kys<__half, 0, 32>(d, x, dy, dw, s) : launch_wgrad_kys<__half, 0, 64>(d, x, dy, dw, s);
  return d->ci <= 32 ? launch_wgrad_kys<__nv_bfloat16, 1, 32>(d, x, dy, dw, s) : launch_wgrad_kys<__nv_bfloat16, 1, 64>(d, x, dy, dw, s);
}

}  // namespace sgb
