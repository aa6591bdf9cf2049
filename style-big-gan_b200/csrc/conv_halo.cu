// tcgen05 implicit-GEMM convolution, persistent "halo tile" kernel (channels_last).
//
// conv_umma.cu re-gathers the A operand once per filter tap (9x the activation traffic through LSU + shared
// memory).  Here the CTA stages the input patch of its output tile ONCE per channel block and the filter taps are
// shared-memory descriptors into the same patch.  Three geometries (template parameter MODE):
//
//   MODE 0  stride 1, conv2d and conv_transpose2d (forward and data gradient of every 3x3 / 1x1 layer)
//           tile = 16 x 8 output pixels (M = 128, row m = ty*8 + tx); patch = (16+kh-1) x (8+kw-1) pixels.
//           A core matrix = 8 consecutive x of one tile row = 8 consecutive patch pixels (16 B apart);
//           SBO = one patch row = next tile row; the tap only moves the start address.
//           Rows are "virtual rows": every image contributes out_h + kh - 1 rows (zero padding included), so tiles
//           may straddle images under one addressing scheme; the kh - 1 junk rows per image are dropped.
//   MODE 1  conv2d stride 2 (D down path after the FIR; data gradient of the G up path)
//           tile = 16 x 8 output pixels; patch = (30+kh) x (14+kw) input pixels stored with the columns
//           DE-INTERLEAVED BY PARITY, so the 8 pixels ox..ox+7 of a tap (input columns 2*ox + kx) are again 16 B
//           apart; SBO = two patch rows.  64 bytes of K per stage (the patch is 4x the tile).
//   MODE 2  conv_transpose2d stride 2 (G up path; data gradient of the D down path)
//           the output splits into 4 parity phases (oy+pad, ox+pad mod 2); a tile = 16 x 8 positions (a, b) of the
//           half-resolution grid, i.e. 4 x 128 output pixels; phase (py, px) only receives the taps with ky = py,
//           kx = px (mod 2), each a stride-1 gather at (a - ky/2, b - kx/2).  Four TMEM accumulators per tile, no
//           multiplications by the inserted zeros (the per-tap kernel spends 4x the MMAs on them).
//
// Persistent: grid = #SMs, each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  Ten warps:
//   warps 0-3  epilogue: tcgen05.ld of an accumulator buffer, demod scale / noise / bias_act, transpose through a
//              per-warp shared-memory staging slab so that every global store instruction writes whole 128-byte
//              lines (4 pixels x 128 B), release the buffer -- overlaps the next tile's MMAs when TMEM allows
//   warp 4     MMA issuer (one lane): tcgen05.mma.cta_group::1; owns TMEM
//   warp 5     weight loader: cp.async.bulk of pre-packed B tiles into a ring of SB stages
//   warps 6-9  patch producers: cp.async 16-byte chunks global -> shared (zero-fill for padding), several patches in
//              flight per thread, optional in-place style scaling, fence.proxy.async, mbarrier arrive
#include "common.cuh"
#include "act.cuh"
#include "umma.cuh"

namespace sgb {

constexpr int TILE_H = 16, TILE_W = 8;
constexpr int HALO_THREADS = 320;
constexpr int MAX_SA = 6, MAX_SB = 8;

struct HaloParams {
  sgb_conv_desc d;
  const void* x; const void* wpack; void* y;
  int VR;               // rows of the tile-row space per image (MODE 0: out_h + kh - 1; else padded to 16)
  int HR, HC;           // patch rows / column slots per row
  int QP;               // MODE 1: slot offset of the odd-column plane
  int top, left;        // patch origin relative to the tile origin (input coordinates)
  int col_tiles, ntiles;
  int64_t total_tiles;
  int taps, cblocks;
  int lbo;              // bytes between channel chunks of the patch (padded)
  int a_stage_bytes;
  int sa, sb;           // stages
  int stg_off;          // byte offset of the epilogue staging area
  int vec_store;
};

__device__ __forceinline__ void decode_tile(const HaloParams& p, int64_t t, int& ntile, int& u0, int& x0) {
  ntile = (int)(t % p.ntiles);
  const int64_t mt = t / p.ntiles;
  x0 = (int)(mt % p.col_tiles) * TILE_W;
  u0 = (int)(mt / p.col_tiles) * TILE_H;
}

template <int N> __device__ __forceinline__ void cp_async_wait_dyn(int n) {
  // wait until at most n of this thread's cp.async groups are pending
  if (n <= 0) cp_async_wait<0>();
  else if (n == 1) cp_async_wait<1>();
  else if (n == 2) cp_async_wait<2>();
  else if (n == 3) cp_async_wait<3>();
  else cp_async_wait<4>();
}

template <class T, int KIND, int BN, int MODE>
__global__ void __launch_bounds__(HALO_THREADS, 1) conv_halo_kernel(HaloParams p) {
  constexpr int TC = 16 / sizeof(T);
  constexpr int CH = (MODE == 1) ? 4 : 8;             // 16-byte channel chunks of K per stage
  constexpr int BK = CH * TC;
  constexpr int B_STAGE_BYTES = BN * CH * 16;
  constexpr int NACC = (MODE == 2) ? 4 : 1;           // accumulators per tile
  constexpr int NBUF = (NACC * BN * 2 <= 512) ? 2 : 1;
  constexpr uint32_t NEED_COLS = NACC * BN * NBUF;
  constexpr uint32_t TMEM_COLS = NEED_COLS <= 32 ? 32 : (NEED_COLS <= 64 ? 64 : (NEED_COLS <= 128 ? 128 : (NEED_COLS <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = make_idesc(KIND, BN);
  constexpr int MAX_SLOTS = (MODE == 1) ? 18 : 12;    // ceil(patch pixels * CH / 128)
  constexpr int PPS = 128 / CH;                       // patch pixels per slot pass
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
      for (int s = 0; s < MAX_SA; s++) { mbar_init(smem_u32(&a_full[s]), 128); mbar_init(smem_u32(&a_empty[s]), 1); }
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
    const int lookahead = SA - 2;
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
    int pa = 0;                                        // patches issued so far (ring position)
    int pub = 0;                                       // patches published so far

    auto publish = [&](int idx) {                      // patch number idx has landed: optional scaling, fence, arrive
      const int sa = idx % SA;
      if (scb) {
        uint8_t* dst = a_base + sa * p.a_stage_bytes;
        int nt_, u0o, x0_;
        decode_tile(p, (int64_t)blockIdx.x + (int64_t)(idx / p.cblocks) * gridDim.x, nt_, u0o, x0_);
        const int co = (idx % p.cblocks) * BK + j * TC;
        if (co < d.ci) {
#pragma unroll
          for (int i = 0; i < MAX_SLOTS; i++) {
            if (hrc[i] >= 0) {
              int n = (MODE == 0) ? (u0o + (hrc[i] >> 16)) / p.VR : u0o / p.VR;
              n = n < d.n ? n : d.n - 1;
              const float* sp = scb + (int64_t)n * d.ci + co;
              uint4* q = (uint4*)(dst + dsl[i]);
              uint4 v = *q;
              if (KIND == 2) {
                const float4 s4 = __ldg((const float4*)sp);
                float* f = (float*)&v;
                f[0] *= s4.x; f[1] *= s4.y; f[2] *= s4.z; f[3] *= s4.w;
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
      mbar_arrive(smem_u32(&a_full[sa]));
    };

    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int ntile, u0, x0;
      decode_tile(p, tile, ntile, u0, x0);
      // per-slot source offsets (elements) of the patch pixels this thread stages; -1 = zero (padding / outside)
      int off[MAX_SLOTS];
      {
        const int n_t = u0 / p.VR, r0 = u0 - n_t * p.VR;          // MODE 1 / 2: tiles never straddle images
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; i++) {
          off[i] = -1;
          if (hrc[i] >= 0) {
            const int hr = hrc[i] >> 16, hc = hrc[i] & 0xffff;
            int n, iy, ix;
            if (MODE == 0) {
              const int u = u0 + hr;
              n = u / p.VR;
              iy = u - n * p.VR - p.top;
              ix = x0 + hc - p.left;
            } else if (MODE == 1) {
              n = n_t; iy = 2 * r0 - p.top + hr; ix = 2 * x0 - p.left + hc;
            } else {
              n = n_t; iy = r0 - p.top + hr; ix = x0 - p.left + hc;
            }
            if (n < d.n && iy >= 0 && iy < d.in_h && ix >= 0 && ix < d.in_w)
              off[i] = (int)(n * d.x_strides[0] + iy * d.x_strides[2] + ix * d.x_strides[3]);
          }
        }
      }
      for (int cb = 0; cb < p.cblocks; cb++, pa++) {
        const int sa = pa % SA;
        const int c = cb * BK + j * TC;
        const bool c_ok = c < d.ci;
        mbar_wait(smem_u32(&a_empty[sa]), ((pa / SA) & 1) ^ 1);
        const uint32_t dst = smem_u32(a_base + sa * p.a_stage_bytes);
#pragma unroll
        for (int i = 0; i < MAX_SLOTS; i++) {
          if (hrc[i] >= 0) {
            const bool ok = c_ok && off[i] >= 0;
            cp_async16(dst + dsl[i], ok ? (const void*)(xb + off[i] + c) : (const void*)xb, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        if (pa - pub >= lookahead) {                   // keep `lookahead` patches in flight, publish the oldest
          cp_async_wait_dyn<0>(lookahead);
          publish(pub++);
        }
      }
    }
    cp_async_wait<0>();
    while (pub < pa) publish(pub++);
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      int pa = 0, kb = 0, li = 0;
      const uint32_t sbo = (uint32_t)((MODE == 1 ? 2 : 1) * p.HC * 16);
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, li++) {
        const int buf = (NBUF == 2) ? (li & 1) : 0;
        const uint32_t eph = (NBUF == 2) ? ((li >> 1) & 1) : (li & 1);
        mbar_wait(smem_u32(&acc_empty[buf]), eph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * (NACC * BN);
        uint32_t started = 0;                              // bit a set: accumulator a has been written in this tile
        for (int cb = 0; cb < p.cblocks; cb++, pa++) {
          const int sa = pa % SA;
          mbar_wait(smem_u32(&a_full[sa]), (pa / SA) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_base + sa * p.a_stage_bytes);
          for (int tap = 0; tap < p.taps; tap++, kb++) {
            const int sb = kb % SB;
            mbar_wait(smem_u32(&b_full[sb]), (kb / SB) & 1);
            tc_fence_after();
            const int ky = tap / d.kw, kx = tap - ky * d.kw;
            int aoff, acc = 0;
            if (MODE == 0) {
              const int pr = d.transposed ? (d.kh - 1 - ky) : ky;      // patch row / col offset of this tap
              const int pc = d.transposed ? (d.kw - 1 - kx) : kx;
              aoff = pr * p.HC + pc;
            } else if (MODE == 1) {
              aoff = ky * p.HC + (kx & 1) * p.QP + (kx >> 1);
            } else {
              aoff = (p.top - (ky >> 1)) * p.HC + (p.left - (kx >> 1));
              acc = (ky & 1) * 2 + (kx & 1);
            }
            const uint32_t a_tap = a_addr + (uint32_t)aoff * 16;
            const uint32_t b_addr = smem_u32(b_base + sb * B_STAGE_BYTES);
            const uint32_t first = (started >> acc) & 1u;
#pragma unroll
            for (int kk = 0; kk < CH / 2; kk++) {
              const uint64_t adesc = make_smem_desc(a_tap + kk * 2 * p.lbo, p.lbo, sbo);
              const uint64_t bdesc = make_smem_desc(b_addr + kk * 2 * (BN * 16), BN * 16, 128);
              umma<KIND>(tmem_d + acc * BN, adesc, bdesc, IDESC, (first | (uint32_t)kk) ? 1u : 0u);
            }
            started |= 1u << acc;
            umma_commit(smem_u32(&b_empty[sb]));
          }
          umma_commit(smem_u32(&a_empty[sa]));
        }
        umma_commit(smem_u32(&acc_full[buf]));
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // =========================== weight loader ===========================
    if (lane == 0) {
      int kb = 0;
      // packed image: [ntile][tap][128-byte channel block][chunk (8)][row (BN)][16 B]; a stage takes CH chunks of it
      constexpr int HALVES = 8 / CH;
      const int cb128 = (p.cblocks + HALVES - 1) / HALVES;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int ntile = (int)(tile % p.ntiles);
        const uint8_t* wsrc = (const uint8_t*)p.wpack + (int64_t)ntile * p.taps * cb128 * (BN * 128);
        for (int cb = 0; cb < p.cblocks; cb++) {
          for (int tap = 0; tap < p.taps; tap++, kb++) {
            const int sb = kb % SB;
            mbar_wait(smem_u32(&b_empty[sb]), ((kb / SB) & 1) ^ 1);
            const uint32_t bar = smem_u32(&b_full[sb]);
            mbar_arrive_expect_tx(bar, B_STAGE_BYTES);
            bulk_copy_g2s(smem_u32(b_base + sb * B_STAGE_BYTES),
                          wsrc + ((int64_t)tap * cb128 + cb / HALVES) * (BN * 128) + (cb % HALVES) * B_STAGE_BYTES,
                          B_STAGE_BYTES, bar);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue (warps 0-3) ===========================
    const int m = threadIdx.x;                         // TMEM lane = tile row
    const int ty = m >> 3, tx = m & 7;
    const float alpha = d.alpha, gain = d.gain, clamp = d.clamp;
    uint8_t* stg = smem + p.stg_off + warp * (32 * PITCH);
    const int q_chunk = lane % LPP;                    // store phase: 16-byte chunk of the slab / first pixel of the warp
    const int q_pix = lane / LPP;
    int li = 0;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, li++) {
      int ntile, u0, x0;
      decode_tile(p, tile, ntile, u0, x0);
      const int buf = (NBUF == 2) ? (li & 1) : 0;
      const uint32_t fph = (NBUF == 2) ? ((li >> 1) & 1) : (li & 1);
      const int o_base = ntile * BN;
      mbar_wait(smem_u32(&acc_full[buf]), fph);
      tc_fence_after();
#pragma unroll 1
      for (int ph = 0; ph < NACC; ph++) {
        const int py = ph >> 1, px = ph & 1;
        // output pixel of tile row mm (this thread's own row for the maths, other rows in the store phase)
        auto out_pixel = [&](int mm, int& n, int& oy, int& ox) -> bool {
          const int u = u0 + (mm >> 3);
          n = u / p.VR;
          const int r = u - n * p.VR, cx = x0 + (mm & 7);
          if (MODE == 2) { oy = 2 * r + py - d.pad_y; ox = 2 * cx + px - d.pad_x; }
          else { oy = r; ox = cx; }
          return n < d.n && oy >= 0 && oy < d.out_h && ox >= 0 && ox < d.out_w;
        };
        int n, oy, ox;
        const bool row_ok = out_pixel(m, n, oy, ox);
        const float* out_scale = (d.out_scale && row_ok) ? (const float*)d.out_scale + (int64_t)n * d.co : nullptr;
        const float nz = (d.noise && row_ok) ? ((const float*)d.noise)[((int64_t)n * d.out_h + oy) * d.out_w + ox] : 0.f;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (NACC * BN) + ph * BN;
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
                Vec16<T> pk;
#pragma unroll
                for (int e = 0; e < TC; e++) pk.v[e] = from_acc<T>(val[g * TC + e]);
                *(uint4*)(stg + lane * PITCH + cs * (int)sizeof(T) + g * 16) = pk.raw;
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
                if (yoff[k] >= 0) {
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

// ---- host side -------------------------------------------------------------------------------------------
// geometry the halo kernel handles; 0 / 1 / 2 = MODE, -1 = not eligible
int conv_halo_mode(const sgb_conv_desc* d) {
  if (d->kh > 3 || d->kw > 3) return -1;
  if ((int64_t)d->n * d->x_strides[0] >= (int64_t)1 << 31) return -1;        // int32 offsets in the producer
  if (d->stride == 1) {
    // geometry must be the "every tap stays inside the padded image" kind (always true for valid descriptors)
    if (!d->transposed) { if (d->out_h != d->in_h + 2 * d->pad_y - d->kh + 1 || d->out_w != d->in_w + 2 * d->pad_x - d->kw + 1) return -1; }
    else { if (d->out_h != d->in_h - 2 * d->pad_y + d->kh - 1 || d->out_w != d->in_w - 2 * d->pad_x + d->kw - 1) return -1;
           if (d->pad_y > d->kh - 1 || d->pad_x > d->kw - 1) return -1; }
    return 0;
  }
  if (d->stride == 2) {
    if (d->kh != 3 || d->kw != 3) return -1;
    return d->transposed ? 2 : 1;
  }
  return -1;
}
bool conv_halo_eligible(const sgb_conv_desc* d) { return conv_halo_mode(d) >= 0; }

// output-channel tile of the tensor-core kernels for this descriptor (also fixes the packed-weight layout)
int conv_bn(const sgb_conv_desc* d) {
  const int bn = pick_bn(d->co);
  if (d->force_simt != 2 && conv_halo_mode(d) == 2 && bn > 128) return 128;   // 4 accumulators must fit 512 TMEM columns
  return bn;
}

template <class T, int KIND, int BN, int MODE>
static int launch_halo(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  constexpr int TC = 16 / sizeof(T);
  constexpr int CH = (MODE == 1) ? 4 : 8;
  constexpr int SLAB = (BN * (int)sizeof(T) >= 128) ? 128 / (int)sizeof(T) : BN;
  constexpr int PITCH = SLAB * (int)sizeof(T) + 16;
  HaloParams p; p.d = *d; p.x = x; p.y = y; p.wpack = d->workspace;
  p.QP = 0;
  int rows_img;                                          // rows of the tile-row space that carry data, per image
  if (MODE == 0) {
    p.VR = d->out_h + d->kh - 1;
    p.HR = TILE_H + d->kh - 1; p.HC = TILE_W + d->kw - 1;
    p.top = d->transposed ? (d->kh - 1 - d->pad_y) : d->pad_y;
    p.left = d->transposed ? (d->kw - 1 - d->pad_x) : d->pad_x;
    rows_img = p.VR;
    p.col_tiles = (d->out_w + TILE_W - 1) / TILE_W;
  } else if (MODE == 1) {
    p.HR = 2 * (TILE_H - 1) + d->kh; p.HC = 2 * (TILE_W - 1) + d->kw;
    p.QP = (p.HC + 1) / 2;
    p.top = d->pad_y; p.left = d->pad_x;
    rows_img = d->out_h;
    p.VR = (rows_img + TILE_H - 1) / TILE_H * TILE_H;
    p.col_tiles = (d->out_w + TILE_W - 1) / TILE_W;
  } else {
    p.top = (d->kh - 1) >> 1; p.left = (d->kw - 1) >> 1;
    p.HR = TILE_H + p.top; p.HC = TILE_W + p.left;
    rows_img = ((d->out_h - 1 + d->pad_y) >> 1) + 1;
    p.VR = (rows_img + TILE_H - 1) / TILE_H * TILE_H;
    p.col_tiles = ((((d->out_w - 1 + d->pad_x) >> 1) + 1) + TILE_W - 1) / TILE_W;
  }
  const int64_t row_tiles = (MODE == 0) ? ceil_div((int64_t)d->n * p.VR, TILE_H) : (int64_t)d->n * (p.VR / TILE_H);
  p.ntiles = (d->co + BN - 1) / BN;
  p.total_tiles = row_tiles * p.col_tiles * p.ntiles;
  p.taps = d->kh * d->kw;
  p.cblocks = (d->ci + CH * TC - 1) / (CH * TC);
  int npix = p.HR * p.HC;
  while (npix % 8 != 1) npix++;                       // chunk planes 16 B (mod 128 B) apart: conflict-free 128-bit stores
  p.lbo = npix * 16;
  p.a_stage_bytes = (CH * p.lbo + 127) / 128 * 128;
  const bool y_al = aligned16(y) && d->y_strides[0] % TC == 0 && d->y_strides[2] % TC == 0 && d->y_strides[3] % TC == 0;
  p.vec_store = (d->co % TC == 0 && y_al) ? 1 : 0;
  SGB_REQUIRE(aligned16(x) && aligned16(d->workspace), "x and workspace must be 16-byte aligned");
  SGB_REQUIRE(p.HR * p.HC * CH <= ((MODE == 1) ? 18 : 12) * 128, "patch too large");
  if (int r = pack_weights_umma(d, w, BN, s)) return r;
  // shared-memory plan: staging slabs, then as many patch stages as useful, the rest for weight stages
  const int budget = 225 * 1024;
  const int stg_bytes = 4 * 32 * PITCH;
  const int b_stage = BN * CH * 16;
  int sa = MAX_SA, sb;
  for (;; sa--) {
    sb = (budget - stg_bytes - sa * p.a_stage_bytes) / b_stage;
    if (sb >= (sa > 3 ? 4 : 2) || sa == 3) break;
  }
  if (sb > MAX_SB) sb = MAX_SB;
  SGB_REQUIRE(sb >= 2, "shared memory budget exceeded");
  p.sa = sa; p.sb = sb;
  p.stg_off = sa * p.a_stage_bytes + sb * b_stage;
  const size_t smem = (size_t)p.stg_off + stg_bytes + 1024;
  auto kern = conv_halo_kernel<T, KIND, BN, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    SGB_REQUIRE(e == cudaSuccess, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  kern<<<(unsigned)grid, HALO_THREADS, smem, s>>>(p);
  SGB_LAUNCH_CHECK();
  return 0;
}

template <class T, int KIND, int MODE>
static int dispatch_halo_bn(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  switch (conv_bn(d)) {
    case 16:  return launch_halo<T, KIND, 16, MODE>(d, x, w, y, s);
    case 32:  return launch_halo<T, KIND, 32, MODE>(d, x, w, y, s);
    case 64:  return launch_halo<T, KIND, 64, MODE>(d, x, w, y, s);
    case 128: return launch_halo<T, KIND, 128, MODE>(d, x, w, y, s);
    default:
      if (MODE == 2) { set_error("conv_halo: BN 256 is not available for the transposed stride-2 geometry"); return 1; }
      return launch_halo<T, KIND, (MODE == 2 ? 128 : 256), MODE>(d, x, w, y, s);
  }
}

template <class T, int KIND>
static int dispatch_halo_mode(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  switch (conv_halo_mode(d)) {
    case 0: return dispatch_halo_bn<T, KIND, 0>(d, x, w, y, s);
    case 1: return dispatch_halo_bn<T, KIND, 1>(d, x, w, y, s);
    case 2: return dispatch_halo_bn<T, KIND, 2>(d, x, w, y, s);
  }
  set_error("conv_halo: geometry not supported");
  return 1;
}

int conv_forward_halo(const sgb_conv_desc* d, const void* x, const void* w, void* y, cudaStream_t s) {
  if (d->dtype == SGB_F16) return dispatch_halo_mode<__half, 0>(d, x, w, y, s);
  if (d->dtype == SGB_BF16) return dispatch_halo_mode<__nv_bfloat16, 1>(d, x, w, y, s);
  return dispatch_halo_mode<float, 2>(d, x, w, y, s);
}

}  // namespace sgb
