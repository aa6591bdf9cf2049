// Backward of the convolution's fused epilogue  y = clamp(lrelu|linear(conv * out_scale[n,c] + noise[n,h,w] + bias[c]) * gain)
// in ONE pass over (dy, y) -- channels_last tensors, C a multiple of one 16-byte vector.
//
// The reference runs this as five activation-sized passes (bias_act backward bias_act.py:161-175 + `dx.sum` for db,
// fma backward fma.py:37-58: dy * dcoefs, sum(dy * x) for d(dcoefs), sum_c(dy) for d(noise)).  Here:
//     dz      = dy * gain * (y > 0 ? 1 : alpha), 0 where |y| >= clamp       (bias_act.cu:141: mask on the stored output)
//     dconv   = dz * out_scale[n,c]                                -> written (the only activation-sized output)
//     dbias   = sum_{n,h,w} dz                                     -> [C]      (red.add after a block reduction)
//     dnoise  = sum_c dz                                           -> [N,H,W]  (red.add, one per thread per pixel)
//     dscale  = sum_{h,w} dz * conv,  conv = (z - noise - bias) / out_scale,  z = y / gain (y > 0) or y / (gain * alpha)
//               -> [N,C]  -- the pre-activation is re-derived from y (exact where it matters: clamped elements have dz = 0),
//               so the convolution output never has to be stored
// All reduction outputs are fp32 and must be zeroed by the caller... no: they are zeroed here (cudaMemsetAsync).
#include "common.cuh"

namespace sgb {
// Blocks of `kern` (256 threads, no dynamic shared memory) that are resident on the whole GPU at once.  The one-pass kernels
// below give every image a whole number of blocks; asking for "4 per SM" when the register count only lets 2 be resident made
// a grid of 2.05 waves run as three rounds (fused_epilogue_bwd at 0.64 of the copy rate, profiles/README.md).  The grid is
// sized to ONE resident wave, rounded DOWN per image.
template <class K>
static int64_t resident_blocks(K kern) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0) != cudaSuccess || occ < 1) { (void)cudaGetLastError(); occ = 2; }
  return (int64_t)occ * num_sms();
}
// blocks per image: floor(resident / n), at least 1, at most one pixel lane per block row
static inline int64_t blocks_per_image(int64_t resident, int n, int64_t hw, int lanes) {
  int64_t want = resident / (n > 0 ? n : 1);
  const int64_t maxb = (hw + lanes - 1) / lanes;
  if (want > maxb) want = maxb;
  if (want < 1) want = 1;
  return want;
}
}  // namespace sgb

namespace sgb {

struct FusedBwdParams {
  const void* dy; const void* y; void* dconv;
  const void* bias; const float* out_scale; const float* noise;
  float* dbias; float* dnoise; float* dscale;
  int n, c, hw;
  float alpha, gain, clamp;
  int ppb;            // pixels per block (all of one image)
  int blocks_per_img;
};

// block = 256 threads: thread -> (channel vector cv = tid % CVT, pixel lane pl = tid / CVT); CVT = c / VEC <= 256
template <class T>
__global__ void __launch_bounds__(256) fused_epilogue_bwd_kernel(FusedBwdParams p) {
  constexpr int VEC = Vec16<T>::N;
  __shared__ float red[256 * 2];
  const int cvt = p.c / VEC;
  const int lanes = 256 / cvt;                        // pixel lanes per block
  const int cv = threadIdx.x % cvt, pl = threadIdx.x / cvt;
  const int n = blockIdx.x / p.blocks_per_img;
  const int p0 = (blockIdx.x - n * p.blocks_per_img) * p.ppb;
  const int p1 = (p0 + p.ppb < p.hw) ? p0 + p.ppb : p.hw;
  const int c0 = cv * VEC;
  const bool active = pl < lanes;
  float sc[VEC], bs[VEC], inv_sc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; j++) {
    sc[j] = p.out_scale ? p.out_scale[(int64_t)n * p.c + c0 + j] : 1.f;
    inv_sc[j] = 1.f / sc[j];
    bs[j] = p.bias ? to_acc<T>(((const T*)p.bias)[c0 + j]) : 0.f;
  }
  // clamp rounded to the storage type, like bias_act.cu, so that saturated 16-bit outputs are recognised
  const float clampr = p.clamp >= 0.f ? to_acc<T>(from_acc<T>(p.clamp)) : -1.f;
  const float inv_g = 1.f / p.gain, inv_ga = (p.alpha != 0.f) ? 1.f / (p.gain * p.alpha) : 0.f;
  float accb[VEC], accs[VEC];
#pragma unroll
  for (int j = 0; j < VEC; j++) { accb[j] = 0.f; accs[j] = 0.f; }
  const int64_t img = (int64_t)n * p.hw;
  // lanes that share a pixel: cvt of them when cvt is a power of two <= 32 (one group = one pixel), else 32 when cvt is a
  // multiple of 32 (a warp = 32 channel vectors of one pixel); otherwise no shuffle reduction
  const bool cvt_pow2 = (cvt & (cvt - 1)) == 0;
  const int grp = (cvt_pow2 && cvt <= 32) ? cvt : 32;
  const bool grp_shfl = cvt_pow2;                         // power of two: <= 32 -> groups inside a warp, > 32 -> whole warps
  const unsigned gmask = grp >= 32 ? 0xffffffffu : (((1u << grp) - 1u) << ((threadIdx.x & 31) & ~(grp - 1)));
  if (active) {
    constexpr int U = 4;                                  // pixels in flight per thread (loads issued before the maths)
    for (int px0 = p0 + pl; px0 < p1; px0 += U * lanes) {
      Vec16<T> dyv[U], yv[U];
      float nzv[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int px = px0 + u * lanes;
        if (px < p1) {
          const int64_t v = (img + px) * cvt + cv;
          dyv[u].raw = ld_stream((const uint4*)p.dy + v);
          if (p.y) yv[u].raw = ld_stream((const uint4*)p.y + v);
          nzv[u] = p.noise ? p.noise[img + px] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int px = px0 + u * lanes;
        if (px >= p1) break;
        const int64_t v = (img + px) * cvt + cv;
        Vec16<T> out;
        float dn = 0.f;
#pragma unroll
        for (int j = 0; j < VEC; j++) {
          const float yy = p.y ? to_acc<T>(yv[u].v[j]) : 1.f;       // y not kept (linear, no clamp, no dscale): slope 1, never masked
          const bool pos = yy > 0.f;
          float dz = to_acc<T>(dyv[u].v[j]) * p.gain * (pos ? 1.f : p.alpha);
          if (clampr >= 0.f && !(yy > -clampr && yy < clampr)) dz = 0.f;
          const float z = yy * (pos ? inv_g : inv_ga);
          accb[j] += dz;
          accs[j] += dz * (z - nzv[u] - bs[j]) * inv_sc[j];
          dn += dz;
          out.v[j] = from_acc<T>(dz * sc[j]);
        }
        st_stream((uint4*)p.dconv + v, out.raw);
        if (p.dnoise) {
          // the cvt threads of a pixel are consecutive lanes with the same trip count: reduce over them with shuffles (group
          // mask), then ONE access per pixel and warp (it was one atomic per thread: 16-way contention at 64 fp32 channels)
          if (grp_shfl) {
            for (int o = grp >> 1; o > 0; o >>= 1) dn += __shfl_xor_sync(gmask, dn, o);
            if ((cv & (grp - 1)) == 0) {
              if (cvt <= 32) p.dnoise[img + px] = dn;                 // the whole pixel is in this group
              else atomicAdd(p.dnoise + img + px, dn);
            }
          } else {
            atomicAdd(p.dnoise + img + px, dn);
          }
        }
      }
    }
  }
  // reduce the per-thread channel sums over the pixel lanes of the block, then one red.add per channel per block
  if (p.dbias || p.dscale) {
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      __syncthreads();
      red[threadIdx.x] = active ? accb[j] : 0.f;
      red[256 + threadIdx.x] = active ? accs[j] : 0.f;
      __syncthreads();
      if (pl == 0) {
        float sb = 0.f, ss = 0.f;
        for (int l = 0; l < lanes; l++) { sb += red[l * cvt + cv]; ss += red[256 + l * cvt + cv]; }
        if (p.dbias) atomicAdd(p.dbias + c0 + j, sb);
        if (p.dscale) atomicAdd(p.dscale + (int64_t)n * p.c + c0 + j, ss);
      }
    }
  }
}

}  // namespace sgb

using namespace sgb;

extern "C" int sgb_fused_epilogue_bwd(const void* dy, const void* y, void* dconv, const void* bias, const void* out_scale,
                                      const void* noise, void* dbias, void* dnoise, void* dscale, int dtype, int n, int c, int hw,
                                      int act, float alpha, float gain, float clamp, void* stream) {
  SGB_REQUIRE(dtype == SGB_F32 || dtype == SGB_F16 || dtype == SGB_BF16, "unsupported dtype");
  SGB_REQUIRE(act == SGB_ACT_LINEAR || act == SGB_ACT_LRELU, "only linear and lrelu epilogues are fused");
  SGB_REQUIRE(n >= 0 && c >= 1 && hw >= 0, "bad sizes");
  const int vec = dtype == SGB_F32 ? 4 : 8;
  SGB_REQUIRE(c % vec == 0 && c / vec <= 256, "channels must be a multiple of one 16-byte vector (and at most 256 vectors)");
  SGB_REQUIRE(gain != 0.f, "gain must not be zero");
  cudaStream_t s = (cudaStream_t)stream;
  if (dbias) SGB_REQUIRE(cudaMemsetAsync(dbias, 0, sizeof(float) * c, s) == cudaSuccess, "memset failed");
  if (dscale) SGB_REQUIRE(cudaMemsetAsync(dscale, 0, sizeof(float) * (size_t)n * c, s) == cudaSuccess, "memset failed");
  if (dnoise) SGB_REQUIRE(cudaMemsetAsync(dnoise, 0, sizeof(float) * (size_t)n * hw, s) == cudaSuccess, "memset failed");
  if ((int64_t)n * hw == 0) return 0;
  SGB_REQUIRE(dy && dconv, "dy and dconv must not be NULL");
  SGB_REQUIRE(y || (act == SGB_ACT_LINEAR && clamp < 0.f && !dscale), "y may only be NULL for a linear epilogue without clamp and without dscale");
  SGB_REQUIRE(aligned16(dy) && aligned16(y) && aligned16(dconv), "tensors must be 16-byte aligned");
  FusedBwdParams p;
  p.dy = dy; p.y = y; p.dconv = dconv; p.bias = bias; p.out_scale = (const float*)out_scale; p.noise = (const float*)noise;
  p.dbias = (float*)dbias; p.dnoise = (float*)dnoise; p.dscale = (float*)dscale;
  p.n = n; p.c = c; p.hw = hw;
  p.alpha = (act == SGB_ACT_LINEAR) ? 1.f : alpha; p.gain = gain; p.clamp = clamp;
  // ~4 blocks per SM in total, whole blocks inside one image
  const int cvt = c / vec, lanes = 256 / cvt;
  static const int64_t res_f32 = resident_blocks(fused_epilogue_bwd_kernel<float>), res_f16 = resident_blocks(fused_epilogue_bwd_kernel<__half>),
                       res_bf16 = resident_blocks(fused_epilogue_bwd_kernel<__nv_bfloat16>);
  const int64_t want = blocks_per_image(dtype == SGB_F32 ? res_f32 : (dtype == SGB_F16 ? res_f16 : res_bf16), n, hw, lanes);
  p.ppb = (int)(((int64_t)hw + want - 1) / want);
  p.blocks_per_img = (hw + p.ppb - 1) / p.ppb;
  const int64_t grid = (int64_t)n * p.blocks_per_img;
  SGB_REQUIRE(grid <= 0x7fffffff, "grid too large");
  switch (dtype) {
    case SGB_F32:  fused_epilogue_bwd_kernel<float><<<(unsigned)grid, 256, 0, s>>>(p); break;
    case SGB_F16:  fused_epilogue_bwd_kernel<__half><<<(unsigned)grid, 256, 0, s>>>(p); break;
    default:       fused_epilogue_bwd_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, s>>>(p); break;
  }
  SGB_LAUNCH_CHECK();
  return 0;
}

// ---- forward of the same epilogue as a stand-alone pass ----------------------------------------------------------
//   y = clamp( lrelu|linear( x * out_scale[n,c] + noise[n,h,w] + bias[c] ) * gain )        channels_last, C % VEC == 0
// One pass instead of the reference's fma (generators.py:83) followed by bias_act (generators.py:328); used for the
// modulated layers whose convolution is followed by a FIR pass (up-sampling) or whose conv epilogue is not fused.
namespace sgb {
struct TailParams {
  const void* x; void* y; const void* bias; const float* out_scale; const float* noise;
  int n, c; int64_t hw;
  float alpha, gain, clamp;
  int ppb, blocks_per_img;      // pixels per block (inside one image) / blocks per image
};

template <class T>
__global__ void __launch_bounds__(256) scale_bias_act_kernel(TailParams p) {
  // thread -> (channel vector cv = tid % cvt, pixel lane pl = tid / cvt) of a chunk of ONE image: the scale and bias of its
  // channels are registers, the index arithmetic is 32-bit and division-free in the loop, four 16-byte loads are in flight per
  // thread.  (The first version decoded every vector with two 64-bit divisions and fetched scale / bias per element: 0.62 of
  // the copy rate in the step, profiles/README.md.)
  constexpr int VEC = Vec16<T>::N;
  const int cvt = p.c / VEC;
  const int lanes = 256 / cvt;
  const int cv = threadIdx.x % cvt, pl = threadIdx.x / cvt;
  if (pl >= lanes) return;
  const int n = blockIdx.x / p.blocks_per_img;
  const int p0 = (blockIdx.x - n * p.blocks_per_img) * p.ppb;
  const int p1 = ((int64_t)p0 + p.ppb < p.hw) ? p0 + p.ppb : (int)p.hw;
  const int c0 = cv * VEC;
  float sc[VEC], bs[VEC];
#pragma unroll
  for (int j = 0; j < VEC; j++) {
    sc[j] = p.out_scale ? p.out_scale[(int64_t)n * p.c + c0 + j] : 1.f;
    bs[j] = p.bias ? to_acc<T>(((const T*)p.bias)[c0 + j]) : 0.f;
  }
  const float e_clamp = p.clamp >= 0.f ? p.clamp : __int_as_float(0x7f800000);
  const int64_t img = (int64_t)n * p.hw;
  constexpr int U = 4;
  for (int px0 = p0 + pl; px0 < p1; px0 += U * lanes) {
    Vec16<T> in[U];
    float nz[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int px = px0 + u * lanes;
      if (px < p1) {
        in[u].raw = ld_stream((const uint4*)p.x + (img + px) * cvt + cv);
        nz[u] = p.noise ? p.noise[img + px] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int px = px0 + u * lanes;
      if (px >= p1) break;
      Vec16<T> out;
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        float t = fmaf(to_acc<T>(in[u].v[j]), sc[j], nz[u] + bs[j]);
        t = (t > 0.f ? t : t * p.alpha) * p.gain;                       // explicit compares: a NaN pre-activation stays NaN, as in
        t = t < -e_clamp ? -e_clamp : (t > e_clamp ? e_clamp : t);      // bias_act.cu and torch.clamp (fmaxf / fminf drop NaN operands)
        out.v[j] = from_acc<T>(t);
      }
      st_stream((uint4*)p.y + (img + px) * cvt + cv, out.raw);
    }
  }
}
}  // namespace sgb

extern "C" int sgb_scale_bias_act(const void* x, const void* bias, const void* out_scale, const void* noise, void* y, int dtype,
                                  int n, int c, int hw, int act, float alpha, float gain, float clamp, void* stream) {
  SGB_REQUIRE(dtype == SGB_F32 || dtype == SGB_F16 || dtype == SGB_BF16, "unsupported dtype");
  SGB_REQUIRE(act == SGB_ACT_LINEAR || (act == SGB_ACT_LRELU && alpha <= 1.f), "only linear and lrelu (alpha <= 1) are fused");
  const int vec = dtype == SGB_F32 ? 4 : 8;
  SGB_REQUIRE(n >= 0 && c >= 1 && hw >= 0 && c % vec == 0, "channels must be a multiple of one 16-byte vector");
  if ((int64_t)n * hw == 0) return 0;
  SGB_REQUIRE(x && y && aligned16(x) && aligned16(y), "x and y must be 16-byte aligned and not NULL");
  sgb::TailParams p;
  p.x = x; p.y = y; p.bias = bias; p.out_scale = (const float*)out_scale; p.noise = (const float*)noise;
  p.n = n; p.c = c; p.hw = hw; p.alpha = (act == SGB_ACT_LINEAR) ? 1.f : alpha; p.gain = gain; p.clamp = clamp;
  SGB_REQUIRE(c / vec <= 256, "at most 256 channel vectors");
  // ~8 blocks per SM in total, whole blocks inside one image
  const int cvt = c / vec, lanes = 256 / cvt;
  static const int64_t res_f32 = sgb::resident_blocks(sgb::scale_bias_act_kernel<float>), res_f16 = sgb::resident_blocks(sgb::scale_bias_act_kernel<__half>),
                       res_bf16 = sgb::resident_blocks(sgb::scale_bias_act_kernel<__nv_bfloat16>);
  const int64_t want = sgb::blocks_per_image(dtype == SGB_F32 ? res_f32 : (dtype == SGB_F16 ? res_f16 : res_bf16), n, hw, lanes);
  p.ppb = (int)(((int64_t)hw + want - 1) / want);
  p.blocks_per_img = (int)((hw + p.ppb - 1) / p.ppb);
  const int64_t blocks = (int64_t)n * p.blocks_per_img;
  SGB_REQUIRE(blocks <= 0x7fffffff, "grid too large");
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case SGB_F32:  sgb::scale_bias_act_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(p); break;
    case SGB_F16:  sgb::scale_bias_act_kernel<__half><<<(unsigned)blocks, 256, 0, s>>>(p); break;
    default:       sgb::scale_bias_act_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>(p); break;
  }
  SGB_LAUNCH_CHECK();
  return 0;
}

// ---- backward of the fused style modulation (conv in_scale) in one pass -------------------------------------------
//   gx[n,c,h,w] = g * s[n,c]           (gradient wrt the un-modulated activations)
//   gs[n,c]     = sum_hw g * x         (gradient wrt the styles)
// g = data gradient of the convolution wrt (x * s).  The reference reads g twice (generators.py:80 backward: mul by
// styles, and the un-broadcast sum of g * x); channels_last, C % VEC == 0, C / VEC <= 256.
namespace sgb {
struct ModBwdParams { const void* g; const void* x; const float* s; void* gx; float* gs; int n, c, hw, ppb, blocks_per_img; };

template <class T>
__global__ void __launch_bounds__(256) mod_bwd_kernel(ModBwdParams p) {
  constexpr int VEC = Vec16<T>::N;
  __shared__ float red[256];
  const int cvt = p.c / VEC, lanes = 256 / cvt;
  const int cv = threadIdx.x % cvt, pl = threadIdx.x / cvt;
  const int n = blockIdx.x / p.blocks_per_img;
  const int p0 = (blockIdx.x - n * p.blocks_per_img) * p.ppb;
  const int p1 = (p0 + p.ppb < p.hw) ? p0 + p.ppb : p.hw;
  const bool active = pl < lanes;
  const int c0 = cv * VEC;
  float sc[VEC], acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; j++) { sc[j] = p.s[(int64_t)n * p.c + c0 + j]; acc[j] = 0.f; }
  const int64_t img = (int64_t)n * p.hw;
  if (active) {
    constexpr int U = 4;
    for (int px0 = p0 + pl; px0 < p1; px0 += U * lanes) {
      Vec16<T> gv[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int px = px0 + u * lanes;
        if (px < p1) {
          const int64_t v = (img + px) * cvt + cv;
          gv[u].raw = ld_stream((const uint4*)p.g + v);
          if (p.gs) xv[u].raw = ld_stream((const uint4*)p.x + v);
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int px = px0 + u * lanes;
        if (px >= p1) break;
        const int64_t v = (img + px) * cvt + cv;
        Vec16<T> out;
#pragma unroll
        for (int j = 0; j < VEC; j++) {
          const float gg = to_acc<T>(gv[u].v[j]);
          if (p.gs) acc[j] += gg * to_acc<T>(xv[u].v[j]);
          out.v[j] = from_acc<T>(gg * sc[j]);
        }
        if (p.gx) st_stream((uint4*)p.gx + v, out.raw);
      }
    }
  }
  if (p.gs) {
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      __syncthreads();
      red[threadIdx.x] = active ? acc[j] : 0.f;
      __syncthreads();
      if (pl == 0) {
        float t = 0.f;
        for (int l = 0; l < lanes; l++) t += red[l * cvt + cv];
        atomicAdd(p.gs + (int64_t)n * p.c + c0 + j, t);
      }
    }
  }
}
}  // namespace sgb

extern "C" int sgb_mod_bwd(const void* g, const void* x, const void* s, void* gx, void* gs, int dtype, int n, int c, int hw, void* stream) {
  SGB_REQUIRE(dtype == SGB_F32 || dtype == SGB_F16 || dtype == SGB_BF16, "unsupported dtype");
  const int vec = dtype == SGB_F32 ? 4 : 8;
  SGB_REQUIRE(n >= 0 && c >= 1 && hw >= 0 && c % vec == 0 && c / vec <= 256, "channels must be a multiple of one 16-byte vector (at most 256 vectors)");
  cudaStream_t st = (cudaStream_t)stream;
  if (gs) SGB_REQUIRE(cudaMemsetAsync(gs, 0, sizeof(float) * (size_t)n * c, st) == cudaSuccess, "memset failed");
  if ((int64_t)n * hw == 0) return 0;
  SGB_REQUIRE(g && s && (gx || gs) && (x || !gs), "g, s and at least one output must not be NULL");
  SGB_REQUIRE(aligned16(g) && aligned16(x) && aligned16(gx), "tensors must be 16-byte aligned");
  sgb::ModBwdParams p;
  p.g = g; p.x = x; p.s = (const float*)s; p.gx = gx; p.gs = (float*)gs; p.n = n; p.c = c; p.hw = hw;
  const int cvt = c / vec, lanes = 256 / cvt;
  static const int64_t res_f32 = sgb::resident_blocks(sgb::mod_bwd_kernel<float>), res_f16 = sgb::resident_blocks(sgb::mod_bwd_kernel<__half>),
                       res_bf16 = sgb::resident_blocks(sgb::mod_bwd_kernel<__nv_bfloat16>);
  const int64_t want = sgb::blocks_per_image(dtype == SGB_F32 ? res_f32 : (dtype == SGB_F16 ? res_f16 : res_bf16), n, hw, lanes);
  p.ppb = (int)(((int64_t)hw + want - 1) / want);
  p.blocks_per_img = (hw + p.ppb - 1) / p.ppb;
  const int64_t grid = (int64_t)n * p.blocks_per_img;
  SGB_REQUIRE(grid <= 0x7fffffff, "grid too large");
  switch (dtype) {
    case SGB_F32:  sgb::mod_bwd_kernel<float><<<(unsigned)grid, 256, 0, st>>>(p); break;
    case SGB_F16:  sgb::mod_bwd_kernel<__half><<<(unsigned)grid, 256, 0, st>>>(p); break;
    default:       sgb::mod_bwd_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>(p); break;
  }
  SGB_LAUNCH_CHECK();
  return 0;
}
