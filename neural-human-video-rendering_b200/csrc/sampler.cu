// Texture lookup (IUV -> learned atlas, soft part blend) and mask/background composite.
// Both are HBM-bound streaming kernels: one thread per pixel (sampler) / per 4 pixels (composite),
// every global access is warp-coalesced along x; the atlas (<= 35 MB) stays L2-resident.
//
// Arithmetic contract (mirrors oracle/texture.py line by line; the integer part is bit-exact):
//   part   = argmax_k logits[k], k in 0..24, lowest index wins ties        (no transcendental)
//   u      = clamp(0.5*U + 0.5, 0, 1)   (exact in fp32: scaling by 0.5 is exact, one rounding)
//   fx     = u * (S-1);  x0 = floor(fx);  x1 = min(x0+1, S-1);  wx = fx - x0      (same for y)
//   sample = (1-wy)*((1-wx)*T[y0][x0] + wx*T[y0][x1]) + wy*((1-wx)*T[y1][x0] + wx*T[y1][x1])
//   tex    = sum_{k=1..24} softmax(logits)[k] * sample_k      (/(1-P0+1e-6) if !use_mask_texture)
#include "common.cuh"
#include "p8.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace nhvr {

extern void note_cuda_error(cudaError_t e);
extern void count_launch();
extern int arch_ok_cached();

constexpr int kParts = 24;

NHVR_DEVINL float uv_act(float t) { return fminf(fmaxf(__fadd_rn(__fmul_rn(t, 0.5f), 0.5f), 0.f), 1.f); }

template <int G, int MINB>   // G = number of float4 channel groups per texel (Ct4 / 4); MINB = resident blocks per SM asked for
__global__ void __launch_bounds__(128, MINB) texture_sample_kernel(const float* __restrict__ uvp, const float4* __restrict__ atlas,
                                                             int N, int H, int W, int S, int Ctex, int use_mask,
                                                             float* __restrict__ tex_out, uint8_t* __restrict__ part_out,
                                                             short2* __restrict__ texel_out) {
  const int64_t HW = (int64_t)H * W;
  const int64_t total = (int64_t)N * HW;
  const float sm1 = (float)(S - 1);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int64_t pix = idx - (int64_t)n * HW;
    const float* base = uvp + (int64_t)n * 73 * HW + pix;

    float lg[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) lg[k] = __ldg(base + (int64_t)k * HW);
    float mx = lg[0];
    int part = 0;
#pragma unroll
    for (int k = 1; k < 25; ++k) {
      if (lg[k] > mx) { mx = lg[k]; part = k; }
    }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 25; ++k) { lg[k] = __expf(lg[k] - mx); den += lg[k]; }
    const float inv_den = 1.f / den;

    float4 acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    short2 texel = make_short2(0, 0);

#pragma unroll
    for (int k = 1; k <= kParts; ++k) {
      const float u = uv_act(__ldg(base + (int64_t)(25 + k - 1) * HW));
      const float v = uv_act(__ldg(base + (int64_t)(49 + k - 1) * HW));
      const float fx = __fmul_rn(u, sm1), fy = __fmul_rn(v, sm1);
      const float x0f = floorf(fx), y0f = floorf(fy);
      const int x0 = (int)x0f, y0 = (int)y0f;
      const int x1 = min(x0 + 1, S - 1), y1 = min(y0 + 1, S - 1);
      const float wx = fx - x0f, wy = fy - y0f;
      if (k == part) texel = make_short2((short)x0, (short)y0);
      const float pk = lg[k] * inv_den;
      const float w00 = (1.f - wy) * (1.f - wx) * pk, w01 = (1.f - wy) * wx * pk;
      const float w10 = wy * (1.f - wx) * pk, w11 = wy * wx * pk;
      const float4* T = atlas + (int64_t)(k - 1) * S * S * G;
      const float4* t00 = T + ((int64_t)y0 * S + x0) * G;
      const float4* t01 = T + ((int64_t)y0 * S + x1) * G;
      const float4* t10 = T + ((int64_t)y1 * S + x0) * G;
      const float4* t11 = T + ((int64_t)y1 * S + x1) * G;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 a = __ldg(t00 + g), b = __ldg(t01 + g), c = __ldg(t10 + g), d = __ldg(t11 + g);
        acc[g].x += w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x;
        acc[g].y += w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y;
        acc[g].z += w00 * a.z + w01 * b.z + w10 * c.z + w11 * d.z;
        acc[g].w += w00 * a.w + w01 * b.w + w10 * c.w + w11 * d.w;
      }
    }
    float norm = 1.f;
    if (!use_mask) norm = 1.f / (1.f - lg[0] * inv_den + 1e-6f);
    float* o = tex_out + (int64_t)n * Ctex * HW + pix;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float vals[4] = {acc[g].x, acc[g].y, acc[g].z, acc[g].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = g * 4 + e;
        if (c < Ctex) o[(int64_t)c * HW] = vals[e] * norm;
      }
    }
    if (part_out) part_out[idx] = (uint8_t)part;
    if (texel_out) texel_out[idx] = texel;
  }
}

// ---- lean variant (Ctex <= 4, 73 * H * W < 2^31; the default there): the same arithmetic for part / texel, written for instruction count.
// cuobjdump counts 2928 SASS instructions per pixel in the kernel above (1017 integer / address, 1027 fp32, 169 loads) against 304
// bytes of algorithmic traffic.  Here 1640: one image per blockIdx.y, 32-bit element indices widened once per load,
// u = fma.rn.sat(U, 0.5, 0.5) (one instruction; identical to clamp(0.5 * U + 0.5) because the product is exact), chained FMAs
// for the blend, no .w channel for Ctex <= 3, the arg-max select chain and the texel select only when the indices are asked for
// (IDX), the texel of the arg-max part recomputed once after the loop.
// Measured (tools/sampler_bench.py, 8 x 512^2, us; profiles/r02c_sampler_lean.md): 44 % fewer instructions buy NOTHING where the lanes
// of a warp scatter over the atlas (white-noise UV 561 vs 557, smooth UV 467 vs 467: the L1 wavefront queue is the limit, ncu
// l1tex__data_pipe_lsu_wavefronts 81 %, ~13 lines per gather) and 21 % where they do not (98 % flat stick-figure UV: 229 vs 290);
// occupancy matters more than instructions (same code at 80 / 96 registers: 264 / 308 us flat), so the 64-register build is the default.
NHVR_DEVINL float fma_sat_half(float t) {
  float r;
  asm("fma.rn.sat.f32 %0, %1, 0f3F000000, 0f3F000000;" : "=f"(r) : "f"(t));
  return r;
}
NHVR_DEVINL float ex2_approx(float t) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}

template <int NCH, bool IDX, int MINB>
__global__ void __launch_bounds__(128, MINB) texture_sample_lean_kernel(const float* __restrict__ uvp, const float4* __restrict__ atlas,
                                                                  int HW, int S, int Ctex, int use_mask, float* __restrict__ tex_out,
                                                                  uint8_t* __restrict__ part_out, short2* __restrict__ texel_out) {
  const int n = blockIdx.y;
  const float* __restrict__ img = uvp + (size_t)n * 73u * (size_t)HW;
  const float sm1 = (float)(S - 1);
  const int SS = S * S;
  for (int pix = blockIdx.x * 128 + threadIdx.x; pix < HW; pix += gridDim.x * 128) {
    const float* __restrict__ base = img + pix;
    float lg[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) lg[k] = __ldg(base + (unsigned)(k * HW));
    float mx = lg[0];
    int part = 0;
    if (IDX) {
#pragma unroll
      for (int k = 1; k < 25; ++k)
        if (lg[k] > mx) { mx = lg[k]; part = k; }
    } else {
#pragma unroll
      for (int k = 1; k < 25; ++k) mx = fmaxf(mx, lg[k]);
    }
    const float kLog2e = 1.4426950408889634f;
    const float nmx = -mx * kLog2e;
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 25; ++k) { lg[k] = ex2_approx(fmaf(lg[k], kLog2e, nmx)); den += lg[k]; }
    const float inv_den = 1.f / den;

    float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
#pragma unroll
    for (int k = 1; k <= kParts; ++k) {
      const float u = fma_sat_half(__ldg(base + (unsigned)((24 + k) * HW)));
      const float v = fma_sat_half(__ldg(base + (unsigned)((48 + k) * HW)));
      const float fx = __fmul_rn(u, sm1), fy = __fmul_rn(v, sm1);
      const float x0f = floorf(fx), y0f = floorf(fy);
      const int x0 = (int)x0f, y0 = (int)y0f;
      const int x1 = min(x0 + 1, S - 1), y1 = min(y0 + 1, S - 1);
      const float wx = fx - x0f, wy = fy - y0f;
      const float pk = lg[k] * inv_den;
      const float c0 = (1.f - wy) * pk, c1 = wy * pk, omx = 1.f - wx;
      const float w00 = c0 * omx, w01 = c0 * wx, w10 = c1 * omx, w11 = c1 * wx;
      // one 32-bit texel index per corner (24 * S * S < 2^31, checked by the host), widened once by the load's address
      const int r0 = y0 * S + (k - 1) * SS, r1 = y1 * S + (k - 1) * SS;
      const float4 a = __ldg(atlas + (unsigned)(r0 + x0)), b = __ldg(atlas + (unsigned)(r0 + x1));
      const float4 c = __ldg(atlas + (unsigned)(r1 + x0)), d = __ldg(atlas + (unsigned)(r1 + x1));
      ax = fmaf(w11, d.x, fmaf(w10, c.x, fmaf(w01, b.x, fmaf(w00, a.x, ax))));
      ay = fmaf(w11, d.y, fmaf(w10, c.y, fmaf(w01, b.y, fmaf(w00, a.y, ay))));
      az = fmaf(w11, d.z, fmaf(w10, c.z, fmaf(w01, b.z, fmaf(w00, a.z, az))));
      if (NCH > 3) aw = fmaf(w11, d.w, fmaf(w10, c.w, fmaf(w01, b.w, fmaf(w00, a.w, aw))));
    }
    float norm = 1.f;
    if (!use_mask) norm = 1.f / (1.f - lg[0] * inv_den + 1e-6f);
    float* o = tex_out + (size_t)n * (size_t)Ctex * (size_t)HW + pix;
    o[0] = ax * norm;
    if (Ctex > 1) o[(unsigned)HW] = ay * norm;
    if (Ctex > 2) o[(unsigned)(2 * HW)] = az * norm;
    if (NCH > 3 && Ctex > 3) o[(unsigned)(3 * HW)] = aw * norm;
    if (IDX) {
      const size_t idx = (size_t)n * (size_t)HW + pix;
      if (part_out) part_out[idx] = (uint8_t)part;
      if (texel_out) {
        short2 texel = make_short2(0, 0);
        if (part > 0) {
          const float u = fma_sat_half(__ldg(base + (unsigned)((24 + part) * HW)));
          const float v = fma_sat_half(__ldg(base + (unsigned)((48 + part) * HW)));
          texel = make_short2((short)(int)floorf(__fmul_rn(u, sm1)), (short)(int)floorf(__fmul_rn(v, sm1)));
        }
        texel_out[idx] = texel;
      }
    }
  }
}

// Measured alternatives (profiles/r02b_sampler.md): a "pair" atlas (texel + right neighbour in one 32-byte record, one LDG.E.256
// per bilinear row) and a "quad" atlas with four lanes per (pixel, part) were built in round 2.  Neither beats this kernel: the
// gathers are bound by distinct L1 line misses (2 per pixel and part either way) and by issue latency (IPC ~1 of 4), not by
// sectors or tag lookups.

// out = m*fg + (1-m)*bg, 4 pixels per thread (float4 along x)
__global__ void __launch_bounds__(256) composite_kernel(const float4* __restrict__ fgm, const float4* __restrict__ bg, int bg_batched,
                                                        int N, int64_t HW4, float4* __restrict__ out) {
  const int64_t total = (int64_t)N * HW4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW4);
    const int64_t p = idx - (int64_t)n * HW4;
    const float4* f = fgm + (int64_t)n * 4 * HW4 + p;
    const float4 m = __ldg(f + 3 * HW4);
    const float4* b = bg + (bg_batched ? (int64_t)n * 3 * HW4 : 0) + p;
    float4* o = out + (int64_t)n * 3 * HW4 + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 fg = __ldg(f + c * HW4);
      const float4 bb = __ldg(b + c * HW4);
      float4 r;
      r.x = m.x * fg.x + (1.f - m.x) * bb.x;
      r.y = m.y * fg.y + (1.f - m.y) * bb.y;
      r.z = m.z * fg.z + (1.f - m.z) * bb.z;
      r.w = m.w * fg.w + (1.f - m.w) * bb.w;
      o[c * HW4] = r;
    }
  }
}

__global__ void __launch_bounds__(256) composite_scalar_kernel(const float* __restrict__ fgm, const float* __restrict__ bg, int bg_batched,
                                                               int N, int64_t HW, float* __restrict__ out) {
  const int64_t total = (int64_t)N * HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int64_t p = idx - (int64_t)n * HW;
    const float m = fgm[((int64_t)n * 4 + 3) * HW + p];
    for (int c = 0; c < 3; ++c) {
      const float fg = fgm[((int64_t)n * 4 + c) * HW + p];
      const float bb = bg[((bg_batched ? (int64_t)n * 3 : 0) + c) * HW + p];
      out[((int64_t)n * 3 + c) * HW + p] = m * fg + (1.f - m) * bb;
    }
  }
}

}  // namespace nhvr

using namespace nhvr;

extern "C" int nhvr_texture_sample(const float* uvp, const float* atlas, int32_t N, int32_t H, int32_t W, int32_t S,
                                   int32_t Ctex, int32_t use_mask_texture, float* tex_out, uint8_t* part_out,
                                   int16_t* texel_out, void* stream) {
  if (!uvp || !atlas || !tex_out) return NHVR_ERR_NULL;
  if (N <= 0 || H <= 0 || W <= 0 || S < 2 || S > 32767 || Ctex <= 0 || Ctex > 20) return NHVR_ERR_SHAPE;
  if (((uintptr_t)atlas & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int G = (Ctex + 3) / 4;
  const int64_t total = (int64_t)N * H * W;
  const float4* a4 = reinterpret_cast<const float4*>(atlas);
  short2* tx = reinterpret_cast<short2*>(texel_out);
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>((total + 127) / 128, (int64_t)148 * 16 * 8);
  // lean variant (see texture_sample_lean_kernel): Ctex <= 4 and 32-bit element indices; NHVR_SAMPLER_LEAN=0 selects the general kernel
  static int lean = -1;
  if (lean < 0) { const char* e = std::getenv("NHVR_SAMPLER_LEAN"); lean = e ? std::atoi(e) : 1; }
  const int64_t HW64 = (int64_t)H * W;
  if (lean && G == 1 && 73 * HW64 < (int64_t)1 << 31 && (int64_t)24 * S * S < (int64_t)1 << 31 && N <= 65535) {
    const int HWi = (int)HW64;
    const dim3 grid((unsigned)std::min<int64_t>((HW64 + 127) / 128, (int64_t)1 << 20), (unsigned)N);
    const bool idx = part_out || texel_out;
    static int lean_minb = -1;
    if (lean_minb < 0) { const char* e = std::getenv("NHVR_SAMPLER_LEAN_MINB"); lean_minb = e ? std::atoi(e) : 8; }
#define NHVR_LAUNCH_LEAN(NCH, IDX) \
    do { if (lean_minb >= 8) texture_sample_lean_kernel<NCH, IDX, 8><<<grid, 128, 0, st>>>(uvp, a4, HWi, S, Ctex, use_mask_texture, tex_out, part_out, tx); \
         else if (lean_minb >= 6) texture_sample_lean_kernel<NCH, IDX, 6><<<grid, 128, 0, st>>>(uvp, a4, HWi, S, Ctex, use_mask_texture, tex_out, part_out, tx); \
         else texture_sample_lean_kernel<NCH, IDX, 5><<<grid, 128, 0, st>>>(uvp, a4, HWi, S, Ctex, use_mask_texture, tex_out, part_out, tx); } while (0)
    if (Ctex <= 3) { if (idx) NHVR_LAUNCH_LEAN(3, true); else NHVR_LAUNCH_LEAN(3, false); }
    else { if (idx) NHVR_LAUNCH_LEAN(4, true); else NHVR_LAUNCH_LEAN(4, false); }
#undef NHVR_LAUNCH_LEAN
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
    return NHVR_OK;
  }
#define NHVR_LAUNCH_SAMPLER(GG) \
  texture_sample_kernel<GG, (GG == 1 ? 6 : 1)><<<blocks, 128, 0, st>>>(uvp, a4, N, H, W, S, Ctex, use_mask_texture, tex_out, part_out, tx)
  // Ctex <= 4: 8 resident blocks per SM asked for = 64 registers, 32 warps / SM.  Measured (tools/sampler_bench.py, us per launch at
  // B = 8, 512^2; white-noise / smooth / 98 % flat UV): 64 registers 557 / 468 / 290, 80 registers 557 / 479 / 336; without any
  // bound ptxas takes 254 registers (8 warps / SM): 457 vs 431 us in the frame step.  NHVR_SAMPLER_MINB=6 selects the 80-register build.
  static int minb = -1;
  if (minb < 0) { const char* e = std::getenv("NHVR_SAMPLER_MINB"); minb = e ? std::atoi(e) : 8; }
  switch (G) {
    case 1:
      if (minb >= 8) texture_sample_kernel<1, 8><<<blocks, 128, 0, st>>>(uvp, a4, N, H, W, S, Ctex, use_mask_texture, tex_out, part_out, tx);
      else NHVR_LAUNCH_SAMPLER(1);
      break;
    case 2: NHVR_LAUNCH_SAMPLER(2); break;
    case 3: NHVR_LAUNCH_SAMPLER(3); break;
    case 4: NHVR_LAUNCH_SAMPLER(4); break;
    default: NHVR_LAUNCH_SAMPLER(5); break;
  }
#undef NHVR_LAUNCH_SAMPLER
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

extern "C" int nhvr_composite(const float* fgm, const float* bg, int32_t bg_batched, int32_t N, int32_t H, int32_t W,
                              float* out, void* stream) {
  if (!fgm || !bg || !out) return NHVR_ERR_NULL;
  if (N <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int64_t HW = (int64_t)H * W;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (HW % 4 == 0) && ((((uintptr_t)fgm | (uintptr_t)bg | (uintptr_t)out) & 15) == 0);
  if (vec) {
    const int64_t total = (int64_t)N * (HW / 4);
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)148 * 8 * 4);
    composite_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(fgm), reinterpret_cast<const float4*>(bg),
                                             bg_batched, N, HW / 4, reinterpret_cast<float4*>(out));
  } else {
    const int64_t total = (int64_t)N * HW;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)148 * 8 * 4);
    composite_scalar_kernel<<<blocks, 256, 0, st>>>(fgm, bg, bg_batched, N, HW, out);
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

// =================================================================================================
// backward of the texture lookup (use_mask_texture variant) and of the composite
// =================================================================================================
namespace nhvr {

// tex_c = sum_{k>=1} P_k * s_kc,  s_kc = bilinear(T_k; fx_k, fy_k)
//   d tex_c / d logit_j = P_j * (s_jc [j>=1] - tex_c)
//   d tex_c / d U_k     = P_k * ds_kc/dfx * (S-1)/2 * [0 <= U/2+1/2 <= 1]          (same for V)
//   d tex_c / d T_k[corner] = P_k * w_corner
// Atlas gradient: vector reductions into a channels-last fp32 buffer [24][S][S][4G] (zeroed by the caller).
template <int G, int MINB>
__global__ void __launch_bounds__(128, MINB) texture_sample_bwd_kernel(const float* __restrict__ uvp, const float4* __restrict__ atlas,
                                                                 const float* __restrict__ gtex, int N, int H, int W, int S, int Ctex,
                                                                 int use_mask, float* __restrict__ guvp, float* __restrict__ gatlas, int agg) {
  const int64_t HW = (int64_t)H * W;
  const int64_t total = (int64_t)N * HW;
  const float sm1 = (float)(S - 1);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int64_t pix = idx - (int64_t)n * HW;
    const float* base = uvp + (int64_t)n * 73 * HW + pix;
    float* gb = guvp + (int64_t)n * 73 * HW + pix;
    const bool full_warp = agg && __activemask() == 0xffffffffu;
    float g[4 * G];
#pragma unroll
    for (int c = 0; c < 4 * G; ++c) g[c] = c < Ctex ? __ldg(gtex + ((int64_t)n * Ctex + c) * HW + pix) : 0.f;

    float lg[25], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 25; ++k) { lg[k] = __ldg(base + (int64_t)k * HW); mx = fmaxf(mx, lg[k]); }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 25; ++k) { lg[k] = __expf(lg[k] - mx); den += lg[k]; }
    const float inv_den = 1.f / den;
    // without --use_mask_texture the blend is renormalised: tex = S / D with S the masked blend and D = 1 - P0 + 1e-6.
    // dL/dS = g / D (applied by scaling g here), dL/dD = -(g . S) / D^2 flows into the logits through P0 (below)
    const float p0 = lg[0] * inv_den;
    const float invD = use_mask ? 1.f : 1.f / (1.f - p0 + 1e-6f);
    if (!use_mask) {
#pragma unroll
      for (int c = 0; c < 4 * G; ++c) g[c] *= invD;
    }
    float tdot = 0.f;           // sum_c g_c * tex_c = sum_k P_k a_k
    float a[25];
    a[0] = 0.f;
#pragma unroll
    for (int k = 1; k <= kParts; ++k) {
      const float tu = __fadd_rn(__fmul_rn(__ldg(base + (int64_t)(24 + k) * HW), 0.5f), 0.5f);
      const float tv = __fadd_rn(__fmul_rn(__ldg(base + (int64_t)(48 + k) * HW), 0.5f), 0.5f);
      const float u = fminf(fmaxf(tu, 0.f), 1.f), v = fminf(fmaxf(tv, 0.f), 1.f);
      const float fx = __fmul_rn(u, sm1), fy = __fmul_rn(v, sm1);
      const float x0f = floorf(fx), y0f = floorf(fy);
      const int x0 = (int)x0f, y0 = (int)y0f;
      const int x1 = min(x0 + 1, S - 1), y1 = min(y0 + 1, S - 1);
      const float wx = fx - x0f, wy = fy - y0f;
      const float pk = lg[k] * inv_den;
      // warp-uniform (U, V) of this part (see the atlas gradient below); only full warps take the aggregated path
      bool wu = false;
      if (full_warp) {
        int pu, pv;
        __match_all_sync(0xffffffffu, __float_as_uint(tu), &pu);
        __match_all_sync(0xffffffffu, __float_as_uint(tv), &pv);
        wu = pu && pv;
      }
      const int64_t tb = (int64_t)(k - 1) * S * S;
      const int64_t i00 = (tb + (int64_t)y0 * S + x0) * G, i01 = (tb + (int64_t)y0 * S + x1) * G;
      const int64_t i10 = (tb + (int64_t)y1 * S + x0) * G, i11 = (tb + (int64_t)y1 * S + x1) * G;
      const float w00 = (1.f - wy) * (1.f - wx), w01 = (1.f - wy) * wx, w10 = wy * (1.f - wx), w11 = wy * wx;
      float ak = 0.f, dfx = 0.f, dfy = 0.f;
#pragma unroll
      for (int q = 0; q < G; ++q) {
        const float4 A = __ldg(atlas + i00 + q), B = __ldg(atlas + i01 + q), Cc = __ldg(atlas + i10 + q), D = __ldg(atlas + i11 + q);
        const float av[4] = {A.x, A.y, A.z, A.w}, bv[4] = {B.x, B.y, B.z, B.w}, cv[4] = {Cc.x, Cc.y, Cc.z, Cc.w}, dv[4] = {D.x, D.y, D.z, D.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float gc = g[q * 4 + e];
          ak += gc * (w00 * av[e] + w01 * bv[e] + w10 * cv[e] + w11 * dv[e]);
          dfx += gc * ((1.f - wy) * (bv[e] - av[e]) + wy * (dv[e] - cv[e]));
          dfy += gc * ((1.f - wx) * (cv[e] - av[e]) + wx * (dv[e] - bv[e]));
        }
        // atlas gradient (vector reduction per corner).  Stick-figure pose maps are flat over most of the image, so whole warps
        // hit the SAME four texels of every part and 2 M same-address reductions per corner serialise in L2: when U and V of this
        // part are bit-identical across a full warp (hence the corner indices and weights too), the warp first sums pk * g over
        // its lanes and one lane issues the four reductions.
        if (wu) {
          float t0 = pk * g[q * 4], t1 = pk * g[q * 4 + 1], t2 = pk * g[q * 4 + 2], t3 = pk * g[q * 4 + 3];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o);
            t2 += __shfl_xor_sync(0xffffffffu, t2, o); t3 += __shfl_xor_sync(0xffffffffu, t3, o);
          }
          if ((threadIdx.x & 31) == 0) {
            float* gw = gatlas;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gw + (i00 + q) * 4), "f"(w00 * t0), "f"(w00 * t1), "f"(w00 * t2), "f"(w00 * t3) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gw + (i01 + q) * 4), "f"(w01 * t0), "f"(w01 * t1), "f"(w01 * t2), "f"(w01 * t3) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gw + (i10 + q) * 4), "f"(w10 * t0), "f"(w10 * t1), "f"(w10 * t2), "f"(w10 * t3) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gw + (i11 + q) * 4), "f"(w11 * t0), "f"(w11 * t1), "f"(w11 * t2), "f"(w11 * t3) : "memory");
          }
          continue;
        }
        const float s00 = pk * w00, s01 = pk * w01, s10 = pk * w10, s11 = pk * w11;
        float* ga = gatlas;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ga + (i00 + q) * 4), "f"(s00 * g[q * 4]), "f"(s00 * g[q * 4 + 1]),
                     "f"(s00 * g[q * 4 + 2]), "f"(s00 * g[q * 4 + 3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ga + (i01 + q) * 4), "f"(s01 * g[q * 4]), "f"(s01 * g[q * 4 + 1]),
                     "f"(s01 * g[q * 4 + 2]), "f"(s01 * g[q * 4 + 3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ga + (i10 + q) * 4), "f"(s10 * g[q * 4]), "f"(s10 * g[q * 4 + 1]),
                     "f"(s10 * g[q * 4 + 2]), "f"(s10 * g[q * 4 + 3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ga + (i11 + q) * 4), "f"(s11 * g[q * 4]), "f"(s11 * g[q * 4 + 1]),
                     "f"(s11 * g[q * 4 + 2]), "f"(s11 * g[q * 4 + 3]) : "memory");
      }
      a[k] = ak;
      tdot += pk * ak;
      const float cu = (tu >= 0.f && tu <= 1.f) ? 0.5f * sm1 * pk : 0.f;
      const float cv2 = (tv >= 0.f && tv <= 1.f) ? 0.5f * sm1 * pk : 0.f;
      gb[(int64_t)(24 + k) * HW] = cu * dfx;
      gb[(int64_t)(48 + k) * HW] = cv2 * dfy;
    }
    // (g' . S) = tdot with the scaled g'; dL/dD = -tdot / D; dD/dlogit_j = P0 P_j - [j == 0] P0
    const float gD = use_mask ? 0.f : -tdot * invD;
#pragma unroll
    for (int k = 0; k < 25; ++k) {
      const float pj = lg[k] * inv_den;
      gb[(int64_t)k * HW] = pj * (a[k] - tdot) + gD * p0 * (pj - (k == 0 ? 1.f : 0.f));
    }
  }
}

// out = m*fg + (1-m)*bg :  g_fg = m*g,  g_m = sum_c g_c (fg_c - bg_c),  g_bg = (1-m)*g (summed over the batch when bg is shared)
__global__ void __launch_bounds__(256) composite_bwd_kernel(const float* __restrict__ fgm, const float* __restrict__ bg, int bg_batched,
                                                            const float* __restrict__ gout, int N, int64_t HW, float* __restrict__ gfgm,
                                                            float* __restrict__ gbg) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    float acc[3] = {0.f, 0.f, 0.f};
    for (int n = 0; n < N; ++n) {
      const float m = fgm[((int64_t)n * 4 + 3) * HW + p];
      float gm = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float go = gout[((int64_t)n * 3 + c) * HW + p];
        const float fg = fgm[((int64_t)n * 4 + c) * HW + p];
        const float bb = bg[((bg_batched ? (int64_t)n * 3 : 0) + c) * HW + p];
        gfgm[((int64_t)n * 4 + c) * HW + p] = m * go;
        gm += go * (fg - bb);
        if (bg_batched) gbg[((int64_t)n * 3 + c) * HW + p] = (1.f - m) * go;
        else acc[c] += (1.f - m) * go;
      }
      gfgm[((int64_t)n * 4 + 3) * HW + p] = gm;
    }
    if (!bg_batched) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gbg[(int64_t)c * HW + p] = acc[c];
    }
  }
}

}  // namespace nhvr

extern "C" int nhvr_texture_sample_bwd(const float* uvp, const float* atlas, const float* grad_tex, int32_t N, int32_t H, int32_t W,
                                       int32_t S, int32_t Ctex, int32_t use_mask_texture, float* grad_uvp, float* grad_atlas, void* stream) {
  if (!uvp || !atlas || !grad_tex || !grad_uvp || !grad_atlas) return NHVR_ERR_NULL;
  if (N <= 0 || H <= 0 || W <= 0 || S < 2 || Ctex <= 0 || Ctex > 20) return NHVR_ERR_SHAPE;
  if ((((uintptr_t)atlas | (uintptr_t)grad_atlas) & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int G = (Ctex + 3) / 4;
  const int64_t total = (int64_t)N * H * W;
  const int blocks = (int)std::min<int64_t>((total + 127) / 128, (int64_t)148 * 16 * 8);
  const float4* a4 = reinterpret_cast<const float4*>(atlas);
  cudaStream_t st = (cudaStream_t)stream;
#define NHVR_LAUNCH_SBWD(GG) texture_sample_bwd_kernel<GG, 1><<<blocks, 128, 0, st>>>(uvp, a4, grad_tex, N, H, W, S, Ctex, use_mask_texture, grad_uvp, grad_atlas, agg)
  static int agg = -1;            // warp-aggregated atlas reductions for warp-uniform (U, V); NHVR_SAMPLER_BWD_AGG=0 disables
  if (agg < 0) { const char* e = std::getenv("NHVR_SAMPLER_BWD_AGG"); agg = e ? std::atoi(e) : 1; }
  static int minb = -1;           // NHVR_SAMPLER_BWD_MINB = 3 (165 registers) | 4 (128, default) | 6 (80) | 8 (64): measured 1.56 / 1.46 / 1.71 / 1.85 ms at 8 x 512^2
  if (minb < 0) { const char* e = std::getenv("NHVR_SAMPLER_BWD_MINB"); minb = e ? std::atoi(e) : 4; }
  switch (G) {
    case 1:
      if (minb >= 8) texture_sample_bwd_kernel<1, 8><<<blocks, 128, 0, st>>>(uvp, a4, grad_tex, N, H, W, S, Ctex, use_mask_texture, grad_uvp, grad_atlas, agg);
      else if (minb >= 6) texture_sample_bwd_kernel<1, 6><<<blocks, 128, 0, st>>>(uvp, a4, grad_tex, N, H, W, S, Ctex, use_mask_texture, grad_uvp, grad_atlas, agg);
      else if (minb >= 4) texture_sample_bwd_kernel<1, 4><<<blocks, 128, 0, st>>>(uvp, a4, grad_tex, N, H, W, S, Ctex, use_mask_texture, grad_uvp, grad_atlas, agg);
      else texture_sample_bwd_kernel<1, 3><<<blocks, 128, 0, st>>>(uvp, a4, grad_tex, N, H, W, S, Ctex, use_mask_texture, grad_uvp, grad_atlas, agg);
      break;
    case 2: NHVR_LAUNCH_SBWD(2); break;
    case 3: NHVR_LAUNCH_SBWD(3); break;
    case 4: NHVR_LAUNCH_SBWD(4); break;
    default: NHVR_LAUNCH_SBWD(5); break;
  }
#undef NHVR_LAUNCH_SBWD
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

extern "C" int nhvr_composite_bwd(const float* fgm, const float* bg, int32_t bg_batched, const float* grad_out, int32_t N, int32_t H,
                                  int32_t W, float* grad_fgm, float* grad_bg, void* stream) {
  if (!fgm || !bg || !grad_out || !grad_fgm || !grad_bg) return NHVR_ERR_NULL;
  if (N <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int64_t HW = (int64_t)H * W;
  composite_bwd_kernel<<<(int)std::min<int64_t>((HW + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(fgm, bg, bg_batched, grad_out, N,
                                                                                                          HW, grad_fgm, grad_bg);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}


// =================================================================================================
// unfold_texture (README.md:64: the initial texture.jpg from the frames and their DensePose IUV) - the ADJOINT of the
// bilinear lookup: every foreground pixel splats its colour into the four texels around (u, v) of its part with the
// lookup's own weights; the atlas is the weighted mean.
// =================================================================================================
namespace nhvr {

// acc: float [24][S][S][Q] with Q = 4 * ceil((C + 1) / 4): C colour sums, then the weight sum
__global__ void __launch_bounds__(256) texture_unfold_kernel(const float* __restrict__ img, const int32_t* __restrict__ dp_i,
                                                             const float* __restrict__ dp_uv, int N, int H, int W, int S, int C, int Q,
                                                             float* __restrict__ acc) {
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * HW;
  const float sm1 = (float)(S - 1);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int part = dp_i[idx];
    if (part < 1 || part > kParts) continue;
    const int n = (int)(idx / HW);
    const int64_t pix = idx - (int64_t)n * HW;
    const float u = fminf(fmaxf(dp_uv[((int64_t)n * 2) * HW + pix], 0.f), 1.f);
    const float v = fminf(fmaxf(dp_uv[((int64_t)n * 2 + 1) * HW + pix], 0.f), 1.f);
    const float fx = __fmul_rn(u, sm1), fy = __fmul_rn(v, sm1);
    const float x0f = floorf(fx), y0f = floorf(fy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int x1 = min(x0 + 1, S - 1), y1 = min(y0 + 1, S - 1);
    const float wx = fx - x0f, wy = fy - y0f;
    const float wts[4] = {(1.f - wy) * (1.f - wx), (1.f - wy) * wx, wy * (1.f - wx), wy * wx};
    const int64_t tb = (int64_t)(part - 1) * S * S;
    const int64_t at[4] = {tb + (int64_t)y0 * S + x0, tb + (int64_t)y0 * S + x1, tb + (int64_t)y1 * S + x0, tb + (int64_t)y1 * S + x1};
    for (int q = 0; q < Q; q += 4) {
      float val[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = q + e;
        val[e] = c < C ? img[((int64_t)n * C + c) * HW + pix] : (c == C ? 1.f : 0.f);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float* dst = acc + at[k] * Q + q;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(wts[k] * val[0]), "f"(wts[k] * val[1]),
                     "f"(wts[k] * val[2]), "f"(wts[k] * val[3]) : "memory");
      }
    }
  }
}

__global__ void __launch_bounds__(256) texture_unfold_finish_kernel(const float* __restrict__ acc, int S, int C, int Q, float min_weight,
                                                                    float* __restrict__ atlas) {
  const int64_t total = (int64_t)kParts * S * S;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t / ((int64_t)S * S));
    const int64_t yx = t - (int64_t)k * S * S;
    const float w = acc[t * Q + C];
    for (int c = 0; c < C; ++c)
      atlas[((int64_t)k * C + c) * S * S + yx] = w > min_weight ? acc[t * Q + c] / w : 0.f;
  }
}

}  // namespace nhvr

extern "C" int nhvr_texture_unfold(const float* img, const int32_t* dp_i, const float* dp_uv, int32_t N, int32_t H, int32_t W, int32_t S,
                                   int32_t C, float* acc, void* stream) {
  if (!img || !dp_i || !dp_uv || !acc) return NHVR_ERR_NULL;
  if (N <= 0 || H <= 0 || W <= 0 || S < 2 || C <= 0 || C > 19) return NHVR_ERR_SHAPE;
  if (((uintptr_t)acc & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int Q = (C + 1 + 3) / 4 * 4;
  const int64_t total = (int64_t)N * H * W;
  texture_unfold_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(img, dp_i, dp_uv, N, H, W, S, C, Q, acc);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

extern "C" int nhvr_texture_unfold_finish(const float* acc, int32_t S, int32_t C, float min_weight, float* atlas, void* stream) {
  if (!acc || !atlas) return NHVR_ERR_NULL;
  if (S < 2 || C <= 0 || C > 19) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int Q = (C + 1 + 3) / 4 * 4;
  const int64_t total = (int64_t)kParts * S * S;
  texture_unfold_finish_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(acc, S, C, Q, min_weight, atlas);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}
