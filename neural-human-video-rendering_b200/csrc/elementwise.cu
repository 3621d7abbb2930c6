// Memory-bound helpers around the conv kernel: NCHW fp32 <-> P8 bf16 packing, and the fused
// InstanceNorm-apply + activation + residual + halo (ReflectionPad2d / zero padding) writer.
// All three are "one thread per 16-byte destination unit" gather kernels: coalesced 16-B stores,
// the halo is produced by index reflection on the read side, never by a second pass.
#include "common.cuh"
#include "p8.cuh"
#include <algorithm>

namespace nhvr {

extern void note_cuda_error(cudaError_t e);
extern void count_launch();
extern int arch_ok_cached();
extern int operand_f16();

// map a padded destination coordinate to its logical source; returns false for "write zeros"
NHVR_DEVINL bool dst_to_src(const ActGeom& g, int yy, int xx, int& y, int& x) {
  y = yy - g.pad_t;
  x = xx - g.pad_l;
  const bool inside = (y >= 0) & (y < g.H) & (x >= 0) & (x < g.W);
  if (inside) return true;
  if (g.halo == NHVR_HALO_ZERO) return false;
  y = reflect_idx(y, g.H);
  x = reflect_idx(x, g.W);
  return (y >= 0) & (y < g.H) & (x >= 0) & (x < g.W);
}

struct PackParams2 {
  const float* src[4];
  int32_t src_c[4];
  int32_t nsrc;
  uint4* dst;
  ActGeom g;
  int32_t f16;
};

__global__ void __launch_bounds__(256) pack_nchw_kernel(const __grid_constant__ PackParams2 P) {
  const ActGeom& g = P.g;
  const int np = blockIdx.y;              // n * C8 + p
  const int n = np / g.C8, p = np - n * g.C8;
  const int64_t HW = (int64_t)g.H * g.W;
  const int total = g.Hp * g.Wp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int yy = i / g.Wp, xx = i - yy * g.Wp;
    int y, x;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (dst_to_src(g, yy, xx, y, x)) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int c = p * 8 + e;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s < P.nsrc) {
            if (c >= 0 && c < P.src_c[s]) {
              v[e] = __ldg(P.src[s] + ((int64_t)n * P.src_c[s] + c) * HW + (int64_t)y * g.W + x);
              c = -1;
            } else if (c >= 0) {
              c -= P.src_c[s];
            }
          }
        }
      }
    }
    uint4 o;
    o.x = pack2(v[0], v[1], P.f16); o.y = pack2(v[2], v[3], P.f16);
    o.z = pack2(v[4], v[5], P.f16); o.w = pack2(v[6], v[7], P.f16);
    P.dst[(int64_t)np * g.plane_units + plane_unit(g, yy, xx)] = o;
  }
}

__global__ void __launch_bounds__(256) unpack_nchw_kernel(const uint4* __restrict__ src, ActGeom g, float* __restrict__ dst,
                                                          int C, int f16) {
  const int np = blockIdx.y;
  const int n = np / g.C8, p = np - n * g.C8;
  const int64_t HW = (int64_t)g.H * g.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int y = i / g.W, x = i - y * g.W;
    const uint4 u = src[(int64_t)np * g.plane_units + plane_unit(g, y + g.pad_t, x + g.pad_l)];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = p * 8 + e;
      if (c < C) dst[((int64_t)n * C + c) * HW + i] = (e & 1) ? unpack_hi(w[e >> 1], f16) : unpack_lo(w[e >> 1], f16);
    }
  }
}

struct ApplyParams {
  const uint4* raw;      // P8 un-padded [N][C8][H][W]
  const float* stats;    // [N][C8*8][2]
  const uint4* res;      // nullable
  uint4* dst;
  ActGeom rg, sg, dg;    // raw, residual, destination geometry
  float eps, inv_hw;
  int32_t act;
  int32_t f16;
};

__global__ void __launch_bounds__(256) in_apply_kernel(const __grid_constant__ ApplyParams P) {
  const ActGeom& g = P.dg;
  const int np = blockIdx.y;
  const int n = np / g.C8, p = np - n * g.C8;
  float scale[8], shift[8];
  {
    const float* st = P.stats + ((int64_t)n * g.C8 + p) * 16;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float mean = st[2 * e] * P.inv_hw;
      const float var = fmaxf(st[2 * e + 1] * P.inv_hw - mean * mean, 0.f);
      const float rstd = rsqrtf(var + P.eps);
      scale[e] = rstd;
      shift[e] = -mean * rstd;
    }
  }
  const int total = g.Hp * g.Wp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int yy = i / g.Wp, xx = i - yy * g.Wp;
    int y, x;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (dst_to_src(g, yy, xx, y, x)) {
      const uint4 r = P.raw[act_unit(P.rg, n, p, y + P.rg.pad_t, x + P.rg.pad_l)];
      const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = (e & 1) ? unpack_hi(rw[e >> 1], P.f16) : unpack_lo(rw[e >> 1], P.f16);
        float t = fmaf(xv, scale[e], shift[e]);
        if (P.act == NHVR_ACT_RELU) t = fmaxf(t, 0.f);
        else if (P.act == NHVR_ACT_LRELU02) t = t > 0.f ? t : 0.2f * t;
        v[e] = t;
      }
      if (P.res) {
        const uint4 s = P.res[act_unit(P.sg, n, p, y + P.sg.pad_t, x + P.sg.pad_l)];
        const uint32_t sw[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += (e & 1) ? unpack_hi(sw[e >> 1], P.f16) : unpack_lo(sw[e >> 1], P.f16);
      }
      o.x = pack2(v[0], v[1], P.f16); o.y = pack2(v[2], v[3], P.f16);
      o.z = pack2(v[4], v[5], P.f16); o.w = pack2(v[6], v[7], P.f16);
    }
    P.dst[(int64_t)np * g.plane_units + plane_unit(g, yy, xx)] = o;
  }
}

static inline int grid_x_for(int64_t work_items, int planes) {
  // ~4 waves of 148 SMs x 8 resident CTAs, split over the plane dimension
  const int64_t want = std::max<int64_t>(1, (int64_t)148 * 8 * 4 / std::max(1, planes));
  return (int)std::max<int64_t>(1, std::min<int64_t>((work_items + 255) / 256, want));
}

}  // namespace nhvr

using namespace nhvr;

extern "C" size_t nhvr_act_bytes(const nhvr_act_desc* d) {
  if (!d) return 0;
  ActGeom g = make_geom(*d);
  return (size_t)(((int64_t)g.N * g.C8 * g.plane_units + kActSlackUnits) * 16);
}

extern "C" int nhvr_pack_nchw(const float* const* src, const int32_t* src_c, int32_t nsrc, void* dst,
                              const nhvr_act_desc* dst_desc, void* stream) {
  if (!src || !src_c || !dst || !dst_desc) return NHVR_ERR_NULL;
  if (nsrc < 1 || nsrc > 4) return NHVR_ERR_SHAPE;
  if (((uintptr_t)dst & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  PackParams2 P;
  int csum = 0;
  for (int i = 0; i < 4; ++i) {
    P.src[i] = i < nsrc ? src[i] : nullptr;
    P.src_c[i] = i < nsrc ? src_c[i] : 0;
    if (i < nsrc) { if (!src[i] || src_c[i] <= 0) return NHVR_ERR_NULL; csum += src_c[i]; }
  }
  P.nsrc = nsrc;
  P.dst = reinterpret_cast<uint4*>(dst);
  P.g = make_geom(*dst_desc);
  P.f16 = operand_f16();
  if (csum > P.g.C8 * 8) return NHVR_ERR_SHAPE;
  const int planes = P.g.N * P.g.C8;
  dim3 grid(grid_x_for((int64_t)P.g.Hp * P.g.Wp, planes), planes);
  pack_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

extern "C" int nhvr_unpack_nchw(const void* src, const nhvr_act_desc* src_desc, float* dst, int32_t C, void* stream) {
  if (!src || !src_desc || !dst) return NHVR_ERR_NULL;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  ActGeom g = make_geom(*src_desc);
  if (C <= 0 || C > g.C8 * 8) return NHVR_ERR_SHAPE;
  const int planes = g.N * ((C + 7) / 8);
  // planes beyond ceil(C/8) hold nothing we need; but blockIdx.y indexes n*C8+p, so launch all
  dim3 grid(grid_x_for((int64_t)g.H * g.W, g.N * g.C8), g.N * g.C8);
  (void)planes;
  unpack_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), g, dst, C, operand_f16());
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

extern "C" int nhvr_in_apply(const void* raw, const nhvr_act_desc* raw_desc, const float* stats, float eps, int32_t act,
                             const void* residual, const nhvr_act_desc* res_desc, void* dst,
                             const nhvr_act_desc* dst_desc, void* stream) {
  if (!raw || !raw_desc || !stats || !dst || !dst_desc) return NHVR_ERR_NULL;
  if (residual && !res_desc) return NHVR_ERR_NULL;
  if ((((uintptr_t)raw | (uintptr_t)dst | (uintptr_t)residual) & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  ApplyParams P;
  P.raw = reinterpret_cast<const uint4*>(raw);
  P.stats = stats;
  P.res = reinterpret_cast<const uint4*>(residual);
  P.dst = reinterpret_cast<uint4*>(dst);
  P.rg = make_geom(*raw_desc);
  P.dg = make_geom(*dst_desc);
  P.sg = residual ? make_geom(*res_desc) : P.dg;
  if (P.rg.N != P.dg.N || P.rg.C8 != P.dg.C8 || P.rg.H != P.dg.H || P.rg.W != P.dg.W) return NHVR_ERR_SHAPE;
  if (residual && (P.sg.N != P.dg.N || P.sg.C8 != P.dg.C8 || P.sg.H != P.dg.H || P.sg.W != P.dg.W)) return NHVR_ERR_SHAPE;
  P.eps = eps;
  P.inv_hw = 1.0f / ((float)P.rg.H * (float)P.rg.W);
  P.act = act;
  P.f16 = operand_f16();
  const int planes = P.dg.N * P.dg.C8;
  dim3 grid(grid_x_for((int64_t)P.dg.Hp * P.dg.Wp, planes), planes);
  in_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}
