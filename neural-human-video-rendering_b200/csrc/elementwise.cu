// Memory-bound helpers around the conv kernel: NCHW fp32 <-> P8 bf16 packing, and the fused
// InstanceNorm-apply + activation + residual + halo (ReflectionPad2d / zero padding) writer.
// All three are "one thread per 16-byte destination unit" gather kernels: coalesced 16-B stores,
// the halo is produced by index reflection on the read side, never by a second pass.
#include "common.cuh"
#include "p8.cuh"
#include <algorithm>
#include <cstdlib>

namespace nhvr {

extern void note_cuda_error(cudaError_t e);
extern void count_launch();
extern int arch_ok_cached();
extern int operand_f16();
extern int32_t* overflow_flag();

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
  const int np = blockIdx.y;              // n * C8 + physical plane
  const int n = np / g.C8, pp = np - n * g.C8;
  // split precision: physical planes come in groups [hi 2g, hi 2g+1, lo 2g, lo 2g+1]
  const int p = g.hilo ? ((pp >> 2) << 1) | (pp & 1) : pp;      // logical plane
  const bool lo_part = g.hilo && (pp & 2);
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
    if (lo_part) {
      uint4 l;
      l.x = pack2(v[0] - unpack_lo(o.x, P.f16), v[1] - unpack_hi(o.x, P.f16), P.f16);
      l.y = pack2(v[2] - unpack_lo(o.y, P.f16), v[3] - unpack_hi(o.y, P.f16), P.f16);
      l.z = pack2(v[4] - unpack_lo(o.z, P.f16), v[5] - unpack_hi(o.z, P.f16), P.f16);
      l.w = pack2(v[6] - unpack_lo(o.w, P.f16), v[7] - unpack_hi(o.w, P.f16), P.f16);
      o = l;
    }
    P.dst[(int64_t)np * g.plane_units + plane_unit(g, yy, xx)] = o;
  }
}

// Row formulation of the pack: one warp per destination row, 4 columns per lane in flight, the hi AND the lo unit of a logical
// plane written by the same thread (the flat kernel above reads every input twice for a split-precision destination and pays two
// integer divisions per unit): the source row is resolved once per row, the 8 channel base pointers once per block.
constexpr int kPackCols = 4;
__global__ void __launch_bounds__(256) pack_nchw_rows_kernel(const __grid_constant__ PackParams2 P) {
  const ActGeom& g = P.g;
  const int CL = g.hilo ? g.C8 >> 1 : g.C8;                 // logical planes
  const int np = blockIdx.y;
  const int n = np / CL, p = np - n * CL;
  const int ph = g.hilo ? hilo_plane(p) : p;
  const int64_t HW = (int64_t)g.H * g.W;
  // base pointer of each of the 8 channels of this logical plane (nullptr: beyond the concatenated sources -> zeros)
  const float* chp[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    int c = p * 8 + e;
    chp[e] = nullptr;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < P.nsrc && c >= 0) {
        if (c < P.src_c[s]) { chp[e] = P.src[s] + ((int64_t)n * P.src_c[s] + c) * HW; c = -1; }
        else c -= P.src_c[s];
      }
    }
  }
  uint4* dst = P.dst + ((int64_t)n * g.C8 + ph) * g.plane_units;
  const int64_t dst_lo = 2 * g.plane_units;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool reflect = g.halo != NHVR_HALO_ZERO;
  const int Hq = g.Hp >> 1, Wq = g.Wp >> 1;
  for (int yy = blockIdx.x * 8 + warp; yy < g.Hp; yy += gridDim.x * 8) {
    int y = yy - g.pad_t;
    bool row_ok = (y >= 0) & (y < g.H);
    if (!row_ok && reflect) { y = reflect_idx(y, g.H); row_ok = (y >= 0) & (y < g.H); }
    const int64_t roff = (int64_t)y * g.W;
    const int64_t d_even = g.split ? ((int64_t)(((yy & 1) << 1) | 0) * Hq + (yy >> 1)) * Wq : (int64_t)yy * g.Wp;
    const int64_t d_odd = g.split ? ((int64_t)(((yy & 1) << 1) | 1) * Hq + (yy >> 1)) * Wq : 0;
    for (int xx0 = lane; xx0 < g.Wp; xx0 += 32 * kPackCols) {
      float v[kPackCols][8];
      bool live[kPackCols];
#pragma unroll
      for (int j = 0; j < kPackCols; ++j) {
        const int xx = xx0 + 32 * j;
        int x = xx - g.pad_l;
        live[j] = xx < g.Wp;
        bool ok = row_ok && live[j];
        if (ok && ((x < 0) | (x >= g.W))) {
          if (reflect) { x = reflect_idx(x, g.W); ok = (x >= 0) & (x < g.W); } else ok = false;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j][e] = (ok && chp[e]) ? __ldg(chp[e] + roff + x) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kPackCols; ++j) {
        if (!live[j]) continue;
        const int xx = xx0 + 32 * j;
        const int64_t du = g.split ? ((xx & 1) ? d_odd : d_even) + (xx >> 1) : d_even + xx;
        uint4 o;
        o.x = pack2(v[j][0], v[j][1], P.f16); o.y = pack2(v[j][2], v[j][3], P.f16);
        o.z = pack2(v[j][4], v[j][5], P.f16); o.w = pack2(v[j][6], v[j][7], P.f16);
        dst[du] = o;
        if (g.hilo) {
          uint4 l;
          l.x = pack2(v[j][0] - unpack_lo(o.x, P.f16), v[j][1] - unpack_hi(o.x, P.f16), P.f16);
          l.y = pack2(v[j][2] - unpack_lo(o.y, P.f16), v[j][3] - unpack_hi(o.y, P.f16), P.f16);
          l.z = pack2(v[j][4] - unpack_lo(o.z, P.f16), v[j][5] - unpack_hi(o.z, P.f16), P.f16);
          l.w = pack2(v[j][6] - unpack_lo(o.w, P.f16), v[j][7] - unpack_hi(o.w, P.f16), P.f16);
          dst[du + dst_lo] = l;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) unpack_nchw_kernel(const uint4* __restrict__ src, ActGeom g, float* __restrict__ dst,
                                                          int C, int f16) {
  const int np = blockIdx.y;
  const int n = np / g.C8, pp = np - n * g.C8;
  if (g.hilo && (pp & 2)) return;                                 // lo planes are read together with their hi plane
  const int p = g.hilo ? ((pp >> 2) << 1) | (pp & 1) : pp;        // logical plane
  const int64_t HW = (int64_t)g.H * g.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int y = i / g.W, x = i - y * g.W;
    const int64_t at = (int64_t)np * g.plane_units + plane_unit(g, y + g.pad_t, x + g.pad_l);
    const uint4 u = src[at];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t wl[4] = {0, 0, 0, 0};
    if (g.hilo) { const uint4 l = src[at + 2 * g.plane_units]; wl[0] = l.x; wl[1] = l.y; wl[2] = l.z; wl[3] = l.w; }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = p * 8 + e;
      float val = (e & 1) ? unpack_hi(w[e >> 1], f16) : unpack_lo(w[e >> 1], f16);
      if (g.hilo) val += (e & 1) ? unpack_hi(wl[e >> 1], f16) : unpack_lo(wl[e >> 1], f16);
      if (c < C) dst[((int64_t)n * C + c) * HW + i] = val;
    }
  }
}

NHVR_DEVINL void unpack8(const uint4& u, float (&v)[8], int f16) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = (e & 1) ? unpack_hi(w[e >> 1], f16) : unpack_lo(w[e >> 1], f16);
}
NHVR_DEVINL uint4 pack8(const float (&v)[8], int f16) {
  uint4 o;
  o.x = pack2(v[0], v[1], f16); o.y = pack2(v[2], v[3], f16);
  o.z = pack2(v[4], v[5], f16); o.w = pack2(v[6], v[7], f16);
  return o;
}

// (sum, sum of squares) in fp64 -> fp32 (rstd, -mean * rstd) of 8 channels.  mean and variance are formed in fp64: with
// |mean| >> std the difference E[x^2] - mean^2 cancels leading digits that fp32 sums do not have.
// Statistics record of one channel: {sum (x - s), sum (x - s)^2, s, unused} with s the (optional) centring shift.
NHVR_DEVINL void norm_params8(const double* __restrict__ st, float inv_hw, float eps, float (&scale)[8], float (&shift)[8]) {
  const double2* s2 = reinterpret_cast<const double2*>(st);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const double2 q = s2[2 * e], r = s2[2 * e + 1];
    const double m0 = q.x * (double)inv_hw;                 // mean of x - s
    const double var = fmax(q.y * (double)inv_hw - m0 * m0, 0.0);
    const float rstd = rsqrtf((float)var + eps);
    scale[e] = rstd;
    shift[e] = -(float)(m0 + r.x) * rstd;
  }
}

struct ApplyParams {
  const uint4* raw;      // P8 un-padded [N][C8][H][W]
  const double* stats;   // [N][C8*8][2] fp64 sums
  const uint4* res;      // nullable
  uint4* dst;
  ActGeom rg, sg, dg;    // raw, residual, destination geometry
  float eps, inv_hw;
  int32_t act;
  int32_t f16;
  int32_t* ovf;          // nullable: OR-ed with 1 when a non-finite raw value is read (nhvr_set_overflow_flag)
};

// Latency-bound gather: the statistics loads, and two destination units per thread and pass, are in flight before
// anything is consumed (software-pipelined: the next pass is loaded while the current one is normalised and stored).
struct ApplyItem {
  uint4 r, s;
  int64_t du;
  bool live, ok;
};

template <bool HAS_RES>
NHVR_DEVINL void apply_fetch(const ApplyParams& P, const ActGeom& g, const uint4* raw, const uint4* res, int i, int total, ApplyItem& it) {
  it.live = i < total;
  const int yy = i / g.Wp, xx = i - yy * g.Wp;
  int y, x;
  it.ok = it.live && dst_to_src(g, yy, xx, y, x);
  it.du = plane_unit(g, yy, xx);
  it.r = make_uint4(0, 0, 0, 0);
  it.s = make_uint4(0, 0, 0, 0);
  if (it.ok) {
    it.r = __ldg(raw + plane_unit(P.rg, y + P.rg.pad_t, x + P.rg.pad_l));
    if (HAS_RES) it.s = __ldg(res + plane_unit(P.sg, y + P.sg.pad_t, x + P.sg.pad_l));
  }
}

template <bool HAS_RES>
NHVR_DEVINL void apply_emit(const ApplyParams& P, const float (&scale)[8], const float (&shift)[8], uint4* dst, const ApplyItem& it) {
  if (!it.live) return;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (it.ok) {
    const uint32_t rw[4] = {it.r.x, it.r.y, it.r.z, it.r.w};
    const uint32_t sw[4] = {it.s.x, it.s.y, it.s.z, it.s.w};
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xv = (e & 1) ? unpack_hi(rw[e >> 1], P.f16) : unpack_lo(rw[e >> 1], P.f16);
      float t = fmaf(xv, scale[e], shift[e]);
      if (P.act == NHVR_ACT_RELU) t = fmaxf(t, 0.f);
      else if (P.act == NHVR_ACT_LRELU02) t = t > 0.f ? t : 0.2f * t;
      if (HAS_RES) t += (e & 1) ? unpack_hi(sw[e >> 1], P.f16) : unpack_lo(sw[e >> 1], P.f16);
      v[e] = t;
    }
    o.x = pack2(v[0], v[1], P.f16); o.y = pack2(v[2], v[3], P.f16);
    o.z = pack2(v[4], v[5], P.f16); o.w = pack2(v[6], v[7], P.f16);
  }
  dst[it.du] = o;
}

template <bool HAS_RES>
__global__ void __launch_bounds__(256, 4) in_apply_kernel(const __grid_constant__ ApplyParams P) {
  const ActGeom& g = P.dg;
  const int np = blockIdx.y;
  const int n = np / g.C8, p = np - n * g.C8;
  const double* st = P.stats + ((int64_t)n * g.C8 + p) * 32;
  const int total = g.Hp * g.Wp;
  const int stride = gridDim.x * blockDim.x;
  const uint4* raw = P.raw + ((int64_t)n * P.rg.C8 + p) * P.rg.plane_units;
  const uint4* res = HAS_RES ? P.res + ((int64_t)n * P.sg.C8 + p) * P.sg.plane_units : nullptr;
  uint4* dst = P.dst + (int64_t)np * g.plane_units;
  int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  ApplyItem a0, a1;
  apply_fetch<HAS_RES>(P, g, raw, res, i0, total, a0);
  apply_fetch<HAS_RES>(P, g, raw, res, i0 + stride, total, a1);
  float scale[8], shift[8];
  norm_params8(st, P.inv_hw, P.eps, scale, shift);
  while (i0 < total) {
    i0 += 2 * stride;
    ApplyItem b0, b1;
    apply_fetch<HAS_RES>(P, g, raw, res, i0, total, b0);
    apply_fetch<HAS_RES>(P, g, raw, res, i0 + stride, total, b1);
    apply_emit<HAS_RES>(P, scale, shift, dst, a0);
    apply_emit<HAS_RES>(P, scale, shift, dst, a1);
    a0 = b0; a1 = b1;
  }
}

// Row formulation: one warp per destination
// row, lanes stride over the columns, four columns per lane in flight.  No division, the source row is resolved once per
// row (mirror / zero halo rows), per unit only the column mirror remains.
// HILO: split-precision activations (nhvr_act_desc.hilo): raw, residual and destination all carry hi + lo planes; the
// value is rebuilt in fp32, normalised, and split again (the residual stream of a ResnetBlock keeps ~22 bits).
// HILO: 0 = plain 16-bit activations; 1, 2, 3 = split precision with 3, 5, 2 columns per lane in flight
template <bool HAS_RES, int HILO>
__global__ void __launch_bounds__(256, HILO == 0 ? 4 : HILO == 3 ? 3 : 2) in_apply_rows_kernel(const __grid_constant__ ApplyParams P) {
  const ActGeom& g = P.dg;
  const int np = blockIdx.y;                                  // n * (logical planes) + logical plane
  const int CL = HILO ? g.C8 >> 1 : g.C8;
  const int n = np / CL, p = np - n * CL;
  const int ph = HILO ? hilo_plane(p) : p;                    // physical (hi) plane; the lo plane is 2 further
  const double* st = P.stats + ((int64_t)n * CL + p) * 32;
  const uint4* raw = P.raw + ((int64_t)n * P.rg.C8 + ph) * P.rg.plane_units;
  const uint4* res = HAS_RES ? P.res + ((int64_t)n * P.sg.C8 + ph) * P.sg.plane_units : nullptr;
  uint4* dst = P.dst + ((int64_t)n * g.C8 + ph) * g.plane_units;
  const int64_t raw_lo = 2 * P.rg.plane_units, res_lo = 2 * P.sg.plane_units, dst_lo = 2 * g.plane_units;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // mean / variance are formed in fp64 (see norm_params8) by 8 lanes, one channel each, and broadcast through shared
  // memory: the fp64 pipe is narrow, 256 threads repeating the same 8 channels cost ~10 % of this kernel
  __shared__ float s_norm[24];
  if (threadIdx.x < 8) {
    const double2 q = reinterpret_cast<const double2*>(st)[2 * threadIdx.x], r = reinterpret_cast<const double2*>(st)[2 * threadIdx.x + 1];
    const double m0 = q.x * (double)P.inv_hw;
    const double var = fmax(q.y * (double)P.inv_hw - m0 * m0, 0.0);
    const float rstd = rsqrtf((float)var + P.eps);
    s_norm[threadIdx.x] = rstd;
    s_norm[8 + threadIdx.x] = -(float)(m0 + r.x) * rstd;
    s_norm[16 + threadIdx.x] = (q.y + r.x * r.x < 4.29e9) ? 0.f : 1.f;     // range guard: see below
  }
  __syncthreads();
  float scale[8], shift[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { scale[e] = s_norm[e]; shift[e] = s_norm[8 + e]; }
  const bool reflect = g.halo != NHVR_HALO_ZERO;
  const uint64_t once = l2_policy_evict_first();      // the raw tensor is dead after this pass
  constexpr int COLS = HILO == 0 ? 5 : HILO == 1 ? 3 : HILO == 2 ? 5 : 2;   // 160 columns per pass: the 130-wide (128^2) rows take one pass, 258 two, 518 four
  const int Hq = g.Hp >> 1, Wq = g.Wp >> 1;
  // Range guard (nhvr_set_overflow_flag): a 16-bit overflow of the conv output needs |x| > 65504, hence a plane whose
  // sum of squares (taken from the un-rounded fp32 accumulators) reaches 65504^2.  Only such planes are scanned for
  // inf / NaN units, in a separate loop so that the common path pays one comparison per block.
  if (P.ovf) {
    bool suspect = false;
#pragma unroll
    for (int e = 0; e < 8; ++e) suspect |= s_norm[16 + e] != 0.f;
    if (suspect) {
      bool bad = false;
      for (int y = blockIdx.x * 8 + warp; y < g.H; y += gridDim.x * 8)
        for (int x = lane; x < g.W; x += 32) bad |= unit_nonfinite(raw[(int64_t)(y + P.rg.pad_t) * P.rg.Wp + P.rg.pad_l + x], P.f16);
      if (bad) atomicOr(P.ovf, 1);
    }
  }
  for (int yy = blockIdx.x * 8 + warp; yy < g.Hp; yy += gridDim.x * 8) {
    int y = yy - g.pad_t;
    bool row_ok = (y >= 0) & (y < g.H);
    if (!row_ok && reflect) { y = reflect_idx(y, g.H); row_ok = (y >= 0) & (y < g.H); }
    const uint4* rrow = raw + (int64_t)(y + P.rg.pad_t) * P.rg.Wp + P.rg.pad_l;
    const uint4* srow = HAS_RES ? res + (int64_t)(y + P.sg.pad_t) * P.sg.Wp + P.sg.pad_l : nullptr;
    // destination row: plain, or the two column-parity planes of the 4-way split format
    const int64_t d_even = g.split ? ((int64_t)(((yy & 1) << 1) | 0) * Hq + (yy >> 1)) * Wq : (int64_t)yy * g.Wp;
    const int64_t d_odd = g.split ? ((int64_t)(((yy & 1) << 1) | 1) * Hq + (yy >> 1)) * Wq : 0;
    for (int xx0 = lane; xx0 < g.Wp; xx0 += 32 * COLS) {
      ApplyItem it[COLS];
      uint4 r2[HILO ? COLS : 1], s2[HILO ? COLS : 1];
#pragma unroll
      for (int j = 0; j < COLS; ++j) {
        const int xx = xx0 + 32 * j;
        int x = xx - g.pad_l;
        bool ok = row_ok && xx < g.Wp;
        if (ok && ((x < 0) | (x >= g.W))) {
          if (reflect) { x = reflect_idx(x, g.W); ok = (x >= 0) & (x < g.W); } else ok = false;
        }
        it[j].live = xx < g.Wp;
        it[j].ok = ok;
        it[j].du = g.split ? ((xx & 1) ? d_odd : d_even) + (xx >> 1) : d_even + xx;
        it[j].r = make_uint4(0, 0, 0, 0);
        it[j].s = make_uint4(0, 0, 0, 0);
        if (HILO) { r2[j] = make_uint4(0, 0, 0, 0); s2[j] = make_uint4(0, 0, 0, 0); }
        if (ok) {
          it[j].r = ld_hint(rrow + x, once);
          if (HILO) r2[j] = ld_hint(rrow + raw_lo + x, once);
          if (HAS_RES) { it[j].s = __ldg(srow + x); if (HILO) s2[j] = __ldg(srow + res_lo + x); }
        }
      }
#pragma unroll
      for (int j = 0; j < COLS; ++j) {
        if (!HILO) { apply_emit<HAS_RES>(P, scale, shift, dst, it[j]); continue; }
        if (!it[j].live) continue;
        uint4 oh = make_uint4(0, 0, 0, 0), ol = oh;
        if (it[j].ok) {
          float xv[8], xl[8], rv[8], rl[8], v[8];
          unpack8(it[j].r, xv, P.f16); unpack8(r2[HILO ? j : 0], xl, P.f16);
          if (HAS_RES) { unpack8(it[j].s, rv, P.f16); unpack8(s2[HILO ? j : 0], rl, P.f16); }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float t = fmaf(xv[e] + xl[e], scale[e], shift[e]);
            if (P.act == NHVR_ACT_RELU) t = fmaxf(t, 0.f);
            else if (P.act == NHVR_ACT_LRELU02) t = t > 0.f ? t : 0.2f * t;
            if (HAS_RES) t += rv[e] + rl[e];
            v[e] = t;
          }
          split_hilo(v[0], v[1], P.f16, oh.x, ol.x); split_hilo(v[2], v[3], P.f16, oh.y, ol.y);
          split_hilo(v[4], v[5], P.f16, oh.z, ol.z); split_hilo(v[6], v[7], P.f16, oh.w, ol.w);
        }
        dst[it[j].du] = oh;
        dst[it[j].du + dst_lo] = ol;
      }
    }
  }
}

// Centring shift of a stem's InstanceNorm statistics: s[n][co] = sum_ci wsum[co][ci] * m[n][ci] with wsum the filter summed
// over its taps and m the (sampled, 256 pixels) per-channel mean of the input image - the value the conv output takes wherever the input is flat.  Stick-
// figure pose maps are ~98 % background, so the stem output is c + small with mean^2 / var up to ~250; centred sums
// keep E[(x-s)^2] - E[x-s]^2 free of that cancellation.  Any s is mathematically valid (it only re-centres the sums).
struct StemShiftParams {
  const float* src[4];
  int32_t src_c[4];
  int32_t nsrc, Cin, Cout, C8out8;
  int64_t HW;
  const float* wsum;   // [Cout][Cin]: the filter summed over its taps
  double* stats;       // [N][C8out8][4]: slot 2 receives s
};
// one block per image: 256 samples per input channel (all loads in flight before the reductions), then a Cout x Cin GEMV
__global__ void __launch_bounds__(256) stem_stat_shift_kernel(const __grid_constant__ StemShiftParams P) {
  __shared__ float m[32];
  __shared__ float red[32][8];
  const int n = blockIdx.x;
  const int64_t stride = P.HW >= 256 ? P.HW / 256 : 1;
  const int64_t at = min((int64_t)threadIdx.x * stride + (stride >> 1), P.HW - 1);
  float v[32];
  int c = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (s < P.nsrc) {
      for (int k = 0; k < P.src_c[s] && c < 32; ++k, ++c) v[c] = __ldg(P.src[s] + ((int64_t)n * P.src_c[s] + k) * P.HW + at);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < P.Cin && k < 32; ++k) {
    float t = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[k][warp] = t;
  }
  __syncthreads();
  if (threadIdx.x < P.Cin && threadIdx.x < 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    m[threadIdx.x] = t * (1.f / 256.f);
  }
  __syncthreads();
  for (int co = threadIdx.x; co < P.Cout; co += 256) {
    float sft = 0.f;
    for (int ci = 0; ci < P.Cin; ++ci) sft += __ldg(P.wsum + (int64_t)co * P.Cin + ci) * m[ci];
    P.stats[((int64_t)n * P.C8out8 + co) * 4 + 2] = (double)sft;
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
  if (P.g.hilo && (P.g.C8 & 3)) return NHVR_ERR_SHAPE;
  if (csum > (P.g.hilo ? P.g.C8 / 2 : P.g.C8) * 8) return NHVR_ERR_SHAPE;
  const int planes = P.g.N * P.g.C8;
  static const char* rows_env = std::getenv("NHVR_PACK_ROWS");
  if (!(rows_env && std::atoi(rows_env) == 0)) {
    const int lplanes = P.g.N * (P.g.hilo ? P.g.C8 / 2 : P.g.C8);
    // ~2 rows per warp while that still leaves >= ~4 waves of blocks
    int gy = std::max(1, (P.g.Hp + 7) / 8);
    const int cap = std::max(1, 148 * 8 * 4 / std::max(1, lplanes));
    if (gy > cap) gy = std::max(cap, (P.g.Hp + 31) / 32);
    pack_nchw_rows_kernel<<<dim3(gy, lplanes), 256, 0, (cudaStream_t)stream>>>(P);
  } else {
    dim3 grid(grid_x_for((int64_t)P.g.Hp * P.g.Wp, planes), planes);
    pack_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

extern "C" int nhvr_stem_stat_shift(const float* wsum, int32_t Cout, int32_t Cin, const float* const* src, const int32_t* src_c,
                                    int32_t nsrc, int32_t N, int32_t H, int32_t W, double* stats, void* stream) {
  if (!wsum || !src || !src_c || !stats) return NHVR_ERR_NULL;
  if (nsrc < 1 || nsrc > 4 || Cin < 1 || Cin > 32 || Cout < 1 || N < 1 || H < 1 || W < 1 || (int64_t)H * W < 256) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  StemShiftParams P;
  int csum = 0;
  for (int i = 0; i < 4; ++i) {
    P.src[i] = i < nsrc ? src[i] : nullptr;
    P.src_c[i] = i < nsrc ? src_c[i] : 0;
    if (i < nsrc) { if (!src[i] || src_c[i] <= 0) return NHVR_ERR_NULL; csum += src_c[i]; }
  }
  if (csum != Cin) return NHVR_ERR_SHAPE;
  P.nsrc = nsrc; P.Cin = Cin; P.Cout = Cout;
  P.C8out8 = ((Cout + 7) / 8 + 1) / 2 * 2 * 8;      // channel count of the statistics record: Cout8 (even) * 8
  P.HW = (int64_t)H * W;
  P.wsum = wsum; P.stats = stats;
  stem_stat_shift_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(P);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

extern "C" int nhvr_unpack_nchw(const void* src, const nhvr_act_desc* src_desc, float* dst, int32_t C, void* stream) {
  if (!src || !src_desc || !dst) return NHVR_ERR_NULL;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  ActGeom g = make_geom(*src_desc);
  if (g.hilo && (g.C8 & 3)) return NHVR_ERR_SHAPE;
  if (C <= 0 || C > (g.hilo ? g.C8 / 2 : g.C8) * 8) return NHVR_ERR_SHAPE;
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

extern "C" int nhvr_in_apply(const void* raw, const nhvr_act_desc* raw_desc, const double* stats, float eps, int32_t act,
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
  const int hilo = P.dg.hilo;
  if (P.rg.hilo != hilo || (residual && P.sg.hilo != hilo) || (hilo && (P.dg.C8 & 3))) return NHVR_ERR_SHAPE;
  P.ovf = overflow_flag();
  P.eps = eps;
  P.inv_hw = 1.0f / ((float)P.rg.H * (float)P.rg.W);
  P.act = act;
  P.f16 = operand_f16();
  const int planes = P.dg.N * (hilo ? P.dg.C8 / 2 : P.dg.C8);     // logical planes
  const int64_t units = (int64_t)P.dg.Hp * P.dg.Wp;
  int gx = (int)((units + 256 * 4 - 1) / (256 * 4));                      // two passes of two units per thread ...
  const int64_t cap = std::max<int64_t>(1, (int64_t)148 * 8 * 4 / std::max(1, planes));
  if (gx > cap) gx = (int)cap;                                            // ... unless that is more than ~4 waves of CTAs
  if (const char* e = std::getenv("NHVR_APPLY_GX")) gx = std::max(1, std::atoi(e));
  dim3 grid(gx, planes);
  static const char* rows_env = std::getenv("NHVR_APPLY_ROWS");
  const bool rows = !P.rg.split && !(residual && P.sg.split) && !(rows_env && std::atoi(rows_env) == 0);     // sources are never split today
  if (rows) {
    // one destination row per warp (8 per block): measured best of 3 / 5 / 9 / 13 / 17 / 33 / 65 row blocks per plane
    int gy = std::max(1, (P.dg.Hp + 7) / 8);
    { static const char* e = std::getenv("NHVR_APPLY_GY"); if (e && std::atoi(e) > 0) gy = std::min(std::atoi(e), gy); }
    dim3 grid_r(gy, planes);
    if (hilo) {
      int variant = 3;      // measured (tools/apply_bench.py): 2 columns in flight at 3 blocks / SM: 4.1-5.4 TB/s vs 3.4-4.9 for 3 columns at 2 blocks
      if (const char* e = std::getenv("NHVR_APPLY_HILO_VARIANT")) variant = std::atoi(e);      // experiments
      cudaStream_t st_ = (cudaStream_t)stream;
      if (variant == 2) { if (P.res) in_apply_rows_kernel<true, 2><<<grid_r, 256, 0, st_>>>(P); else in_apply_rows_kernel<false, 2><<<grid_r, 256, 0, st_>>>(P); }
      else if (variant != 1) { if (P.res) in_apply_rows_kernel<true, 3><<<grid_r, 256, 0, st_>>>(P); else in_apply_rows_kernel<false, 3><<<grid_r, 256, 0, st_>>>(P); }
      else { if (P.res) in_apply_rows_kernel<true, 1><<<grid_r, 256, 0, st_>>>(P); else in_apply_rows_kernel<false, 1><<<grid_r, 256, 0, st_>>>(P); }
    } else if (P.res) in_apply_rows_kernel<true, 0><<<grid_r, 256, 0, (cudaStream_t)stream>>>(P);
    else in_apply_rows_kernel<false, 0><<<grid_r, 256, 0, (cudaStream_t)stream>>>(P);
  } else if (hilo) {
    return NHVR_ERR_UNSUPPORTED;
  } else if (P.res) in_apply_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  else in_apply_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

// =================================================================================================
// backward of  x_next = pad( act( InstanceNorm(raw) ) [+ residual] )
// =================================================================================================
// dX is the gradient w.r.t. the PADDED tensor x_next (un-padded P8 buffer whose logical size is the padded
// extent, as written by the dgrad conv), `fold` describes that padding: the gradient of an interior pixel is
// the sum of dX over every padded position that mirrors it (ReflectionPad2d) or just its own (zero padding).
// Pass 1 reduces  s1 = sum g_z,  s2 = sum g_z * z  per (n, c);  pass 2 writes
//   g_raw = rstd * (g_z - s1/HW - z * s2/HW)      with  z = (raw - mean) * rstd,  g_z = dY * act'(z)
// into the gradient format shared by this layer's dgrad and wgrad (zero halo), and optionally the folded
// total dY (needed later as the skip-connection gradient of a ResnetBlock).
namespace nhvr {

struct FoldGeom {
  int32_t H, W;            // interior size
  int32_t pad_t, pad_l;    // interior origin inside dX
  int32_t Hp, Wp;          // dX logical size
  int32_t reflect;         // 1: mirror halo contributes, 0: zero padding (halo gradient dropped)
};

struct InBwdParams {
  const uint4* dx;         // P8 [N][C8][Hp][Wp] un-padded buffer
  const uint4* skip;       // nullable, P8 [N][C8][H][W]
  const uint4* raw;        // P8 [N][C8][H][W]
  const double* stats;     // [N][C8*8][2] forward sums (fp64)
  float* sums;             // [N][C8*8][2] backward sums (pass 1 out, pass 2 in)
  uint4* g;                // pass 2 out: gradient format
  uint4* dy_out;           // pass 2 out (nullable): folded dY, P8 [N][C8][H][W]
  FoldGeom f;
  ActGeom gg;              // gradient format geometry
  int32_t N, C8;
  float eps, inv_hw;
  int32_t act, f16;
  int32_t* ovf;            // nullable overflow flag (nhvr_set_overflow_flag)
  uint32_t* sync;          // single-launch variant: per-plane arrival counters [N*C8], behind `sums`, zeroed with them
};

// folded gradient of interior pixel (y, x) of plane np
NHVR_DEVINL void fold_load(const InBwdParams& P, int64_t np, int y, int x, float (&dy)[8]) {
  const FoldGeom& f = P.f;
  int ys[2], xs[2], ny = 1, nx = 1;
  ys[0] = y + f.pad_t; xs[0] = x + f.pad_l;
  if (f.reflect) {
    // padded rows that mirror row y: pad_t - y (top halo, 1 <= y <= pad_t) and 2(H-1) - y + pad_t (bottom halo)
    if (y >= 1 && y <= f.pad_t) ys[ny++] = f.pad_t - y;
    else { const int yb = 2 * (f.H - 1) - y + f.pad_t; if (y <= f.H - 2 && yb < f.Hp) ys[ny++] = yb; }
    if (x >= 1 && x <= f.pad_l) xs[nx++] = f.pad_l - x;
    else { const int xb = 2 * (f.W - 1) - x + f.pad_l; if (x <= f.W - 2 && xb < f.Wp) xs[nx++] = xb; }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) dy[e] = 0.f;
  const uint4* base = P.dx + np * (int64_t)f.Hp * f.Wp;
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      float t[8];
      unpack8(base[(int64_t)ys[a] * f.Wp + xs[b]], t, P.f16);
#pragma unroll
      for (int e = 0; e < 8; ++e) dy[e] += t[e];
    }
  if (P.skip) {
    float t[8];
    unpack8(P.skip[np * (int64_t)f.H * f.W + (int64_t)y * f.W + x], t, P.f16);
#pragma unroll
    for (int e = 0; e < 8; ++e) dy[e] += t[e];
  }
}

NHVR_DEVINL void fwd_norm_params(const InBwdParams& P, int64_t np, float (&scale)[8], float (&shift)[8]) {
  norm_params8(P.stats + np * 32, P.inv_hw, P.eps, scale, shift);
}

NHVR_DEVINL float act_grad(float z, int act) {
  if (act == NHVR_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == NHVR_ACT_LRELU02) return z > 0.f ? 1.f : 0.2f;
  return 1.f;
}

__global__ void __launch_bounds__(256) in_bwd_reduce_kernel(const __grid_constant__ InBwdParams P) {
  const int64_t np = blockIdx.y;
  float scale[8], shift[8];
  fwd_norm_params(P, np, scale, shift);
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
  const int total = P.f.H * P.f.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / P.f.W, x = i - y * P.f.W;
    float dy[8], r[8];
    fold_load(P, np, y, x, dy);
    unpack8(P.raw[np * (int64_t)total + i], r, P.f16);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float z = fmaf(r[e], scale[e], shift[e]);
      const float gz = dy[e] * act_grad(z, P.act);
      s1[e] += gz;
      s2[e] += gz * z;
    }
  }
  __shared__ float sh[16][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float a = s1[e], b = s2[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { sh[2 * e][warp] = a; sh[2 * e + 1][warp] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
    atomicAdd(P.sums + np * 16 + threadIdx.x, t);
  }
}

__global__ void __launch_bounds__(256) in_bwd_apply_kernel(const __grid_constant__ InBwdParams P) {
  const int64_t np = blockIdx.y;
  const int n = (int)(np / P.C8), p = (int)(np - (int64_t)n * P.C8);
  float scale[8], shift[8], m1[8], m2[8];
  fwd_norm_params(P, np, scale, shift);
#pragma unroll
  for (int e = 0; e < 8; ++e) { m1[e] = P.sums[np * 16 + 2 * e] * P.inv_hw; m2[e] = P.sums[np * 16 + 2 * e + 1] * P.inv_hw; }
  const ActGeom& g = P.gg;
  const int total = g.Hp * g.Wp;       // every position of the gradient format is written (halo = 0)
  const int HW = P.f.H * P.f.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int yy = i / g.Wp, xx = i - yy * g.Wp;
    const int y = yy - g.pad_t, x = xx - g.pad_l;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < P.f.H && x >= 0 && x < P.f.W) {
      float dy[8], r[8], v[8];
      fold_load(P, np, y, x, dy);
      unpack8(P.raw[np * (int64_t)HW + (int64_t)y * P.f.W + x], r, P.f16);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float z = fmaf(r[e], scale[e], shift[e]);
        const float gz = dy[e] * act_grad(z, P.act);
        v[e] = scale[e] * (gz - m1[e] - z * m2[e]);
      }
      o = pack8(v, P.f16);
      if (P.dy_out) P.dy_out[np * (int64_t)HW + (int64_t)y * P.f.W + x] = pack8(dy, P.f16);
    }
    P.g[((int64_t)n * g.C8 + p) * g.plane_units + plane_unit(g, yy, xx)] = o;
  }
}

// Row formulation of both passes (PASS 0: the two reductions, PASS 1: apply): one warp per interior row y, the padded
// rows that fold onto it are resolved once per row, lanes stride over x with kBwdCols columns in flight (primary dX tap,
// raw and skip loads are issued before anything is consumed; the mirrored extra taps only exist on the border).

// One pass over the rows this block owns.  PASS 0 accumulates (s1, s2), PASS 1 applies with (m1, m2).
template <int PASS, int kBwdCols>
NHVR_DEVINL void in_bwd_rows_pass(const InBwdParams& P, int64_t np, int n, int p, const float (&scale)[8], const float (&shift)[8],
                                  const float (&m1)[8], const float (&m2)[8], float (&s1)[8], float (&s2)[8], bool& bad, int row_blocks) {
  const FoldGeom& f = P.f;
  const ActGeom& g = P.gg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4* dxp = P.dx + np * (int64_t)f.Hp * f.Wp;
  const uint4* rawp = P.raw + np * (int64_t)f.H * f.W;
  const uint4* skp = P.skip ? P.skip + np * (int64_t)f.H * f.W : nullptr;
  uint4* gp = PASS == 1 ? P.g + ((int64_t)n * g.C8 + p) * g.plane_units : nullptr;
  uint4* dyo = (PASS == 1 && P.dy_out) ? P.dy_out + np * (int64_t)f.H * f.W : nullptr;
  const int Hq = g.Hp >> 1, Wq = g.Wp >> 1;
  // PASS 0 walks the interior rows, PASS 1 every row of the gradient format (halo rows are written as zeros)
  const int nrows = PASS == 1 ? g.Hp : f.H;
  for (int rr = blockIdx.x * 8 + warp; rr < nrows; rr += row_blocks * 8) {
    const int y = PASS == 1 ? rr - g.pad_t : rr;
    const bool row_in = (y >= 0) & (y < f.H);
    // padded rows of dX that fold onto interior row y
    int y0 = y + f.pad_t, y1 = -1;
    if (row_in && f.reflect) {
      if (y >= 1 && y <= f.pad_t) y1 = f.pad_t - y;
      else { const int yb = 2 * (f.H - 1) - y + f.pad_t; if (y <= f.H - 2 && yb < f.Hp) y1 = yb; }
    }
    const uint4* d0 = dxp + (int64_t)y0 * f.Wp;
    const uint4* d1 = dxp + (int64_t)(y1 < 0 ? 0 : y1) * f.Wp;
    const uint4* rrow = rawp + (int64_t)y * f.W;
    const uint4* srow = skp ? skp + (int64_t)y * f.W : nullptr;
    int64_t o_even = 0, o_odd = 0;
    if (PASS == 1) {
      o_even = g.split ? ((int64_t)(((rr & 1) << 1) | 0) * Hq + (rr >> 1)) * Wq : (int64_t)rr * g.Wp;
      o_odd = g.split ? ((int64_t)(((rr & 1) << 1) | 1) * Hq + (rr >> 1)) * Wq : 0;
    }
    const int ncols = PASS == 1 ? g.Wp : f.W;
    for (int c0 = lane; c0 < ncols; c0 += 32 * kBwdCols) {
      uint4 a[kBwdCols], r[kBwdCols], sk[kBwdCols];
      int xs[kBwdCols], x1s[kBwdCols];
      bool ok[kBwdCols], live[kBwdCols];
#pragma unroll
      for (int j = 0; j < kBwdCols; ++j) {
        const int cc = c0 + 32 * j;
        const int x = PASS == 1 ? cc - g.pad_l : cc;
        live[j] = cc < ncols;
        ok[j] = live[j] && row_in && x >= 0 && x < f.W;
        xs[j] = x;
        x1s[j] = -1;
        a[j] = make_uint4(0, 0, 0, 0); r[j] = a[j]; sk[j] = a[j];
        if (ok[j]) {
          if (f.reflect) {
            if (x >= 1 && x <= f.pad_l) x1s[j] = f.pad_l - x;
            else { const int xb = 2 * (f.W - 1) - x + f.pad_l; if (x <= f.W - 2 && xb < f.Wp) x1s[j] = xb; }
          }
          a[j] = __ldg(d0 + x + f.pad_l);
          r[j] = __ldg(rrow + x);
          if (srow) sk[j] = __ldg(srow + x);
        }
      }
#pragma unroll
      for (int j = 0; j < kBwdCols; ++j) {
        if (!live[j]) continue;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (ok[j]) {
          float dy[8], t[8], rv[8];
          if (PASS == 1) bad |= unit_nonfinite(a[j], P.f16);
          unpack8(a[j], dy, P.f16);
          if (x1s[j] >= 0) { unpack8(__ldg(d0 + x1s[j]), t, P.f16);
#pragma unroll
            for (int e = 0; e < 8; ++e) dy[e] += t[e]; }
          if (y1 >= 0) {
            unpack8(__ldg(d1 + xs[j] + f.pad_l), t, P.f16);
#pragma unroll
            for (int e = 0; e < 8; ++e) dy[e] += t[e];
            if (x1s[j] >= 0) { unpack8(__ldg(d1 + x1s[j]), t, P.f16);
#pragma unroll
              for (int e = 0; e < 8; ++e) dy[e] += t[e]; }
          }
          if (srow) { unpack8(sk[j], t, P.f16);
#pragma unroll
            for (int e = 0; e < 8; ++e) dy[e] += t[e]; }
          unpack8(r[j], rv, P.f16);
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float z = fmaf(rv[e], scale[e], shift[e]);
            const float gz = dy[e] * act_grad(z, P.act);
            if (PASS == 0) { s1[e] += gz; s2[e] += gz * z; }
            else v[e] = scale[e] * (gz - m1[e] - z * m2[e]);
          }
          if (PASS == 1) {
            o = pack8(v, P.f16);
            if (dyo) dyo[(int64_t)y * f.W + xs[j]] = pack8(dy, P.f16);
          }
        }
        if (PASS == 1) {
          const int cc = c0 + 32 * j;
          gp[g.split ? ((cc & 1) ? o_odd : o_even) + (cc >> 1) : o_even + cc] = o;
        }
      }
    }
  }
}

// PASS 0 / 1: the two launches of the two-kernel path.  PASS 2: BOTH passes in one launch - the blocks of a plane reduce
// their rows, meet at a per-plane arrival counter (blocks are dispatched in index order and a plane's blocks fit the GPU at
// once, checked by the host), and apply to the same rows, which are then served by L2 instead of HBM: 3 tensor passes of DRAM
// traffic instead of 5.  A wait that exceeds ~2 s traps.
template <int PASS, int COLS, int MINB>
__global__ void __launch_bounds__(256, MINB) in_bwd_rows_kernel(const __grid_constant__ InBwdParams P) {
  const int64_t np = blockIdx.y;
  const int n = (int)(np / P.C8), p = (int)(np - (int64_t)n * P.C8);
  float scale[8], shift[8], m1[8], m2[8];
  // (rstd, -mean * rstd) of the plane's 8 channels are formed in fp64 (norm_params8) by 8 lanes, one channel each, and broadcast through
  // shared memory like in_apply_rows_kernel does (16 % fewer instructions).  Measured: no change in time (7.5 ms per end-to-end training
  // step either way; two rows per warp in the apply pass: none either) - ncu (profiles/r02c_inbwd_metrics.csv) shows issue slots 58-74 % busy at
  // 17-44 % of DRAM throughput with 27 warps resident, yet neither fewer instructions nor more loads in flight per lane (variants
  // above) move it: the passes behave latency-bound on their few-KB-per-warp row segments
  __shared__ float s_np[32];
  if (threadIdx.x < 8) {
    const double2* s2 = reinterpret_cast<const double2*>(P.stats + np * 32);
    const double2 q = s2[2 * threadIdx.x], r = s2[2 * threadIdx.x + 1];
    const double m0 = q.x * (double)P.inv_hw;
    const double var = fmax(q.y * (double)P.inv_hw - m0 * m0, 0.0);
    const float rstd = rsqrtf((float)var + P.eps);
    s_np[threadIdx.x] = rstd;
    s_np[8 + threadIdx.x] = -(float)(m0 + r.x) * rstd;
  } else if (PASS == 1 && threadIdx.x >= 32 && threadIdx.x < 48) {
    s_np[threadIdx.x - 16] = P.sums[np * 16 + (threadIdx.x - 32)] * P.inv_hw;      // [16 + 2e] = m1[e], [17 + 2e] = m2[e]
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) { scale[e] = s_np[e]; shift[e] = s_np[8 + e]; m1[e] = 0.f; m2[e] = 0.f; }
  if (PASS == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { m1[e] = s_np[16 + 2 * e]; m2[e] = s_np[17 + 2 * e]; }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
  bool bad = false;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (PASS == 0 || PASS == 2) {
    in_bwd_rows_pass<0, COLS>(P, np, n, p, scale, shift, m1, m2, s1, s2, bad, gridDim.x);
    __shared__ float sh[16][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = s1[e], b = s2[e];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if (lane == 0) { sh[2 * e][warp] = a; sh[2 * e + 1][warp] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
      atomicAdd(P.sums + np * 16 + threadIdx.x, t);
    }
  }
  if (PASS == 2) {
    __shared__ float s_m[16];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t* cnt = P.sync + np;
      red_release_gpu_add(cnt, 1u);
      const long long t0 = clock64();
      while (ld_acquire_gpu(cnt) < gridDim.x) {
        __nanosleep(100);
        if (clock64() - t0 > (1ll << 32)) __trap();
      }
    }
    __syncthreads();
    if (threadIdx.x < 16) s_m[threadIdx.x] = __ldcg(P.sums + np * 16 + threadIdx.x) * P.inv_hw;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) { m1[e] = s_m[2 * e]; m2[e] = s_m[2 * e + 1]; }
  }
  if (PASS == 1 || PASS == 2) {
    in_bwd_rows_pass<1, COLS>(P, np, n, p, scale, shift, m1, m2, s1, s2, bad, gridDim.x);
    if (bad && P.ovf) atomicOr(P.ovf, 1);
  }
}

// backward of  x_next = pad(act(conv + bias))  (a layer WITHOUT normalisation, e.g. the discriminator's first):
// g = dY * act'(y) with y read back from the stored activation; dbias[c] += sum g  (scaled like g)
struct ActBwdExtra {
  const uint4* yact;   // stored activation, P8 in the consumer's format
  ActGeom yg;
  float* dbias;        // [C8*8], zeroed by the caller
};

__global__ void __launch_bounds__(256) act_bwd_kernel(const __grid_constant__ InBwdParams P, const __grid_constant__ ActBwdExtra X) {
  const int64_t np = blockIdx.y;
  const int n = (int)(np / P.C8), p = (int)(np - (int64_t)n * P.C8);
  const ActGeom& g = P.gg;
  const int total = g.Hp * g.Wp;
  float bs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) bs[e] = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int yy = i / g.Wp, xx = i - yy * g.Wp;
    const int y = yy - g.pad_t, x = xx - g.pad_l;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < P.f.H && x >= 0 && x < P.f.W) {
      float dy[8], a[8], v[8];
      fold_load(P, np, y, x, dy);
      unpack8(X.yact[act_unit(X.yg, n, p, y + X.yg.pad_t, x + X.yg.pad_l)], a, P.f16);
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[e] = dy[e] * act_grad(a[e], P.act); bs[e] += v[e]; }
      o = pack8(v, P.f16);
    }
    P.g[((int64_t)n * g.C8 + p) * g.plane_units + plane_unit(g, yy, xx)] = o;
  }
  __shared__ float sh[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float t = bs[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sh[e][warp] = t;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
    atomicAdd(X.dbias + p * 8 + threadIdx.x, t);
  }
}

// g_pre = grad_out * act'(out) * scale  (fp32 NCHW), the output layer has no norm
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ out, const float* __restrict__ gout, int64_t n_per_c,
                                                       int C, int64_t total, int act, float scale, float* __restrict__ gpre) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / n_per_c) % C);
    const float o = out[i];
    float d = 1.f;
    if (act == NHVR_ACT_TANH || (act == NHVR_ACT_TANH_SIGMOID_LAST && c != C - 1)) d = 1.f - o * o;
    else if (act == NHVR_ACT_TANH_SIGMOID_LAST) d = o * (1.f - o);
    else if (act == NHVR_ACT_RELU) d = o > 0.f ? 1.f : 0.f;
    else if (act == NHVR_ACT_LRELU02) d = o > 0.f ? 1.f : 0.2f;
    gpre[i] = gout[i] * d * scale;
  }
}

// db[c] (+)= scale * sum_{n,h,w} g[n][c][h][w]
// grid (C, chunks): block (c, k) sums chunk k of every image's plane c (float4 loads, fp64 across threads) and adds its scaled
// partial to db[c] (zeroed by the host wrapper unless accumulating).  One block per channel - the first version - took 314 us
// for the 4-channel RGB + mask head at 8 x 512^2 (4 resident blocks); ncu r02b_ncu_launches_train.csv.
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ g, int N, int C, int64_t HW, float scale,
                                                        float* __restrict__ db) {
  const int c = blockIdx.x;
  const int64_t per = (HW + gridDim.y - 1) / gridDim.y;
  const int64_t i0 = (int64_t)blockIdx.y * per, i1 = min(HW, i0 + per);
  double s = 0.0;
  for (int n = 0; n < N; ++n) {
    const float* p = g + ((int64_t)n * C + c) * HW;
    float part = 0.f;
    if ((((uintptr_t)(p + i0)) & 15) == 0) {
      const int64_t n4 = (i1 - i0) >> 2;
      const float4* p4 = reinterpret_cast<const float4*>(p + i0);
      for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) { const float4 v = __ldg(p4 + i); part += (v.x + v.y) + (v.z + v.w); }
      for (int64_t i = i0 + (n4 << 2) + threadIdx.x; i < i1; i += blockDim.x) part += p[i];
    } else {
      for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) part += p[i];
    }
    s += (double)part;
  }
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    atomicAdd(db + c, (float)(t * (double)scale));
  }
}

// folded input gradient -> NCHW fp32 (first C channels), scaled: the chain's dL/d(input)
__global__ void __launch_bounds__(256) fold_unpack_kernel(const __grid_constant__ InBwdParams P, float* __restrict__ dst, int C, float scale) {
  const int64_t np = blockIdx.y;
  const int n = (int)(np / P.C8), p = (int)(np - (int64_t)n * P.C8);
  const int total = P.f.H * P.f.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / P.f.W, x = i - y * P.f.W;
    float dy[8];
    fold_load(P, np, y, x, dy);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = p * 8 + e;
      if (c < C) dst[((int64_t)n * C + c) * total + i] = dy[e] * scale;
    }
  }
}

}  // namespace nhvr

static int make_in_bwd(InBwdParams& P, const void* dx, int32_t dx_H, int32_t dx_W, int32_t pad_t, int32_t pad_l, int32_t reflect,
                       const void* skip, const void* raw, const nhvr_act_desc* raw_desc, const double* stats, float* sums, float eps,
                       int32_t act) {
  if (!dx || !raw_desc) return NHVR_ERR_NULL;
  const ActGeom rg = make_geom(*raw_desc);
  if (rg.pad_t || rg.pad_l || rg.Hp != rg.H || rg.Wp != rg.W || rg.split) return NHVR_ERR_SHAPE;   // raw / skip are un-padded
  if (dx_H < rg.H + pad_t || dx_W < rg.W + pad_l) return NHVR_ERR_SHAPE;
  P.dx = reinterpret_cast<const uint4*>(dx);
  P.skip = reinterpret_cast<const uint4*>(skip);
  P.raw = reinterpret_cast<const uint4*>(raw);
  P.stats = stats; P.sums = sums;
  P.g = nullptr; P.dy_out = nullptr;
  P.f.H = rg.H; P.f.W = rg.W; P.f.pad_t = pad_t; P.f.pad_l = pad_l; P.f.Hp = dx_H; P.f.Wp = dx_W; P.f.reflect = reflect;
  P.N = rg.N; P.C8 = rg.C8;
  P.eps = eps; P.inv_hw = 1.0f / ((float)rg.H * (float)rg.W);
  P.act = act; P.f16 = operand_f16();
  P.ovf = overflow_flag();
  if (rg.hilo) return NHVR_ERR_UNSUPPORTED;          // split precision is a forward-only (inference) format
  return NHVR_OK;
}

#define NHVR_POST(...)                                                    \
  count_launch();                                                         \
  {                                                                       \
    cudaError_t e_ = cudaGetLastError();                                  \
    if (e_ != cudaSuccess) { note_cuda_error(e_); return NHVR_ERR_CUDA; } \
  }

extern "C" int nhvr_in_bwd(const void* dx, int32_t dx_H, int32_t dx_W, int32_t pad_t, int32_t pad_l, int32_t reflect,
                           const void* skip, const void* raw, const nhvr_act_desc* raw_desc, const double* stats, float eps,
                           int32_t act, float* sums, void* g, const nhvr_act_desc* g_desc, void* dy_out, void* stream) {
  if (!raw || !stats || !sums || !g || !g_desc) return NHVR_ERR_NULL;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  InBwdParams P;
  int st = make_in_bwd(P, dx, dx_H, dx_W, pad_t, pad_l, reflect, skip, raw, raw_desc, stats, sums, eps, act);
  if (st != NHVR_OK) return st;
  P.gg = make_geom(*g_desc);
  if (P.gg.N != P.N || P.gg.C8 != P.C8 || P.gg.H != P.f.H || P.gg.W != P.f.W) return NHVR_ERR_SHAPE;
  P.g = reinterpret_cast<uint4*>(g);
  P.dy_out = reinterpret_cast<uint4*>(dy_out);
  cudaStream_t s = (cudaStream_t)stream;
  const int planes = P.N * P.C8;
  P.sync = reinterpret_cast<uint32_t*>(sums + (size_t)planes * 16);
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)planes * 17 * sizeof(float), s);      // sums + arrival counters
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  static const char* rows_env = std::getenv("NHVR_BWD_ROWS");
  static const char* one_env = std::getenv("NHVR_BWD_ONE_LAUNCH");
  // measured (profiles/r02b_in_fused.md): the single launch LOSES - 9.7 vs 8.2 ms per end-to-end training step, 2.5 vs 2.0 ms per UV
  // pre-train step: blocks idle at the meeting point cost more than the L2-served second pass saves - so it is opt-in
  if (!(rows_env && std::atoi(rows_env) == 0) && (one_env && std::atoi(one_env) != 0)) {
    // Both passes in ONE launch: one row per warp and pass, so that few planes are in flight at a time (296 resident blocks /
    // 65 blocks per 512-row plane = 4.5 planes x 12.6 MB of dX + raw + g: L2-resident); a plane's blocks (<= 148: half of
    // the resident blocks of this kernel) meet at an arrival counter between the passes and the second pass re-reads its
    // rows from L2.
    int sms = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int gx = std::max(1, std::min((P.gg.Hp + 7) / 8, sms));
    in_bwd_rows_kernel<2, 4, 2><<<dim3(gx, planes), 256, 0, s>>>(P);
    NHVR_POST();
    return NHVR_OK;
  }
  if (!(rows_env && std::atoi(rows_env) == 0)) {
    // reduce: ~4 rows per warp (fewer atomics per plane), apply: one row per warp
    // NHVR_BWD_VARIANT = 0 (4 columns in flight per lane, 2 blocks / SM: 122-124 registers) | 1 (2 columns, 3 blocks) | 2 (2, 4) |
    // 3 (1 column, 4 blocks: 64 registers).  Measured per end-to-end training step / UV pre-train step (ms of in_bwd):
    // 8.14 / 1.98, 7.24 / 1.71, 7.77 / 1.72, 7.26 / 1.60 -> occupancy beats loads in flight per thread; default 3
    static int variant = -1;
    if (variant < 0) { const char* ev = std::getenv("NHVR_BWD_VARIANT"); variant = ev ? std::atoi(ev) : 3; }
    // reduce pass: every block does the same work and 4 blocks fit an SM, so the launch runs in rounds of 148 * 4 blocks.  Row blocks per
    // plane are chosen to minimise rounds x rows per warp (128^2 x 192 planes: 4 row blocks = 768 blocks = 2 rounds of 4 rows; 16 row blocks
    // = 3072 blocks = 6 rounds of 1 row); ties go to the finer split.  NHVR_BWD_GX0 forces a value.
    int gx0 = std::max(1, (P.f.H + 31) / 32);
    {
      static int force = -1;
      if (force < 0) { const char* e = std::getenv("NHVR_BWD_GX0"); force = e ? std::atoi(e) : 0; }
      if (force > 0) gx0 = force;
      else {
        const int slots = 148 * 4, gmax = std::max(1, (P.f.H + 7) / 8);
        long best = -1;
        for (int gx = 1; gx <= gmax; ++gx) {
          const long rounds = ((long)gx * planes + slots - 1) / slots, rows = (P.f.H + gx * 8 - 1) / (gx * 8);
          const long cost = rounds * rows * 64 + rounds;          // + a little per round (block prologue, atomics)
          if (best < 0 || cost <= best) { best = cost; gx0 = gx; }
        }
      }
    }
    const dim3 g0(gx0, planes), g1(std::max(1, (P.gg.Hp + 7) / 8), planes);
    if (variant == 1) { in_bwd_rows_kernel<0, 2, 3><<<g0, 256, 0, s>>>(P); NHVR_POST(); in_bwd_rows_kernel<1, 2, 3><<<g1, 256, 0, s>>>(P); }
    else if (variant == 2) { in_bwd_rows_kernel<0, 2, 4><<<g0, 256, 0, s>>>(P); NHVR_POST(); in_bwd_rows_kernel<1, 2, 4><<<g1, 256, 0, s>>>(P); }
    else if (variant == 3) { in_bwd_rows_kernel<0, 1, 4><<<g0, 256, 0, s>>>(P); NHVR_POST(); in_bwd_rows_kernel<1, 1, 4><<<g1, 256, 0, s>>>(P); }
    else { in_bwd_rows_kernel<0, 4, 2><<<g0, 256, 0, s>>>(P); NHVR_POST(); in_bwd_rows_kernel<1, 4, 2><<<g1, 256, 0, s>>>(P); }
    NHVR_POST();
    return NHVR_OK;
  }
  in_bwd_reduce_kernel<<<dim3(grid_x_for((int64_t)P.f.H * P.f.W, planes), planes), 256, 0, s>>>(P);
  NHVR_POST();
  in_bwd_apply_kernel<<<dim3(grid_x_for((int64_t)P.gg.Hp * P.gg.Wp, planes), planes), 256, 0, s>>>(P);
  NHVR_POST();
  return NHVR_OK;
}

extern "C" int nhvr_act_bwd(const void* dx, int32_t dx_H, int32_t dx_W, int32_t pad_t, int32_t pad_l, int32_t reflect,
                            const void* skip, const void* yact, const nhvr_act_desc* y_desc, int32_t act, void* g,
                            const nhvr_act_desc* g_desc, float* dbias, void* stream) {
  if (!yact || !y_desc || !g || !g_desc || !dbias) return NHVR_ERR_NULL;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  InBwdParams P;
  ActBwdExtra X;
  X.yact = reinterpret_cast<const uint4*>(yact);
  X.yg = make_geom(*y_desc);
  X.dbias = dbias;
  nhvr_act_desc interior = *y_desc;
  interior.pad_t = interior.pad_l = interior.pad_b = interior.pad_r = 0; interior.split = 0;
  int st = make_in_bwd(P, dx, dx_H, dx_W, pad_t, pad_l, reflect, skip, yact, &interior, nullptr, nullptr, 0.f, act);
  if (st != NHVR_OK) return st;
  P.gg = make_geom(*g_desc);
  if (P.gg.N != P.N || P.gg.C8 != P.C8 || P.gg.H != P.f.H || P.gg.W != P.f.W) return NHVR_ERR_SHAPE;
  P.g = reinterpret_cast<uint4*>(g);
  const int planes = P.N * P.C8;
  act_bwd_kernel<<<dim3(grid_x_for((int64_t)P.gg.Hp * P.gg.Wp, planes), planes), 256, 0, (cudaStream_t)stream>>>(P, X);
  NHVR_POST();
  return NHVR_OK;
}

extern "C" int nhvr_fold_unpack(const void* dx, int32_t dx_H, int32_t dx_W, int32_t pad_t, int32_t pad_l, int32_t reflect,
                                const nhvr_act_desc* interior_desc, float* dst, int32_t C, float scale, void* stream) {
  if (!dst) return NHVR_ERR_NULL;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  InBwdParams P;
  int st = make_in_bwd(P, dx, dx_H, dx_W, pad_t, pad_l, reflect, nullptr, dx, interior_desc, nullptr, nullptr, 0.f, 0);
  if (st != NHVR_OK) return st;
  const int planes = P.N * P.C8;
  fold_unpack_kernel<<<dim3(grid_x_for((int64_t)P.f.H * P.f.W, planes), planes), 256, 0, (cudaStream_t)stream>>>(P, dst, C, scale);
  NHVR_POST();
  return NHVR_OK;
}

extern "C" int nhvr_head_bwd(const float* out, const float* grad_out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t act,
                             float scale, float* g_pre, void* stream) {
  if (!out || !grad_out || !g_pre) return NHVR_ERR_NULL;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int64_t total = (int64_t)N * C * H * W;
  head_bwd_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(out, grad_out, (int64_t)H * W, C,
                                                                                                          total, act, scale, g_pre);
  NHVR_POST();
  return NHVR_OK;
}

extern "C" int nhvr_bias_grad(const float* g, int32_t N, int32_t C, int32_t H, int32_t W, float scale, int32_t accumulate, float* db,
                              void* stream) {
  if (!g || !db) return NHVR_ERR_NULL;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!accumulate) {
    cudaError_t e0 = cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), (cudaStream_t)stream);
    if (e0 != cudaSuccess) { note_cuda_error(e0); return NHVR_ERR_CUDA; }
  }
  const int64_t HW = (int64_t)H * W;
  // ~4 blocks per SM in total, at least 4096 elements per block and image
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>((148 * 4 + C - 1) / C, HW / 4096));
  bias_grad_kernel<<<dim3(C, chunks), 256, 0, (cudaStream_t)stream>>>(g, N, C, HW, scale, db);
  NHVR_POST();
  return NHVR_OK;
}
