// P8 activation geometry shared by host and device (see include/nhvr.h for the layout definition).
#pragma once
#include <stdint.h>
#include "../../include/nhvr.h"

namespace nhvr {

constexpr int64_t kActSlackUnits = 2048;  // 32 KB tail: conv tiles may over-read past the last plane

struct ActGeom {
  int32_t N, C8, H, W;
  int32_t pad_t, pad_l;
  int32_t Hp, Wp;       // padded (and, if split, even-rounded) extents
  int32_t split, halo;
  int32_t hilo;         // split precision: C8 counts physical planes in groups [hi, hi, lo, lo]
  int64_t plane_units;  // 16-byte units per (n, plane)
};

__host__ __device__ inline ActGeom make_geom(const nhvr_act_desc& d) {
  ActGeom g;
  g.N = d.N; g.C8 = d.C8; g.H = d.H; g.W = d.W;
  g.pad_t = d.pad_t; g.pad_l = d.pad_l;
  g.Hp = d.H + d.pad_t + d.pad_b;
  g.Wp = d.W + d.pad_l + d.pad_r;
  g.split = d.split; g.halo = d.halo; g.hilo = d.hilo;
  if (d.split) { g.Hp += g.Hp & 1; g.Wp += g.Wp & 1; }
  g.plane_units = (int64_t)g.Hp * g.Wp;
  return g;
}

// unit index of padded coordinate (yy, xx) inside one plane
__host__ __device__ inline int64_t plane_unit(const ActGeom& g, int32_t yy, int32_t xx) {
  if (g.split) {
    const int32_t Hq = g.Hp >> 1, Wq = g.Wp >> 1;
    return ((int64_t)(((yy & 1) << 1) | (xx & 1)) * Hq + (yy >> 1)) * Wq + (xx >> 1);
  }
  return (int64_t)yy * g.Wp + xx;
}

__host__ __device__ inline int64_t act_unit(const ActGeom& g, int32_t n, int32_t p, int32_t yy, int32_t xx) {
  return ((int64_t)n * g.C8 + p) * g.plane_units + plane_unit(g, yy, xx);
}

// split precision: physical plane of logical plane lp (hi part; the lo part is 2 planes further)
__host__ __device__ inline int32_t hilo_plane(int32_t lp) { return ((lp >> 1) << 2) | (lp & 1); }

// ReflectionPad2d index map: logical coordinate (may be outside [0, n)) -> source inside [0, n)
__host__ __device__ inline int32_t reflect_idx(int32_t i, int32_t n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

}  // namespace nhvr
