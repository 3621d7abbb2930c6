// Training-side reductions (fp32 inputs, fp64 accumulation: the 1e-3 relative bound on losses is met with
// orders of magnitude to spare) and the discriminator's inter-scale AvgPool2d(3, s2, p1,
// count_include_pad=False).  All are streaming, coalesced, one pass over their inputs.
//
// Every loss kernel ADDS partial sums into a caller-zeroed double accumulator array; the host divides
// by the element count (the reference's torch losses are means; evidence for the terms:
// train_start/pretrain_start.sh:31-37 --lambda_L2/--lambda_UV/--lambda_Prob/--lambda_Temp,
// pix2pixHD GANLoss / feature matching).
#include "common.cuh"
#include <cmath>
#include "p8.cuh"
#include <algorithm>

namespace nhvr {

extern void note_cuda_error(cudaError_t e);
extern void count_launch();
extern int arch_ok_cached();

NHVR_DEVINL double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-reduce K partial sums (one double each per thread) and add them to acc[0..K)
template <int K>
NHVR_DEVINL void block_accumulate(const double (&v)[K], double* acc) {
  __shared__ double sh[K][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double w = warp_sum(v[k]);
    if (lane == 0) sh[k][warp] = w;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, t);
  }
}

// mode 0: sum (a-b)^2   mode 1: sum |a-b|   mode 2: sum (a-c)^2 (b unused, c = target constant)
__global__ void __launch_bounds__(256) loss_pair_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                        int mode, float c, double* acc) {
  double s[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float part = 0.f;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float x = __ldg(a + i);
    const float d = x - (mode == 2 ? c : __ldg(b + i));
    part += (mode == 1) ? fabsf(d) : d * d;
    if (++cnt == 64) { s[0] += part; part = 0.f; cnt = 0; }   // bounded fp32 run, fp64 carry
  }
  s[0] += part;
  block_accumulate<1>(s, acc);
}

// uvp [N,73,H,W]; dp_i int32 [N,H,W] in 0..24; dp_uv [N,2,H,W]
// acc[0] += sum_fg |u_k - U| + |v_k - V| (k = ground-truth part)   acc[1] += #fg   acc[2] += sum CE
__global__ void __launch_bounds__(256) loss_uv_prob_kernel(const float* __restrict__ uvp, const int32_t* __restrict__ dp_i,
                                                           const float* __restrict__ dp_uv, int N, int64_t HW, double* acc) {
  double s[3] = {0.0, 0.0, 0.0};
  const int64_t total = (int64_t)N * HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int64_t pix = idx - (int64_t)n * HW;
    const float* base = uvp + (int64_t)n * 73 * HW + pix;
    const int part = dp_i[idx];
    float lg[25], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 25; ++k) { lg[k] = __ldg(base + (int64_t)k * HW); mx = fmaxf(mx, lg[k]); }
    float den = 0.f, tgt = 0.f;
#pragma unroll
    for (int k = 0; k < 25; ++k) { den += expf(lg[k] - mx); if (k == part) tgt = lg[k]; }
    s[2] += (double)(logf(den) + mx - tgt);
    if (part > 0) {
      const float u = fminf(fmaxf(__ldg(base + (int64_t)(24 + part) * HW) * 0.5f + 0.5f, 0.f), 1.f);
      const float v = fminf(fmaxf(__ldg(base + (int64_t)(48 + part) * HW) * 0.5f + 0.5f, 0.f), 1.f);
      const float U = __ldg(dp_uv + (int64_t)n * 2 * HW + pix), V = __ldg(dp_uv + ((int64_t)n * 2 + 1) * HW + pix);
      s[0] += (double)(fabsf(u - U) + fabsf(v - V));
      s[1] += 1.0;
    }
  }
  block_accumulate<3>(s, acc);
}

// acc[0] += sum |cur - warp(prev, flow)| over [N,C,H,W]; warp = bilinear sample of prev at (x+fx, y+fy),
// border-clamped (grid_sample(padding_mode='border', align_corners=True) in pixel units)
__global__ void __launch_bounds__(256) loss_temporal_kernel(const float* __restrict__ cur, const float* __restrict__ prev,
                                                            const float* __restrict__ flow, int N, int C, int H, int W, double* acc) {
  double s[1] = {0.0};
  const int64_t HW = (int64_t)H * W;
  const int64_t total = (int64_t)N * HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int64_t pix = idx - (int64_t)n * HW;
    const int y = (int)(pix / W), x = (int)(pix - (int64_t)y * W);
    float sx = (float)x + __ldg(flow + (int64_t)n * 2 * HW + pix);
    float sy = (float)y + __ldg(flow + ((int64_t)n * 2 + 1) * HW + pix);
    sx = fminf(fmaxf(sx, 0.f), (float)(W - 1));
    sy = fminf(fmaxf(sy, 0.f), (float)(H - 1));
    const float x0f = floorf(sx), y0f = floorf(sy);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float wx = sx - x0f, wy = sy - y0f;
    for (int c = 0; c < C; ++c) {
      const float* p = prev + ((int64_t)n * C + c) * HW;
      const float w = (1.f - wy) * ((1.f - wx) * __ldg(p + (int64_t)y0 * W + x0) + wx * __ldg(p + (int64_t)y0 * W + x1)) +
                      wy * ((1.f - wx) * __ldg(p + (int64_t)y1 * W + x0) + wx * __ldg(p + (int64_t)y1 * W + x1));
      s[0] += (double)fabsf(__ldg(cur + ((int64_t)n * C + c) * HW + pix) - w);
    }
  }
  block_accumulate<1>(s, acc);
}

// gradient of  w_uv * uv_loss + w_prob * prob_loss  w.r.t. uvp (all 73 channels written), times *gscale.
// acc3 are the forward sums (acc3[1] = foreground count) so no host round trip is needed.
__global__ void __launch_bounds__(256) loss_uv_prob_bwd_kernel(const float* __restrict__ uvp, const int32_t* __restrict__ dp_i,
                                                               const float* __restrict__ dp_uv, int N, int64_t HW, const double* acc3,
                                                               float w_uv, float w_prob, const float* gscale, float* __restrict__ grad) {
  const int64_t total = (int64_t)N * HW;
  const float gs = gscale ? *gscale : 1.f;
  const float cuv = gs * w_uv / (float)fmax(acc3[1], 1.0);
  const float cpr = gs * w_prob / (float)total;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int64_t pix = idx - (int64_t)n * HW;
    const float* base = uvp + (int64_t)n * 73 * HW + pix;
    float* gb = grad + (int64_t)n * 73 * HW + pix;
    const int part = dp_i[idx];
    float lg[25], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 25; ++k) { lg[k] = __ldg(base + (int64_t)k * HW); mx = fmaxf(mx, lg[k]); }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 25; ++k) { lg[k] = expf(lg[k] - mx); den += lg[k]; }
    const float inv = 1.f / den;
#pragma unroll
    for (int k = 0; k < 25; ++k) gb[(int64_t)k * HW] = cpr * (lg[k] * inv - (k == part ? 1.f : 0.f));
    for (int k = 1; k <= 24; ++k) {
      float gu = 0.f, gv = 0.f;
      if (k == part) {
        const float xu = __ldg(base + (int64_t)(24 + k) * HW), xv = __ldg(base + (int64_t)(48 + k) * HW);
        const float tu = xu * 0.5f + 0.5f, tv = xv * 0.5f + 0.5f;
        const float u = fminf(fmaxf(tu, 0.f), 1.f), v = fminf(fmaxf(tv, 0.f), 1.f);
        const float du = u - __ldg(dp_uv + (int64_t)n * 2 * HW + pix), dv = v - __ldg(dp_uv + ((int64_t)n * 2 + 1) * HW + pix);
        // d|t|/dt = sign(t) (0 at 0); clamp passes the gradient on the closed interval, as torch.clamp does
        if (tu >= 0.f && tu <= 1.f) gu = 0.5f * cuv * (du > 0.f ? 1.f : (du < 0.f ? -1.f : 0.f));
        if (tv >= 0.f && tv <= 1.f) gv = 0.5f * cuv * (dv > 0.f ? 1.f : (dv < 0.f ? -1.f : 0.f));
      }
      gb[(int64_t)(24 + k) * HW] = gu;
      gb[(int64_t)(48 + k) * HW] = gv;
    }
  }
}

// gradient of coef * mean|cur - warp(prev, flow)| w.r.t. cur (prev is the detached previous frame)
__global__ void __launch_bounds__(256) loss_temporal_bwd_kernel(const float* __restrict__ cur, const float* __restrict__ prev,
                                                                const float* __restrict__ flow, int N, int C, int H, int W, float coef,
                                                                const float* gscale, float* __restrict__ gcur) {
  const int64_t HW = (int64_t)H * W;
  const int64_t total = (int64_t)N * HW;
  const float k = coef * (gscale ? *gscale : 1.f) / (float)(total * C);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int64_t pix = idx - (int64_t)n * HW;
    const int y = (int)(pix / W), x = (int)(pix - (int64_t)y * W);
    float sx = (float)x + __ldg(flow + (int64_t)n * 2 * HW + pix);
    float sy = (float)y + __ldg(flow + ((int64_t)n * 2 + 1) * HW + pix);
    sx = fminf(fmaxf(sx, 0.f), (float)(W - 1));
    sy = fminf(fmaxf(sy, 0.f), (float)(H - 1));
    const float x0f = floorf(sx), y0f = floorf(sy);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const float wx = sx - x0f, wy = sy - y0f;
    for (int c = 0; c < C; ++c) {
      const float* p = prev + ((int64_t)n * C + c) * HW;
      const float w = (1.f - wy) * ((1.f - wx) * __ldg(p + (int64_t)y0 * W + x0) + wx * __ldg(p + (int64_t)y0 * W + x1)) +
                      wy * ((1.f - wx) * __ldg(p + (int64_t)y1 * W + x0) + wx * __ldg(p + (int64_t)y1 * W + x1));
      const float d = __ldg(cur + ((int64_t)n * C + c) * HW + pix) - w;
      gcur[((int64_t)n * C + c) * HW + pix] = d > 0.f ? k : (d < 0.f ? -k : 0.f);
    }
  }
}

// gradient of the mean reductions w.r.t. a:  mode 0: 2(a-b)/n   mode 1: sign(a-b)/n   mode 2: 2(a-c)/n ; times coef * *gscale
__global__ void __launch_bounds__(256) loss_pair_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, int mode,
                                                            float c, float coef, const float* gscale, int accumulate, float* __restrict__ ga) {
  const float k = coef * (gscale ? *gscale : 1.f) / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - (mode == 2 ? c : b[i]);
    const float g = (mode == 1) ? (d > 0.f ? k : (d < 0.f ? -k : 0.f)) : 2.f * k * d;
    ga[i] = accumulate ? ga[i] + g : g;
  }
}

// backward of AvgPool2d(3, s2, p1, count_include_pad=False): gin (+)= sum over the windows containing the pixel of gout / count
__global__ void __launch_bounds__(256) avgpool3s2_bwd_kernel(const float* __restrict__ gout, int64_t planes, int H, int W, int Ho, int Wo,
                                                             int accumulate, float* __restrict__ gin) {
  const int64_t total = planes * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int64_t pl = idx / ((int64_t)W * H);
    float s = 0.f;
    // windows (yo, xo) with |2*yo - y| <= 1
    const int yo0 = max((y - 1 + 1) / 2, 0), yo1 = min((y + 1) / 2, Ho - 1);
    const int xo0 = max((x - 1 + 1) / 2, 0), xo1 = min((x + 1) / 2, Wo - 1);
    for (int yo = yo0; yo <= yo1; ++yo)
      for (int xo = xo0; xo <= xo1; ++xo) {
        const int cy = min(2 * yo + 1, H - 1) - max(2 * yo - 1, 0) + 1;
        const int cx = min(2 * xo + 1, W - 1) - max(2 * xo - 1, 0) + 1;
        s += gout[(pl * Ho + yo) * Wo + xo] / (float)(cy * cx);
      }
    gin[idx] = accumulate ? gin[idx] + s : s;
  }
}

// AvgPool2d(kernel 3, stride 2, padding 1, count_include_pad=False) on NCHW fp32
__global__ void __launch_bounds__(256) avgpool3s2_kernel(const float* __restrict__ in, int64_t planes, int H, int W, int Ho, int Wo,
                                                         float* __restrict__ out) {
  const int64_t total = planes * Ho * Wo;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % Wo);
    const int yo = (int)((idx / Wo) % Ho);
    const int64_t pl = idx / ((int64_t)Wo * Ho);
    const float* p = in + pl * H * W;
    float s = 0.f;
    int cnt = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = 2 * yo + dy;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = 2 * xo + dx;
        if (x < 0 || x >= W) continue;
        s += __ldg(p + (int64_t)y * W + x);
        ++cnt;
      }
    }
    out[idx] = s / (float)cnt;
  }
}

static inline int blocks_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8)); }

}  // namespace nhvr

using namespace nhvr;

#define NHVR_POST_LAUNCH()                                         \
  count_launch();                                                  \
  {                                                                \
    cudaError_t e_ = cudaGetLastError();                           \
    if (e_ != cudaSuccess) { note_cuda_error(e_); return NHVR_ERR_CUDA; } \
  }                                                                \
  return NHVR_OK

extern "C" int nhvr_loss_sum_sq_diff(const float* a, const float* b, int64_t n, double* acc, void* stream) {
  if (!a || !b || !acc) return NHVR_ERR_NULL;
  if (n <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  loss_pair_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, 0, 0.f, acc);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_loss_sum_abs_diff(const float* a, const float* b, int64_t n, double* acc, void* stream) {
  if (!a || !b || !acc) return NHVR_ERR_NULL;
  if (n <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  loss_pair_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, 1, 0.f, acc);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_loss_sum_sq_const(const float* a, float target, int64_t n, double* acc, void* stream) {
  if (!a || !acc) return NHVR_ERR_NULL;
  if (n <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  loss_pair_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(a, nullptr, n, 2, target, acc);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_loss_uv_prob(const float* uvp, const int32_t* dp_i, const float* dp_uv, int32_t N, int32_t H, int32_t W,
                                 double* acc3, void* stream) {
  if (!uvp || !dp_i || !dp_uv || !acc3) return NHVR_ERR_NULL;
  if (N <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  loss_uv_prob_kernel<<<blocks_for((int64_t)N * H * W), 256, 0, (cudaStream_t)stream>>>(uvp, dp_i, dp_uv, N, (int64_t)H * W, acc3);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_loss_uv_prob_bwd(const float* uvp, const int32_t* dp_i, const float* dp_uv, int32_t N, int32_t H, int32_t W,
                                     const double* acc3, float w_uv, float w_prob, const float* grad_scale, float* grad,
                                     void* stream) {
  if (!uvp || !dp_i || !dp_uv || !acc3 || !grad) return NHVR_ERR_NULL;
  if (N <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  loss_uv_prob_bwd_kernel<<<blocks_for((int64_t)N * H * W), 256, 0, (cudaStream_t)stream>>>(uvp, dp_i, dp_uv, N, (int64_t)H * W, acc3,
                                                                                            w_uv, w_prob, grad_scale, grad);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_loss_temporal(const float* cur, const float* prev, const float* flow, int32_t N, int32_t C, int32_t H,
                                  int32_t W, double* acc, void* stream) {
  if (!cur || !prev || !flow || !acc) return NHVR_ERR_NULL;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  loss_temporal_kernel<<<blocks_for((int64_t)N * H * W), 256, 0, (cudaStream_t)stream>>>(cur, prev, flow, N, C, H, W, acc);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_loss_temporal_bwd(const float* cur, const float* prev, const float* flow, int32_t N, int32_t C, int32_t H,
                                      int32_t W, float coef, const float* grad_scale, float* grad_cur, void* stream) {
  if (!cur || !prev || !flow || !grad_cur) return NHVR_ERR_NULL;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  loss_temporal_bwd_kernel<<<blocks_for((int64_t)N * H * W), 256, 0, (cudaStream_t)stream>>>(cur, prev, flow, N, C, H, W, coef, grad_scale,
                                                                                             grad_cur);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_loss_pair_bwd(const float* a, const float* b, int64_t n, int32_t mode, float target, float coef,
                                  const float* grad_scale, int32_t accumulate, float* grad_a, void* stream) {
  if (!a || !grad_a || (mode != 2 && !b)) return NHVR_ERR_NULL;
  if (n <= 0 || mode < 0 || mode > 2) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  loss_pair_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, mode, target, coef, grad_scale, accumulate, grad_a);
  NHVR_POST_LAUNCH();
}
// MaxPool2d(2, 2) of the VGG19 feature stack (the reference's perceptual loss, pix2pixHD VGGLoss; README.md:101).  fp32 NCHW, floor
// output size; the backward routes each output gradient to the FIRST maximum of its window in row-major order (torch semantics).
namespace nhvr {
__global__ void __launch_bounds__(256) maxpool2_kernel(const float* __restrict__ in, int64_t planes, int H, int W, int Ho, int Wo,
                                                       float* __restrict__ out) {
  const int64_t total = planes * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho);
    const int64_t pl = i / ((int64_t)Wo * Ho);
    const float* p = in + (pl * H + 2 * yo) * W + 2 * xo;
    out[i] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[W], p[W + 1]));
  }
}
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ in, const float* __restrict__ gout, int64_t planes, int H,
                                                           int W, int Ho, int Wo, float* __restrict__ gin) {
  const int64_t total = planes * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int64_t pl = i / ((int64_t)W * H);
    const int yo = y >> 1, xo = x >> 1;
    float g = 0.f;
    if (yo < Ho && xo < Wo) {
      const float* p = in + (pl * H + 2 * yo) * W + 2 * xo;
      const float v[4] = {p[0], p[1], p[W], p[W + 1]};
      int best = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k] > v[best]) best = k;
      if (best == ((y & 1) << 1 | (x & 1))) g = gout[(pl * Ho + yo) * Wo + xo];
    }
    gin[i] = g;
  }
}
}  // namespace nhvr

extern "C" int nhvr_maxpool2(const float* in, int32_t N, int32_t C, int32_t H, int32_t W, float* out, void* stream) {
  if (!in || !out) return NHVR_ERR_NULL;
  if (N <= 0 || C <= 0 || H < 2 || W < 2) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int Ho = H / 2, Wo = W / 2;
  maxpool2_kernel<<<blocks_for((int64_t)N * C * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(in, (int64_t)N * C, H, W, Ho, Wo, out);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_maxpool2_bwd(const float* in, const float* grad_out, int32_t N, int32_t C, int32_t H, int32_t W, float* grad_in,
                                 void* stream) {
  if (!in || !grad_out || !grad_in) return NHVR_ERR_NULL;
  if (N <= 0 || C <= 0 || H < 2 || W < 2) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int Ho = H / 2, Wo = W / 2;
  maxpool2_bwd_kernel<<<blocks_for((int64_t)N * C * H * W), 256, 0, (cudaStream_t)stream>>>(in, grad_out, (int64_t)N * C, H, W, Ho, Wo, grad_in);
  NHVR_POST_LAUNCH();
}

extern "C" int nhvr_avgpool3s2_bwd(const float* grad_out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t accumulate, float* grad_in,
                                   void* stream) {
  if (!grad_out || !grad_in) return NHVR_ERR_NULL;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  avgpool3s2_bwd_kernel<<<blocks_for((int64_t)N * C * H * W), 256, 0, (cudaStream_t)stream>>>(grad_out, (int64_t)N * C, H, W, Ho, Wo,
                                                                                              accumulate, grad_in);
  NHVR_POST_LAUNCH();
}
extern "C" int nhvr_avgpool3s2(const float* in, int32_t N, int32_t C, int32_t H, int32_t W, float* out, void* stream) {
  if (!in || !out) return NHVR_ERR_NULL;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  avgpool3s2_kernel<<<blocks_for((int64_t)N * C * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(in, (int64_t)N * C, H, W, Ho, Wo, out);
  NHVR_POST_LAUNCH();
}


// =================================================================================================
// Adam over a flat fp32 parameter bucket (pix2pixHD: torch.optim.Adam(lr 2e-4, betas (0.5, 0.999)), no weight decay):
// one launch per bucket instead of four foreach launches per step of the stock optimiser.
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps),   g = grad * grad_scale
// =================================================================================================
namespace nhvr {
__global__ void __launch_bounds__(256) adam_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, int64_t n4, float b1, float b2, float eps, float step_size,
                                                        float inv_sqrt_bc2, float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 P = p[i], G = g[i], M = m[i], V = v[i];
    float* pp = reinterpret_cast<float*>(&P); float* gg = reinterpret_cast<float*>(&G);
    float* mm = reinterpret_cast<float*>(&M); float* vv = reinterpret_cast<float*>(&V);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = gg[e] * grad_scale;
      mm[e] = b1 * mm[e] + (1.f - b1) * gr;
      vv[e] = b2 * vv[e] + (1.f - b2) * gr * gr;
      pp[e] -= step_size * mm[e] / (sqrtf(vv[e]) * inv_sqrt_bc2 + eps);
    }
    p[i] = P; m[i] = M; v[i] = V;
  }
}
}  // namespace nhvr

extern "C" int nhvr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                              float grad_scale, int32_t step, void* stream) {
  if (!p || !g || !m || !v) return NHVR_ERR_NULL;
  if (n <= 0 || (n & 3) || step < 1) return NHVR_ERR_SHAPE;
  if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step), bc2 = 1.0 - std::pow((double)beta2, (double)step);
  adam_step_kernel<<<blocks_for(n / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                                                        reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n / 4, beta1, beta2,
                                                                        eps, (float)((double)lr / bc1), (float)(1.0 / std::sqrt(bc2)), grad_scale);
  NHVR_POST_LAUNCH();
}
