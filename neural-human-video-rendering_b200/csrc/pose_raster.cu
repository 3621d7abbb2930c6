// Keypoints -> pose maps on the GPU: the step just before the hot path.
// [T][25][3] OpenPose BODY_25 keypoints (x, y, confidence; REF keypoints/frame000NN_keypoints.json) are drawn as the
// 3-channel stick figure of nhvr_b200/pose.py (`--input_nc 3`, start.sh:24) straight into the fp32 NCHW tensor the UV
// generator's pack kernel reads; channels beyond 3 (LaplaceProj, `--use_laplace`: no data in the fixtures) are zero.
// One thread per pixel, limbs in OpenPose order, the last limb that covers a pixel wins - the arithmetic mirrors
// pose.rasterize operation by operation in IEEE fp32 (no FMA contraction), so the maps are bit-identical to the host's.
#include "common.cuh"
#include "p8.cuh"
#include <algorithm>

namespace nhvr {

extern void note_cuda_error(cudaError_t e);
extern void count_launch();
extern int arch_ok_cached();

constexpr int kLimbs = 24;
__constant__ int c_pairs[kLimbs][2] = {{1, 8},  {1, 2},   {1, 5},   {2, 3},   {3, 4},   {5, 6},   {6, 7},   {8, 9},
                                       {9, 10}, {10, 11}, {8, 12},  {12, 13}, {13, 14}, {1, 0},   {0, 15},  {15, 17},
                                       {0, 16}, {16, 18}, {14, 19}, {19, 20}, {14, 21}, {11, 22}, {22, 23}, {11, 24}};

struct RasterParams {
  const float* kps;      // [T][25][3]
  float* out;            // [T][pose_nc][size][size]
  int32_t T, size, pose_nc;
  float scale;           // size / src_size
  float half;            // half line width in output pixels
  float conf_thresh;
  uint8_t colors[kLimbs][3];
};

__global__ void __launch_bounds__(256) pose_raster_kernel(const __grid_constant__ RasterParams P) {
  __shared__ float seg[kLimbs][6];        // ax, ay, dx, dy, L2, drawn
  const int t = blockIdx.y;
  if (threadIdx.x < kLimbs) {
    const float* k = P.kps + (int64_t)t * 75;
    const int a = c_pairs[threadIdx.x][0], b = c_pairs[threadIdx.x][1];
    const bool ok = !(k[a * 3 + 2] < P.conf_thresh || k[b * 3 + 2] < P.conf_thresh);
    const float ax = __fmul_rn(k[a * 3], P.scale), ay = __fmul_rn(k[a * 3 + 1], P.scale);
    const float bx = __fmul_rn(k[b * 3], P.scale), by = __fmul_rn(k[b * 3 + 1], P.scale);
    const float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay);
    seg[threadIdx.x][0] = ax; seg[threadIdx.x][1] = ay; seg[threadIdx.x][2] = dx; seg[threadIdx.x][3] = dy;
    seg[threadIdx.x][4] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    seg[threadIdx.x][5] = ok ? 1.f : 0.f;
  }
  __syncthreads();
  const int64_t HW = (int64_t)P.size * P.size;
  const float h2 = __fmul_rn(P.half, P.half);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / P.size), x = (int)(i - (int64_t)y * P.size);
    const float px = (float)x, py = (float)y;
    int hit = -1;
#pragma unroll 1
    for (int l = 0; l < kLimbs; ++l) {
      if (seg[l][5] == 0.f) continue;
      const float ax = seg[l][0], ay = seg[l][1], dx = seg[l][2], dy = seg[l][3], L2 = seg[l][4];
      const float rx = __fsub_rn(px, ax), ry = __fsub_rn(py, ay);
      float tt = 0.f;
      if (L2 > 1e-6f) {
        tt = __fdiv_rn(__fadd_rn(__fmul_rn(rx, dx), __fmul_rn(ry, dy)), L2);
        tt = fminf(fmaxf(tt, 0.f), 1.f);
      }
      const float ex = __fsub_rn(px, __fadd_rn(ax, __fmul_rn(tt, dx)));
      const float ey = __fsub_rn(py, __fadd_rn(ay, __fmul_rn(tt, dy)));
      const float d2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
      if (d2 <= h2) hit = l;
    }
    float* o = P.out + (int64_t)t * P.pose_nc * HW + i;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = hit >= 0 ? (float)P.colors[hit][c] : 0.f;
      if (c < P.pose_nc) o[c * HW] = __fsub_rn(__fdiv_rn(v, 127.5f), 1.0f);
    }
    for (int c = 3; c < P.pose_nc; ++c) o[c * HW] = 0.f;
  }
}

}  // namespace nhvr

using namespace nhvr;

extern "C" int nhvr_pose_rasterize(const float* kps, int32_t T, int32_t size, float src_size, float thickness, float conf_thresh,
                                   int32_t pose_nc, const uint8_t* limb_colors_host, float* out, void* stream) {
  if (!kps || !out || !limb_colors_host) return NHVR_ERR_NULL;
  if (T < 1 || size < 8 || pose_nc < 1 || src_size <= 0.f) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  RasterParams P;
  P.kps = kps; P.out = out; P.T = T; P.size = size; P.pose_nc = pose_nc;
  P.scale = (float)((double)size / (double)src_size);
  P.half = (float)(std::max((double)thickness * size / 512.0, 1.0) * 0.5);
  P.conf_thresh = conf_thresh;
  for (int l = 0; l < kLimbs; ++l)
    for (int c = 0; c < 3; ++c) P.colors[l][c] = limb_colors_host[l * 3 + c];
  const int64_t HW = (int64_t)size * size;
  dim3 grid((unsigned)std::min<int64_t>((HW + 255) / 256, 148 * 8), T);
  pose_raster_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}
