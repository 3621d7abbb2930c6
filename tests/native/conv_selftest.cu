// Native self-test of libnhvr_sm100.so through its C-ABI only (no torch): every conv lowering
// (3x3 s1 reflect, 7x7, 3x3 s2 zero, 4x4 s2/s1 p2, transposed 3x3 s2) against a scalar CPU loop on
// the same bf16-rounded operands.  Prints one line per case; exit code = number of failures.
// This is test infrastructure; the reference arithmetic here is a restatement of nn.Conv2d /
// nn.ConvTranspose2d / nn.ReflectionPad2d / nn.InstanceNorm2d semantics (PyTorch docs).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include "../../include/nhvr.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(99); } } while (0)
#define NK(x) do { int s_ = (x); if (s_ != 0) { printf("nhvr error %d (%s / %s) at %s:%d\n", s_, nhvr_strerror(s_), nhvr_last_cuda_error(), __FILE__, __LINE__); return 1; } } while (0)

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }
static uint32_t rng_state = 12345u;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f; }
static int refl(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * (n - 1) - i; return i; }

struct Case { const char* name; int kind, Cin, Cout, k, stride, pad, N, H, W, halo, epi, act; };

static int run_case(const Case& c) {
  nhvr_conv_desc d{};
  d.kind = c.kind; d.Cin = c.Cin; d.Cout = c.Cout; d.kh = d.kw = c.k; d.stride = c.stride; d.pad = c.pad;
  d.N = c.N; d.H = c.H; d.W = c.W; d.halo = c.halo; d.epilogue = c.epi; d.act = c.act;
  nhvr_conv_plan* plan = nullptr;
  NK(nhvr_conv_plan_create(&d, &plan));
  nhvr_act_desc in_desc;
  NK(nhvr_conv_input_desc(plan, &in_desc));
  int Ho, Wo, Cout8;
  NK(nhvr_conv_output_dims(plan, &Ho, &Wo, &Cout8));

  const size_t in_elems = (size_t)c.N * c.Cin * c.H * c.W;
  const size_t w_elems = (size_t)c.Cin * c.Cout * c.k * c.k;
  std::vector<float> h_in(in_elems), h_w(w_elems), h_b(c.Cout);
  for (auto& v : h_in) v = bf16r(frand());
  for (auto& v : h_w) v = bf16r(frand() * 0.1f);
  for (auto& v : h_b) v = frand() * 0.5f;

  float *d_in, *d_w, *d_b, *d_out32; double* d_stats;
  void *d_p8, *d_wp, *d_raw;
  CK(cudaMalloc(&d_in, in_elems * 4)); CK(cudaMalloc(&d_w, w_elems * 4)); CK(cudaMalloc(&d_b, c.Cout * 4));
  CK(cudaMemcpy(d_in, h_in.data(), in_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, h_w.data(), w_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, h_b.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  const size_t p8_bytes = nhvr_act_bytes(&in_desc);
  CK(cudaMalloc(&d_p8, p8_bytes)); CK(cudaMemset(d_p8, 0xFF, p8_bytes));   // NaN-poison: halo/slack must not leak
  CK(cudaMalloc(&d_wp, nhvr_conv_weight_bytes(plan)));
  const size_t out_elems = (size_t)c.N * c.Cout * Ho * Wo;
  CK(cudaMalloc(&d_out32, out_elems * 4)); CK(cudaMemset(d_out32, 0, out_elems * 4));
  nhvr_act_desc raw_desc{}; raw_desc.N = c.N; raw_desc.C8 = Cout8; raw_desc.H = Ho; raw_desc.W = Wo;
  CK(cudaMalloc(&d_raw, nhvr_act_bytes(&raw_desc))); CK(cudaMemset(d_raw, 0, nhvr_act_bytes(&raw_desc)));
  CK(cudaMalloc(&d_stats, (size_t)c.N * Cout8 * 8 * 4 * 8)); CK(cudaMemset(d_stats, 0, (size_t)c.N * Cout8 * 8 * 4 * 8));

  const float* srcs[1] = {d_in}; int32_t sc[1] = {c.Cin};
  NK(nhvr_pack_nchw(srcs, sc, 1, d_p8, &in_desc, 0));
  NK(nhvr_conv_pack_weights(plan, d_w, d_wp, 0));
  CK(cudaDeviceSynchronize());

  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 5;
  float ms = 0.f;
  for (int it = 0; it < iters + 1; ++it) {
    if (c.epi == NHVR_EPI_RAW_STATS) CK(cudaMemsetAsync(d_stats, 0, (size_t)c.N * Cout8 * 8 * 4 * 8, 0));
    if (it == 1) CK(cudaEventRecord(e0, 0));
    if (c.epi == NHVR_EPI_RAW_STATS) NK(nhvr_conv_forward(plan, d_p8, d_wp, nullptr, d_raw, nullptr, d_stats, 0));
    else if (c.epi == NHVR_EPI_BIAS_ACT_F32) NK(nhvr_conv_forward(plan, d_p8, d_wp, d_b, d_out32, nullptr, nullptr, 0));
    else NK(nhvr_conv_forward(plan, d_p8, d_wp, d_b, d_raw, &raw_desc, nullptr, 0));
  }
  CK(cudaEventRecord(e1, 0));
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("%-28s KERNEL FAILED: %s\n", c.name, cudaGetErrorString(se)); return 1; }
  CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;

  std::vector<float> h_out(out_elems);
  std::vector<double> h_stats((size_t)c.N * Cout8 * 8 * 4);
  if (c.epi == NHVR_EPI_BIAS_ACT_F32) {
    CK(cudaMemcpy(h_out.data(), d_out32, out_elems * 4, cudaMemcpyDeviceToHost));
  } else {
    NK(nhvr_unpack_nchw(d_raw, &raw_desc, d_out32, c.Cout, 0));
    CK(cudaMemcpy(h_out.data(), d_out32, out_elems * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_stats.data(), d_stats, h_stats.size() * 8, cudaMemcpyDeviceToHost));
  }

  // ---- CPU reference
  std::vector<float> ref(out_elems, 0.f);
  auto IN = [&](int n, int ci, int y, int x) -> float {
    if (y < 0 || y >= c.H || x < 0 || x >= c.W) {
      if (c.halo == NHVR_HALO_ZERO) return 0.f;
      y = refl(y, c.H); x = refl(x, c.W);
    }
    return h_in[(((size_t)n * c.Cin + ci) * c.H + y) * c.W + x];
  };
  if (c.kind == NHVR_CONV) {
#pragma omp parallel for collapse(2)
    for (int n = 0; n < c.N; ++n)
      for (int co = 0; co < c.Cout; ++co)
        for (int y = 0; y < Ho; ++y)
          for (int x = 0; x < Wo; ++x) {
            double acc = 0.0;
            for (int ci = 0; ci < c.Cin; ++ci)
              for (int r = 0; r < c.k; ++r)
                for (int s = 0; s < c.k; ++s)
                  acc += (double)IN(n, ci, y * c.stride + r - c.pad, x * c.stride + s - c.pad) *
                         h_w[(((size_t)co * c.Cin + ci) * c.k + r) * c.k + s];
            ref[(((size_t)n * c.Cout + co) * Ho + y) * Wo + x] = (float)acc;
          }
  } else {
    std::vector<double> accd(out_elems, 0.0);
    for (int n = 0; n < c.N; ++n)
      for (int ci = 0; ci < c.Cin; ++ci)
        for (int i = 0; i < c.H; ++i)
          for (int j = 0; j < c.W; ++j) {
            const float xv = h_in[(((size_t)n * c.Cin + ci) * c.H + i) * c.W + j];
            for (int co = 0; co < c.Cout; ++co)
              for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                  const int Y = 2 * i - 1 + ky, X = 2 * j - 1 + kx;
                  if (Y < 0 || Y >= Ho || X < 0 || X >= Wo) continue;
                  accd[(((size_t)n * c.Cout + co) * Ho + Y) * Wo + X] += (double)xv * h_w[(((size_t)ci * c.Cout + co) * 3 + ky) * 3 + kx];
                }
          }
    for (size_t i = 0; i < out_elems; ++i) ref[i] = (float)accd[i];
  }
  double max_err = 0.0, max_ref = 0.0;
  size_t bad = 0;
  const bool bf16_out = c.epi != NHVR_EPI_BIAS_ACT_F32;
  for (int n = 0; n < c.N; ++n)
    for (int co = 0; co < c.Cout; ++co)
      for (size_t i = 0; i < (size_t)Ho * Wo; ++i) {
        const size_t idx = ((size_t)n * c.Cout + co) * Ho * Wo + i;
        float r = ref[idx];
        if (c.epi != NHVR_EPI_RAW_STATS) {
          r += h_b[co];
          if (c.act == NHVR_ACT_TANH) r = tanhf(r);
          else if (c.act == NHVR_ACT_LRELU02) r = r > 0 ? r : 0.2f * r;
          else if (c.act == NHVR_ACT_RELU) r = r > 0 ? r : 0.f;
        }
        const double err = fabs((double)h_out[idx] - r);
        const double tol = bf16_out ? 1e-2 * fmax(1.0, fabs(r)) : 2e-3 * fmax(1.0, fabs(r));
        if (!(err <= tol)) ++bad;
        if (err > max_err || std::isnan(h_out[idx])) max_err = std::isnan(h_out[idx]) ? 1e30 : err;
        if (fabs(r) > max_ref) max_ref = fabs(r);
      }
  double stat_err = 0.0;
  if (c.epi == NHVR_EPI_RAW_STATS) {
    for (int n = 0; n < c.N; ++n)
      for (int co = 0; co < c.Cout; ++co) {
        double s = 0, ss = 0;
        for (size_t i = 0; i < (size_t)Ho * Wo; ++i) { const double v = ref[((size_t)n * c.Cout + co) * Ho * Wo + i]; s += v; ss += v * v; }
        const double gs = h_stats[((size_t)n * Cout8 * 8 + co) * 4], gss = h_stats[((size_t)n * Cout8 * 8 + co) * 4 + 1];
        const double e1 = fabs(gs - s) / (1.0 + fabs(s)) , e2 = fabs(gss - ss) / (1.0 + fabs(ss));
        stat_err = fmax(stat_err, fmax(e1, e2));
      }
    if (stat_err > 2e-3) ++bad;
  }
  int32_t info[16];
  nhvr_conv_plan_info(plan, info, 16);
  const double gflop = nhvr_conv_flops(plan) * 1e-9;
  printf("%-28s %s max_err=%.3e (max|ref|=%.2f) stat_err=%.1e bad=%zu  %.3f ms %.1f TFLOP/s  [kcp=%d chunks=%d jobs=%d runs=%d acc=%d slab=%d N=%d bpb=%d SA=%d SB=%d tmem=%d smem=%d tiles=%d]\n",
         c.name, bad ? "FAIL" : "ok  ", max_err, max_ref, stat_err, bad, ms, gflop / ms, info[0], info[1], info[2], info[3], info[4],
         info[5], info[6], info[7], info[9], info[10], info[11], info[12], info[13]);
  nhvr_conv_plan_destroy(plan);
  cudaFree(d_in); cudaFree(d_w); cudaFree(d_b); cudaFree(d_out32); cudaFree(d_stats); cudaFree(d_p8); cudaFree(d_wp); cudaFree(d_raw);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (nhvr_arch_ok() != 0) { printf("not an sm_100 device: %s\n", nhvr_strerror(nhvr_arch_ok())); return 98; }
  const int R = NHVR_HALO_REFLECT, Z = NHVR_HALO_ZERO;
  std::vector<Case> cases = {
      {"c3s1 16->16 tiny", NHVR_CONV, 16, 16, 3, 1, 1, 1, 12, 20, R, NHVR_EPI_BIAS_ACT_F32, NHVR_ACT_NONE},
      {"c3s1 16->16 w126 (merged)", NHVR_CONV, 16, 16, 3, 1, 1, 2, 9, 126, R, NHVR_EPI_BIAS_ACT_F32, NHVR_ACT_NONE},
      {"c3s1 32->48 w200 (runs)", NHVR_CONV, 32, 48, 3, 1, 1, 1, 10, 200, Z, NHVR_EPI_BIAS_ACT_F32, NHVR_ACT_TANH},
      {"c3s1 192->192 stats", NHVR_CONV, 192, 192, 3, 1, 1, 2, 24, 40, R, NHVR_EPI_RAW_STATS, 0},
      {"c7s1 6->48 stem", NHVR_CONV, 6, 48, 7, 1, 3, 1, 40, 72, R, NHVR_EPI_RAW_STATS, 0},
      {"c7s1 48->4 head", NHVR_CONV, 48, 4, 7, 1, 3, 1, 40, 72, R, NHVR_EPI_BIAS_ACT_F32, NHVR_ACT_TANH},
      {"c7s1 64->73 uvhead", NHVR_CONV, 64, 73, 7, 1, 3, 1, 24, 40, R, NHVR_EPI_BIAS_ACT_F32, NHVR_ACT_NONE},
      {"c3s2 48->96 down", NHVR_CONV, 48, 96, 3, 2, 1, 2, 32, 48, Z, NHVR_EPI_RAW_STATS, 0},
      {"c3s2 96->192 odd", NHVR_CONV, 96, 192, 3, 2, 1, 1, 17, 23, Z, NHVR_EPI_RAW_STATS, 0},
      {"ct3s2 192->96 up", NHVR_CONV_TRANSPOSE, 192, 96, 3, 2, 1, 2, 16, 24, Z, NHVR_EPI_RAW_STATS, 0},
      {"ct3s2 96->48 up", NHVR_CONV_TRANSPOSE, 96, 48, 3, 2, 1, 1, 20, 36, Z, NHVR_EPI_RAW_STATS, 0},
      {"c4s2p2 6->64 D0 lrelu p8", NHVR_CONV, 6, 64, 4, 2, 2, 1, 33, 41, Z, NHVR_EPI_BIAS_ACT_P8, NHVR_ACT_LRELU02},
      {"c4s2p2 64->128 D1", NHVR_CONV, 64, 128, 4, 2, 2, 1, 17, 21, Z, NHVR_EPI_RAW_STATS, 0},
      {"c4s1p2 256->512 D3 split", NHVR_CONV, 256, 512, 4, 1, 2, 1, 9, 11, Z, NHVR_EPI_RAW_STATS, 0},
      {"c4s1p2 512->1 D4", NHVR_CONV, 512, 1, 4, 1, 2, 1, 10, 12, Z, NHVR_EPI_BIAS_ACT_F32, NHVR_ACT_NONE},
  };
  int only = argc > 1 ? atoi(argv[1]) : -1;
  int fails = 0;
  for (size_t i = 0; i < cases.size(); ++i) {
    if (only >= 0 && (int)i != only) continue;
    fails += run_case(cases[i]);
  }
  // perf-only shapes (no CPU check): the generator bottleneck conv at 128x128
  if (only < 0 || only >= 100) {
    struct Perf { const char* name; Case c; };
    std::vector<Case> perf = {
        {"PERF c3s1 192->192 128^2 b1", NHVR_CONV, 192, 192, 3, 1, 1, 1, 128, 128, R, NHVR_EPI_RAW_STATS, 0},
        {"PERF c3s1 192->192 128^2 b8", NHVR_CONV, 192, 192, 3, 1, 1, 8, 128, 128, R, NHVR_EPI_RAW_STATS, 0},
        {"PERF c3s1 256->256 128^2 b8", NHVR_CONV, 256, 256, 3, 1, 1, 8, 128, 128, R, NHVR_EPI_RAW_STATS, 0},
        {"PERF c7s1 64->73 512^2 b1", NHVR_CONV, 64, 73, 7, 1, 3, 1, 512, 512, R, NHVR_EPI_BIAS_ACT_F32, 0},
        {"PERF c7s1 16->48 512^2 b1", NHVR_CONV, 16, 48, 7, 1, 3, 1, 512, 512, R, NHVR_EPI_RAW_STATS, 0},
        {"PERF c7s1 48->4 512^2 b1", NHVR_CONV, 48, 4, 7, 1, 3, 1, 512, 512, R, NHVR_EPI_BIAS_ACT_F32, NHVR_ACT_TANH_SIGMOID_LAST},
        {"PERF c3s2 48->96 512^2 b1", NHVR_CONV, 48, 96, 3, 2, 1, 1, 512, 512, Z, NHVR_EPI_RAW_STATS, 0},
        {"PERF ct 96->48 256^2 b1", NHVR_CONV_TRANSPOSE, 96, 48, 3, 2, 1, 1, 256, 256, Z, NHVR_EPI_RAW_STATS, 0},
        {"PERF c7s1 64->73 512^2 b8", NHVR_CONV, 64, 73, 7, 1, 3, 8, 512, 512, R, NHVR_EPI_BIAS_ACT_F32, 0},
        {"PERF c7s1 16->64 512^2 b8", NHVR_CONV, 16, 64, 7, 1, 3, 8, 512, 512, R, NHVR_EPI_RAW_STATS, 0},
        {"PERF c3s2 64->128 512^2 b8", NHVR_CONV, 64, 128, 3, 2, 1, 8, 512, 512, Z, NHVR_EPI_RAW_STATS, 0},
        {"PERF c3s2 128->256 256^2 b8", NHVR_CONV, 128, 256, 3, 2, 1, 8, 256, 256, Z, NHVR_EPI_RAW_STATS, 0},
        {"PERF ct 256->128 128^2 b8", NHVR_CONV_TRANSPOSE, 256, 128, 3, 2, 1, 8, 128, 128, Z, NHVR_EPI_RAW_STATS, 0},
        {"PERF ct 128->64 256^2 b8", NHVR_CONV_TRANSPOSE, 128, 64, 3, 2, 1, 8, 256, 256, Z, NHVR_EPI_RAW_STATS, 0},
    };
    int pidx = -1;
    for (auto& c : perf) {
      ++pidx;
      if (only > 100 && pidx != only - 101) continue;   // 101 + i selects perf case i only
      // run without CPU reference: reuse run_case machinery would be too slow; time only
      nhvr_conv_desc d{}; d.kind = c.kind; d.Cin = c.Cin; d.Cout = c.Cout; d.kh = d.kw = c.k; d.stride = c.stride; d.pad = c.pad;
      d.N = c.N; d.H = c.H; d.W = c.W; d.halo = c.halo; d.epilogue = c.epi; d.act = c.act;
      nhvr_conv_plan* plan = nullptr;
      if (nhvr_conv_plan_create(&d, &plan) != 0) { printf("%s plan failed\n", c.name); ++fails; continue; }
      nhvr_act_desc in_desc; nhvr_conv_input_desc(plan, &in_desc);
      int Ho, Wo, Cout8; nhvr_conv_output_dims(plan, &Ho, &Wo, &Cout8);
      void *d_p8, *d_wp, *d_out; double* d_stats;
      CK(cudaMalloc(&d_p8, nhvr_act_bytes(&in_desc))); CK(cudaMemset(d_p8, 0, nhvr_act_bytes(&in_desc)));
      CK(cudaMalloc(&d_wp, nhvr_conv_weight_bytes(plan))); CK(cudaMemset(d_wp, 0, nhvr_conv_weight_bytes(plan)));
      const size_t ob = (size_t)c.N * Cout8 * 8 * Ho * Wo * 4 + (1 << 20);
      CK(cudaMalloc(&d_out, ob)); CK(cudaMalloc(&d_stats, (size_t)c.N * Cout8 * 32 * 8)); CK(cudaMemset(d_stats, 0, (size_t)c.N * Cout8 * 32 * 8));
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      const int iters = (only > 100) ? 3 : 20;
      for (int it = 0; it < iters + 3; ++it) {
        if (it == 3) CK(cudaEventRecord(e0, 0));
        int s = nhvr_conv_forward(plan, d_p8, d_wp, nullptr, d_out, nullptr, d_stats, 0);
        if (s) { printf("%s launch failed %d %s\n", c.name, s, nhvr_last_cuda_error()); break; }
      }
      CK(cudaEventRecord(e1, 0));
      cudaError_t se = cudaDeviceSynchronize();
      float ms = 0; if (se == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
      ms /= iters;
      int32_t info[16];
      nhvr_conv_plan_info(plan, info, 16);
      printf("%-30s %s %.4f ms  %.1f TFLOP/s  [kcp=%d chunks=%d jobs=%d runs=%d slab=%d N=%d bpb=%d SA=%d SB=%d smem=%d tiles=%d]\n", c.name,
             se == cudaSuccess ? "" : cudaGetErrorString(se), ms, nhvr_conv_flops(plan) * 1e-9 / ms, info[0], info[1], info[2], info[3], info[5],
             info[6], info[7], info[9], info[10], info[12], info[13]);
      if (se != cudaSuccess) { ++fails; break; }
      nhvr_conv_plan_destroy(plan);
      cudaFree(d_p8); cudaFree(d_wp); cudaFree(d_out); cudaFree(d_stats);
    }
  }
  printf("selftest: %d failure(s)\n", fails);
  return fails;
}
