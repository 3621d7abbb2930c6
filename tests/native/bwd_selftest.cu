// Native self-test of the backward conv kernels through the C-ABI: weight gradient (tcgen05, MN-major from
// P8) for stride-1 / stride-2 / transposed convs, and the stride-1 input gradient (NHVR_CONV_DGRAD_S1) fed
// from the SAME gradient buffer.  Reference: scalar CPU loops on the same bf16-rounded operands
// (nn.Conv2d / nn.ConvTranspose2d gradient definitions).  Exit code = number of failures.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../include/nhvr.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(99); } } while (0)
#define NK(x) do { int s_ = (x); if (s_ != 0) { printf("nhvr error %d (%s / %s) at %s:%d\n", s_, nhvr_strerror(s_), nhvr_last_cuda_error(), __FILE__, __LINE__); return 1; } } while (0)

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }
static uint32_t rng_state = 777u;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f; }
static int refl(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * (n - 1) - i; return i; }

struct Case { const char* name; int kind, Cin, Cout, k, stride, pad, N, H, W, halo; };

static int run_case(const Case& c) {
  nhvr_conv_desc d{};
  d.kind = c.kind; d.Cin = c.Cin; d.Cout = c.Cout; d.kh = d.kw = c.k; d.stride = c.stride; d.pad = c.pad;
  d.N = c.N; d.H = c.H; d.W = c.W; d.halo = c.halo; d.epilogue = NHVR_EPI_RAW_P8;
  nhvr_conv_plan* fplan = nullptr;
  NK(nhvr_conv_plan_create(&d, &fplan));
  nhvr_act_desc x_desc; NK(nhvr_conv_input_desc(fplan, &x_desc));
  int Ho, Wo, Cout8; NK(nhvr_conv_output_dims(fplan, &Ho, &Wo, &Cout8));
  nhvr_wgrad_plan* wplan = nullptr;
  NK(nhvr_wgrad_plan_create(&d, &wplan));
  nhvr_act_desc g_desc; NK(nhvr_wgrad_grad_desc(wplan, &g_desc));

  const size_t xn = (size_t)c.N * c.Cin * c.H * c.W, gn = (size_t)c.N * c.Cout * Ho * Wo, wn = (size_t)c.Cin * c.Cout * c.k * c.k;
  std::vector<float> hx(xn), hg(gn), hw(wn);
  for (auto& v : hx) v = bf16r(frand());
  for (auto& v : hg) v = bf16r(frand() * 0.25f);
  for (auto& v : hw) v = bf16r(frand() * 0.1f);
  float *dx, *dg, *dw, *dwg; void *px, *pg, *ws;
  CK(cudaMalloc(&dx, xn * 4)); CK(cudaMalloc(&dg, gn * 4)); CK(cudaMalloc(&dw, wn * 4)); CK(cudaMalloc(&dwg, wn * 4));
  CK(cudaMemcpy(dx, hx.data(), xn * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dg, hg.data(), gn * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), wn * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&px, nhvr_act_bytes(&x_desc))); CK(cudaMemset(px, 0, nhvr_act_bytes(&x_desc)));
  CK(cudaMalloc(&pg, nhvr_act_bytes(&g_desc))); CK(cudaMemset(pg, 0xFF, nhvr_act_bytes(&g_desc)));   // poison: pack must zero the halo
  CK(cudaMalloc(&ws, nhvr_wgrad_workspace_bytes(wplan)));
  const float* sx[1] = {dx}; int32_t cx[1] = {c.Cin};
  const float* sg[1] = {dg}; int32_t cg[1] = {c.Cout};
  NK(nhvr_pack_nchw(sx, cx, 1, px, &x_desc, 0));
  NK(nhvr_pack_nchw(sg, cg, 1, pg, &g_desc, 0));
  // poison everything after the last plane's data except the documented slack (zero): emulate torch.zeros alloc
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  NK(nhvr_wgrad(wplan, px, pg, ws, dwg, 1.0f, 0, 0));
  CK(cudaEventRecord(e0, 0));
  for (int it = 0; it < 5; ++it) NK(nhvr_wgrad(wplan, px, pg, ws, dwg, 1.0f, 0, 0));
  CK(cudaEventRecord(e1, 0));
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("%-26s WGRAD KERNEL FAILED: %s\n", c.name, cudaGetErrorString(se)); return 1; }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
  std::vector<float> got(wn); CK(cudaMemcpy(got.data(), dwg, wn * 4, cudaMemcpyDeviceToHost));

  // ---- CPU wgrad
  std::vector<double> ref(wn, 0.0);
  auto X = [&](int n, int ci, int y, int x) -> float {
    if (y < 0 || y >= c.H || x < 0 || x >= c.W) {
      if (c.halo == NHVR_HALO_ZERO) return 0.f;
      y = refl(y, c.H); x = refl(x, c.W);
    }
    return hx[(((size_t)n * c.Cin + ci) * c.H + y) * c.W + x];
  };
  if (c.kind == NHVR_CONV) {
#pragma omp parallel for collapse(2)
    for (int co = 0; co < c.Cout; ++co)
      for (int ci = 0; ci < c.Cin; ++ci)
        for (int r = 0; r < c.k; ++r)
          for (int s = 0; s < c.k; ++s) {
            double a = 0;
            for (int n = 0; n < c.N; ++n)
              for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x)
                  a += (double)hg[(((size_t)n * c.Cout + co) * Ho + y) * Wo + x] * X(n, ci, y * c.stride + r - c.pad, x * c.stride + s - c.pad);
            ref[(((size_t)co * c.Cin + ci) * c.k + r) * c.k + s] = a;
          }
  } else {
#pragma omp parallel for collapse(2)
    for (int ci = 0; ci < c.Cin; ++ci)
      for (int co = 0; co < c.Cout; ++co)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            double a = 0;
            for (int n = 0; n < c.N; ++n)
              for (int i = 0; i < c.H; ++i)
                for (int j = 0; j < c.W; ++j) {
                  const int Y = 2 * i - 1 + ky, Xc = 2 * j - 1 + kx;
                  if (Y < 0 || Y >= Ho || Xc < 0 || Xc >= Wo) continue;
                  a += (double)hx[(((size_t)n * c.Cin + ci) * c.H + i) * c.W + j] * hg[(((size_t)n * c.Cout + co) * Ho + Y) * Wo + Xc];
                }
            ref[(((size_t)ci * c.Cout + co) * 3 + ky) * 3 + kx] = a;
          }
  }
  double werr = 0, wmax = 0; size_t bad = 0;
  for (size_t i = 0; i < wn; ++i) { wmax = fmax(wmax, fabs(ref[i])); }
  for (size_t i = 0; i < wn; ++i) {
    const double e = fabs(got[i] - ref[i]);
    if (!(e <= 2e-3 * wmax + 1e-4)) ++bad;
    if (e > werr || std::isnan(got[i])) werr = std::isnan(got[i]) ? 1e30 : e;
  }

  if (getenv("BWD_VERBOSE")) {
    const int kk = c.k * c.k;
    for (int tap = 0; tap < kk; ++tap) {
      double te = 0, tr = 0;
      for (size_t i = tap; i < wn; i += kk) { te = fmax(te, fabs(got[i] - ref[i])); tr = fmax(tr, fabs(ref[i])); }
      printf("   tap %d (ky=%d kx=%d): max_err %.3e  max|ref| %.2f   got[0]=%.4f ref[0]=%.4f\n", tap, tap / c.k, tap % c.k, te, tr, got[tap], ref[tap]);
    }
  }
  // ---- dgrad of a stride-1 conv from the same gradient buffer
  double derr = 0, dmax = 0; size_t dbad = 0;
  if (c.kind == NHVR_CONV && c.stride == 1) {
    nhvr_conv_desc dd = d; dd.kind = NHVR_CONV_DGRAD_S1; dd.epilogue = NHVR_EPI_RAW_P8;
    nhvr_conv_plan* dplan = nullptr;
    NK(nhvr_conv_plan_create(&dd, &dplan));
    NK(nhvr_conv_plan_set_input_desc(dplan, &g_desc));
    int Hd, Wd, C8d; NK(nhvr_conv_output_dims(dplan, &Hd, &Wd, &C8d));
    void *wp, *pout; float* dout;
    CK(cudaMalloc(&wp, nhvr_conv_weight_bytes(dplan)));
    NK(nhvr_conv_pack_weights(dplan, dw, wp, 0));
    nhvr_act_desc od{}; od.N = c.N; od.C8 = C8d; od.H = Hd; od.W = Wd;
    CK(cudaMalloc(&pout, nhvr_act_bytes(&od))); CK(cudaMemset(pout, 0, nhvr_act_bytes(&od)));
    const size_t on = (size_t)c.N * c.Cin * Hd * Wd;
    CK(cudaMalloc(&dout, on * 4));
    NK(nhvr_conv_forward(dplan, pg, wp, nullptr, pout, nullptr, nullptr, 0));
    NK(nhvr_unpack_nchw(pout, &od, dout, c.Cin, 0));
    se = cudaDeviceSynchronize();
    if (se != cudaSuccess) { printf("%-26s DGRAD KERNEL FAILED: %s\n", c.name, cudaGetErrorString(se)); return 1; }
    std::vector<float> gd(on); CK(cudaMemcpy(gd.data(), dout, on * 4, cudaMemcpyDeviceToHost));
    if (Hd != c.H + 2 * c.pad || Wd != c.W + 2 * c.pad) { printf("dgrad dims wrong\n"); ++dbad; }
#pragma omp parallel for collapse(2) reduction(max : derr, dmax) reduction(+ : dbad)
    for (int n = 0; n < c.N; ++n)
      for (int ci = 0; ci < c.Cin; ++ci)
        for (int u = 0; u < Hd; ++u)
          for (int v = 0; v < Wd; ++v) {
            double a = 0;
            for (int co = 0; co < c.Cout; ++co)
              for (int r = 0; r < c.k; ++r)
                for (int s = 0; s < c.k; ++s) {
                  const int y = u - r, x = v - s;
                  if (y < 0 || y >= Ho || x < 0 || x >= Wo) continue;
                  a += (double)hg[(((size_t)n * c.Cout + co) * Ho + y) * Wo + x] * hw[(((size_t)co * c.Cin + ci) * c.k + r) * c.k + s];
                }
            const double e = fabs(gd[(((size_t)n * c.Cin + ci) * Hd + u) * Wd + v] - a);
            if (!(e <= 1e-2 * fmax(1.0, fabs(a)))) ++dbad;
            derr = fmax(derr, e); dmax = fmax(dmax, fabs(a));
          }
    nhvr_conv_plan_destroy(dplan); cudaFree(wp); cudaFree(pout); cudaFree(dout);
  }
  const double gflop = 2.0 * c.k * c.k * c.Cin * c.Cout * (c.kind == NHVR_CONV_TRANSPOSE ? (double)c.H * c.W : (double)Ho * Wo) * c.N * 1e-9;
  printf("%-26s %s wgrad max_err=%.3e (max|ref|=%.2f) bad=%zu  %.3f ms %.1f TFLOP/s | dgrad max_err=%.3e (max|ref|=%.2f) bad=%zu\n", c.name,
         (bad || dbad) ? "FAIL" : "ok  ", werr, wmax, bad, ms, gflop / ms, derr, dmax, dbad);
  nhvr_conv_plan_destroy(fplan); nhvr_wgrad_plan_destroy(wplan);
  cudaFree(dx); cudaFree(dg); cudaFree(dw); cudaFree(dwg); cudaFree(px); cudaFree(pg); cudaFree(ws);
  return (bad || dbad) ? 1 : 0;
}

int main(int argc, char** argv) {
  if (nhvr_arch_ok() != 0) { printf("not an sm_100 device\n"); return 98; }
  const int R = NHVR_HALO_REFLECT, Z = NHVR_HALO_ZERO;
  std::vector<Case> cases = {
      {"c3s1 16->16 tiny", NHVR_CONV, 16, 16, 3, 1, 1, 1, 12, 20, R},
      {"c3s1 32->48 w200", NHVR_CONV, 32, 48, 3, 1, 1, 2, 10, 200, Z},
      {"c3s1 192->192", NHVR_CONV, 192, 192, 3, 1, 1, 2, 24, 40, R},
      {"c7s1 6->48 stem", NHVR_CONV, 6, 48, 7, 1, 3, 1, 40, 72, R},
      {"c7s1 9->48 stem 2 planes", NHVR_CONV, 9, 48, 7, 1, 3, 2, 33, 150, R},
      {"c5s1 3->20 fold k5", NHVR_CONV, 3, 20, 5, 1, 2, 1, 30, 41, Z},
      {"c7s1 48->4 head", NHVR_CONV, 48, 4, 7, 1, 3, 1, 40, 72, R},
      {"c7s1 64->73 uvhead", NHVR_CONV, 64, 73, 7, 1, 3, 1, 24, 40, R},
      {"c3s2 48->96 down", NHVR_CONV, 48, 96, 3, 2, 1, 2, 32, 48, Z},
      {"c3s2 96->192", NHVR_CONV, 96, 192, 3, 2, 1, 1, 16, 24, Z},
      {"ct3s2 192->96 up", NHVR_CONV_TRANSPOSE, 192, 96, 3, 2, 1, 2, 16, 24, Z},
      {"ct3s2 96->48 up", NHVR_CONV_TRANSPOSE, 96, 48, 3, 2, 1, 1, 20, 36, Z},
      {"c4s1p2 64->128 D", NHVR_CONV, 64, 128, 4, 1, 2, 1, 9, 11, Z},
      {"c4s2p2 16->64 D", NHVR_CONV, 16, 64, 4, 2, 2, 1, 18, 22, Z},
  };
  int only = argc > 1 ? atoi(argv[1]) : -1, fails = 0;
  for (size_t i = 0; i < cases.size(); ++i) {
    if (only >= 0 && (int)i != only) continue;
    fails += run_case(cases[i]);
  }
  if (only < 0 || only >= 100) {
    // wgrad timing at training shapes (no CPU check)
    std::vector<Case> perf = {
        {"PERF wgrad c3 256->256 64^2 b16", NHVR_CONV, 256, 256, 3, 1, 1, 16, 64, 64, R},
        {"PERF wgrad c3 192->192 128^2 b8", NHVR_CONV, 192, 192, 3, 1, 1, 8, 128, 128, R},
        {"PERF wgrad c7 64->73 256^2 b16", NHVR_CONV, 64, 73, 7, 1, 3, 16, 256, 256, R},
        {"PERF wgrad c3s2 64->128 256^2 b16", NHVR_CONV, 64, 128, 3, 2, 1, 16, 256, 256, Z},
        {"PERF wgrad ct 128->64 128^2 b16", NHVR_CONV_TRANSPOSE, 128, 64, 3, 2, 1, 16, 128, 128, Z},
        {"PERF wgrad c7 3->64 512^2 b8 (stem)", NHVR_CONV, 3, 64, 7, 1, 3, 8, 512, 512, R},
        {"PERF wgrad c7 9->48 512^2 b8 (stem)", NHVR_CONV, 9, 48, 7, 1, 3, 8, 512, 512, R},
        {"PERF wgrad c7 48->4 512^2 b8 (head)", NHVR_CONV, 48, 4, 7, 1, 3, 8, 512, 512, R},
    };
    int pi = 100;
    for (auto& c : perf) {
      ++pi;
      if (only > 100 && only != pi) continue;
      nhvr_conv_desc d{};
      d.kind = c.kind; d.Cin = c.Cin; d.Cout = c.Cout; d.kh = d.kw = c.k; d.stride = c.stride; d.pad = c.pad;
      d.N = c.N; d.H = c.H; d.W = c.W; d.halo = c.halo; d.epilogue = NHVR_EPI_RAW_P8;
      nhvr_conv_plan* fplan = nullptr; nhvr_wgrad_plan* wplan = nullptr;
      if (nhvr_conv_plan_create(&d, &fplan) || nhvr_wgrad_plan_create(&d, &wplan)) { printf("%s plan failed\n", c.name); ++fails; continue; }
      nhvr_act_desc x_desc, g_desc; nhvr_conv_input_desc(fplan, &x_desc); nhvr_wgrad_grad_desc(wplan, &g_desc);
      void *px, *pg, *ws; float* dwg;
      CK(cudaMalloc(&px, nhvr_act_bytes(&x_desc))); CK(cudaMemset(px, 0, nhvr_act_bytes(&x_desc)));
      CK(cudaMalloc(&pg, nhvr_act_bytes(&g_desc))); CK(cudaMemset(pg, 0, nhvr_act_bytes(&g_desc)));
      CK(cudaMalloc(&ws, nhvr_wgrad_workspace_bytes(wplan)));
      CK(cudaMalloc(&dwg, (size_t)c.Cin * c.Cout * c.k * c.k * 4));
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      const int iters = (only > 100) ? 2 : 10;
      for (int it = 0; it < iters + 2; ++it) {
        if (it == 2) CK(cudaEventRecord(e0, 0));
        if (nhvr_wgrad(wplan, px, pg, ws, dwg, 1.0f, 0, 0)) { printf("launch failed\n"); break; }
      }
      CK(cudaEventRecord(e1, 0));
      cudaError_t se = cudaDeviceSynchronize();
      float ms = 0; if (se == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
      ms /= iters;
      int Ho, Wo, C8; nhvr_conv_output_dims(fplan, &Ho, &Wo, &C8);
      const double gflop = 2.0 * c.k * c.k * c.Cin * c.Cout * (c.kind == NHVR_CONV_TRANSPOSE ? (double)c.H * c.W : (double)Ho * Wo) * c.N * 1e-9;
      printf("%-36s %s %.4f ms  %.1f TFLOP/s\n", c.name, se == cudaSuccess ? "" : cudaGetErrorString(se), ms, gflop / ms);
      nhvr_conv_plan_destroy(fplan); nhvr_wgrad_plan_destroy(wplan);
      cudaFree(px); cudaFree(pg); cudaFree(ws); cudaFree(dwg);
    }
  }
  printf("bwd selftest: %d failure(s)\n", fails);
  return fails;
}
