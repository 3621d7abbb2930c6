// Weight gradient of every conv lowering, on tcgen05 tensor cores, straight from P8 activations.
//
//   dW_t[co][ci] = sum_n sum_q  G[n][co][q + g_off_t] * X[n][ci][q + x_off_t]          (t = filter tap)
//
// q runs over the same linearised output positions as the forward shift-GEMM, x_off_t is the forward
// tap shift and G is the output gradient stored in the dgrad conv's input format (zero wherever the forward
// produced no output), so both offsets are constants.  This is a GEMM with K = positions: a P8 slab
// [plane][position][8 channels] *is* the un-swizzled MN-major core-matrix layout of tcgen05 (8 channels
// contiguous in 16 B along M/N, 8 positions 16 B apart along K, LBO = 128 B between position groups,
// SBO = plane stride), so A = G^T and B = X^T are fed from the staged slabs with no transpose and the taps
// are, again, only descriptor start-address shifts.
//
// One CTA: 128 output channels (M) x `nci` input channels (N) x a group of taps (one TMEM accumulator per
// tap), summing over its share of the (image, 128-position chunk) list; partial sums are added to a packed
// fp32 workspace [tap][Cout_pad][Cin_pad] with vector reductions.  A small finalise kernel scales and adds
// them into the parameter-layout gradient.
#include "conv_plan.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace nhvr {

extern void note_cuda_error(cudaError_t e);
extern void count_launch();
extern int arch_ok_cached();
extern int operand_f16();

constexpr int kWgMaxTaps = 52;
constexpr int kWgThreads = 256;   // warps 0,2,3 producers (2 also TMEM alloc), warp1 MMA, warps 4-7 epilogue

struct WgTap {
  int32_t x_off;    // shift (units) inside the X plane slab
  int32_t g_off;    // shift (units) inside the G plane slab
  int32_t tap;      // filter tap index r*kw+s
};

struct WgParams {
  const uint4* x;
  const uint4* g;
  float* ws;                       // [ntaps_total][CoutP][CinP] fp32
  int64_t x_plane_units, g_plane_units;
  int32_t C8x, C8g;                // planes per image of X / G
  int32_t CoutP, CinP;             // padded channel counts of the workspace
  int32_t nci, nci8;               // input channels (planes) per CTA
  int32_t ntaps_total, ntaps_grp;  // taps, taps per group (TMEM accumulators)
  int32_t nxruns, ngruns;
  int32_t xslab_units, gslab_units;
  int32_t tiles_per_img, nchunks_total, chunks_per_cta;
  int32_t n_ci_blocks;
  int32_t S;                       // stages
  int32_t tmem_cols;
  int32_t f16;
  // Tap folding (stride-1 convs with <= 16 input channels, the 7x7 stems): N-group j of the B descriptor is the SAME
  // 8-channel plane shifted by j positions (SBO = 16 bytes, overlapping reads), so ONE MMA of N = 64 covers the kw
  // filter columns of a filter row; taps[].tap = r * 16 + plane.
  int32_t fold, kw;
  // Row folding on the M side (<= 8 output channels, the RGB+mask head): M group j of A is gradient plane 0 shifted UP
  // by j rows (a separate bulk copy per group), so ONE MMA per X plane covers all kh x kw taps: D[(r,co)][(s,ci)].
  int32_t mfold, kh, g_row_units;
  ConvRun xruns[kMaxRuns];
  ConvRun gruns[4];
  WgTap taps[kWgMaxTaps];
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform role index
  const int split = blockIdx.x;
  const int co_blk = blockIdx.y / P.n_ci_blocks, ci_blk = blockIdx.y % P.n_ci_blocks;
  const int grp = blockIdx.z;
  const int tap0 = grp * P.ntaps_grp;
  const int ntaps = min(P.ntaps_grp, P.ntaps_total - tap0);

  const uint32_t g_stage_bytes = 16u * (uint32_t)P.gslab_units * 16u;            // 16 planes = 128 output channels
  const uint32_t x_stage_bytes = (uint32_t)P.nci8 * (uint32_t)P.xslab_units * 16u;
  const uint32_t stage_bytes = g_stage_bytes + x_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + P.S;
  uint64_t* acc_full = empty + P.S;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int c_begin = split * P.chunks_per_cta;
  const int c_end = min(c_begin + P.chunks_per_cta, P.nchunks_total);
  const int g_planes = P.mfold ? P.kh : min(16, P.C8g - co_blk * 16);      // planes that exist; the rest of the slab stays stale
                                                           // (rows of D that are never written out)
  if (warp == 0 || warp == 2 || warp == 3) {
    // ------------------------------------------------------------------ producers: G and X slabs of a chunk
    // The slabs are many small plane runs (2-6 KB each); one thread issuing them all is latency-bound on the
    // per-copy address arithmetic, so three warps (0, 2 after its TMEM allocation, 3) each issue every third copy.
    // Warp 0 alone arms the transaction count; copies of the other warps may complete before that (the
    // transaction count is signed within a phase, the pending arrival keeps the phase open).
    const int pid = (warp == 0) ? 0 : warp - 1;          // 0, 1, 2
    int st = 0;
    uint32_t ph = 0;
    const uint32_t bytes = (uint32_t)g_planes * P.gslab_units * 16u + x_stage_bytes;
    const int ncopy_g = g_planes * P.ngruns, ncopy_x = P.nci8 * P.nxruns;
    for (int c = c_begin; c < c_end; ++c) {
      const int n = c / P.tiles_per_img;
      const int64_t q0 = (int64_t)(c - n * P.tiles_per_img) * kTileM;
      mbar_wait(&empty[st], ph ^ 1u);
      if (elect_one()) {
        if (pid == 0) mbar_arrive_expect_tx(&full[st], bytes);
        uint8_t* gdst = smem + (size_t)st * stage_bytes;
        uint8_t* xdst = gdst + g_stage_bytes;
        const uint4* gimg = P.g + ((int64_t)n * P.C8g + co_blk * 16) * P.g_plane_units + q0;
        const uint4* ximg = P.x + ((int64_t)n * P.C8x + ci_blk * P.nci8) * P.x_plane_units + q0;
        for (int i = pid; i < ncopy_g; i += 3) {
          const int pl = i / P.ngruns, r = i - pl * P.ngruns;
          const int64_t poff = P.mfold ? -(int64_t)pl * P.g_row_units : (int64_t)pl * P.g_plane_units;
          bulk_g2s(gdst + ((size_t)pl * P.gslab_units + P.gruns[r].s_off) * 16, gimg + poff + P.gruns[r].g_off,
                   (uint32_t)P.gruns[r].len * 16u, &full[st]);
        }
        for (int i = pid; i < ncopy_x; i += 3) {
          const int pl = i / P.nxruns, r = i - pl * P.nxruns;
          bulk_g2s(xdst + ((size_t)pl * P.xslab_units + P.xruns[r].s_off) * 16, ximg + (int64_t)pl * P.x_plane_units + P.xruns[r].g_off,
                   (uint32_t)P.xruns[r].len * 16u, &full[st]);
        }
      }
      __syncwarp();
      if (++st == P.S) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // instruction descriptor: D f32, A/B 16-bit, BOTH MN-major (bits 15, 16), N = nci, M = 128
    const int Nf = P.fold ? 64 : P.nci;                      // accumulator columns per tap
    const uint32_t idesc = make_idesc_16(kTileM, (uint32_t)Nf, P.f16) | (1u << 15) | (1u << 16);
    // MN-major un-swizzled descriptor: LBO = 128 B (next group of 8 positions along K), SBO = plane stride
    const uint32_t g_hi = (((uint32_t)P.gslab_units) & 0x3FFFu) | (1u << 14);
    const uint32_t x_hi = ((P.fold ? 1u : (uint32_t)P.xslab_units) & 0x3FFFu) | (1u << 14);   // fold: next N group = next position
    const uint32_t lbo_field = (128u >> 4) << 16;
    const bool leader = elect_one();
    int st = 0;
    uint32_t ph = 0;
    bool first_chunk = true;
    for (int c = c_begin; c < c_end; ++c) {
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t g_lo0 = ((smem_u32(smem + (size_t)st * stage_bytes) & 0x3FFFFu) >> 4) | lbo_field;
      const uint32_t x_lo0 = (((smem_u32(smem + (size_t)st * stage_bytes) + g_stage_bytes) & 0x3FFFFu) >> 4) | lbo_field;
      for (int t = 0; t < ntaps; ++t) {
        const WgTap tp = P.taps[tap0 + t];
        const uint32_t d_tmem = tmem_base + (uint32_t)(t * Nf);
#pragma unroll
        for (int k = 0; k < kTileM / 16; ++k) {            // 16 positions (K) per MMA
          const uint64_t adesc = ((uint64_t)g_hi << 32) | (g_lo0 + (uint32_t)tp.g_off + (uint32_t)(k * 16));
          const uint64_t bdesc = ((uint64_t)x_hi << 32) | (x_lo0 + (uint32_t)tp.x_off + (uint32_t)(k * 16));
          if (leader) umma_bf16(d_tmem, adesc, bdesc, idesc, (first_chunk && k == 0) ? 0u : 1u);
        }
      }
      if (leader) umma_commit(&empty[st]);
      first_chunk = false;
      if (++st == P.S) { st = 0; ph ^= 1u; }
    }
    if (leader) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: D -> fp32 workspace (vector red)
    const int we = warp & 3;
    const int m = we * 32 + lane;
    const int co = co_blk * kTileM + m;
    const uint32_t t_lane = tmem_base + ((uint32_t)(we * 32) << 16);
    mbar_wait_warp(acc_full, 0);
    tc_fence_after();
    if (c_end > c_begin) {
      for (int t = 0; t < ntaps; ++t) {
        if (P.fold) {
          // columns n = s * 8 + c of fold tap (r, plane): workspace entry [tap r*kw+s][co][plane*8 + c]
          const int code = P.taps[tap0 + t].tap, pl = code & 15;
          const int r = P.mfold ? (m >> 3) : (code >> 4);                 // mfold: D row m = r * 8 + co
          const int co = P.mfold ? (m & 7) : co_blk * kTileM + m;
          const bool row_ok = !P.mfold || r < P.kh;          // (the TMEM loads below are warp-collective: no early exit)
          for (int gcol = 0; gcol < 64; gcol += 16) {
            uint32_t vr[16];
            tmem_ld16(t_lane + (uint32_t)(t * 64 + gcol), vr);
            tmem_ld_wait();
            if (co < P.CoutP && row_ok) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int sft = (gcol >> 3) + h;
                if (sft < P.kw) {
                  float* dst = P.ws + ((int64_t)(r * P.kw + sft) * P.CoutP + co) * P.CinP + pl * 8;
#pragma unroll
                  for (int i = 0; i < 8; i += 4) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(vr[h * 8 + i])),
                                 "f"(__uint_as_float(vr[h * 8 + i + 1])), "f"(__uint_as_float(vr[h * 8 + i + 2])),
                                 "f"(__uint_as_float(vr[h * 8 + i + 3]))
                                 : "memory");
                  }
                }
              }
            }
          }
          continue;
        }
        float* dst = P.ws + ((int64_t)P.taps[tap0 + t].tap * P.CoutP + co) * P.CinP + ci_blk * P.nci;   // filter-tap order
        for (int gcol = 0; gcol < P.nci; gcol += 16) {
          uint32_t vr[16];
          tmem_ld16(t_lane + (uint32_t)(t * P.nci + gcol), vr);
          tmem_ld_wait();
          if (co < P.CoutP) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + gcol + i), "f"(__uint_as_float(vr[i])),
                           "f"(__uint_as_float(vr[i + 1])), "f"(__uint_as_float(vr[i + 2])), "f"(__uint_as_float(vr[i + 3]))
                           : "memory");
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// workspace [tap][CoutP][CinP] -> dw (parameter layout) : dw (+)= scale * ws
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin,
                                                             int CoutP, int CinP, int kk, int transposed, float scale, int accumulate) {
  const int64_t total = (int64_t)Cout * Cin * kk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kk);
    const int64_t t = i / kk;
    int co, ci;
    if (transposed) { co = (int)(t % Cout); ci = (int)(t / Cout); }     // [Cin][Cout][kh][kw]
    else            { ci = (int)(t % Cin);  co = (int)(t / Cin); }      // [Cout][Cin][kh][kw]
    const float v = scale * ws[((int64_t)tap * CoutP + co) * CinP + ci];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

}  // namespace nhvr

using namespace nhvr;

struct nhvr_wgrad_plan {
  nhvr_conv_desc d;
  nhvr_act_desc x_desc, g_desc;
  WgParams wp;
  int32_t nsplit_k, n_co_blocks, n_tap_groups;
  size_t smem_bytes, ws_bytes;
};

static inline int wg_round_up(int a, int b) { return (a + b - 1) / b * b; }

extern "C" int nhvr_wgrad_plan_create(const nhvr_conv_desc* fwd, nhvr_wgrad_plan** out) {
  if (!fwd || !out) return NHVR_ERR_NULL;
  if (fwd->kind != NHVR_CONV && fwd->kind != NHVR_CONV_TRANSPOSE) return NHVR_ERR_UNSUPPORTED;
  // the forward plan gives the X slab program (runs + tap shifts) and the q-space
  nhvr_conv_desc fd = *fwd;
  fd.epilogue = NHVR_EPI_RAW_P8;
  fd.in_extra_rows = fd.in_extra_cols = 0;
  fd.flags |= 1;            // the tap program of the plain lowering is what wgrad mirrors
  nhvr_conv_plan* fp = nullptr;
  int st = nhvr_conv_plan_create(&fd, &fp);
  if (st != NHVR_OK) return st;
  nhvr_wgrad_plan* p = new nhvr_wgrad_plan();
  std::memset(p, 0, sizeof(*p));
  p->d = *fwd;
  p->x_desc = fp->in_desc;
  WgParams& W = p->wp;
  const ConvKParams& K = fp->kp;
  const ActGeom xg = make_geom(fp->in_desc);

  // ---- gradient format = input format of the matching dgrad conv (both consume one buffer)
  nhvr_act_desc& g = p->g_desc;
  std::memset(&g, 0, sizeof(g));
  g.N = fwd->N; g.C8 = wg_round_up((fwd->Cout + 7) / 8, 2); g.H = fp->Ho; g.W = fp->Wo; g.halo = NHVR_HALO_ZERO;
  std::vector<int> g_off_of_acc(4, 0);
  // K runs over 128-position tiles: the round-up of the last tile must read ZERO gradient, so every (parity)
  // plane of the gradient buffer carries `extra` additional zero rows below the dgrad conv's own bottom halo
  const int extra = (kTileM - 1 + K.Wrow - 1) / K.Wrow;
  // row folding on the M side (see WgParams::mfold): the position loop runs kh-1 rows further, so the gradient buffer
  // carries kh-1 more zero rows at the bottom (the dgrad plan accepts the taller descriptor)
  const bool mfold = fwd->kind == NHVR_CONV && fwd->stride == 1 && fwd->Cout <= 8 && fwd->Cin <= 64 && fwd->kh <= 16 &&
                     fwd->kw >= 2 && fwd->kw <= 8 && !(std::getenv("NHVR_WGRAD_FOLD") && std::atoi(std::getenv("NHVR_WGRAD_FOLD")) == 0);
  if (fwd->kind == NHVR_CONV && fwd->stride == 1) {
    g.pad_t = fwd->kh - 1; g.pad_b = fwd->kh - 1 + extra + (mfold ? fwd->kh - 1 : 0); g.pad_l = fwd->kw - 1; g.pad_r = 0;
    const int pitch = fp->Wo + fwd->kw - 1;
    if (pitch != K.Wrow) { nhvr_conv_plan_destroy(fp); delete p; return NHVR_ERR_SHAPE; }
    g_off_of_acc[0] = (fwd->kh - 1) * pitch + (fwd->kw - 1);
  } else if (fwd->kind == NHVR_CONV) {          // stride 2: dgrad is the transposed conv, input pad (0,0,1,1)
    g.pad_b = 1 + extra; g.pad_r = K.Wrow - fp->Wo;
    if (g.pad_r < 0) { nhvr_conv_plan_destroy(fp); delete p; return NHVR_ERR_SHAPE; }
    g_off_of_acc[0] = 0;
  } else {                                      // forward transposed conv: dgrad is the s2 conv, split pad-1 format
    g.pad_t = g.pad_l = g.pad_r = 1; g.pad_b = 1 + 2 * extra; g.split = 1;
    const ActGeom gg = make_geom(g);
    const int Hq = gg.Hp / 2, Wq = gg.Wp / 2;
    if (Wq != K.Wrow) { nhvr_conv_plan_destroy(fp); delete p; return NHVR_ERR_SHAPE; }
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int py = (a + 1) & 1, px = (b + 1) & 1;
        g_off_of_acc[a * 2 + b] = ((py * 2 + px) * Hq + ((a + 1) >> 1)) * Wq + ((b + 1) >> 1);
      }
  }
  const ActGeom gg = make_geom(g);

  // ---- taps
  W.ntaps_total = fp->njobs_h;
  if (W.ntaps_total > kWgMaxTaps) { nhvr_conv_plan_destroy(fp); delete p; return NHVR_ERR_UNSUPPORTED; }
  // distinct G offsets -> G runs of 128 positions each
  std::vector<int> goffs;
  for (int j = 0; j < fp->njobs_h; ++j) {
    const int go = g_off_of_acc[fp->jobs_h[j].acc];
    if (std::find(goffs.begin(), goffs.end(), go) == goffs.end()) goffs.push_back(go);
  }
  std::sort(goffs.begin(), goffs.end());
  W.ngruns = (int)goffs.size();
  int gslab = 0;
  for (int i = 0; i < W.ngruns; ++i) { W.gruns[i] = ConvRun{goffs[i], kTileM, gslab}; gslab += kTileM; }
  W.gslab_units = gslab;
  for (int j = 0; j < fp->njobs_h; ++j) {
    const int go = g_off_of_acc[fp->jobs_h[j].acc];
    const int ri = (int)(std::find(goffs.begin(), goffs.end(), go) - goffs.begin());
    W.taps[j].x_off = fp->jobs_h[j].a_off;
    W.taps[j].g_off = W.gruns[ri].s_off;
    W.taps[j].tap = fp->pp.job_tap[j];
  }
  W.nxruns = K.nruns;
  for (int i = 0; i < K.nruns; ++i) W.xruns[i] = K.runs[i];
  W.xslab_units = K.slab_units;
  W.x_plane_units = xg.plane_units;
  W.g_plane_units = gg.plane_units;
  W.C8x = xg.C8; W.C8g = gg.C8;
  W.CoutP = wg_round_up(fwd->Cout, kTileM);
  W.CinP = xg.C8 * 8;
  W.fold = 0; W.kw = fwd->kw;

  // ---- tap folding for narrow inputs (the 7x7 stems: 3 or 9 input channels).  The plain lowering issues kh*kw MMAs of
  // N = 16 per 16 positions (issue-bound, 13-76 TFLOP/s measured) and needs several passes over the data because
  // kh*kw accumulators do not fit TMEM; folded, a filter row is ONE MMA of N = 64 per input plane.
  const int real_planes = (fwd->Cin + 7) / 8;
  const bool fold = fwd->kind == NHVR_CONV && fwd->stride == 1 && real_planes <= 2 && fwd->kw >= 5 && fwd->kw <= 8 &&
                    !(std::getenv("NHVR_WGRAD_FOLD") && std::atoi(std::getenv("NHVR_WGRAD_FOLD")) == 0);
  int nci = 0, best_groups = 1, best_S = 1;
  W.mfold = 0; W.kh = fwd->kh; W.g_row_units = K.Wrow;
  if (mfold) {
    // A: kh row-shifted copies of gradient plane 0 (one 128-position run each); B: X row 0 with the kw column shifts folded
    // into N (as below); one accumulator of 64 columns per X plane.  The position loop runs kh-1 rows past the last valid
    // one (G shifted up by r rows pairs X row y with G row y - r), which the gradient buffer covers with kh-1 more zero rows.
    const int run_len = kTileM + 8;
    W.nxruns = 1;
    W.xruns[0] = ConvRun{0, run_len, 0};
    W.xslab_units = run_len;
    W.ngruns = 1;
    W.gruns[0] = ConvRun{g_off_of_acc[0], kTileM, 0};
    W.gslab_units = kTileM;
    W.ntaps_total = real_planes;
    for (int pl = 0; pl < real_planes; ++pl) {
      WgTap& t = W.taps[pl];
      t.x_off = pl * W.xslab_units; t.g_off = 0; t.tap = pl;
    }
    W.fold = 1; W.mfold = 1;
    nci = real_planes * 8;
    best_groups = 1;
    const size_t stage = (size_t)16 * W.gslab_units * 16 + (size_t)real_planes * W.xslab_units * 16;
    best_S = (int)std::min<size_t>(4, (size_t)(210 * 1024) / stage);
  } else if (fold) {
    // own X slab: one run per filter row, 8 extra positions for the shifted N groups (the 8th group is discarded)
    const int run_len = kTileM + 8;
    W.nxruns = fwd->kh;
    for (int r = 0; r < fwd->kh; ++r) W.xruns[r] = ConvRun{r * K.Wrow, run_len, r * run_len};
    W.xslab_units = fwd->kh * run_len;
    W.ntaps_total = fwd->kh * real_planes;
    const int g_off = W.taps[0].g_off;            // stride 1: a single G offset
    for (int r = 0; r < fwd->kh; ++r)
      for (int pl = 0; pl < real_planes; ++pl) {
        WgTap& t = W.taps[r * real_planes + pl];
        t.x_off = r * run_len + pl * W.xslab_units;
        t.g_off = g_off;
        t.tap = r * 16 + pl;
      }
    W.fold = 1;
    nci = real_planes * 8;                        // X planes staged per CTA
    best_groups = (W.ntaps_total * 64 + 511) / 512;
    const size_t stage = (size_t)16 * gslab * 16 + (size_t)real_planes * W.xslab_units * 16;
    best_S = (int)std::min<size_t>(4, (size_t)(210 * 1024) / stage);
    if (best_S < 1) { nhvr_conv_plan_destroy(fp); delete p; return NHVR_ERR_SMEM; }
  } else {
  // ---- N (input channels per CTA), tap groups (ntaps_grp * nci <= 512 TMEM columns) and pipeline depth.
  // Cost model per (chunk, all input channels): tensor time ~ ntaps * (CinP/nci) * 8 MMAs * cycles(nci), operand
  // staging ~ ngroups * (CinP/nci) * stage_bytes / ~24 B per cycle; a stage must fit at least twice.
  double best_cost = 1e30;
  int force_nci = 0;                      // experiments: NHVR_WGRAD_NCI=<input channels per CTA>; measured on the end-to-end step: 64 (two tap
                                          // groups, two stages) 13.0 ms of wgrad vs 12.5 ms for the model's choice (32 at Cin = 256), 48: 12.5
  if (const char* e = std::getenv("NHVR_WGRAD_NCI")) force_nci = std::atoi(e);
  for (int cand : {128, 96, 64, 48, 32, 16}) {
    if (W.CinP % cand) continue;
    if (cand > 64 && force_nci != cand) continue;
    if (force_nci && force_nci != cand && W.CinP % force_nci == 0 && force_nci <= W.CinP) continue;
    const int ngroups = (W.ntaps_total * cand + 511) / 512;
    const size_t stage = (size_t)16 * gslab * 16 + (size_t)(cand / 8) * W.xslab_units * 16;
    const int S = (int)std::min<size_t>(4, (size_t)(210 * 1024) / stage);
    if (S < 1) continue;
    const double cyc = cand == 128 ? 66.0 : cand == 96 ? 56.0 : cand == 64 ? 48.0 : cand == 48 ? 44.0 : cand == 32 ? 40.0 : 39.0;
    const double blocks = (double)W.CinP / cand;
    const double t_mma = W.ntaps_total * blocks * 8.0 * cyc;
    const double t_ld = ngroups * blocks * (double)stage / 24.0;
    double cost = std::max(t_mma, t_ld) + 0.25 * std::min(t_mma, t_ld);
    if (S < 2) cost *= 2.0;                     // no overlap of staging and MMAs
    if (cost < best_cost) { best_cost = cost; nci = cand; best_groups = ngroups; best_S = S; }
  }
  }
  if (!nci) { nhvr_conv_plan_destroy(fp); delete p; return NHVR_ERR_SMEM; }
  W.nci = nci; W.nci8 = nci / 8;
  p->n_tap_groups = best_groups;
  W.ntaps_grp = (W.ntaps_total + best_groups - 1) / best_groups;        // balanced groups
  int cols = 32; while (cols < W.ntaps_grp * (W.fold ? 64 : nci)) cols <<= 1;
  W.tmem_cols = cols;
  W.n_ci_blocks = W.fold ? 1 : W.CinP / nci;
  p->n_co_blocks = W.CoutP / kTileM;

  // ---- stages and K split: whole waves of 148 CTAs (one CTA per SM: TMEM / shared memory bound)
  const size_t stage = (size_t)16 * gslab * 16 + (size_t)W.nci8 * W.xslab_units * 16;
  W.S = best_S;
  p->smem_bytes = best_S * stage + (2 * best_S + 1) * 8 + 16 + 128;
  W.tiles_per_img = fp->tiles_per_img;
  if (W.mfold) {
    const int64_t last_q = (int64_t)(K.Hv - 1) * K.Wrow + K.Wv + (int64_t)(fwd->kh - 1) * K.Wrow;
    W.tiles_per_img = (int)((last_q + kTileM - 1) / kTileM);
  }
  W.nchunks_total = W.tiles_per_img * fwd->N;
  const int ctas_other = p->n_co_blocks * W.n_ci_blocks * p->n_tap_groups;
  int best_split = 1;
  double best_t = 1e30;
  for (int sp = 1; sp <= std::min(W.nchunks_total, 64); ++sp) {
    const int cpc = (W.nchunks_total + sp - 1) / sp;
    const int nsp = (W.nchunks_total + cpc - 1) / cpc;
    const int waves = (ctas_other * nsp + 147) / 148;
    const double t = waves * (cpc + 6.0);      // + per-CTA prologue / epilogue in chunk units
    if (t < best_t) { best_t = t; best_split = nsp; }
  }
  W.chunks_per_cta = (W.nchunks_total + best_split - 1) / best_split;
  p->nsplit_k = (W.nchunks_total + W.chunks_per_cta - 1) / W.chunks_per_cta;
  p->ws_bytes = (size_t)(W.fold ? fwd->kh * fwd->kw : W.ntaps_total) * W.CoutP * W.CinP * sizeof(float);
  nhvr_conv_plan_destroy(fp);
  *out = p;
  return NHVR_OK;
}

extern "C" void nhvr_wgrad_plan_destroy(nhvr_wgrad_plan* p) { delete p; }
extern "C" int nhvr_wgrad_grad_desc(const nhvr_wgrad_plan* p, nhvr_act_desc* g_desc) {
  if (!p || !g_desc) return NHVR_ERR_NULL;
  *g_desc = p->g_desc;
  return NHVR_OK;
}
extern "C" size_t nhvr_wgrad_workspace_bytes(const nhvr_wgrad_plan* p) { return p ? p->ws_bytes : 0; }

extern "C" int nhvr_wgrad(const nhvr_wgrad_plan* p, const void* x, const void* g, void* workspace, float* dw, float scale,
                          int32_t accumulate, void* stream) {
  if (!p || !x || !g || !workspace || !dw) return NHVR_ERR_NULL;
  if ((((uintptr_t)x | (uintptr_t)g | (uintptr_t)workspace) & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  cudaStream_t st = (cudaStream_t)stream;
  WgParams W = p->wp;
  W.x = reinterpret_cast<const uint4*>(x);
  W.g = reinterpret_cast<const uint4*>(g);
  W.ws = reinterpret_cast<float*>(workspace);
  W.f16 = operand_f16();
  cudaError_t e = cudaMemsetAsync(workspace, 0, p->ws_bytes, st);
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid(p->nsplit_k, p->n_co_blocks * W.n_ci_blocks, p->n_tap_groups);
  wgrad_kernel<<<grid, kWgThreads, p->smem_bytes, st>>>(W);
  count_launch();
  const nhvr_conv_desc& d = p->d;
  const int64_t total = (int64_t)d.Cout * d.Cin * d.kh * d.kw;
  wgrad_finalize_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(
      W.ws, dw, d.Cout, d.Cin, W.CoutP, W.CinP, d.kh * d.kw, d.kind == NHVR_CONV_TRANSPOSE, scale, accumulate);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}
