// Shared between the forward/dgrad conv kernel and the wgrad kernel: the "shift program" of one conv.
#pragma once
#include "common.cuh"
#include "p8.cuh"

namespace nhvr {

constexpr int kMaxJobs = 52;
constexpr int kMaxMma = 208;   // jobs x k-steps per chunk
constexpr int kMaxRuns = 16;
constexpr int kTileM = 128;

struct ConvJob {
  int32_t a_off;   // shift (16-B units) inside a plane slab
  int16_t acc;     // accumulator index
  int16_t first;   // first job of its accumulator (overwrites instead of accumulating)
};
struct ConvMma {     // one weight block of a chunk and the tcgen05.mma(s) it feeds: precomputed so the issue loop has no arithmetic chains
  int32_t a_off;     // (job shift + k-step plane offset) in 16-B units inside the chunk slab
  uint32_t meta;     // bits 0-15: accumulator column offset in TMEM; bit 16: first MMA of its accumulator (overwrites
                     // instead of accumulating when executed in the first chunk).  32-bit fields keep the reads uniform.
  int32_t a_off2;    // split precision: the same weight block (w_hi) feeds a SECOND MMA whose A operand is the lo planes at this
                     // offset (x_hi*w_hi and x_lo*w_hi share one streamed block); < 0: none
};
struct ConvRun {
  int32_t g_off;   // offset (units) from the plane base + q0
  int32_t len;     // units
  int32_t s_off;   // offset (units) inside the plane slab
};

struct ConvKParams {
  const uint4* in;
  const uint4* w;
  const float* bias;
  void* out;
  double* stats;           // [N][Cout8*8][2] sum / sum of squares (fp64: E[x^2] - mean^2 must survive |mean| >> std)
  const float* acc_scale;  // split precision: device scalar 2^-s multiplied into the accumulators (nullptr: none)
  long long* trace;        // NHVR_CONV_TRACE: per-CTA cycle counters (nullptr in production)
  int64_t in_plane_units;
  int64_t w_split_units;   // packed-weight units per N-split
  int32_t C8in, kcp, nchunks, njobs, nruns, nacc;
  int32_t slab_units, Npad, bpb, nbstages, nblocks;
  int32_t SA, SB;
  int32_t Wrow, Hv, Wv, oys, oxs;
  int32_t oy[4], ox[4];
  int32_t Ho, Wo, Cout, Cout8;
  int32_t epilogue, act;
  int32_t tmem_cols;
  int32_t f16;             // operand element type: 0 bf16, 1 fp16
  int32_t debug;           // NHVR_CONV_DEBUG experiments (results are wrong when set): 8 no epilogue, 16 no statistics, 64 no global stores
  ActGeom og;              // BIAS_ACT_P8 destination
  int32_t mmas_per_chunk, stages_per_chunk;
  int32_t tile_step;       // linear positions a CTA advances by: 128, or 128-(kw-1) in row mode
  int32_t xstep;           // stacked tiles: columns a tile advances by (128, or 128-(kw-1) in row mode)
  int32_t rowmode, Cp, kw; // row mode: accumulator column n = s*Cp + co, outputs = shifted sums over s (epilogue)
  // M replication: one CTA owns `mrep` 128-position blocks that share every weight block (one smem B tile feeds
  // mrep MMAs), `q_mstride` linear positions / `a_mstride` slab units apart.  xtiles > 0: the blocks are the same
  // 128-pixel row segment of mrep consecutive output rows ("stacked"); xtiles == 0: mrep*128 consecutive positions.
  int32_t mrep, a_mstride, q_mstride, xtiles, acc_mstride;
  int32_t a_lbo_units;     // K-group stride of the A descriptor: the slab plane stride, or 1 (tap pairing: next pixel)
  int32_t stat_centred;    // RAW_STATS sums are centred on stats[n][c][2] (conv desc flag bit 4)
  int32_t out_hilo;        // RAW outputs are written as a split-precision (hilo) activation
  int32_t pair;            // 1: launched as clusters of two CTAs running cta_group::2 MMAs (weights packed per CTA half)
  // NHVR_EPI_IN_FUSED (nhvr_conv_forward_in_fused): the InstanceNorm is finished inside this kernel.  The accumulators stay
  // in TMEM while the CTAs of an image meet at a per-image arrival counter; then every CTA normalises its own tile and
  // writes it (activation, + residual, mirrored halo) straight into the consumer's P8 buffer `out` with geometry `og`.
  const uint4* res;        // residual (skip) activation, nullable; geometry sg
  uint32_t* sync;          // [N + 1] zero-initialised: per-image arrivals, CTAs past their wait (the last one re-zeroes)
  ActGeom sg;
  float eps, inv_hw;
  int32_t res_bulk;        // residual tile staged in shared memory by bulk copies (skip activation linearises like the input)
  int32_t res_off;         // units from (plane base + q0) to the tile's first residual unit
  int32_t start_delay;     // cycles the launch's second image waits before it starts (0: none)
  int32_t res_sp;          // epilogue steps (pairs of 16-channel groups) per staging phase
  ConvRun runs[kMaxRuns];
  ConvMma mma[kMaxMma + 1];   // +1: the issue loop prefetches one entry ahead
};
static_assert(sizeof(ConvKParams) <= 4096, "ConvKParams is passed as a __grid_constant__ kernel parameter");


}  // namespace nhvr

// weight packing parameters (conv_pack_weights_kernel)
namespace nhvr {
struct PackParams {
  const float* w;
  uint4* dst;
  float* tail;                 // split precision: {max|w| (bits), 2^-s} behind the packed blocks
  int32_t Cin, Cout, kh, kw, transposed;
  int32_t kcp, nchunks, njobs, Npad, nsplit, nblocks_padded;
  int32_t f16;
  int32_t flip;                // dgrad of a stride-1 conv: taps mirrored (r,s) -> (kh-1-r, kw-1-s)
  int32_t rowmode, Cp;         // row mode: job_tap = r*8 + accumulator, column n = s*Cp + co
  int32_t kfold;               // tap pairing: K group kp of a block = filter column job_tap%kw + kp, channels 0..7
  int32_t khalf;               // split precision, single logical input plane: K group 1 = the lo plane (see the plan builder)
  int32_t split3;              // split precision: blocks (hi*w_hi, hi*w_lo, lo*w_hi) per group of four physical planes
  int32_t nowlo;               // split precision without the w_lo blocks (conv desc flag bit 6): x_hi*w_hi + x_lo*w_hi only
  int32_t pair, bpb;           // CTA-pair layout: [stage of bpb blocks][rank][bpb][2][Npad/2][8]
  int16_t job_tap[kMaxJobs];   // r*kw + s of each job
};
}  // namespace nhvr

struct nhvr_conv_plan {
  nhvr_conv_desc d;
  nhvr_act_desc in_desc;
  nhvr::ConvKParams kp;          // pointers filled at launch
  nhvr::PackParams pp;
  int32_t Ho, Wo, Cout8, nsplit, tiles_per_img;
  size_t smem_bytes;
  size_t weight_bytes;
  // host copies of the shift program (wgrad reuses them)
  int32_t njobs_h;
  nhvr::ConvJob jobs_h[nhvr::kMaxJobs];
};
