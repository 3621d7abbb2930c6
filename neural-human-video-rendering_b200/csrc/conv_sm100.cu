// Shift-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Every Conv2d / ConvTranspose2d of the rendering path is lowered to ONE kernel shape:
//
//   acc[a][q0+m][:] += In8[plane][off_j + q0 + m][0:16] * W_j[0:16][:]      m = 0..127
//
// The P8 layout (include/nhvr.h) stores 8 channels of one pixel in one 16-byte unit and pixels of a
// (halo-padded) image linearly, so a filter tap is a constant shift `off_j` of the linear pixel
// index.  A tile is 128 consecutive linear output positions; its input "slab" is staged ONCE in
// shared memory by 1-D bulk async copies (UBLKCP) and every tap is just a different start address
// of a K-major, un-swizzled tcgen05 shared-memory descriptor (rows 16 B apart, SBO = 128 B,
// LBO = plane stride).  No im2col expansion, no padding logic, no per-tap reload.
// Weights are pre-packed in consumption order and streamed through a second ring.
//
// Roles (384 threads): warp0 = slab producer, warp1 = weight producer, warp2 = MMA issuer (one
// thread), warp3 = TMEM allocator, warps 4-11 = epilogue (TMEM -> registers -> global, InstanceNorm
// statistics by warp-shuffle column reduction, or bias + activation).
#include "conv_plan.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>
#include <algorithm>

namespace nhvr {

extern void note_cuda_error(cudaError_t e);
extern void count_launch();
extern int arch_ok_cached();
extern int operand_f16();

// -------------------------------------------------------------------------------------------------
// warp-level column sums: every lane holds 16 values (one row, 16 columns); on return lane l holds
// the sum over the warp's 32 rows of column (l >> 1).  16 shuffles instead of 16*5.
NHVR_DEVINL float warp_colsum16(const float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2], a1;
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float keep = h4 ? v[i + 8] : v[i];
    float send = h4 ? v[i] : v[i + 8];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float keep = h3 ? a8[i + 4] : a8[i];
    float send = h3 ? a8[i] : a8[i + 4];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float keep = h2 ? a4[i + 2] : a4[i];
    float send = h2 ? a4[i] : a4[i + 2];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    float keep = h1 ? a2[1] : a2[0];
    float send = h1 ? a2[0] : a2[1];
    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}

// bias + activation on 16 consecutive output channels.  The switch sits OUTSIDE the per-channel loops: a per-element
// switch compiles to an indirect branch per value and serialises the epilogue (measured: 13.5 K cycles per 128 x 80
// block of the UV head, profiles/r01_conv_ablation.md).
// tanh / sigmoid through one ex2 + one reciprocal (absolute error ~1e-7: cancellation only where the result is ~0)
NHVR_DEVINL float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }
NHVR_DEVINL float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

NHVR_DEVINL void bias_act16(const float (&v)[16], const float* sb, int act, int c0, int Cout, float (&t)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = v[i] + sb[i];
  if (act == NHVR_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = fmaxf(t[i], 0.f);
  } else if (act == NHVR_ACT_LRELU02) {
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = t[i] > 0.f ? t[i] : 0.2f * t[i];
  } else if (act == NHVR_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = tanh_fast(t[i]);
  } else if (act == NHVR_ACT_TANH_SIGMOID_LAST) {
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = (c0 + i == Cout - 1) ? sigmoid_fast(t[i]) : tanh_fast(t[i]);
  }
}

// one 16-channel group of one output pixel: raw P8 store (+ InstanceNorm statistics), or bias + activation.
// sb: this group's 16 bias values in shared memory (zeros beyond Cout / without bias).
// V: kernel variant (0 = standard layers; 1 = also split-precision (hilo) raw outputs and centred statistics)
template <int V>
NHVR_DEVINL void emit16(const ConvKParams& P, const float (&v)[16], int c0, bool valid, int n, int Y, int X, int g_local, int lane,
                        float* s_stats, const float* sb) {
  if (P.debug & 64) valid = valid && (v[0] == 123456.f);
  if (P.epilogue == NHVR_EPI_RAW_STATS || P.epilogue == NHVR_EPI_RAW_P8) {
    if (valid) {
      uint4* o = reinterpret_cast<uint4*>(P.out);
      const int64_t u0 = (((int64_t)n * P.Cout8 + (c0 >> 3)) * P.Ho + Y) * P.Wo + X;
      const int64_t pstride = (int64_t)P.Ho * P.Wo;
      uint4 lo, hi;
      lo.x = pack2(v[0], v[1], P.f16);  lo.y = pack2(v[2], v[3], P.f16);
      lo.z = pack2(v[4], v[5], P.f16);  lo.w = pack2(v[6], v[7], P.f16);
      hi.x = pack2(v[8], v[9], P.f16);  hi.y = pack2(v[10], v[11], P.f16);
      hi.z = pack2(v[12], v[13], P.f16); hi.w = pack2(v[14], v[15], P.f16);
      const uint64_t keep = l2_policy_evict_last();      // read back once by the IN-apply / gradient pass that follows
      if (V && P.out_hilo) {
        // split-precision raw output: planes [hi 2g, hi 2g+1, lo 2g, lo 2g+1] of channel group g = c0 / 16
        // (Cout8 is even, so both logical planes of the group exist); `lo`, `hi` above are the rounded halves
        uint4 l0, l1;
        l0.x = pack2(v[0] - unpack_lo(lo.x, P.f16), v[1] - unpack_hi(lo.x, P.f16), P.f16);
        l0.y = pack2(v[2] - unpack_lo(lo.y, P.f16), v[3] - unpack_hi(lo.y, P.f16), P.f16);
        l0.z = pack2(v[4] - unpack_lo(lo.z, P.f16), v[5] - unpack_hi(lo.z, P.f16), P.f16);
        l0.w = pack2(v[6] - unpack_lo(lo.w, P.f16), v[7] - unpack_hi(lo.w, P.f16), P.f16);
        l1.x = pack2(v[8] - unpack_lo(hi.x, P.f16), v[9] - unpack_hi(hi.x, P.f16), P.f16);
        l1.y = pack2(v[10] - unpack_lo(hi.y, P.f16), v[11] - unpack_hi(hi.y, P.f16), P.f16);
        l1.z = pack2(v[12] - unpack_lo(hi.z, P.f16), v[13] - unpack_hi(hi.z, P.f16), P.f16);
        l1.w = pack2(v[14] - unpack_lo(hi.w, P.f16), v[15] - unpack_hi(hi.w, P.f16), P.f16);
        if ((c0 >> 3) < P.Cout8) {
          const int64_t uh = (((int64_t)n * 2 * P.Cout8 + ((c0 >> 4) << 2)) * P.Ho + Y) * P.Wo + X;
          st_hint(o + uh, lo, keep);
          st_hint(o + uh + pstride, hi, keep);
          st_hint(o + uh + 2 * pstride, l0, keep);
          st_hint(o + uh + 3 * pstride, l1, keep);
        }
      } else {
        if ((c0 >> 3) < P.Cout8) st_hint(o + u0, lo, keep);
        if ((c0 >> 3) + 1 < P.Cout8) st_hint(o + u0 + pstride, hi, keep);
      }
    }
    if (P.epilogue == NHVR_EPI_RAW_STATS && !(P.debug & 16)) {
      float s[16], ss[16];
#pragma unroll
      if (V && P.stat_centred) {               // first layers only: sums centred on the shift held in sb (nhvr_stem_stat_shift)
#pragma unroll
        for (int i = 0; i < 16; ++i) { s[i] = valid ? v[i] - sb[i] : 0.f; ss[i] = s[i] * s[i]; }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { s[i] = valid ? v[i] : 0.f; ss[i] = s[i] * s[i]; }
      }
      const float cs = warp_colsum16(s, lane);
      const float css = warp_colsum16(ss, lane);
      const int col = g_local * 16 + (lane >> 1);
      atomicAdd(&s_stats[col * 2 + (lane & 1)], (lane & 1) ? css : cs);
    }
  } else if (P.epilogue == NHVR_EPI_BIAS_ACT_F32) {
    if (valid) {
      float t[16];
      bias_act16(v, sb, P.act, c0, P.Cout, t);
      const int64_t plane = (int64_t)P.Ho * P.Wo;
      float* o = reinterpret_cast<float*>(P.out) + (((int64_t)n * P.Cout + c0) * P.Ho + Y) * P.Wo + X;
      const int nc = P.Cout - c0;            // channels of this group that exist (warp-uniform)
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nc) o[i * plane] = t[i];
    }
  } else {  // NHVR_EPI_BIAS_ACT_P8
    if (valid) {
      uint4* o = reinterpret_cast<uint4*>(P.out);
      float t[16];
      bias_act16(v, sb, P.act, c0, P.Cout, t);
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = (c0 + i < P.Cout) ? t[i] : 0.f;
      uint4 lo, hi;
      lo.x = pack2(t[0], t[1], P.f16);  lo.y = pack2(t[2], t[3], P.f16);
      lo.z = pack2(t[4], t[5], P.f16);  lo.w = pack2(t[6], t[7], P.f16);
      hi.x = pack2(t[8], t[9], P.f16);  hi.y = pack2(t[10], t[11], P.f16);
      hi.z = pack2(t[12], t[13], P.f16); hi.w = pack2(t[14], t[15], P.f16);
      const int p0 = c0 >> 3;
      if (p0 < P.og.C8) o[act_unit(P.og, n, p0, Y + P.og.pad_t, X + P.og.pad_l)] = lo;
      if (p0 + 1 < P.og.C8) o[act_unit(P.og, n, p0 + 1, Y + P.og.pad_t, X + P.og.pad_l)] = hi;
    }
  }
}

constexpr int kThreads = 384;       // warps 0-3: roles, warps 4-11: epilogue

// PAIR: two CTAs of a cluster (the two SMs of a TPC) run ONE M = 256 tcgen05.mma.cta_group::2 per weight block: each
// CTA stages its own 128-position input slab and only HALF of the weight block's N rows, so the shared-memory traffic
// per MMA (the measured ceiling of the single-CTA kernel: profiles/r01_conv_ablation.md) drops from
// A + 2*B to A + B bytes.  The leader (rank 0) issues the MMAs; the peer's MMA warp relays its local "stage full"
// events to the leader's barriers; commits are multicast to both CTAs; each CTA drains its own 128 TMEM lanes.
// V = 1 adds the rarely used paths (accumulator scale + hilo raw output; centred stem statistics; the shuffle epilogue of the
// row-mode RGB head; the fused InstanceNorm epilogue) so that their registers and code do not weigh on the standard layers;
// V = 2 = V = 1 + the split-precision issue loop (a w_hi weight block feeds the x_hi and the x_lo MMA).
template <bool PAIR, int V>
__global__ void __launch_bounds__(kThreads, V ? 2 : 1) conv_shiftgemm_kernel(const __grid_constant__ ConvKParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  // warp index through a shuffle: provably warp-uniform, so the role branches and everything inside them (ring
  // counters, descriptors, table reads) stay on the uniform datapath instead of R2UR round trips per MMA
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n = blockIdx.y, split = blockIdx.z;
  const long long t_entry = P.trace ? clock64() : 0;
  long long* trace = P.trace ? P.trace + 16 * ((int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
  if (trace && threadIdx.x == 0) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); trace[8] = sm; trace[9] = clock64(); }
  int ty0 = 0, tx0 = 0;                 // stacked tiles: first output row / column of the tile
  int64_t q0;
  if (P.xtiles) {
    const int yg = blockIdx.x / P.xtiles;
    ty0 = yg * P.mrep; tx0 = (blockIdx.x - yg * P.xtiles) * P.xstep;
    q0 = (int64_t)ty0 * P.Wrow + tx0;
  } else {
    q0 = (int64_t)blockIdx.x * P.tile_step;
  }
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // K-chunk order is rotated per CTA: neighbouring CTAs stream different parts of the (shared) packed
  // weights at any instant instead of all hitting the same L2 lines in lock-step.
  const int rot = (int)((PAIR ? blockIdx.x >> 1 : blockIdx.x) % (unsigned)P.nchunks);   // one rotation per CTA pair

  const uint32_t a_stage_bytes = (uint32_t)P.kcp * P.slab_units * 16u;
  const uint32_t b_block_bytes = (uint32_t)P.Npad * (PAIR ? 16u : 32u);   // PAIR: this CTA's N/2 rows of the block
  const uint32_t b_stage_bytes = b_block_bytes * P.bpb;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = a_smem + (size_t)P.SA * a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)P.SB * b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + P.SA;
  uint64_t* b_full = a_empty + P.SA;
  uint64_t* b_empty = b_full + P.SB;
  uint64_t* acc_full = b_empty + P.SB;
  uint64_t* a_peer = acc_full + 1;           // PAIR, leader only: the peer's slab / weight stage is full
  uint64_t* b_peer = a_peer + P.SA;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_peer + P.SB);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_ptr + 2);   // fused InstanceNorm epilogue: residual tile landed in shared memory
  float* s_stats = reinterpret_cast<float*>(res_bar + 1);    // [Npad][2]

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < P.SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    for (int i = 0; i < P.SA; ++i) mbar_init(&a_peer[i], 1);
    for (int i = 0; i < P.SB; ++i) mbar_init(&b_peer[i], 1);
    mbar_init(res_bar, 1);
    fence_mbar_init();
  }
  if (warp == 3) {
    if (PAIR) { tmem_alloc2(tmem_ptr, (uint32_t)P.tmem_cols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_ptr, (uint32_t)P.tmem_cols); tmem_relinquish(); }
  }
  float* s_bias = s_stats + P.Npad * 2 * P.nacc;             // [Npad] bias of this CTA's output channels
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < P.Npad * 2 * P.nacc; i += 256) s_stats[i] = 0.f;
    for (int i = threadIdx.x - 128; i < P.Npad; i += 256) {
      const int c = split * P.Npad + i;
      // RAW_STATS: no bias (it cancels in the InstanceNorm); the slot holds the statistics shift of this channel
      // (stats[n][c][2], 0 unless the caller centred the sums, see nhvr_stem_stat_shift)
      if (P.epilogue == NHVR_EPI_RAW_STATS) s_bias[i] = (V && P.stat_centred && c < P.Cout8 * 8) ? (float)P.stats[((int64_t)n * P.Cout8 * 8 + c) * 4 + 2] : 0.f;
      else s_bias[i] = (P.bias && c < P.Cout) ? __ldg(P.bias + c) : 0.f;
    }
  }
  // Fused InstanceNorm epilogue: the image-wide meeting point keeps all CTAs of an image in phase; starting the SECOND image
  // of the launch late puts the two CTAs of an SM half a period apart for the rest of the launch (every later image starts
  // when the image two before it leaves), so that one runs its MMA loop while the other is in its epilogue.
  if (V && P.epilogue == NHVR_EPI_IN_FUSED && P.start_delay > 0 && blockIdx.y == 1 && threadIdx.x == 0) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)P.start_delay) __nanosleep(200);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();     // PAIR: the peer's barriers must be initialised before remote arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ slab producer
    {
      const uint4* img = P.in + (int64_t)n * P.C8in * P.in_plane_units + q0;
      int st = 0;
      uint32_t ph = 0;
      int cc = rot;                                   // chunk order rotated per CTA (see kernel header)
      for (int c = 0; c < P.nchunks; ++c, st = (st + 1 == P.SA) ? 0 : st + 1, ph ^= (st == 0) ? 1u : 0u) {
        mbar_wait(&a_empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[st], a_stage_bytes);
          uint8_t* dst = a_smem + (size_t)st * a_stage_bytes;
          for (int pl = 0; pl < P.kcp; ++pl) {
            const uint4* plane = img + (int64_t)(cc * P.kcp + pl) * P.in_plane_units;
            for (int r = 0; r < P.nruns; ++r) {
              bulk_g2s(dst + ((size_t)pl * P.slab_units + P.runs[r].s_off) * 16, plane + P.runs[r].g_off,
                       (uint32_t)P.runs[r].len * 16u, &a_full[st]);
            }
          }
        }
        __syncwarp();
        if (++cc == P.nchunks) cc = 0;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ weight producer
    {
      const uint4* wbase = P.w + (int64_t)split * P.w_split_units;
      const uint32_t stage_units = b_stage_bytes >> 4;
      int gs = rot * P.stages_per_chunk;              // global stage index, rotated start
      // PAIR weights are packed [stage][rank][bpb][2][N/2][8]: each CTA's half of a stage is one contiguous copy
      const uint32_t src_step = PAIR ? 2u * stage_units : stage_units;
      const uint4* wsrc = wbase + (int64_t)gs * src_step + (PAIR ? rank * stage_units : 0u);
      int st = 0;
      uint32_t ph = 0;
      for (int s = 0; s < P.nbstages; ++s) {
        mbar_wait(&b_empty[st], ph ^ 1u);
        if (trace && s == P.nbstages - 1 && lane == 0) trace[6] = clock64() - t_entry;   // last weight stage requested
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[st], b_stage_bytes);
          bulk_g2s(b_smem + (size_t)st * b_stage_bytes, wsrc, b_stage_bytes, &b_full[st]);
        }
        __syncwarp();
        wsrc += src_step;
        if (++gs == P.nbstages) { gs = 0; wsrc = wbase + (PAIR ? rank * stage_units : 0u); }
        if (++st == P.SB) { st = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 2 && PAIR && rank != 0) {
    // ------------------------------------------------------------------ peer of a CTA pair: relay "stage full" to the leader
    {
      const int SA = P.SA, SB = P.SB, nchunks = P.nchunks, spc = P.stages_per_chunk;
      int ast = 0, bst = 0;
      uint32_t aph = 0, bph = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&a_full[ast], aph);
        if (elect_one()) mbar_arrive_remote(&a_peer[ast], 0);
        __syncwarp();
        for (int s = 0; s < spc; ++s) {
          mbar_wait(&b_full[bst], bph);
          if (elect_one()) mbar_arrive_remote(&b_peer[bst], 0);
          __syncwarp();
          if (++bst == SB) { bst = 0; bph ^= 1u; }
        }
        if (++ast == SA) { ast = 0; aph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer
    // One thread feeds the tensor core, so this loop is kept to a handful of integer ops per MMA:
    // descriptors are advanced incrementally in 16-byte units, ring positions by counters (no div/mod).
    {
      const uint32_t idesc = make_idesc_16(PAIR ? 2 * kTileM : kTileM, (uint32_t)P.Npad, P.f16);
      const uint32_t a_lbo_u = (uint32_t)P.a_lbo_units;           // K-group stride: plane stride (or 1: tap pairing)
      const uint32_t b_lbo_u = PAIR ? (uint32_t)P.Npad >> 1 : (uint32_t)P.Npad;   // K-group stride = rows held by this CTA
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);          // SBO = 128 B, version 1, no swizzle
      const uint64_t dhi = (uint64_t)desc_hi << 32;
      const uint32_t a_lo0 = ((smem_u32(a_smem) & 0x3FFFFu) >> 4) | (a_lbo_u << 16);
      const uint32_t b_lo0 = ((smem_u32(b_smem) & 0x3FFFFu) >> 4) | (b_lbo_u << 16);
      const uint32_t a_stage_u = a_stage_bytes >> 4, b_block_u = b_block_bytes >> 4;
      const int bpb = P.bpb, SA = P.SA, SB = P.SB, nchunks = P.nchunks, spc = P.stages_per_chunk;
      const int mrep = P.mrep;
      const uint32_t a_mstride = (uint32_t)P.a_mstride, acc_mstride = (uint32_t)P.acc_mstride;
      int ast = 0, bst = 0;
      uint32_t aph = 0, bph = 0;
      uint32_t b_lo = b_lo0;
      const bool leader = elect_one();
      // chunk -> weight stage -> MMA.  All ring bookkeeping sits at stage granularity (bpb divides the MMAs of a
      // chunk by construction); the innermost loop is: load table entry, two adds, tcgen05.mma.
      long long wa = 0, wb = 0;
      const long long t_mma0 = trace ? clock64() : 0;
      // The barrier wait of the NEXT weight stage (and slab chunk) is issued before the LAST MMA of the current stage:
      // a try_wait costs ~100-200 cycles even on a completed phase, and the tensor pipe only buffers about two MMAs
      // behind the running one - waiting between stages drained it once per stage (profiles/r01_conv_ablation.md).
      auto wait_a = [&](int st, uint32_t ph) {
        const long long c0 = trace ? clock64() : 0;
        mbar_wait(&a_full[st], ph);
        if (PAIR) mbar_wait_cluster(&a_peer[st], ph);
        if (trace) wa += clock64() - c0;
      };
      auto wait_b = [&](int st, uint32_t ph) {
        const long long c0 = trace ? clock64() : 0;
        mbar_wait(&b_full[st], ph);
        if (PAIR) mbar_wait_cluster(&b_peer[st], ph);
        if (trace) wb += clock64() - c0;
      };
      wait_a(0, 0);
      wait_b(0, 0);
      tc_fence_after();
      for (int c = 0; c < nchunks; ++c) {
        // single slab stage: refilled only after the commit below.  The first weight stage of the chunk is waited for HERE too, not
        // before the previous chunk's last block: a CTA pair's relay warp forwards "slab full" before "weights full" of a chunk,
        // and the slab of chunk c only refills after chunk c-1 is committed - pre-waiting the weights would close a cycle
        if (SA == 1 && c > 0) { wait_a(0, aph); wait_b(bst, bph); }
        const uint32_t a_st_lo = a_lo0 + (uint32_t)ast * a_stage_u;
        const uint32_t keep_mask = (c == 0) ? 0u : 1u;          // chunk 0: the first MMA of an accumulator overwrites
        const int nast = (ast + 1 == SA) ? 0 : ast + 1;
        const uint32_t naph = aph ^ ((ast + 1 == SA) ? 1u : 0u);
        const ConvMma* e = P.mma;
        ConvMma cur = e[0];                                   // software-pipelined table read: the constant-bank
        for (int s = 0; s < spc; ++s) {                       // load of entry i+1 overlaps the issue of MMA i
          const int nbst = (bst + 1 == SB) ? 0 : bst + 1;
          const uint32_t nbph = bph ^ ((bst + 1 == SB) ? 1u : 0u);
          // one table entry = one weight block: mrep MMAs.  Operand arithmetic stays outside the leader-guarded
          // statements so that the guard compiles to a predicate on the UTCHMMA instead of a divergent branch.
          auto issue = [&](auto multi) {
            ++e;
            const ConvMma nxt = *e;                           // one entry past the end is still inside ConvKParams
            const uint32_t acc_flag = keep_mask | ((cur.meta >> 16) ^ 1u);
            uint32_t a_lo = a_st_lo + (uint32_t)cur.a_off, d_col = tmem_base + (cur.meta & 0xffffu);
            const uint64_t adesc = dhi | a_lo, bdesc = dhi | b_lo;
            if (leader) { if (PAIR) umma2_bf16(d_col, adesc, bdesc, idesc, acc_flag); else umma_bf16(d_col, adesc, bdesc, idesc, acc_flag); }
            if (decltype(multi)::value) {
              for (int i = 1; i < mrep; ++i) {               // the same weight block feeds the other M blocks
                a_lo += a_mstride; d_col += acc_mstride;
                const uint64_t adesc_i = dhi | a_lo;
                if (leader) { if (PAIR) umma2_bf16(d_col, adesc_i, bdesc, idesc, acc_flag); else umma_bf16(d_col, adesc_i, bdesc, idesc, acc_flag); }
              }
            }
            if (V == 2 && cur.a_off2 >= 0) {                 // split precision: x_lo * w_hi on the block that just fed x_hi * w_hi
              uint32_t a2 = a_st_lo + (uint32_t)cur.a_off2, d2 = tmem_base + (cur.meta & 0xffffu);
              const uint64_t adesc2 = dhi | a2;
              if (leader) { if (PAIR) umma2_bf16(d2, adesc2, bdesc, idesc, 1u); else umma_bf16(d2, adesc2, bdesc, idesc, 1u); }
              if (decltype(multi)::value) {
                for (int i = 1; i < mrep; ++i) {
                  a2 += a_mstride; d2 += acc_mstride;
                  const uint64_t adesc_i = dhi | a2;
                  if (leader) { if (PAIR) umma2_bf16(d2, adesc_i, bdesc, idesc, 1u); else umma_bf16(d2, adesc_i, bdesc, idesc, 1u); }
                }
              }
            }
            b_lo += b_block_u;
            cur = nxt;
          };
          if (mrep == 1) {
#pragma unroll 1
            for (int k = 0; k < bpb - 1; ++k) issue(std::false_type{});
          } else {
#pragma unroll 1
            for (int k = 0; k < bpb - 1; ++k) issue(std::true_type{});
          }
          // pre-wait what the next stage needs, then the last block of this stage
          if (s + 1 < spc) wait_b(nbst, nbph);
          else if (c + 1 < nchunks && SA > 1) { wait_a(nast, naph); wait_b(nbst, nbph); }
          if (mrep == 1) issue(std::false_type{}); else issue(std::true_type{});
          if (leader) { if (PAIR) umma2_commit(&b_empty[bst]); else umma_commit(&b_empty[bst]); }
          bst = nbst; bph = nbph;
          if (bst == 0) b_lo = b_lo0;
        }
        if (leader) { if (PAIR) umma2_commit(&a_empty[ast]); else umma_commit(&a_empty[ast]); }
        ast = nast; aph = naph;
      }
      __syncwarp();
      if (elect_one()) { if (PAIR) umma2_commit(acc_full); else umma_commit(acc_full); }
      if (trace && lane == 0) { trace[0] = t_mma0 - t_entry; trace[1] = wa; trace[2] = wb; trace[3] = clock64() - t_entry; }
    }
    __syncwarp();
  } else if (warp == 3 && V && P.epilogue == NHVR_EPI_IN_FUSED && P.res_bulk) {
    // ------------------------------------------------------------------ fused InstanceNorm epilogue: residual tile stager
    // The residual tile is 128 consecutive units per plane (the skip activation linearises like this conv's input): bulk
    // async copies bring it into the shared memory the operand rings no longer need (every MMA has retired), in phases of
    // res_sp epilogue steps = 2 * res_sp channel groups, one plane per lane, while the statistics are reduced and the
    // image's CTAs meet.  Later phases start when the 256 epilogue threads have released the previous one (barrier 3).
    const ActGeom& sg = P.sg;
    const int hilo = sg.hilo, npl = hilo ? 4 : 2;
    const int ngroups = P.Npad >> 4, cout_off = split * P.Npad;
    mbar_wait_warp(acc_full, 0);
    const uint4* rbase = P.res + (int64_t)n * sg.C8 * sg.plane_units + q0 + P.res_off;
    for (int phase = 0; 2 * phase * P.res_sp < ngroups; ++phase) {
      if (phase) asm volatile("bar.sync 3, 288;" ::: "memory");
      const int g0 = 2 * phase * P.res_sp, g1 = min(ngroups, g0 + 2 * P.res_sp);
      int planes = 0;
      for (int g = g0; g < g1; ++g) {
        const int pb = hilo ? ((cout_off + g * 16) >> 4) << 2 : (cout_off + g * 16) >> 3;
        if (pb + npl <= sg.C8) planes += npl;
      }
      if (lane == 0) mbar_arrive_expect_tx(res_bar, (uint32_t)planes * 2048u);
      __syncwarp();
      for (int i = lane; i < (g1 - g0) * npl; i += 32) {
        const int g = g0 + i / npl, j = i - (i / npl) * npl;
        const int pb = hilo ? ((cout_off + g * 16) >> 4) << 2 : (cout_off + g * 16) >> 3;
        if (pb + npl <= sg.C8)
          bulk_g2s(a_smem + (size_t)i * 2048, rbase + (int64_t)(pb + j) * sg.plane_units, 2048u, res_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int we = warp & 3;                  // TMEM lane quarter == warp % 4
    const int half = (warp - 4) >> 2;         // warps 4-7 take the even column groups, 8-11 the odd ones
    const int m = we * 32 + lane;
    int y, x;
    if (P.xtiles) { y = ty0; x = tx0 + m; }
    else { const int64_t q = q0 + m; y = (int)(q / P.Wrow); x = (int)(q - (int64_t)y * P.Wrow); }
    const bool valid_m = (y < P.Hv) && (x < P.Wv) && (P.xtiles || P.rowmode || m < P.tile_step);   // tiles may own fewer than 128 positions
    const uint32_t t_lane = tmem_base + ((uint32_t)(we * 32) << 16);
    const int cout_off = split * P.Npad;
    const int ngroups = P.Npad >> 4;

    // split precision: the packed weights carry a power-of-two scale (conv_pack_weights_kernel); undo it on the accumulators
    const bool scaled = V && P.acc_scale != nullptr;
    const float accs = scaled ? __ldg(P.acc_scale) : 1.f;
    mbar_wait_warp(acc_full, 0);
    tc_fence_after();
    if (trace && threadIdx.x == 128) trace[4] = clock64() - t_entry;

    if (V && P.rowmode && P.Cp == 8) {
      // ---- row mode, <= 8 output channels (the RGB + mask head): accumulator column s*8 + co of block `rep` holds
      // Z[m][s][co] (filter column s un-shifted); output Y[m][co] = sum_s Z[m+s][s][co].  Row m+s lives in lane
      // lane+s of the same warp (register shuffle) or, past lane 31, in the first kw-1 lanes of the next warp, which
      // publish those values through a small double-buffered shared-memory window.  The two epilogue halves (warps 4-7 /
      // 8-11) take alternate M blocks of the tile.
      if (!(P.debug & 8)) {
        const int kw = P.kw;                                 // 5..8 -> at most 7 neighbour rows
        const int nco = min(P.Cout, 8);
        const int xsz = 3 * 7 * 7 * nco;                     // [warp 3][lane 7][s-1 7][co nco] floats per buffer
        float* X = s_bias + P.Npad + half * 2 * xsz;         // two buffers per epilogue half
        const bool valid_x = (x < P.Wv) && (m < P.xstep);
        int it = 0;
        for (int rep = half; rep < P.mrep; rep += 2, ++it) {
          float* Xb = X + (it & 1) * xsz;
          float acc[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = 0.f;
          // phase A: own column 0, in-warp neighbours by shuffle, and the values the previous warp needs go to shared memory
          if (nco <= 4) {
            // <= 4 real output channels: seven 4-column TMEM loads in flight, one wait
            uint32_t zr[8][4];
#pragma unroll
            for (int sft = 0; sft < 8; ++sft)
              if (sft < kw) tmem_ld4(t_lane + (uint32_t)(rep * P.acc_mstride + sft * 8), zr[sft]);
            tmem_ld_wait();
#pragma unroll
            for (int sft = 0; sft < 8; ++sft) {
              if (sft < kw) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  if (c < nco) {
                    const float zv = __uint_as_float(zr[sft][c]) * accs;
                    if (sft == 0) acc[c] += zv;
                    else {
                      const float up = __shfl_down_sync(0xffffffffu, zv, sft);
                      if (lane + sft <= 31) acc[c] += up;
                      if (we > 0 && lane < sft) Xb[(((we - 1) * 7 + lane) * 7 + (sft - 1)) * nco + c] = zv;
                    }
                  }
                }
              }
            }
          } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j * 2 < kw) {
              uint32_t vr[16];
              tmem_ld16(t_lane + (uint32_t)(rep * P.acc_mstride + j * 16), vr);
              tmem_ld_wait();
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const int sft = 2 * j + h2;                    // filter column of these 8 accumulator columns
                if (sft < kw) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    if (c < nco) {
                      const float zv = __uint_as_float(vr[h2 * 8 + c]) * accs;
                      if (sft == 0) acc[c] += zv;
                      else {
                        const float up = __shfl_down_sync(0xffffffffu, zv, sft);
                        if (lane + sft <= 31) acc[c] += up;
                        if (we > 0 && lane < sft) Xb[(((we - 1) * 7 + lane) * 7 + (sft - 1)) * nco + c] = zv;
                      }
                    }
                  }
                }
              }
            }
          }
          }
          if (half == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
          // phase B: rows past lane 31 come from the next warp's first lanes
          if (we < 3) {
            for (int sft = 32 - lane; sft < kw; ++sft) {       // empty unless lane > 32 - kw
              if (sft >= 1) {
                const float* xs = Xb + ((we * 7 + (lane + sft - 32)) * 7 + (sft - 1)) * nco;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  if (c < nco) acc[c] += xs[c];
              }
            }
          }
          const int yr = y + rep;
          if (P.epilogue == NHVR_EPI_BIAS_ACT_F32) {           // the RGB (+ mask) head: only the nco real channels are finished
            if (valid_x && yr < P.Hv) {
              const int64_t plane = (int64_t)P.Ho * P.Wo;
              float* o = reinterpret_cast<float*>(P.out) + ((int64_t)n * P.Cout * P.Ho + yr) * P.Wo + x;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                if (c < nco) {
                  float t = acc[c] + s_bias[c];
                  if (P.act == NHVR_ACT_RELU) t = fmaxf(t, 0.f);
                  else if (P.act == NHVR_ACT_LRELU02) t = t > 0.f ? t : 0.2f * t;
                  else if (P.act == NHVR_ACT_TANH) t = tanh_fast(t);
                  else if (P.act == NHVR_ACT_TANH_SIGMOID_LAST) t = (c == P.Cout - 1) ? sigmoid_fast(t) : tanh_fast(t);
                  o[c * plane] = t;
                }
              }
            }
          } else {
            emit16<V>(P, acc, 0, valid_x && yr < P.Hv, n, yr, x, 0, lane, s_stats, s_bias);
          }
        }
      }
    } else if (P.rowmode) {
      // ---- row mode, generic (Cp >= 16, one M block): accumulator column n = s*Cp + co holds Z[m][s][co]; output Y[m][co] = sum_s Z[m+s][s][co].
      // Warps 4-7 exchange one 16-column chunk at a time through shared memory (row m reads row m+s).
      if (half == 0 && !(P.debug & 8)) {
        float* S = s_bias + P.Npad;                         // [128][17] floats
        const bool valid = valid_m && (m < P.tile_step);
        const int ngrp = (P.Cp + 15) >> 4;
        for (int cg = 0; cg < ngrp; ++cg) {
          float acc[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = 0.f;
          const int nchunk = (P.Cp >= 16) ? P.kw : (P.kw * P.Cp + 15) / 16;
          for (int j = 0; j < nchunk; ++j) {
            const int n0 = (P.Cp >= 16) ? (j * P.Cp + cg * 16) : j * 16;     // first accumulator column of this chunk
            const int a = n0 / P.Npad;
            uint32_t vr[16];
            tmem_ld16(t_lane + (uint32_t)(a * P.Npad + (n0 - a * P.Npad)), vr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) S[m * 17 + i] = __uint_as_float(vr[i]) * accs;
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (P.Cp >= 16) {
              if (m + j < kTileM) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] += S[(m + j) * 17 + i];
              }
            } else {                                          // Cp == 8: two filter columns per chunk
              const int s0 = 2 * j, s1 = 2 * j + 1;
              if (m + s0 < kTileM) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += S[(m + s0) * 17 + i];
              }
              if (s1 < P.kw && m + s1 < kTileM) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += S[(m + s1) * 17 + 8 + i];
              }
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
          }
          emit16<V>(P, acc, cg * 16, valid, n, y, x, cg, lane, s_stats, s_bias + cg * 16);
        }
      }
    } else if (V && P.epilogue == NHVR_EPI_IN_FUSED) {
      // ---- conv + InstanceNorm + activation (+ residual) + halo in one kernel (one M block, one accumulator per CTA).
      const ActGeom& og = P.og;
      const ActGeom& sg = P.sg;
      const int hilo = og.hilo;
      const int npl = hilo ? 4 : 2;                                   // physical planes per 16-channel group
      // The residual tile is 128 consecutive units per plane (the skip activation has this conv's own input geometry):
      // bulk async copies bring it into the shared memory the operand rings no longer need (every MMA has retired), in
      // phases of `res_sp` steps = 2 * res_sp groups, while the statistics are reduced and the image's CTAs meet.
      // Pass 1: per-channel sum / sum of squares of this tile straight from TMEM (nothing is stored).
      // (the TMEM load of the next group is in flight while the current one is reduced)
      uint32_t vr[16];
      if (half < ngroups) tmem_ld16(t_lane + (uint32_t)(half * 16), vr);
      for (int g = half; g < ngroups; g += 2) {
        tmem_ld_wait();
        float s[16], ss[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { s[i] = valid_m ? __uint_as_float(vr[i]) * accs : 0.f; ss[i] = s[i] * s[i]; }
        if (g + 2 < ngroups) tmem_ld16(t_lane + (uint32_t)((g + 2) * 16), vr);
        const float cs = warp_colsum16(s, lane);
        const float css = warp_colsum16(ss, lane);
        atomicAdd(&s_stats[(g * 16 + (lane >> 1)) * 2 + (lane & 1)], (lane & 1) ? css : cs);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (trace && threadIdx.x == 128) trace[1] = clock64() - t_entry;          // (fused) pass 1 done
      const int nch = min(P.Npad, P.Cout8 * 8 - cout_off);         // channels of this CTA that exist in the record
      // the record's four slots hold TWO replicas of {sum, sum of squares} here (no centring shift on this path): even /
      // odd tiles add to different addresses, halving the same-address contention of an image's 130 CTAs
      double* gs = P.stats + ((int64_t)n * P.Cout8 * 8 + cout_off) * 4 + 2 * (blockIdx.x & 1);
      for (int i = threadIdx.x - 128; i < nch * 2; i += 256) atomicAdd(gs + (i >> 1) * 4 + (i & 1), (double)s_stats[i]);
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // The image's CTAs meet: CTAs are dispatched in block-index order and an image's CTAs fit the GPU at once (checked by
      // the host), so everything this wait depends on is already running.  A protocol bug traps instead of hanging.
      if (threadIdx.x == 128) {
        if (trace) trace[2] = clock64() - t_entry;                               // (fused) statistics merged, arriving
        red_release_gpu_add(P.sync + n, 1u);
        const uint32_t total = gridDim.x * gridDim.z;
        const long long t0 = clock64();
        while (ld_acquire_gpu(P.sync + n) < total) {
          __nanosleep(128);
          if (clock64() - t0 > (1ll << 32)) __trap();
        }
        if (trace) trace[7] = clock64() - t_entry;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // per-channel (rstd, -mean * rstd): fp64 mean / variance exactly as nhvr_in_apply forms them
      for (int i = threadIdx.x - 128; i < P.Npad; i += 256) {
        float sc = 0.f, sh = 0.f;
        if (i < nch) {
          const double2* rec = reinterpret_cast<const double2*>(P.stats + ((int64_t)n * P.Cout8 * 8 + cout_off + i) * 4);
          const double2 q = __ldcg(rec), r = __ldcg(rec + 1);
          const double m0 = (q.x + r.x) * (double)P.inv_hw;
          const double var = fmax((q.y + r.y) * (double)P.inv_hw - m0 * m0, 0.0);
          sc = rsqrtf((float)var + P.eps);
          sh = -(float)m0 * sc;
        }
        reinterpret_cast<float2*>(s_stats)[i] = make_float2(sc, sh);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (trace && threadIdx.x == 128) trace[6] = clock64() - t_entry;          // (fused) scale / shift ready
      // Pass 2: normalise from the fp32 accumulators, activation, + residual, and write the consumer's format: the pixel
      // itself plus its mirror images in the halo (ReflectionPad2d: up to 3 rows x 3 columns for tiny images, 1 x 1 inside).
      int tyv[3], txv[3], nty = 0, ntx = 0;
      if (valid_m) {
        tyv[nty++] = y + og.pad_t;
        txv[ntx++] = x + og.pad_l;
        if (og.halo != NHVR_HALO_ZERO) {
          if (y >= 1 && y <= og.pad_t) tyv[nty++] = og.pad_t - y;
          if (y <= og.H - 2 && og.pad_t + 2 * (og.H - 1) - y < og.Hp) tyv[nty++] = og.pad_t + 2 * (og.H - 1) - y;
          if (x >= 1 && x <= og.pad_l) txv[ntx++] = og.pad_l - x;
          if (x <= og.W - 2 && og.pad_l + 2 * (og.W - 1) - x < og.Wp) txv[ntx++] = og.pad_l + 2 * (og.W - 1) - x;
        }
      }
      const int64_t u_main = valid_m ? plane_unit(og, y + og.pad_t, x + og.pad_l) : 0;
      const bool mirrors = nty * ntx > 1;
      const int64_t res_u = (P.res && valid_m && !P.res_bulk) ? plane_unit(sg, y + sg.pad_t, x + sg.pad_l) : 0;
      uint4* o = reinterpret_cast<uint4*>(P.out);
      int step = 0, phase = 0;
      if (half < ngroups) tmem_ld16(t_lane + (uint32_t)(half * 16), vr);
      for (int g = half; g < ngroups; g += 2, ++step) {
        const int c0 = cout_off + g * 16;
        const int pb = hilo ? (c0 >> 4) << 2 : c0 >> 3;               // first physical plane of this 16-channel group
        const bool have = pb + npl <= og.C8;                          // Cout8 is even: both logical planes exist or neither
        uint4 rr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = make_uint4(0, 0, 0, 0);
        if (P.res_bulk) {
          if (step == (phase + 1) * P.res_sp) {                       // next phase: everyone is done with the previous tile part
            ++phase;
            asm volatile("bar.sync 3, 288;" ::: "memory");             // releases warp 3, which stages the next part
          }
          if (step == phase * P.res_sp) mbar_wait_warp(res_bar, (uint32_t)(phase & 1));
          if (have) {
            const uint4* rs = reinterpret_cast<const uint4*>(a_smem + (size_t)(g - 2 * phase * P.res_sp) * npl * 2048) + m;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < npl) rr[j] = rs[j * 128];
          }
        } else if (P.res && valid_m && have) {
          const uint4* rp = P.res + ((int64_t)n * sg.C8 + pb) * sg.plane_units + res_u;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < npl) rr[j] = __ldg(rp + (int64_t)j * sg.plane_units);
        }
        tmem_ld_wait();
        float t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 nrm = reinterpret_cast<const float2*>(s_stats)[g * 16 + i];
          t[i] = fmaf(__uint_as_float(vr[i]) * accs, nrm.x, nrm.y);
        }
        if (g + 2 < ngroups) tmem_ld16(t_lane + (uint32_t)((g + 2) * 16), vr);   // next group's accumulators in flight
        if (!(valid_m && have)) continue;
        if (P.act == NHVR_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) t[i] = fmaxf(t[i], 0.f);
        } else if (P.act == NHVR_ACT_LRELU02) {
#pragma unroll
          for (int i = 0; i < 16; ++i) t[i] = t[i] > 0.f ? t[i] : 0.2f * t[i];
        }
        if (P.res) {
          float a[8], b[8];
          unpack8f(rr[0], a, P.f16); unpack8f(rr[1], b, P.f16);
#pragma unroll
          for (int i = 0; i < 8; ++i) { t[i] += a[i]; t[8 + i] += b[i]; }
          if (hilo) {
            unpack8f(rr[2], a, P.f16); unpack8f(rr[3], b, P.f16);
#pragma unroll
            for (int i = 0; i < 8; ++i) { t[i] += a[i]; t[8 + i] += b[i]; }
          }
        }
        uint4 u[4];
        if (hilo) {
          split_hilo(t[0], t[1], P.f16, u[0].x, u[2].x);   split_hilo(t[2], t[3], P.f16, u[0].y, u[2].y);
          split_hilo(t[4], t[5], P.f16, u[0].z, u[2].z);   split_hilo(t[6], t[7], P.f16, u[0].w, u[2].w);
          split_hilo(t[8], t[9], P.f16, u[1].x, u[3].x);   split_hilo(t[10], t[11], P.f16, u[1].y, u[3].y);
          split_hilo(t[12], t[13], P.f16, u[1].z, u[3].z); split_hilo(t[14], t[15], P.f16, u[1].w, u[3].w);
        } else {
          u[0].x = pack2(t[0], t[1], P.f16);  u[0].y = pack2(t[2], t[3], P.f16);
          u[0].z = pack2(t[4], t[5], P.f16);  u[0].w = pack2(t[6], t[7], P.f16);
          u[1].x = pack2(t[8], t[9], P.f16);  u[1].y = pack2(t[10], t[11], P.f16);
          u[1].z = pack2(t[12], t[13], P.f16); u[1].w = pack2(t[14], t[15], P.f16);
          u[2] = u[3] = make_uint4(0, 0, 0, 0);
        }
        uint4* ob = o + ((int64_t)n * og.C8 + pb) * og.plane_units;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < npl) ob[u_main + (int64_t)j * og.plane_units] = u[j];
        if (mirrors) {                                                // border pixels only
#pragma unroll
          for (int iy = 0; iy < 3; ++iy) {
            if (iy >= nty) break;
#pragma unroll
            for (int ix = 0; ix < 3; ++ix) {
              if (ix >= ntx) break;
              if (iy + ix == 0) continue;
              uint4* od = ob + plane_unit(og, tyv[iy], txv[ix]);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < npl) od[(int64_t)j * og.plane_units] = u[j];
            }
          }
        }
      }
    } else {
    for (int rep = 0; rep < ((P.debug & 8) ? 0 : P.mrep); ++rep) {
      int yr = y, xr = x;
      if (rep) {
        if (P.xtiles) { yr = y + rep; }
        else { const int64_t q = q0 + (int64_t)rep * P.q_mstride + m; yr = (int)(q / P.Wrow); xr = (int)(q - (int64_t)yr * P.Wrow); }
      }
      const bool valid_r = (yr < P.Hv) && (xr < P.Wv);
      for (int a = 0; a < P.nacc; ++a) {
        const int Y = yr * P.oys + P.oy[a];
        const int X = xr * P.oxs + P.ox[a];
        const bool valid = valid_r && Y < P.Ho && X < P.Wo;
        for (int g = half; g < ngroups; g += 2) {
          uint32_t vr[16];
          tmem_ld16(t_lane + (uint32_t)(rep * P.acc_mstride + a * P.Npad + g * 16), vr);
          tmem_ld_wait();
          float v[16];
          if (scaled) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(vr[i]) * accs;
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(vr[i]);
          }
          emit16<V>(P, v, cout_off + g * 16, valid, n, Y, X, g, lane, s_stats, s_bias + g * 16);
        }
      }
    }
    }
    if (P.epilogue == NHVR_EPI_RAW_STATS) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // per-CTA partial sums (fp32, <= a few hundred values each) are merged in fp64: the variance E[x^2] - mean^2 of a
      // nearly constant channel (stick-figure pose maps: mean^2 / var ~ 250 at the stem) cancels 2-3 digits
      double* gs = P.stats + ((int64_t)n * P.Cout8 * 8 + cout_off) * 4;       // {sum, sum of squares, shift, -} per channel
      const int lim = (P.rowmode ? P.Cout8 * 8 : min(P.Npad, P.Cout8 * 8 - cout_off)) * 2;
      for (int i = threadIdx.x - 128; i < lim; i += 256) atomicAdd(gs + (i >> 1) * 4 + (i & 1), (double)s_stats[i]);
    }
  }

  if (trace && threadIdx.x == 128) trace[5] = clock64() - t_entry;
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();     // PAIR: the leader's MMAs read the peer's shared memory and TMEM
  if (warp == 3) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)P.tmem_cols); else tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// -------------------------------------------------------------------------------------------------
// weight packing: OIHW (or IOHW for transposed) fp32 -> per-MMA blocks [2][Npad][8] bf16 in
// consumption order (chunk, job, k-step); see header comment.
// split precision: |w| max over the weight tensor (float bits compare like unsigned integers for non-negative values)
__global__ void conv_weight_absmax_kernel(const float* __restrict__ w, int64_t n, uint32_t* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m < INFINITY) atomicMax(out, __float_as_uint(m));
}

NHVR_DEVINL void conv_pack_weights_body(const PackParams& P, int64_t u0, int64_t ustride) {
  const int64_t total = (int64_t)P.nsplit * P.nblocks_padded * 2 * P.Npad;
  // split precision: the lo parts of N(0, 0.02)-sized weights would be fp16 subnormals (18-19 significant bits instead
  // of 22), so the whole tensor is scaled by a power of two that puts max|w| into [8192, 16384); the conv epilogue
  // multiplies the accumulators by the inverse (exact).  tail = {max|w| bits, 2^-s} behind the packed blocks.
  float wscale = 1.f;
  if (P.split3) {
    const float amax = __uint_as_float(*reinterpret_cast<const uint32_t*>(P.tail));
    const int sh = amax > 0.f ? 13 - ilogbf(amax) : 0;
    wscale = ldexpf(1.f, sh);
    if (u0 == 0) P.tail[1] = ldexpf(1.f, -sh);
  }
  // weight blocks per (chunk, tap): kcp/2 plane pairs; split precision: 2 per group of four physical planes - w_hi (feeds the
  // x_hi and the x_lo MMA) and w_lo (feeds x_hi)
  // split precision without w_lo (PackParams::nowlo): only the w_hi block of every group of four physical planes
  const int qsteps = P.kfold ? 1 : (P.split3 ? (P.nowlo ? 1 : 2) * (P.kcp >> 2) : P.kcp >> 1);
  const int nblocks = P.nchunks * P.njobs * qsteps;
  for (int64_t u = u0; u < total; u += ustride) {
    const int nrow = (int)(u % P.Npad);
    int64_t t = u / P.Npad;
    const int kp = (int)(t & 1); t >>= 1;
    const int blk = (int)(t % P.nblocks_padded);
    const int z = (int)(t / P.nblocks_padded);
    uint32_t packed[4] = {0, 0, 0, 0};
    if (blk < nblocks) {
      const int qq = blk % qsteps;
      const int j = (blk / qsteps) % P.njobs;
      const int c = blk / (qsteps * P.njobs);
      int tap = P.flip ? (P.kh * P.kw - 1 - P.job_tap[j]) : P.job_tap[j];
      int co = z * P.Npad + nrow;
      if (P.rowmode) {                       // job_tap = r*8 + accumulator; column n = s*Cp + co
        const int r = P.job_tap[j] >> 3, ncol = (P.job_tap[j] & 7) * P.Npad + nrow;
        const int sfl = ncol / P.Cp;
        co = (sfl < P.kw) ? (ncol - sfl * P.Cp) : P.Cout;       // columns beyond kw*Cp are padding
        tap = r * P.kw + min(sfl, P.kw - 1);
      }
      if (P.kfold) {                         // K group kp = the next filter column of the same 8-channel plane
        if (P.job_tap[j] % P.kw + kp >= P.kw) co = P.Cout;
        else tap += kp;
      }
      float vals[8];
      // split precision: block qq = 2*g + t of a chunk covers logical planes 2*(c*kcp/4 + g) + kp; t == 1 carries w_lo
      // khalf (split precision, <= 8 input channels): K group 1 of every MMA is the LO plane of the same 8 channels, so block 0
      // carries w_hi in both K groups (x_hi*w_hi + x_lo*w_hi in ONE MMA) and block 1 carries w_lo in K group 0 only
      const int lplane = P.khalf ? 0 : P.split3 ? 2 * (c * (P.kcp >> 2) + (P.nowlo ? qq : qq / 2)) + kp : c * P.kcp + 2 * qq + kp;
      const bool w_lo = P.split3 && !P.nowlo && (qq & 1);
      if (P.khalf && w_lo && kp == 1) co = P.Cout;          // zero weights: the lo plane meets no w_lo
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ci = P.kfold ? e : lplane * 8 + e;
        float val = 0.f;
        if (ci < P.Cin && co < P.Cout) {
          const int64_t idx = P.transposed ? (((int64_t)ci * P.Cout + co) * (P.kh * P.kw) + tap)
                                           : (((int64_t)co * P.Cin + ci) * (P.kh * P.kw) + tap);
          val = P.w[idx] * wscale;
        }
        vals[e] = val;
      }
      packed[0] = pack2(vals[0], vals[1], P.f16); packed[1] = pack2(vals[2], vals[3], P.f16);
      packed[2] = pack2(vals[4], vals[5], P.f16); packed[3] = pack2(vals[6], vals[7], P.f16);
      if (w_lo) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          packed[i] = pack2(vals[2 * i] - unpack_lo(packed[i], P.f16), vals[2 * i + 1] - unpack_hi(packed[i], P.f16), P.f16);
      }
    }
    int64_t du = u;
    if (P.pair) {              // [split][stage][rank][block in stage][k-group][N/2 rows]
      const int half = P.Npad >> 1;
      const int r = nrow >= half ? 1 : 0;
      const int stage = blk / P.bpb, j = blk - stage * P.bpb;
      du = (int64_t)z * P.nblocks_padded * 2 * P.Npad + ((int64_t)(stage * 2 + r) * P.bpb + j) * P.Npad + (int64_t)kp * half + (nrow - r * half);
    }
    P.dst[du] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

__global__ void conv_pack_weights_kernel(const __grid_constant__ PackParams P) {
  conv_pack_weights_body(P, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// every layer of a network in ONE launch (blockIdx.y = layer; parameter records in device memory): a training step re-packs all
// weights of all networks for the forward and the dgrad convs after every optimiser update - 216 launches of ~6 us each
// (profiles/r02b_ncu_launches_train.csv) otherwise
__global__ void conv_pack_weights_batched_kernel(const PackParams* __restrict__ arr) {
  const PackParams& P = arr[blockIdx.y];
  conv_pack_weights_body(P, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

}  // namespace nhvr

// =================================================================================================
// host side: plan builder + C-ABI
// =================================================================================================
using namespace nhvr;

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline int next_pow2_cols(int c) { int p = 32; while (p < c) p <<= 1; return p; }
// dynamic shared memory per CTA: two co-resident CTAs need 2 * (bytes + 1 KB reserved) <= 228 KB; one CTA may use 227 KB
constexpr long kSmemTwoPerSm = (228 * 1024) / 2 - 1024 - 256;
constexpr long kSmemOnePerSm = 227 * 1024 - 1024;

extern "C" int nhvr_conv_plan_create(const nhvr_conv_desc* d, nhvr_conv_plan** out) {
  if (!d || !out) return NHVR_ERR_NULL;
  if (d->Cin <= 0 || d->Cout <= 0 || d->N <= 0 || d->H <= 0 || d->W <= 0) return NHVR_ERR_SHAPE;
  nhvr_conv_plan* p = new nhvr_conv_plan();
  std::memset(p, 0, sizeof(*p));
  p->d = *d;
  ConvKParams& K = p->kp;
  int gemm_k = d->Cin, gemm_n = d->Cout;          // dgrad swaps them
  if (d->kind == NHVR_CONV_DGRAD_S1) { gemm_k = d->Cout; gemm_n = d->Cin; }
  int C8 = round_up((gemm_k + 7) / 8, 2);   // K = 16 per MMA -> an even number of planes
  // split precision (flag bit 3): hilo input, C8 counts PHYSICAL planes (groups of four), three MMAs per K step
  const bool split3 = (d->flags & 8) != 0;
  const bool nowlo = split3 && (d->flags & 64) != 0;           // flag bit 6: two MMAs per K step (x_hi*w_hi + x_lo*w_hi), no w_lo blocks
  if (split3 && d->kind == NHVR_CONV_DGRAD_S1) { delete p; return NHVR_ERR_UNSUPPORTED; }
  if (split3) C8 *= 2;
  // Tap pairing (flag bit 2, <= 8 input channels, e.g. the 3-channel pose stem): ONE plane; K group 1 of every MMA is the
  // same plane one pixel to the right (LBO = 16 bytes), so an MMA covers two filter columns: kh*ceil(kw/2) MMAs instead
  // of kh*kw on half the slab bytes.
  bool kfold = false;
  {
    const int Cp_row = d->Cout <= 8 ? 8 : round_up(d->Cout, 16);
    const bool row_ok = !(d->flags & 1) && d->kw >= 5 && d->kw <= 8 && d->kw * Cp_row <= 128 && !std::getenv("NHVR_NO_ROWMODE") &&
                        d->epilogue != NHVR_EPI_BIAS_ACT_P8;
    kfold = (d->flags & 4) && !split3 && !(d->flags & 1) && d->kind == NHVR_CONV && d->stride == 1 && d->Cin <= 8 && d->kw >= 2 && !row_ok &&
            !(std::getenv("NHVR_CONV_KFOLD") && std::atoi(std::getenv("NHVR_CONV_KFOLD")) == 0);
    if (kfold) C8 = 1;
  }
  K.C8in = C8;
  nhvr_act_desc& in = p->in_desc;
  in.N = d->N; in.C8 = C8; in.H = d->H; in.W = d->W; in.halo = d->halo; in.split = 0;
  in.hilo = split3 ? 1 : 0;

  struct Tap { int run_key; int shift; int acc; int tap; };
  std::vector<Tap> taps;
  std::vector<std::pair<int, int>> run_specs;   // (g_off, len) before merging, indexed by run_key
  int nacc = 1;
  int Ho = 0, Wo = 0;
  bool rowmode = false;
  int row_npad = 0;
  int key_stride = 1;                            // run-key distance between consecutive input rows (stacked tiles)

  // ---- geometry: the shift program for tiles of `mrep` M blocks (stacked rows or consecutive positions)
  auto build_geometry = [&](int mrep, bool stacked) -> int {
    taps.clear(); run_specs.clear();
    nacc = 1; rowmode = false; row_npad = 0; key_stride = 1;
    in.N = d->N; in.C8 = C8; in.H = d->H; in.W = d->W; in.halo = d->halo; in.split = 0;
    K.rowmode = 0; K.Cp = 0; K.kw = 0;
    const int L = stacked ? kTileM : kTileM * mrep;     // consecutive positions one run has to cover
    const int xrows = stacked ? mrep - 1 : 0;           // extra input rows below the first block's
    K.tile_step = kTileM * mrep;
    if (d->kind == NHVR_CONV && d->stride == 1) {
      in.pad_t = in.pad_b = in.pad_l = in.pad_r = d->pad;
      const int Wp = d->W + 2 * d->pad;
      Ho = d->H + 2 * d->pad - d->kh + 1;
      Wo = d->W + 2 * d->pad - d->kw + 1;
      if (Ho <= 0 || Wo <= 0) return NHVR_ERR_SHAPE;
      // Row mode (wide kernels, few output channels: the 7x7 stems / RGB head): instead of kh*kw MMAs of N = Cout,
      // issue kh MMAs of N = kw*Cp whose column n = s*Cp + co accumulates filter column s un-shifted; the epilogue
      // adds Z[m+s][s][co] over s.  kw x fewer MMAs and A-operand reads; tiles advance by 128-(kw-1) positions.
      const int Cp_row = d->Cout <= 8 ? 8 : round_up(d->Cout, 16);
      // measured (profiles/r01_selftest_v5_rowmode.log): 2x on the 48->4 head; the 16->48 stem loses 2x (two 176-column
      // accumulators need all 512 TMEM columns -> one CTA per SM, epilogue-bound), so only narrow outputs qualify
      rowmode = !(d->flags & 1) && d->kw >= 5 && d->kw <= 8 && d->kw * Cp_row <= 128 && !std::getenv("NHVR_NO_ROWMODE") &&
                d->epilogue != NHVR_EPI_BIAS_ACT_P8;
      if (rowmode) {
        // M replication in row mode: `mrep` consecutive output rows of the same 128-column segment ("stacked") share
        // their input rows (kh + mrep - 1 slab rows instead of mrep * kh); shuffle epilogue, <= 8 output channels only
        if (mrep != 1 && !(stacked && Cp_row == 8)) return NHVR_ERR_UNSUPPORTED;
        const int ntot = d->kw * Cp_row;
        nacc = (ntot + 255) / 256;
        row_npad = round_up((ntot + nacc - 1) / nacc, 16);
        for (int r = 0; r < d->kh + xrows; ++r) run_specs.push_back({r * Wp, kTileM});
        for (int r = 0; r < d->kh; ++r)
          for (int a = 0; a < nacc; ++a) taps.push_back({r, 0, a, r * 8 + a});
        K.rowmode = 1; K.Cp = Cp_row; K.kw = d->kw;
        K.tile_step = kTileM - (d->kw - 1);
      } else {
        for (int r = 0; r < d->kh + xrows; ++r) run_specs.push_back({r * Wp, L + d->kw - 1 + (kfold ? 1 : 0)});
        for (int r = 0; r < d->kh; ++r)
          for (int s = 0; s < d->kw; s += (kfold ? 2 : 1)) taps.push_back({r, s, 0, r * d->kw + s});
      }
      K.Wrow = Wp; K.Hv = Ho; K.Wv = Wo; K.oys = K.oxs = 1;
    } else if (d->kind == NHVR_CONV_DGRAD_S1) {
      // d describes the FORWARD conv (Cin, Cout, k, pad, input H x W).  dX over the padded input extent is the
      // correlation of the output gradient (stored with k-1 zero rows above/below and k-1 zero columns LEFT of
      // every row: consecutive rows share that gap in the linearised image) with the mirrored, transposed taps.
      if (d->stride != 1) return NHVR_ERR_UNSUPPORTED;
      const int Hof = d->H + 2 * d->pad - d->kh + 1, Wof = d->W + 2 * d->pad - d->kw + 1;
      if (Hof <= 0 || Wof <= 0) return NHVR_ERR_SHAPE;
      in.H = Hof; in.W = Wof;
      in.C8 = round_up((d->Cout + 7) / 8, 2);
      in.pad_t = in.pad_b = d->kh - 1; in.pad_l = d->kw - 1; in.pad_r = 0;
      in.halo = NHVR_HALO_ZERO;
      const int Wpi = Wof + d->kw - 1;                 // == W + 2*pad: same pitch as the forward input
      Ho = d->H + 2 * d->pad; Wo = d->W + 2 * d->pad;  // gradient w.r.t. the PADDED forward input
      for (int r = 0; r < d->kh + xrows; ++r) run_specs.push_back({r * Wpi, L + d->kw - 1});
      for (int r = 0; r < d->kh; ++r)
        for (int s = 0; s < d->kw; ++s) taps.push_back({r, s, 0, r * d->kw + s});
      K.Wrow = Wpi; K.Hv = Ho; K.Wv = Wo; K.oys = K.oxs = 1;
    } else if (d->kind == NHVR_CONV && d->stride == 2) {
      in.pad_t = in.pad_b = in.pad_l = in.pad_r = d->pad;
      in.pad_b += d->in_extra_rows;
      in.split = 1;
      ActGeom g = make_geom(in);
      const int Hq = g.Hp / 2, Wq = g.Wp / 2;
      Ho = (d->H + 2 * d->pad - d->kh) / 2 + 1;
      Wo = (d->W + 2 * d->pad - d->kw) / 2 + 1;
      if (Ho <= 0 || Wo <= 0) return NHVR_ERR_SHAPE;
      const int rr_n = (d->kh - 1) / 2 + 1 + xrows;
      // run key = parity-plane * rr_n + (r >> 1)
      for (int pp = 0; pp < 4; ++pp)
        for (int rr = 0; rr < rr_n; ++rr) run_specs.push_back({pp * Hq * Wq + rr * Wq, L + (d->kw - 1) / 2});
      for (int r = 0; r < d->kh; ++r)
        for (int s = 0; s < d->kw; ++s) {
          const int pp = ((r & 1) << 1) | (s & 1);
          taps.push_back({pp * rr_n + (r >> 1), s >> 1, 0, r * d->kw + s});
        }
      K.Wrow = Wq; K.Hv = Ho; K.Wv = Wo; K.oys = K.oxs = 1;
    } else if (d->kind == NHVR_CONV_TRANSPOSE) {
      if (mrep != 1) return NHVR_ERR_UNSUPPORTED;
      const bool k3 = (d->kh == 3 && d->kw == 3 && d->pad == 1), k4 = (d->kh == 4 && d->kw == 4 && d->pad == 2);
      if (d->stride != 2 || !(k3 || k4)) return NHVR_ERR_UNSUPPORTED;
      in.pad_t = in.pad_l = 0; in.pad_b = 1; in.pad_r = 1 + d->in_extra_cols;
      in.halo = NHVR_HALO_ZERO;
      const int Wp = d->W + in.pad_r;
      // out[2i - pad + ky] += x[i] * w[ky]:  k3 p1 -> 2H (output_padding 1);  k4 p2 -> 2H-2 (+ output_padding via out_h)
      Ho = d->out_h > 0 ? d->out_h : (k3 ? 2 * d->H : 2 * d->H - 2);
      Wo = d->out_w > 0 ? d->out_w : (k3 ? 2 * d->W : 2 * d->W - 2);
      if (Ho < 1 || Wo < 1 || Ho > 2 * d->H || Wo > 2 * d->W) return NHVR_ERR_SHAPE;
      run_specs.push_back({0, kTileM + 1});
      run_specs.push_back({Wp, kTileM + 1});
      nacc = 4;
      // phase a (output row parity), M index t <-> output row 2t+a, input row t+di:
      //   k3 p1: a=0 -> (di 0, ky 1);          a=1 -> (0, 2), (1, 0)
      //   k4 p2: a=0 -> (di 1, ky 0), (0, 2);  a=1 -> (1, 1), (0, 3)
      int n_opt[2], o_d[2][2], o_k[2][2];
      if (k3) { n_opt[0] = 1; n_opt[1] = 2; o_d[0][0] = 0; o_k[0][0] = 1; o_d[0][1] = 0; o_k[0][1] = 1; o_d[1][0] = 0; o_k[1][0] = 2; o_d[1][1] = 1; o_k[1][1] = 0; }
      else    { n_opt[0] = 2; n_opt[1] = 2; o_d[0][0] = 1; o_k[0][0] = 0; o_d[0][1] = 0; o_k[0][1] = 2; o_d[1][0] = 1; o_k[1][0] = 1; o_d[1][1] = 0; o_k[1][1] = 3; }
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
          for (int ia = 0; ia < n_opt[a]; ++ia)
            for (int ib = 0; ib < n_opt[b]; ++ib)
              taps.push_back({o_d[a][ia], o_d[b][ib], a * 2 + b, o_k[a][ia] * d->kw + o_k[b][ib]});
      K.Wrow = Wp; K.Hv = (Ho + 1) / 2; K.Wv = (Wo + 1) / 2; K.oys = K.oxs = 2;
      for (int a = 0; a < 4; ++a) { K.oy[a] = a >> 1; K.ox[a] = a & 1; }
    } else {
      return NHVR_ERR_UNSUPPORTED;
    }
    if ((int)taps.size() > kMaxJobs) return NHVR_ERR_UNSUPPORTED;
    K.xstep = rowmode ? kTileM - (d->kw - 1) : kTileM;
    return NHVR_OK;
  };

  // ---- runs: drop unused, sort by offset, merge neighbours that touch / nearly touch.  With stacked tiles block i
  // of a tap reads run (key + i): the slab distance between consecutive rows must be one constant (a_mstride).
  std::vector<int> key_soff;
  int slab = 0;
  auto layout_runs = [&](int mrep, bool stacked) -> int {
    std::vector<int> used(run_specs.size(), 0);
    for (auto& t : taps)
      for (int i = 0; i < (stacked ? mrep : 1); ++i) {
        if (t.run_key + i * key_stride >= (int)run_specs.size()) return NHVR_ERR_UNSUPPORTED;
        used[t.run_key + i * key_stride] = 1;
      }
    std::vector<int> order;
    for (size_t i = 0; i < run_specs.size(); ++i) if (used[i]) order.push_back((int)i);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return run_specs[a].first < run_specs[b].first; });
    std::vector<ConvRun> runs;
    key_soff.assign(run_specs.size(), 0);
    slab = 0;
    for (int k : order) {
      const int g = run_specs[k].first, len = run_specs[k].second;
      if (!runs.empty() && g <= runs.back().g_off + runs.back().len + 16) {
        ConvRun& r = runs.back();
        key_soff[k] = r.s_off + (g - r.g_off);
        const int new_len = std::max(r.len, g + len - r.g_off);
        slab += new_len - r.len;
        r.len = new_len;
      } else {
        ConvRun r{g, len, slab};
        key_soff[k] = slab;
        slab += len;
        runs.push_back(r);
      }
    }
    if ((int)runs.size() > kMaxRuns) return NHVR_ERR_UNSUPPORTED;
    K.nruns = (int)runs.size();
    for (int i = 0; i < K.nruns; ++i) K.runs[i] = runs[i];
    K.slab_units = slab;
    K.mrep = mrep;
    K.xtiles = 0;
    K.a_mstride = kTileM; K.q_mstride = kTileM;
    if (stacked && mrep > 1) {
      const int k0 = taps[0].run_key;
      K.a_mstride = key_soff[k0 + key_stride] - key_soff[k0];
      K.q_mstride = K.Wrow;
      for (auto& t : taps)
        for (int i = 1; i < mrep; ++i)
          if (key_soff[t.run_key + i * key_stride] - key_soff[t.run_key] != i * K.a_mstride) return NHVR_ERR_UNSUPPORTED;
    }
    K.xstep = rowmode ? kTileM - (d->kw - 1) : kTileM;
    if (stacked) K.xtiles = (K.Wv + K.xstep - 1) / K.xstep;
    return NHVR_OK;
  };

  // ---- N (Cout) tiling, TMEM columns, shared-memory rings
  int Npad = 0, nsplit = 1;
  auto tile_n = [&](int mrep, int want_split) {
    Npad = rowmode ? row_npad : round_up(gemm_n, 16);
    nsplit = 1;
    while (!rowmode && (Npad > std::min(256, 512 / nacc) || nsplit < want_split)) {
      nsplit *= 2; Npad = round_up((gemm_n + nsplit - 1) / nsplit, 16);
    }
    K.Npad = Npad;
    K.acc_mstride = nacc * Npad;
    K.tmem_cols = next_pow2_cols(mrep * nacc * Npad);
  };
  int kcp = 0, SA = 0, bpb = 0, SB = 0;
  // exact dynamic shared memory of a configuration (must equal what the kernel carves up)
  int pair = 0;
  auto smem_need = [&](int kcp_, int sa, int dv, int sb) -> long {
    return (long)sa * kcp_ * slab * 16 + (long)sb * dv * Npad * (pair ? 16 : 32) + (long)(3 * sa + 3 * sb + 1) * 8 + 16 + (long)Npad * 8 * nacc +
           (long)Npad * 4 + (rowmode ? (K.Cp == 8 ? 2L * 2 * 3 * 7 * 7 * std::min(gemm_n, 8) * 4 : 8704L) : 0) + 128;
  };
  auto try_fit = [&](long limit) -> bool {
    const int b_block = Npad * 32;
    long best_score = -1;
    for (int cand = kfold ? 1 : 8; cand >= (kfold ? 1 : 2); cand -= 2) {
      if (C8 % cand) continue;
      if (split3 && (cand & 3)) continue;                     // hi/lo groups of four physical planes stay in one chunk
      const int nch = C8 / cand;
      const int bpc = kfold ? (int)taps.size() : split3 ? (int)taps.size() * (nowlo ? 1 : 2) * (cand / 4) : (int)taps.size() * cand / 2;   // weight blocks per chunk
      if (bpc > kMaxMma) continue;
      // split precision doubles the slab bytes of a chunk (hi + lo planes) while tripling its MMAs: a single slab stage
      // (the co-resident CTA covers the refill) is allowed there when two stages do not leave room for the weight ring
      bool found = false;
      for (int sa = std::min(2, nch); sa >= (split3 ? 1 : std::min(2, nch)) && !found; --sa) {
        for (int dv = 8; dv >= 1; --dv) {                      // blocks per B stage: a divisor of bpc, stage <= 24 KB
          if (bpc % dv || (long)dv * b_block > (pair ? 49152 : 24576)) continue;
          int sb = 6;
          while (sb >= 2 && smem_need(cand, sa, dv, sb) > limit) --sb;
          if (sb < 2) continue;
          // score: weight bytes in flight, mild preference for >= 3 stages and for fewer, larger chunks
          const long score = (long)std::min(sb, 4) * dv * b_block + (sb >= 3 ? 4096 : 0) + cand * 64 - ((sa < 2 && nch > 1) ? 16384 : 0);
          if (score > best_score) { best_score = score; kcp = cand; SA = sa; bpb = dv; SB = sb; }
          found = true;
          break;
        }
      }
    }
    return best_score >= 0;
  };
  auto count_tiles = [&](int mrep, bool stacked) -> int {
    if (stacked) return ((K.Hv + mrep - 1) / mrep) * ((K.Wv + K.xstep - 1) / K.xstep);
    const int64_t last_q = (int64_t)(K.Hv - 1) * K.Wrow + K.Wv;   // one past the last valid linear position
    return (int)((last_q + K.tile_step - 1) / K.tile_step);
  };

  int st0 = build_geometry(1, false);
  if (st0 != NHVR_OK) { delete p; return st0; }
  // M replication (profiles/r01_conv_ablation.md): a tile of mrep M blocks streams the layer's packed weights once
  // instead of once per 128 positions and amortises the per-CTA prologue.  Candidates keep two CTAs per SM
  // (<= 256 TMEM columns, <= 112 KB).
  int mrep = 1;
  bool stacked = false;
  bool ok = false;
  const bool plain = (d->flags & 1) || rowmode || d->kind == NHVR_CONV_TRANSPOSE || std::getenv("NHVR_CONV_TUNE");
  int force_m = 0, force_split = 0;
  if (const char* e = std::getenv("NHVR_CONV_MREP")) std::sscanf(e, "%d,%d", &force_m, &force_split);
  if (rowmode && K.Cp == 8 && force_m != 1 && !std::getenv("NHVR_CONV_TUNE")) {
    // row mode (RGB + mask head): 2-4 stacked output rows per tile, two CTAs per SM
    int mmax = force_m > 1 ? force_m : 4;
    while (mmax > 1 && (int64_t)(mmax - 1) * K.Wrow + 2 * kTileM > kActSlackUnits) --mmax;
    for (int m = mmax; m >= 2 && !ok; --m) {
      if (build_geometry(m, true) != NHVR_OK) continue;
      if (m * nacc * row_npad > 256) continue;
      K.xstep = kTileM - (d->kw - 1);
      if (!force_m && (int64_t)count_tiles(m, true) * d->N < 148) continue;
      if (layout_runs(m, true) != NHVR_OK) continue;
      tile_n(m, 1);
      ok = try_fit(kSmemTwoPerSm);
      if (ok) { mrep = m; stacked = true; }
    }
  }
  if (!plain && force_m != 1) {
    const int n16 = round_up(gemm_n, 16);
    // measured (profiles/r01_conv_ablation.md): splitting wide outputs over two CTAs to make room for a second M block
    // loses (N = 96 MMAs are shared-memory bound, the input slab is fetched twice) - replicate only where N <= 128
    int ws = 1;
    (void)n16;
    if (force_split > 0) ws = force_split;
    tile_n(1, ws);
    int mmax = std::min(4, 256 / (nacc * Npad));
    if (force_m > 1) mmax = force_m;
    // reads of the last tile run past the image into the tail slack of the buffer
    while (mmax > 1 && (int64_t)(mmax - 1) * K.Wrow + 2 * kTileM > kActSlackUnits) --mmax;
    // experiment: NHVR_CONV_PAIR=2 also pairs M-replicated layers (N >= 64): each CTA then reads A + B/2 per MMA
    // Split precision: on by default (N >= 64) - the weight stream per MMA pair is twice as long there: UV head 2.51 -> 2.34 ms,
    // stride-2 64 -> 128 278 -> 268 us (fp16: +2 % on the UV head, -3...-10 % on stems and stride-2 layers, so opt-in there).
    { const char* pe = std::getenv("NHVR_CONV_PAIR");
      const bool want = pe ? std::atoi(pe) == 2 : split3;
      pair = (want && nacc == 1 && Npad >= 64 && (Npad % 16) == 0) ? 1 : 0; }
    for (int m = mmax; m >= 2 && !ok; --m) {
      const int xt = (K.Wv + kTileM - 1) / kTileM;
      const bool prefer_stk = d->kind != NHVR_CONV_DGRAD_S1 && xt * kTileM * 10 <= K.Wv * 11;
      for (int alt = 0; alt < (prefer_stk ? 2 : 1) && !ok; ++alt) {
        const bool stk = prefer_stk && alt == 0;
        if (m * nacc * Npad > 512) continue;
        if (build_geometry(m, stk) != NHVR_OK) continue;
        if (!force_m && (int64_t)count_tiles(m, stk) * d->N * std::max(ws, 1) < 148) continue;   // keep every SM busy
        if (layout_runs(m, stk) != NHVR_OK) continue;
        tile_n(m, ws);
        if (K.tmem_cols <= 256) ok = try_fit(kSmemTwoPerSm);
        else if (force_m) ok = try_fit(kSmemOnePerSm);
        if (ok) { mrep = m; stacked = stk; }
      }
    }
  }
  if (!ok) {
    pair = 0;
    mrep = 1; stacked = false;
    st0 = build_geometry(1, false);
    if (st0 == NHVR_OK) st0 = layout_runs(1, false);
    if (st0 != NHVR_OK) { delete p; return st0; }
    // transposed convs hold 4 phase accumulators: wide outputs are split over CTAs so that 4*N <= 256 TMEM columns and
    // two CTAs share an SM (the epilogue of 4 x 128 x N outputs is otherwise fully exposed); 256->128: 419 -> 448 TFLOP/s
    int tsplit = 1;
    if (d->kind == NHVR_CONV_TRANSPOSE && !(d->flags & 1) && !(std::getenv("NHVR_CONVT_SPLIT") && std::atoi(std::getenv("NHVR_CONVT_SPLIT")) == 0))
      while (nacc * round_up((gemm_n + tsplit - 1) / tsplit, 16) > 256 && round_up((gemm_n + tsplit - 1) / tsplit, 16) > 32) tsplit *= 2;
    if (force_m == 1 && force_split > tsplit) tsplit = force_split;          // experiments: NHVR_CONV_MREP=1,<nsplit>
    tile_n(1, tsplit);
    // CTA pairs (cta_group::2) for the wide layers: one accumulator, N >= 96 (below that the A operand dominates the
    // shared-memory traffic and M replication is the better tool), not for the plain lowering wgrad mirrors
    {
      // measured (profiles/r01_conv_ablation.md): +4 % at N = 192, -14 % at N = 256 (the relay hop costs more than the
      // halved weight stream saves once the MMA is 128 cycles long) -> default only for 128 < N <= 192;
      // NHVR_CONV_PAIR=1 forces every eligible layer, =0 disables
      const char* pe = std::getenv("NHVR_CONV_PAIR");
      const bool eligible = nacc == 1 && !rowmode && !(d->flags & 1) && Npad >= 96 && (Npad % 16) == 0;
      // split precision: also at N = 256 (the weight stream is twice as long per MMA block pair: 435 -> 419 us per 256 -> 256 layer)
      const int pair_max = split3 ? 256 : 192;
      pair = eligible && (pe ? std::atoi(pe) != 0 : (Npad > 128 && Npad <= pair_max)) ? 1 : 0;
    }
    // flag bit 5 (fused InstanceNorm epilogue): shrink the tile step so that an image has exactly one tile per SM (148) -
    // whole images then fill the resident slots (2 per SM) and no CTA waits a full round for the rest of its image
    if ((d->flags & 32) && !rowmode && d->kind == NHVR_CONV && d->stride == 1 && !std::getenv("NHVR_NO_TILE_ALIGN")) {
      const int64_t last_q = (int64_t)(K.Hv - 1) * K.Wrow + K.Wv;
      const int sms = 148;
      const int step = (int)((last_q + sms - 1) / sms);
      if (step <= kTileM && step >= 96) K.tile_step = step;
    }
    const int b_block = Npad * 32;
    if (const char* tune = split3 ? nullptr : std::getenv("NHVR_CONV_TUNE")) {   // experiments: "kcp,SA,bpb,SB"
      int a, b, c, e;
      if (std::sscanf(tune, "%d,%d,%d,%d", &a, &b, &c, &e) == 4 && a >= 2 && (a % 2) == 0 && C8 % a == 0 && b >= 1 && c >= 1 && e >= 2 &&
          ((int)taps.size() * a / 2) % c == 0 && (int)taps.size() * a / 2 <= kMaxMma) {
        const long need = (long)std::min(b, C8 / a) * a * slab * 16 + (long)e * c * b_block + 2048 + (long)Npad * 8;
        if (need <= 227 * 1024) { kcp = a; SA = std::min(b, C8 / a); bpb = c; SB = e; ok = true; }
      }
    }
    if (const char* lim = std::getenv("NHVR_CONV_SMEM_LIMIT")) { if (!ok) ok = try_fit(std::atol(lim)); }   // experiments
    if (!ok && K.tmem_cols <= 256) ok = try_fit(kSmemTwoPerSm);
    if (!ok) ok = try_fit(kSmemOnePerSm);
    if (!ok) { delete p; return NHVR_ERR_SMEM; }
  }
  // ---- jobs (grouped by accumulator so that "first" is well defined)
  std::vector<Tap> staps = taps;
  std::stable_sort(staps.begin(), staps.end(), [](const Tap& a, const Tap& b) { return a.acc < b.acc; });
  K.njobs = (int)staps.size();
  std::vector<ConvJob> jobs(K.njobs);
  int prev_acc = -1;
  for (int j = 0; j < K.njobs; ++j) {
    jobs[j].a_off = key_soff[staps[j].run_key] + staps[j].shift;
    jobs[j].acc = (int16_t)staps[j].acc;
    jobs[j].first = (staps[j].acc != prev_acc) ? 1 : 0;
    prev_acc = staps[j].acc;
    p->pp.job_tap[j] = (int16_t)staps[j].tap;
  }
  K.nacc = nacc;
  p->njobs_h = K.njobs;
  for (int j = 0; j < K.njobs; ++j) p->jobs_h[j] = jobs[j];

  p->nsplit = nsplit;
  p->Ho = Ho; p->Wo = Wo;
  // P8 outputs carry an even number of planes (zero channels beyond Cout) so that they can feed the next
  // conv directly: one MMA consumes K = 16 channels = 2 planes
  p->Cout8 = round_up((gemm_n + 7) / 8, 2);
  K.Ho = Ho; K.Wo = Wo; K.Cout = gemm_n; K.Cout8 = p->Cout8;
  K.epilogue = d->epilogue; K.act = d->act;

  K.kcp = kcp; K.SA = SA; K.bpb = bpb; K.SB = SB;
  K.pair = pair;
  K.nchunks = C8 / kcp;
  // split precision with a single logical input plane (<= 8 channels: the pose stem): the second K group of an MMA would be a zero
  // plane; address the LO plane there instead (A descriptor K-group stride = two slab planes) - two MMAs per tap instead of three
  const bool khalf = split3 && !kfold && !rowmode && gemm_k <= 8 && C8 == 4 && kcp == 4 && !std::getenv("NHVR_NO_KHALF");
  const int ksteps = kfold ? 1 : split3 ? (nowlo ? 1 : 2) * (kcp / 4) : kcp / 2;      // weight blocks per (chunk, tap)
  K.mmas_per_chunk = K.njobs * ksteps;
  K.a_lbo_units = kfold ? 1 : khalf ? 2 * slab : slab;
  K.stages_per_chunk = K.mmas_per_chunk / bpb;           // bpb divides mmas_per_chunk by construction
  K.nblocks = K.nchunks * K.mmas_per_chunk;
  K.nbstages = K.nchunks * K.stages_per_chunk;
  const int nblocks_padded = K.nblocks;
  for (int j = 0; j < K.njobs; ++j)
    for (int q = 0; q < ksteps; ++q) {
      ConvMma& m = K.mma[j * ksteps + q];
      // A planes of step q inside the chunk slab: plane pair q, or (split precision) the hi planes of group q / 2; the w_hi
      // block (q even) also feeds the lo planes (second MMA)
      const int aplane = split3 ? 4 * (nowlo ? q : q / 2) : 2 * q;
      m.a_off = jobs[j].a_off + aplane * slab;
      m.a_off2 = (split3 && !khalf && (nowlo || (q & 1) == 0)) ? jobs[j].a_off + (aplane + 2) * slab : -1;
      m.meta = (uint32_t)(jobs[j].acc * Npad) | ((jobs[j].first && q == 0) ? 0x10000u : 0u);
    }
  K.w_split_units = (int64_t)nblocks_padded * 2 * Npad;
  p->weight_bytes = (size_t)nsplit * K.w_split_units * 16 + (split3 ? 16 : 0);   // split precision: + {max|w|, 2^-s} tail
  p->smem_bytes = (size_t)smem_need(kcp, SA, bpb, SB);

  if (!(d->kind == NHVR_CONV && d->stride == 2)) in.pad_b += d->in_extra_rows;   // plain formats: only the plane stride grows
  ActGeom gin = make_geom(in);
  K.in_plane_units = gin.plane_units;
  p->tiles_per_img = count_tiles(mrep, stacked);

  PackParams& PP = p->pp;
  PP.Cin = gemm_k; PP.Cout = gemm_n; PP.kh = d->kh; PP.kw = d->kw;
  // weight tensor layout [GEMM-K channel][GEMM-N channel][kh][kw]: ConvTranspose2d weights, and the forward
  // Conv2d weight [Cout][Cin] seen from its dgrad (K = Cout, N = Cin)
  PP.transposed = (d->kind == NHVR_CONV_TRANSPOSE || d->kind == NHVR_CONV_DGRAD_S1);
  PP.flip = (d->kind == NHVR_CONV_DGRAD_S1);
  PP.rowmode = rowmode ? 1 : 0; PP.Cp = K.Cp;
  PP.kcp = kcp; PP.nchunks = K.nchunks; PP.njobs = K.njobs; PP.Npad = Npad; PP.nsplit = nsplit;
  PP.nblocks_padded = nblocks_padded;
  PP.pair = pair; PP.bpb = bpb;
  PP.kfold = kfold ? 1 : 0;
  PP.split3 = split3 ? 1 : 0;
  PP.nowlo = nowlo ? 1 : 0;
  PP.khalf = khalf ? 1 : 0;
  K.stat_centred = (d->flags & 16) ? 1 : 0;
  K.out_hilo = (split3 && (d->epilogue == NHVR_EPI_RAW_STATS || d->epilogue == NHVR_EPI_RAW_P8)) ? 1 : 0;
  *out = p;
  return NHVR_OK;
}

extern "C" void nhvr_conv_plan_destroy(nhvr_conv_plan* p) { delete p; }

extern "C" int nhvr_conv_input_desc(const nhvr_conv_plan* p, nhvr_act_desc* in_desc) {
  if (!p || !in_desc) return NHVR_ERR_NULL;
  *in_desc = p->in_desc;
  return NHVR_OK;
}
// Accept a caller-supplied input descriptor that is layout-compatible with the plan's own (same pitch, top /
// left halo, split and channel planes) but taller at the bottom: gradient buffers shared with a wgrad plan carry
// extra zero rows (see nhvr_wgrad_grad_desc).  Only the plane stride changes.
extern "C" int nhvr_conv_plan_set_input_desc(nhvr_conv_plan* p, const nhvr_act_desc* desc) {
  if (!p || !desc) return NHVR_ERR_NULL;
  const ActGeom a = make_geom(p->in_desc), b = make_geom(*desc);
  if (a.N != b.N || a.C8 != b.C8 || a.Wp != b.Wp || a.pad_t != b.pad_t || a.pad_l != b.pad_l || a.split != b.split ||
      a.H != b.H || a.W != b.W || b.Hp < a.Hp)
    return NHVR_ERR_SHAPE;
  if (a.split && b.Hp != a.Hp) return NHVR_ERR_SHAPE;   // parity-plane offsets depend on Hp: create the plan with in_extra_rows
  p->in_desc = *desc;
  p->kp.in_plane_units = b.plane_units;
  return NHVR_OK;
}

extern "C" int nhvr_conv_output_dims(const nhvr_conv_plan* p, int32_t* Ho, int32_t* Wo, int32_t* Cout8) {
  if (!p) return NHVR_ERR_NULL;
  if (Ho) *Ho = p->Ho;
  if (Wo) *Wo = p->Wo;
  if (Cout8) *Cout8 = p->Cout8;
  return NHVR_OK;
}
extern "C" size_t nhvr_conv_weight_bytes(const nhvr_conv_plan* p) { return p ? p->weight_bytes : 0; }
extern "C" double nhvr_conv_flops(const nhvr_conv_plan* p) {
  if (!p) return 0.0;
  const nhvr_conv_desc& d = p->d;
  const double px = d.kind == NHVR_CONV_TRANSPOSE ? (double)d.H * d.W
                    : d.kind == NHVR_CONV_DGRAD_S1 ? (double)(d.H + 2 * d.pad - d.kh + 1) * (d.W + 2 * d.pad - d.kw + 1)
                                                   : (double)p->Ho * p->Wo;
  return 2.0 * d.kh * d.kw * d.Cin * d.Cout * px * d.N;
}
// introspection used by tests / DESIGN.md tables
extern "C" int nhvr_conv_plan_info(const nhvr_conv_plan* p, int32_t* info, int32_t n) {
  if (!p || !info) return NHVR_ERR_NULL;
  const ConvKParams& K = p->kp;
  const int32_t vals[] = {K.kcp, K.nchunks, K.njobs, K.nruns, K.nacc, K.slab_units, K.Npad, K.bpb, K.nbstages,
                          K.SA, K.SB, K.tmem_cols, (int32_t)p->smem_bytes, p->tiles_per_img, p->nsplit, K.nblocks};
  for (int i = 0; i < n && i < (int)(sizeof(vals) / sizeof(vals[0])); ++i) info[i] = vals[i];
  return NHVR_OK;
}

extern "C" int nhvr_conv_pack_weights(const nhvr_conv_plan* p, const float* w, void* packed, void* stream) {
  if (!p || !w || !packed) return NHVR_ERR_NULL;
  if (((uintptr_t)packed & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  PackParams PP = p->pp;
  PP.w = w;
  PP.f16 = operand_f16();
  PP.dst = reinterpret_cast<uint4*>(packed);
  const int64_t total = (int64_t)PP.nsplit * PP.nblocks_padded * 2 * PP.Npad;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
  PP.tail = nullptr;
  if (PP.split3) {
    PP.tail = reinterpret_cast<float*>(reinterpret_cast<uint4*>(packed) + total);
    cudaError_t e0 = cudaMemsetAsync(PP.tail, 0, 16, (cudaStream_t)stream);
    if (e0 != cudaSuccess) { note_cuda_error(e0); return NHVR_ERR_CUDA; }
    const int64_t nw = (int64_t)p->d.Cin * p->d.Cout * p->d.kh * p->d.kw;
    conv_weight_absmax_kernel<<<(int)std::min<int64_t>((nw + 255) / 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(w, nw, reinterpret_cast<uint32_t*>(PP.tail));
    count_launch();
  }
  conv_pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(PP);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

static cudaError_t conv_set_attrs() {
  static bool attr_set = false;
  if (attr_set) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(conv_shiftgemm_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_shiftgemm_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_shiftgemm_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_shiftgemm_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_shiftgemm_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_shiftgemm_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) attr_set = true;
  return e;
}

static inline dim3 conv_grid(const nhvr_conv_plan* p) {
  return dim3(p->kp.pair ? (p->tiles_per_img + 1) & ~1 : p->tiles_per_img, p->d.N, p->nsplit);   // pairs: an even number of tiles
}

// launch with the plan's lowering (CTA pair or not; `variant` selects the kernel instantiation, see the kernel header)
static cudaError_t conv_launch(const nhvr_conv_plan* p, const ConvKParams& K, int variant, cudaStream_t stream) {
  const dim3 grid = conv_grid(p);
  if (!K.pair) {
    if (variant == 2) conv_shiftgemm_kernel<false, 2><<<grid, kThreads, p->smem_bytes, stream>>>(K);
    else if (variant == 1) conv_shiftgemm_kernel<false, 1><<<grid, kThreads, p->smem_bytes, stream>>>(K);
    else conv_shiftgemm_kernel<false, 0><<<grid, kThreads, p->smem_bytes, stream>>>(K);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = p->smem_bytes; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (variant == 2) return cudaLaunchKernelEx(&cfg, conv_shiftgemm_kernel<true, 2>, K);
  if (variant == 1) return cudaLaunchKernelEx(&cfg, conv_shiftgemm_kernel<true, 1>, K);
  return cudaLaunchKernelEx(&cfg, conv_shiftgemm_kernel<true, 0>, K);
}

// NHVR_CONV_TRACE diagnostics: per-CTA cycle breakdown printed to stderr (synchronises)
static void conv_launch_traced(const nhvr_conv_plan* p, ConvKParams& K, int variant, cudaStream_t stream) {
  const dim3 grid = conv_grid(p);
  const size_t nct = (size_t)grid.x * grid.y * grid.z;
  cudaMalloc(&K.trace, nct * 128);
  cudaMemset(K.trace, 0, nct * 128);
  conv_launch(p, K, variant, stream);
  cudaDeviceSynchronize();
  std::vector<long long> h(nct * 16);
  cudaMemcpy(h.data(), K.trace, nct * 128, cudaMemcpyDeviceToHost);
  cudaFree(K.trace);
  double m[16] = {0};
  for (size_t i = 0; i < nct; ++i) for (int j = 0; j < 16; ++j) m[j] += (double)h[i * 16 + j] / nct;
  if (K.epilogue == NHVR_EPI_IN_FUSED)
    std::fprintf(stderr, "[conv trace fused] acc_full@%.0f pass1_done@%.0f stats_merged@%.0f image_met@%.0f norm_ready@%.0f epi_done@%.0f\n", m[4], m[1], m[2], m[7], m[6], m[5]);
  std::fprintf(stderr, "[conv trace] ctas=%zu pair=%d mrep=%d N=%d  first_wait@%.0f  wait_a=%.0f wait_b=%.0f  mma_issued@%.0f  last_b_req@%.0f  acc_full@%.0f  image_met@%.0f  epi_done@%.0f cycles (mean per CTA)\n",
               nct, K.pair, K.mrep, K.Npad, m[0], m[1], m[2], m[3], m[6], m[4], m[7], m[5]);
  if (std::getenv("NHVR_CONV_TRACE_MAP")) {      // placement: SM and start / end time (us at ~1.9 GHz, relative to the first CTA) of every CTA
    long long tmin = h[9];
    for (size_t i = 0; i < nct; ++i) tmin = std::min(tmin, h[i * 16 + 9]);
    for (size_t i = 0; i < nct; ++i)
      std::fprintf(stderr, "[cta] %zu img=%zu sm=%lld start=%.1f acc_full=%.1f end=%.1f\n", i, i / grid.x % grid.y, h[i * 16 + 8], (h[i * 16 + 9] - tmin) / 1900.0,
                   (h[i * 16 + 9] - tmin + h[i * 16 + 4]) / 1900.0, (h[i * 16 + 9] - tmin + h[i * 16 + 5]) / 1900.0);
  }
}

extern "C" size_t nhvr_conv_pack_record_bytes(void) { return sizeof(PackParams); }

extern "C" int nhvr_conv_pack_record_fill(const nhvr_conv_plan* p, const float* w, void* packed, void* record_host, int64_t* units) {
  if (!p || !w || !packed || !record_host) return NHVR_ERR_NULL;
  if (((uintptr_t)packed & 15) != 0) return NHVR_ERR_ALIGN;
  if (p->pp.split3) return NHVR_ERR_UNSUPPORTED;          // the split-precision pack needs the |w| maximum first: per-layer call
  PackParams PP = p->pp;
  PP.w = w;
  PP.f16 = operand_f16();
  PP.dst = reinterpret_cast<uint4*>(packed);
  PP.tail = nullptr;
  std::memcpy(record_host, &PP, sizeof(PP));
  if (units) *units = (int64_t)PP.nsplit * PP.nblocks_padded * 2 * PP.Npad;
  return NHVR_OK;
}

extern "C" int nhvr_conv_pack_weights_batched(const void* records_dev, int32_t n, int64_t max_units, void* stream) {
  if (!records_dev) return NHVR_ERR_NULL;
  if (n <= 0 || max_units <= 0) return NHVR_ERR_SHAPE;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((max_units + 255) / 256, 148 * 8 / std::max(1, std::min(n, 8))));
  conv_pack_weights_batched_kernel<<<dim3(blocks, n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackParams*>(records_dev));
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

extern "C" int nhvr_conv_forward(const nhvr_conv_plan* p, const void* in, const void* packed_w, const float* bias,
                                 void* out, const nhvr_act_desc* out_desc, double* stats, void* stream) {
  if (!p || !in || !packed_w || !out) return NHVR_ERR_NULL;
  if ((((uintptr_t)in | (uintptr_t)packed_w | (uintptr_t)out) & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  ConvKParams K = p->kp;
  if (K.epilogue == NHVR_EPI_RAW_STATS && !stats) return NHVR_ERR_NULL;
  if (K.epilogue == NHVR_EPI_BIAS_ACT_P8) {
    if (!out_desc) return NHVR_ERR_NULL;
    K.og = make_geom(*out_desc);
    if (K.og.H != p->Ho || K.og.W != p->Wo || K.og.N != p->d.N) return NHVR_ERR_SHAPE;
  }
  K.f16 = operand_f16();
  { const char* dbg = std::getenv("NHVR_CONV_DEBUG"); K.debug = dbg ? std::atoi(dbg) : 0; }
  K.in = reinterpret_cast<const uint4*>(in);
  K.w = reinterpret_cast<const uint4*>(packed_w);
  K.bias = bias;
  K.out = out;
  K.stats = stats;
  K.res = nullptr; K.sync = nullptr; K.start_delay = 0;
  K.acc_scale = p->pp.split3 ? reinterpret_cast<const float*>(reinterpret_cast<const uint4*>(packed_w) + (int64_t)p->nsplit * K.w_split_units) + 1 : nullptr;
  { cudaError_t e = conv_set_attrs(); if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; } }
  // variant 1 = the layers that need the special epilogue paths (see the kernel header)
  const bool special = K.acc_scale != nullptr || K.out_hilo || K.stat_centred || (K.rowmode && K.Cp == 8);
  const int variant = p->pp.split3 ? 2 : special ? 1 : 0;
  K.trace = nullptr;
  if (std::getenv("NHVR_CONV_TRACE")) {
    conv_launch_traced(p, K, variant, (cudaStream_t)stream);
    count_launch();
    return NHVR_OK;
  }
  cudaError_t e = conv_launch(p, K, variant, (cudaStream_t)stream);
  count_launch();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}

// ---- conv + InstanceNorm + activation (+ residual) + halo in one kernel (include/nhvr.h) -----------------------------
// structural conditions of the fused epilogue (no device needed)
static bool in_fused_shape_ok(const nhvr_conv_plan* p) {
  const ConvKParams& K = p->kp;
  return K.epilogue == NHVR_EPI_RAW_STATS && K.mrep == 1 && K.nacc == 1 && !K.rowmode && !K.xtiles && !K.stat_centred &&
         K.oys == 1 && K.oxs == 1 && p->d.kind == NHVR_CONV && !std::getenv("NHVR_NO_IN_FUSED");
}

extern "C" int nhvr_conv_in_fused_supported(const nhvr_conv_plan* p) {
  if (!p || !in_fused_shape_ok(p) || !arch_ok_cached()) return 0;
  if (conv_set_attrs() != cudaSuccess) return 0;
  int dev = 0, sms = 0, smem_sm = 0, regs_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev) != cudaSuccess)
    return 0;
  cudaFuncAttributes fa;
  cudaError_t e;
  if (p->pp.split3) e = p->kp.pair ? cudaFuncGetAttributes(&fa, conv_shiftgemm_kernel<true, 2>) : cudaFuncGetAttributes(&fa, conv_shiftgemm_kernel<false, 2>);
  else e = p->kp.pair ? cudaFuncGetAttributes(&fa, conv_shiftgemm_kernel<true, 1>) : cudaFuncGetAttributes(&fa, conv_shiftgemm_kernel<false, 1>);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  // Resident CTAs per SM of kernel variant 1 with this plan's shared memory.  Computed from the device limits (1 KB of
  // shared memory is reserved per CTA; registers are allocated per warp in units of 256; 512 TMEM columns per SM) rather
  // than cudaOccupancyMaxActiveBlocksPerMultiprocessor, which answers for the DEFAULT shared-memory carve-out (1 CTA)
  // while launches of this kernel run with the maximum one (ncu: launch__occupancy_limit_shared_mem = 2,
  // sm__warps_active 22-23 of 24; profiles/r02_ncu_uvres_raw.csv).
  const long regs_cta = (long)((fa.numRegs * 32 + 255) / 256 * 256) * (kThreads / 32);
  const long by_smem = smem_sm / (long)(p->smem_bytes + fa.sharedSizeBytes + 1024);
  const long by_regs = regs_sm / regs_cta;
  const long by_tmem = 512 / p->kp.tmem_cols;
  const long per_sm = std::min(by_smem, std::min(by_regs, by_tmem));
  const long slots = per_sm * sms;
  const dim3 grid = conv_grid(p);
  if (std::getenv("NHVR_DEBUG_FUSED"))
    std::fprintf(stderr, "[in_fused] pair=%d grid=(%u,%u,%u) smem=%zu regs=%d tmem=%d per_sm=%ld (smem %ld regs %ld tmem %ld) slots=%ld\n", p->kp.pair,
                 grid.x, grid.y, grid.z, p->smem_bytes, fa.numRegs, p->kp.tmem_cols, per_sm, by_smem, by_regs, by_tmem, slots);
  // an image's CTAs must be resident together, also while a second fused kernel runs on a concurrent stream
  return (long)grid.x * grid.z <= slots / 2 ? 1 : 0;
}

extern "C" int nhvr_conv_forward_in_fused(const nhvr_conv_plan* p, const void* in, const void* packed_w, double* stats, float eps,
                                          int32_t act, const void* residual, const nhvr_act_desc* res_desc, void* dst,
                                          const nhvr_act_desc* dst_desc, uint32_t* sync, void* stream) {
  if (!p || !in || !packed_w || !stats || !dst || !dst_desc || !sync) return NHVR_ERR_NULL;
  if (residual && !res_desc) return NHVR_ERR_NULL;
  if ((((uintptr_t)in | (uintptr_t)packed_w | (uintptr_t)dst | (uintptr_t)residual) & 15) != 0) return NHVR_ERR_ALIGN;
  if (!arch_ok_cached()) return NHVR_ERR_ARCH;
  if (!in_fused_shape_ok(p)) return NHVR_ERR_UNSUPPORTED;
  if (act != NHVR_ACT_NONE && act != NHVR_ACT_RELU && act != NHVR_ACT_LRELU02) return NHVR_ERR_UNSUPPORTED;
  ConvKParams K = p->kp;
  const int hilo = p->pp.split3 ? 1 : 0;
  K.og = make_geom(*dst_desc);
  if (K.og.N != p->d.N || K.og.H != p->Ho || K.og.W != p->Wo || K.og.hilo != hilo || K.og.C8 != p->Cout8 * (hilo ? 2 : 1)) return NHVR_ERR_SHAPE;
  K.sg = K.og;
  if (residual) {
    K.sg = make_geom(*res_desc);
    if (K.sg.N != K.og.N || K.sg.H != K.og.H || K.sg.W != K.og.W || K.sg.C8 != K.og.C8 || K.sg.hilo != hilo || K.sg.split) return NHVR_ERR_SHAPE;
  }
  K.epilogue = NHVR_EPI_IN_FUSED;
  K.act = act;
  K.f16 = operand_f16();
  K.debug = 0;
  K.in = reinterpret_cast<const uint4*>(in);
  K.w = reinterpret_cast<const uint4*>(packed_w);
  K.bias = nullptr;
  K.out = dst;
  K.stats = stats;
  K.res = reinterpret_cast<const uint4*>(residual);
  // residual tile by bulk copy: the skip activation must linearise like this conv's input (same row pitch, not split), so
  // that a tile's 128 positions are 128 consecutive units of every residual plane
  K.res_bulk = 0; K.res_off = 0; K.res_sp = 1;
  if (residual && !K.sg.split && K.sg.Wp == K.Wrow && !std::getenv("NHVR_NO_RES_BULK")) {
    const long avail = (long)K.SA * K.kcp * K.slab_units * 16 + (long)K.SB * K.bpb * K.Npad * (K.pair ? 16 : 32);
    const int sp = (int)(avail / (2L * (hilo ? 4 : 2) * 2048));
    const int ngroups = K.Npad >> 4;
    // several phases need both epilogue halves to take the same number of steps (they meet at a barrier between phases)
    if (sp >= 1 && (2 * sp >= ngroups || (ngroups & 1) == 0)) { K.res_bulk = 1; K.res_sp = sp; K.res_off = K.sg.pad_t * K.sg.Wp + K.sg.pad_l; }
  }
  K.sync = sync;
  // experiment knob (profiles/r02b_in_fused.md): NHVR_FUSED_DELAY = cycles the launch's second image starts late.  Measured: the
  // delay adds to the launch time one for one - a CTA's MMA loop is latency-bound, two in phase use the tensor pipe better
  // than one alone - so the default is none
  { static int delay = -1;
    if (delay < 0) { const char* e = std::getenv("NHVR_FUSED_DELAY"); delay = e ? std::atoi(e) : 0; }
    K.start_delay = p->d.N >= 2 ? delay : 0; }
  K.eps = eps;
  K.inv_hw = 1.0f / ((float)p->Ho * (float)p->Wo);
  K.acc_scale = p->pp.split3 ? reinterpret_cast<const float*>(reinterpret_cast<const uint4*>(packed_w) + (int64_t)p->nsplit * K.w_split_units) + 1 : nullptr;
  { cudaError_t e = conv_set_attrs(); if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; } }
  K.trace = nullptr;
  const int variant = p->pp.split3 ? 2 : 1;
  if (std::getenv("NHVR_CONV_TRACE")) {
    conv_launch_traced(p, K, variant, (cudaStream_t)stream);
    count_launch();
    return NHVR_OK;
  }
  cudaError_t e = conv_launch(p, K, variant, (cudaStream_t)stream);
  count_launch();
  if (e != cudaSuccess) { note_cuda_error(e); return NHVR_ERR_CUDA; }
  return NHVR_OK;
}
