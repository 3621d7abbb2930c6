// Shared device helpers for the sm_100a kernels: mbarrier, 1-D bulk async copy (UBLKCP),
// tcgen05 (TMEM alloc / MMA / commit / ld) wrappers written as inline PTX.
// Nothing here is library code: these are the raw instructions the conv kernel is built from.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define NHVR_DEVINL __device__ __forceinline__

namespace nhvr {

NHVR_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// one lane of a converged warp (the role loops run warp-uniformly and predicate single-thread instructions)
NHVR_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
NHVR_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
NHVR_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
NHVR_DEVINL void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
NHVR_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
NHVR_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait WITH a suspend-time hint: the waiting thread is descheduled by the hardware until the phase
// completes (or the hint expires) instead of busy-polling and stealing issue slots from the MMA warp.
NHVR_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trap (sticky error on the host), never a hung GPU.
NHVR_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) { __trap(); }
  }
}
// Long waits of whole warps (epilogue waiting for the accumulator): one lane polls, the others park at a
// warp barrier, so at most one thread per warp competes for issue slots.
NHVR_DEVINL void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
  // every lane observes the completed phase once (acquire) before touching the data it guards
  while (!mbar_try_wait(bar, parity)) {}
}

// ------------------------------------------------------------------ bulk async copy (1-D TMA, SASS UBLKCP)
// gmem -> smem, completion signalled on an mbarrier as transaction bytes. src/dst/bytes 16-B aligned.
NHVR_DEVINL void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
NHVR_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
NHVR_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
NHVR_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
NHVR_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
NHVR_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread.
NHVR_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every MMA issued so far by this thread has retired.
NHVR_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------ CTA pair (cta_group::2): two CTAs of one cluster on the
// two SMs of a TPC run ONE M=256 MMA; each supplies its own 128 rows of A and HALF of the N columns of B from its own
// shared memory (same offsets in both CTAs), and holds its own 128 accumulator rows in its own TMEM.
NHVR_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
NHVR_DEVINL void cluster_sync_all() {   // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of this cluster
NHVR_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}
// wait with cluster-scope acquire: the phase was completed by an arrive from the peer CTA
NHVR_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 22)) { __trap(); }
  }
}
NHVR_DEVINL void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair, same smem offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
NHVR_DEVINL void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
NHVR_DEVINL void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA of the pair, issued by ONE thread of the leader CTA (rank 0)
NHVR_DEVINL void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once every MMA issued so far has retired
NHVR_DEVINL void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

// ------------------------------------------------------------------ L2 eviction-priority hints
// A conv's raw output is read exactly once, by the IN-apply that follows it: the conv stores it with evict_last, the
// apply loads it with evict_first (measured: IN-apply 1.67 -> 1.62 ms per step; hinting the apply's stores as well
// gives the gain back).
NHVR_DEVINL uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
NHVR_DEVINL uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
NHVR_DEVINL void st_hint(uint4* addr, const uint4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}
NHVR_DEVINL uint4 ld_hint(const uint4* addr, uint64_t pol) {
  uint4 w;
  asm volatile("ld.global.nc.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(addr), "l"(pol));
  return w;
}

// ------------------------------------------------------------------ inter-CTA arrival counter (fused InstanceNorm epilogue)
NHVR_DEVINL void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
NHVR_DEVINL uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

NHVR_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i).
NHVR_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 4 consecutive fp32 columns (no wait: several loads can be in flight before one tmem_ld_wait)
NHVR_DEVINL void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleaved" canonical layout):
//   element (row r, 16-byte K-chunk k) lives at  start + (r%8)*16 + (r/8)*SBO + k*LBO.
// Bit layout (sm_100 tcgen05 "matrix descriptor"): [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version=1, [49,52) base offset=0, [52] lbo mode=0, [61,64) layout type (0 = no swizzle).
NHVR_DEVINL uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Instruction descriptor for kind::f16: D=f32, A=B=bf16 or fp16, both K-major, dense.
// [4,6) c_format=1(F32), [7,10) a_format (0 F16 / 1 BF16), [10,13) b_format, [15] a_major=0, [16] b_major=0,
// [17,23) N>>3, [24,29) M>>4.
__host__ __device__ inline uint32_t make_idesc_16(uint32_t M, uint32_t N, int f16) {
  const uint32_t fmt = f16 ? 0u : 1u;   // F16 = 0, BF16 = 1
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// 16-bit operand element type of the whole path: bf16 (f16 == 0) or IEEE fp16 (f16 != 0).  Both feed
// tcgen05.mma kind::f16 at the same rate; fp16 keeps 3 more mantissa bits.
NHVR_DEVINL uint32_t pack2(float lo, float hi, int f16) {
  uint32_t r;
  if (f16) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));     // no saturation: overflow -> inf, flagged downstream
  else     asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
NHVR_DEVINL float unpack_lo(uint32_t v, int f16) {
  if (f16) { __half2 h = *reinterpret_cast<__half2*>(&v); return __low2float(h); }
  return __uint_as_float(v << 16);
}
NHVR_DEVINL float unpack_hi(uint32_t v, int f16) {
  if (f16) { __half2 h = *reinterpret_cast<__half2*>(&v); return __high2float(h); }
  return __uint_as_float(v & 0xFFFF0000u);
}

// eight 16-bit values of one P8 unit -> fp32
NHVR_DEVINL void unpack8f(const uint4& u, float (&v)[8], int f16) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = (e & 1) ? unpack_hi(w[e >> 1], f16) : unpack_lo(w[e >> 1], f16);
}

// a 16-byte unit of eight 16-bit values holds an inf / NaN (exponent field all ones)?
NHVR_DEVINL bool unit_nonfinite(const uint4& u, int f16) {
  const uint32_t m = f16 ? 0x7C007C00u : 0x7F807F80u;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t e = w[i] & m;
    bad |= ((e & 0xFFFFu) == (m & 0xFFFFu)) | ((e >> 16) == (m >> 16));
  }
  return bad;
}

// split precision: (a, b) -> hi = rn16 pair, lo = rn16 of the remainders (v = hi + lo to ~22 bits with fp16)
NHVR_DEVINL void split_hilo(float a, float b, int f16, uint32_t& hi, uint32_t& lo) {
  hi = pack2(a, b, f16);
  lo = pack2(a - unpack_lo(hi, f16), b - unpack_hi(hi, f16), f16);
}

}  // namespace nhvr
