// Micro-benchmark: back-to-back tcgen05.mma from resident shared memory (no loads), to separate the tensor
// pipe / operand-fetch rate from the operand pipeline.  Varies N, operand layout (un-swizzled K-major as the
// conv kernel uses it vs SWIZZLE_128B K-major), and A start alignment (shifted views).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../neural-human-video-rendering_b200/csrc/common.cuh"
using namespace nhvr;

__global__ void __launch_bounds__(128, 1) mma_bench(int N, int iters, int mode, int shift_units, int two_acc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_16(128, (uint32_t)N, 0);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
    uint64_t adesc, bdesc;
    if (mode == 0) {   // un-swizzled K-major: rows 16 B apart, SBO 128 B, LBO = plane stride (here 8 KB)
      adesc = make_desc_nosw(a_base + shift_units * 16, 8192, 128);
      bdesc = make_desc_nosw(b_base, (uint32_t)N * 16, 128);
    } else {           // SWIZZLE_128B K-major: rows 128 B, 8-row atoms of 1024 B (SBO), layout type 2
      adesc = ((uint64_t)((a_base & 0x3FFFF) >> 4)) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      bdesc = ((uint64_t)((b_base & 0x3FFFF) >> 4)) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    }
    __syncwarp();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) {
        umma_bf16(tmem + ((two_acc && (i & 1)) ? 256 : 0), adesc + (uint64_t)((i & 3) * 2), bdesc, idesc, 1u);
      }
    }
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (leader && blockIdx.x == 0) cycles[0] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// Cadence experiment: the conv kernel commits to an mbarrier after every weight stage (grp MMAs) and waits on the
// next stage's full barrier.  cadence 1: commit only; 2: + wait on an already-completed barrier + fence;
// 3: full handshake with a second warp that re-arms the stage as soon as the commit arrives (no copies).
__global__ void __launch_bounds__(128, 1) mma_cadence(int N, int iters, int grp, int cadence, int nst, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8], done;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const int nstage_iters = iters / grp;
  if (warp == 2 && cadence == 3) {          // "producer": stage s is full again as soon as it was released
    int st = 0; uint32_t ph = 0;
    for (int s = 0; s < nstage_iters; ++s) {
      mbar_wait(&empty[st], ph ^ 1u);
      if (elect_one()) mbar_arrive(&full[st]);
      __syncwarp();
      if (++st == nst) { st = 0; ph ^= 1u; }
    }
  }
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_16(128, (uint32_t)N, 0);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
    const uint64_t adesc = make_desc_nosw(a_base, 8192, 128);
    const uint64_t bdesc = make_desc_nosw(b_base, (uint32_t)N * 16, 128);
    __syncwarp();
    long long t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int s = 0; s < nstage_iters; ++s) {
      if (cadence == 3) { mbar_wait(&full[st], ph); tc_fence_after(); }
      else if (cadence == 2) { mbar_wait(&done, 1); tc_fence_after(); }      // parity-1 wait on a fresh barrier returns at once
      for (int k = 0; k < grp; ++k)
        if (leader) umma_bf16(tmem, adesc + (uint64_t)((k & 3) * 2), bdesc + (uint64_t)(k * 64), idesc, 1u);
      if (cadence >= 1 && leader) umma_commit(&empty[st]);
      if (++st == nst) { st = 0; ph ^= 1u; }
    }
    if (leader) umma_commit(&done);
    mbar_wait(&done, 0);
    long long t1 = clock64();
    if (leader && blockIdx.x == 0) cycles[0] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// CTA-pair experiment (cta_group::2): M = 256 over two SMs, each CTA supplies its 128 rows of A and half of B.
// Numeric probe: A rows of CTA r hold (r+1), B rows of CTA r hold (r+1) -> D[m][n] = 16*iters*(rank of m + 1)*(half of n + 1).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_pair(int N, int iters, long long* cycles, float* probe) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const __nv_bfloat16 val = __float2bfloat16((float)(rank + 1));
  for (int i = threadIdx.x; i < 160 * 1024 / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(smem)[i] = val;
  if (threadIdx.x == 0) { mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc2(&tmem_ptr, 512); tmem_relinquish2(); }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1 && rank == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_16(256, (uint32_t)N, 0);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
    const uint64_t adesc = make_desc_nosw(a_base, 8192, 128);
    const uint64_t bdesc = make_desc_nosw(b_base, (uint32_t)(N / 2) * 16, 128);   // this CTA's N/2 rows of B
    __syncwarp();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
      if (leader) umma2_bf16(tmem, adesc + (uint64_t)((i & 3) * 2), bdesc, idesc, i ? 1u : 0u);
    if (leader) umma2_commit(&done);
    mbar_wait(&done, 0);
    long long t1 = clock64();
    if (leader && blockIdx.x == 0) cycles[0] = t1 - t0;
    __syncwarp();
  } else {
    if (warp == 1) mbar_wait(&done, 0);      // the peer learns about completion through the multicast commit
  }
  __syncthreads();
  tc_fence_after();
  if (blockIdx.x < 2 && probe) {             // rows 0 and 127 of each CTA, columns 0 and N-1
    if (warp == 0) {
      uint32_t v[16];
      tmem_ld16(tmem, v); tmem_ld_wait();
      if (threadIdx.x == 0) probe[rank * 4 + 0] = __uint_as_float(v[0]);
      tmem_ld16(tmem + (uint32_t)(N - 16), v); tmem_ld_wait();
      if (threadIdx.x == 0) probe[rank * 4 + 1] = __uint_as_float(v[15]);
    }
    if (warp == 3) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)96 << 16), v); tmem_ld_wait();
      if (threadIdx.x == 127) probe[rank * 4 + 2] = __uint_as_float(v[0]);
      tmem_ld16(tmem + ((uint32_t)96 << 16) + (uint32_t)(N - 16), v); tmem_ld_wait();
      if (threadIdx.x == 127) probe[rank * 4 + 3] = __uint_as_float(v[15]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem, 512); }
}

// Does a concurrent bulk-copy stream into shared memory slow the MMAs down?  warp 0 copies `chunk` bytes at a time from
// global memory into a ring (as the conv kernel's weight producer does, `depth` copies in flight), warp 1 issues MMAs on
// resident operands (mode 0) or on B blocks walking through the ring that is being written (mode 1, data irrelevant).
__global__ void __launch_bounds__(128, 1) mma_with_tma(int N, int iters, int chunk, int depth, int mode, int a_lbo, const uint4* __restrict__ src,
                                                       long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], done;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1);
    mbar_init(&done, 1);
    stop = 0;
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  uint8_t* ring = smem + 64 * 1024;             // 128 KB ring region; A at 0, resident B at 32 KB
  if (warp == 2 && chunk > 0) {
    const uint4* g = src + (size_t)blockIdx.x * (1 << 16);     // 1 MB window per CTA (L2 resident)
    int st = 0; uint32_t ph = 0; long long n = 0;
    // prime `depth` copies, then keep the ring full until the MMA warp says stop
    for (int i = 0; i < depth; ++i) {
      if (elect_one()) { mbar_arrive_expect_tx(&full[i], chunk); bulk_g2s(ring + (size_t)i * chunk, g + ((n * chunk / 16) & 0x7fff), chunk, &full[i]); }
      ++n;
    }
    __syncwarp();
    while (!stop) {
      mbar_wait(&full[st], ph);
      if (elect_one()) { mbar_arrive_expect_tx(&full[st], chunk); bulk_g2s(ring + (size_t)st * chunk, g + ((n * chunk / 16) & 0x7fff), chunk, &full[st]); }
      __syncwarp();
      ++n;
      if (++st == depth) { st = 0; ph ^= 1u; }
    }
    for (int i = 0; i < depth; ++i) { mbar_wait(&full[st], ph); if (++st == depth) { st = 0; ph ^= 1u; } }   // drain
    if (blockIdx.x == 0 && elect_one()) cycles[1] = n * chunk;
  }
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_16(128, (uint32_t)N, 0);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 32 * 1024), r_base = smem_u32(ring);
    const uint64_t adesc = make_desc_nosw(a_base, (uint32_t)a_lbo, 128);
    const uint32_t bblk = (uint32_t)N * 32;
    const int nring = (depth * chunk) / (int)bblk > 0 ? (depth * chunk) / (int)bblk : 1;
    __syncwarp();
    long long t0 = clock64();
    int rb = 0;
    for (int i = 0; i < iters; ++i) {
      const uint64_t bdesc = make_desc_nosw((mode ? r_base + (uint32_t)rb * bblk : b_base), (uint32_t)N * 16, 128);
      if (leader) umma_bf16(tmem, adesc + (uint64_t)((i % 9) * 3), bdesc, idesc, 1u);
      if (++rb == nring) rb = 0;
    }
    if (leader) umma_commit(&done);
    mbar_wait(&done, 0);
    long long t1 = clock64();
    stop = 1;
    if (leader && blockIdx.x == 0) cycles[0] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  printf("%-10s %-6s %-6s %-8s %10s %12s\n", "layout", "N", "shift", "accs", "cyc/MMA", "TFLOP/s@148");
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 48, 80, 96, 192, 256})
      for (int shift : {0, 1})
        for (int two : {0, 1}) {
          if (mode == 1 && shift) continue;
          if (two && N > 256) continue;
          long long h = 0;
          for (int rep = 0; rep < 2; ++rep) {
            mma_bench<<<148, 128, 200 * 1024>>>(N, iters, mode, shift, two, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          const double cyc = (double)h / iters;
          printf("%-10s %-6d %-6d %-8d %10.1f %12.1f\n", mode ? "sw128" : "nosw", N, shift, two + 1, cyc,
                 2.0 * 128 * N * 16 / cyc * 1.9e9 * 148 / 1e12);
        }
  {
    cudaFuncSetAttribute(mma_with_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    uint4* src; cudaMalloc(&src, (size_t)148 << 20); cudaMemset(src, 0, (size_t)148 << 20);
    long long* d2; cudaMalloc(&d2, 16);
    printf("\nMMA (M=128) with a concurrent bulk-copy stream into shared memory: N, chunk bytes x depth, mode (1 = B walks the ring)\n");
    for (int N : {192, 256})
     for (int a_lbo : {6240, 6272, 8192})
      for (int chunk : {0, 18432})
        for (int depth : {2, 6})
          for (int mode : {0, 1}) {
            if (chunk == 0 && (depth != 2 || mode)) continue;
            const int it2 = 8192;
            long long h[2] = {0, 0};
            cudaMemset(d2, 0, 16);
            for (int rep = 0; rep < 2; ++rep) {
              mma_with_tma<<<148, 128, 200 * 1024>>>(N, it2, chunk, depth, mode, a_lbo, src, d2);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("tma error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d2, 16, cudaMemcpyDeviceToHost);
            printf("N=%d a_lbo=%d chunk=%5d depth=%d mode=%d  %.1f cyc/MMA   copy %.1f B/clk\n", N, a_lbo, chunk, depth, mode, (double)h[0] / it2, (double)h[1] / (double)h[0]);
          }
  }
  {
    cudaFuncSetAttribute(mma_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    float* probe; cudaMalloc(&probe, 64); cudaMemset(probe, 0, 64);
    printf("\nCTA pair (cta_group::2), M=256: cyc/MMA and probe D values [rank][row0col0,row0colN-1,row127col0,row127colN-1]\n");
    for (int N : {96, 192, 256}) {
      const int it2 = 1024;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        mma_pair<<<148, 128, 200 * 1024>>>(N, it2, d, probe);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pair error %s\n", cudaGetErrorString(e)); return 1; }
      }
      float hp[8];
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(hp, probe, 32, cudaMemcpyDeviceToHost);
      printf("N=%d  %.1f cyc/MMA (M=256: %.1f TFLOP/s @148)  probe/(16*iters): r0 [%.2f %.2f %.2f %.2f] r1 [%.2f %.2f %.2f %.2f]\n", N, (double)h / it2,
             2.0 * 256 * N * 16 / ((double)h / it2) * 1.9e9 * 74 / 1e12, hp[0] / (16.f * it2), hp[1] / (16.f * it2), hp[2] / (16.f * it2),
             hp[3] / (16.f * it2), hp[4] / (16.f * it2), hp[5] / (16.f * it2), hp[6] / (16.f * it2), hp[7] / (16.f * it2));
    }
  }
  cudaFuncSetAttribute(mma_cadence, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("\n%-8s %-5s %-5s %-8s %-4s %10s\n", "cadence", "N", "grp", "stages", "", "cyc/MMA");
  for (int N : {96, 192, 256})
    for (int grp : {1, 3, 6, 9})
      for (int cad : {0, 1, 2, 3})
        for (int nst : {2, 4}) {
          if (cad != 3 && nst != 4) continue;
          const int it2 = 4096 / grp * grp;
          long long h = 0;
          for (int rep = 0; rep < 2; ++rep) {
            mma_cadence<<<148, 128, 200 * 1024>>>(N, it2, grp, cad, nst, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          printf("%-8d %-5d %-5d %-8d %-4s %10.1f\n", cad, N, grp, nst, "", (double)h / it2);
        }
  return 0;
}
