// Micro-benchmark 2: which pipeline feature of the shared-MLP GEMM slows a 3xTF32 k-block (12 x tcgen05.mma 128x128x8)
// down from its 768-cycle floor (tools/mma_floor.cu) to the ~2000 cycles seen inside the real kernels
// (profiles/r2_tile_trace.txt)?  The skeleton of the kernel without global activations, features switched by a bit mask:
//   1  random operand data instead of zeros
//   2  two tcgen05.commit per k-block (stage-release barriers), nobody waits on them
//   4  weights through a 4-stage ring refilled by cp.async.bulk (32 KB per k-block, loader thread, full/empty barriers)
//   8  A operand written into tensor memory by 16 warps per k-block (tcgen05.st, full_a/empty_a hand-shake), TS form
//  16  TS form with a fixed A (no producers)
//  32  the MMA warp stays CONVERGED: all 32 lanes run the loop on warp-uniform values, the MMAs / commits are issued
//      under elect.sync (instead of the whole loop living inside `if (lane == 0)`)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_floor2 tools/mma_floor2.cu && tools/mma_floor2
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../omni-pq_b200/csrc/pn2_sm100.cuh"

using namespace pn2::sm100;

constexpr int TILE = 128 * 32 * 4;  // 16 KB
constexpr int NB = 4, NA = 6;

__device__ __forceinline__ void expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__global__ void __launch_bounds__(576, 1) k(int iters, int feat, const float *bimg, unsigned long long *out) {
  extern __shared__ unsigned char raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char *ring_b = tiles;                 // NB x 32 KB
  unsigned char *a_sm = tiles + NB * 2 * TILE;   // A hi/lo for the SS form
  __shared__ uint64_t full_b[NB], empty_b[NB], full_a[NA], empty_a[NA], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool rnd = feat & 1, commits = feat & 2, bring = feat & 4, prod = feat & 8, ts = (feat & 16) || prod;
  for (int i = tid; i < (NB * 2 + 2) * TILE / 4; i += blockDim.x) {
    unsigned h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    reinterpret_cast<float *>(tiles)[i] = rnd ? (static_cast<float>(h >> 8) / 16777216.f - 0.5f) : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < NB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < NA; ++s) { mbar_init(&full_a[s], 16); mbar_init(&empty_a[s], 1); }
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const uint32_t idesc = idesc_tf32(128, 128);
  if (warp == 16 && (feat & 32)) {  // converged MMA warp
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int sb = bring ? it % NB : 0, sa = it % NA;
      if (prod) mbar_wait(&full_a[sa], (it / NA) & 1);
      if (bring) mbar_wait(&full_b[sb], (it / NB) & 1);
      tc_fence_after_sync();
      const uint32_t bbase = smem_addr(ring_b + sb * 2 * TILE);
      const uint64_t bh = smem_desc_sw128(bbase), bl = smem_desc_sw128(bbase + TILE);
      const uint64_t ah = smem_desc_sw128(smem_addr(a_sm)), al = smem_desc_sw128(smem_addr(a_sm + TILE));
      const uint32_t ta_hi = tmem_d + 128 + (prod ? sa : 0) * 64, ta_lo = ta_hi + 32;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = 2 * ks;
          if (!ts) {
            mma_tf32(tmem_d, ah + adv, bh + adv, idesc, it > 0 || ks > 0);
            mma_tf32(tmem_d, ah + adv, bl + adv, idesc, true);
            mma_tf32(tmem_d, al + adv, bh + adv, idesc, true);
          } else {
            mma_tf32_ts(tmem_d, ta_hi + 8 * ks, bh + adv, idesc, it > 0 || ks > 0);
            mma_tf32_ts(tmem_d, ta_hi + 8 * ks, bl + adv, idesc, true);
            mma_tf32_ts(tmem_d, ta_lo + 8 * ks, bh + adv, idesc, true);
          }
        }
        if (commits || prod) mma_commit(&empty_a[sa]);
        if (commits || bring) mma_commit(&empty_b[sb]);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
  } else if (warp == 16 && lane == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int sb = bring ? it % NB : 0, sa = it % NA;
      if (prod) mbar_wait(&full_a[sa], (it / NA) & 1);
      if (bring) mbar_wait(&full_b[sb], (it / NB) & 1);
      tc_fence_after_sync();
      const uint32_t bbase = smem_addr(ring_b + sb * 2 * TILE);
      const uint64_t bh = smem_desc_sw128(bbase), bl = smem_desc_sw128(bbase + TILE);
      const uint64_t ah = smem_desc_sw128(smem_addr(a_sm)), al = smem_desc_sw128(smem_addr(a_sm + TILE));
      const uint32_t ta_hi = tmem_d + 128 + (prod ? sa : 0) * 64, ta_lo = ta_hi + 32;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = 2 * ks;
        if (!ts) {
          mma_tf32(tmem_d, ah + adv, bh + adv, idesc, it > 0 || ks > 0);
          mma_tf32(tmem_d, ah + adv, bl + adv, idesc, true);
          mma_tf32(tmem_d, al + adv, bh + adv, idesc, true);
        } else {
          mma_tf32_ts(tmem_d, ta_hi + 8 * ks, bh + adv, idesc, it > 0 || ks > 0);
          mma_tf32_ts(tmem_d, ta_hi + 8 * ks, bl + adv, idesc, true);
          mma_tf32_ts(tmem_d, ta_lo + 8 * ks, bh + adv, idesc, true);
        }
      }
      if (commits || prod) mma_commit(&empty_a[sa]);
      if (commits || bring) mma_commit(&empty_b[sb]);
    }
    mma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
  } else if (warp == 17 && lane == 0 && bring) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % NB;
      if (it >= NB) mbar_wait(&empty_b[s], ((it / NB) - 1) & 1);
      expect_tx(&full_b[s], 2 * TILE);
      bulk_g2s(ring_b + s * 2 * TILE, bimg + static_cast<size_t>(it % 8) * (2 * TILE / 4), 2 * TILE, &full_b[s]);
    }
  } else if (warp < 16 && prod) {
    const int quarter = warp & 3, cgrp = warp >> 2;
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + 128 + cgrp * 8;
    float hi[8], lo[8];
    for (int j = 0; j < 8; ++j) { hi[j] = rnd ? 0.37f * (j + lane) : 0.f; lo[j] = rnd ? 1e-4f * (j + warp) : 0.f; }
    for (int it = 0; it < iters; ++it) {
      const int sa = it % NA;
      if (it >= NA) {
        mbar_wait(&empty_a[sa], ((it / NA) - 1) & 1);
        tc_fence_after_sync();
      }
      tmem_st8(lane_base + sa * 64, hi);
      tmem_st8(lane_base + sa * 64 + 32, lo);
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive(&full_a[sa]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_d);
}

int main() {
  unsigned long long *d;
  float *bimg;
  cudaMalloc(&d, 8);
  cudaMalloc(&bimg, 8 * 2 * TILE);
  cudaMemset(bimg, 0, 8 * 2 * TILE);
  const int smem = (NB * 2 + 2) * TILE + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  const int feats[] = {0, 32, 16, 48, 7, 39, 15, 47};
  for (int grid : {1, 140}) {
    for (int f : feats) {
      for (int rep = 0; rep < 2; ++rep) k<<<grid, 576, smem>>>(iters, f, bimg, d);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long cyc = 0;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("grid %3d feat %2d [%s%s%s%s%s%s] %s cycles/k-block %8.1f\n", grid, f, f & 1 ? "data " : "zeros ", f & 2 ? "commits " : "",
             f & 4 ? "bulkB " : "", f & 8 ? "sttmA " : "", f & 16 ? "TS " : "", f & 32 ? "CONVERGED " : "", e == cudaSuccess ? "ok" : cudaGetErrorString(e),
             static_cast<double>(cyc) / iters);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
