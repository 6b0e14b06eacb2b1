// Micro-benchmark: issue rate of tcgen05.mma.kind::tf32 on B200 for the operand forms the shared-MLP GEMM can use.
// One elected thread issues `iters` groups of 12 MMAs (the 3xTF32 pattern of one 32-wide k-block) on fixed
// shared-memory / tensor-memory operands and waits for the last commit; no global traffic at all, so the number is
// the tensor pipe + operand fetch floor of a k-block.  Optional "noise" warps hammer shared memory with 128-bit
// stores the way the operand producers do.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_floor tools/mma_floor.cu && ./mma_floor
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "../omni-pq_b200/csrc/pn2_sm100.cuh"

using namespace pn2::sm100;

constexpr int TILE = 128 * 32 * 4;  // 16 KB: 128 rows x 32 fp32, K-major SW128

// form 0: SS (A,B in smem)   form 1: TS (A in tmem, B in smem)
template <int N, int FORM>
__global__ void __launch_bounds__(288, 1) floor_kernel(int iters, int noise_warps, unsigned long long *out) {
  extern __shared__ unsigned char raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  // layout: A_hi, A_lo (16 KB each), B_hi, B_lo (N/128 * 16 KB each), noise area 32 KB
  unsigned char *a_hi = tiles, *a_lo = tiles + TILE;
  unsigned char *b_hi = tiles + 2 * TILE, *b_lo = b_hi + (N / 128) * TILE;
  unsigned char *noise = b_lo + (N / 128) * TILE;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (2 + 2 * (N / 128)) * TILE / 4; i += blockDim.x) reinterpret_cast<float *>(tiles)[i] = 0.f;
  if (tid == 0) {
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = idesc_tf32(128, N);
  if (warp == 8 && lane == 0) {
    const uint64_t ah = smem_desc_sw128(smem_addr(a_hi)), al = smem_desc_sw128(smem_addr(a_lo));
    const uint64_t bh = smem_desc_sw128(smem_addr(b_hi)), bl = smem_desc_sw128(smem_addr(b_lo));
    const uint32_t ta_hi = tmem_d + 256, ta_lo = tmem_d + 256 + 32;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = 2 * ks;
        if (FORM == 0) {
          mma_tf32(tmem_d, ah + adv, bh + adv, idesc, true);
          mma_tf32(tmem_d, ah + adv, bl + adv, idesc, true);
          mma_tf32(tmem_d, al + adv, bh + adv, idesc, true);
        } else {
          mma_tf32_ts(tmem_d, ta_hi + 8 * ks, bh + adv, idesc, true);
          mma_tf32_ts(tmem_d, ta_hi + 8 * ks, bl + adv, idesc, true);
          mma_tf32_ts(tmem_d, ta_lo + 8 * ks, bh + adv, idesc, true);
        }
      }
    }
    mma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
    *reinterpret_cast<volatile int *>(noise + 32768 - 4) = 1;  // stop flag for the noise warps
  } else if (warp < noise_warps) {
    // producers' traffic: each thread two 128-bit stores per "k-block", spinning until the MMA thread is done
    volatile int *stop = reinterpret_cast<volatile int *>(noise + 32768 - 4);
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    unsigned off = (tid * 16) & 16383;
    while (*stop == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr(noise + ((off + j * 2048) & 16383))), "f"(v.x),
                     "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_d);
}

template <int N, int FORM>
void run(const char *name, int grid, int noise_warps) {
  unsigned long long *d;
  cudaMalloc(&d, 8);
  const int smem = (2 + 2 * (N / 128)) * TILE + 32768 + 1024;
  cudaFuncSetAttribute(floor_kernel<N, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) floor_kernel<N, FORM><<<grid, 288, smem>>>(iters, noise_warps, d);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long cyc = 0;
  cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  const double per_mma = static_cast<double>(cyc) / (iters * 12.0);
  const double flops = 2.0 * 128 * N * 8;
  printf("%-28s grid=%3d noise_warps=%d  %s  cycles/MMA=%7.1f  cycles/k-block(12 MMA)=%8.1f  flops/cycle/SM=%7.0f\n", name, grid,
         noise_warps, e == cudaSuccess ? "ok " : cudaGetErrorString(e), per_mma, per_mma * 12, flops / per_mma);
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    for (int nw : {0, 8}) {
      run<128, 0>("SS  M128 N128 tf32", grid, nw);
      run<256, 0>("SS  M128 N256 tf32", grid, nw);
      run<128, 1>("TS  M128 N128 tf32", grid, nw);
      run<256, 1>("TS  M128 N256 tf32", grid, nw);
    }
  }
  return 0;
}
