// Shared helpers for the sm_100a kernels behind include/pn2_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pn2_b200.h"

#define PN2_EXPORT extern "C" __attribute__((visibility("default")))

namespace pn2 {

// Per-thread last error text (pn2_last_error()).
void set_error(const char *fmt, ...);
int check_launch(const char *what);  // cudaGetLastError() -> return code, never exit()
int sm_count();

#define PN2_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      pn2::set_error(__VA_ARGS__);              \
      return PN2_ERR_INVALID_ARG;               \
    }                                           \
  } while (0)

// a*a + b*b + c*c with the reference's contraction: FMUL b*b; FFMA a,a; FFMA c,c
// (sampling_gpu.cu:102,108-109, ball_query_gpu.cu:36-37, interpolate_gpu.cu:37 as compiled by
// nvcc -O2; the integer outputs depend on it, so it is spelt with intrinsics that the compiler
// may not re-associate or re-contract).
__device__ __forceinline__ float sq3(float a, float b, float c) {
  return __fmaf_rn(c, c, __fmaf_rn(a, a, __fmul_rn(b, b)));
}
__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  return sq3(__fsub_rn(ax, bx), __fsub_rn(ay, by), __fsub_rn(az, bz));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

}  // namespace pn2
