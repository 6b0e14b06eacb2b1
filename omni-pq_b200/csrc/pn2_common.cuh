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

// ---- programmatic dependent launch ----------------------------------------------------------------------------
// A step is ~160 small-to-medium kernels in two dependency chains; between two dependent kernels the GPU otherwise
// idles for the launch latency of the second one.  Every kernel of the library starts with pdl_prologue(): it lets
// the NEXT kernel of the stream be scheduled already (griddepcontrol.launch_dependents; that kernel's CTAs take
// whatever SM resources are free) and then waits until the PREVIOUS kernel has completed and its writes are
// visible (griddepcontrol.wait) before touching memory.  Launches go through pn2::launch(), which sets the
// programmatic-stream-serialization attribute when PN2_PDL=1 (the device instructions are no-ops otherwise).
// Measured on the backbone step it is SLOWER (4.21 vs 3.94 ms in round 1; restricted to the small finalize / reduce /
// layout kernels in round 2: 3.64 vs 3.56 ms): the early-scheduled CTAs of the next kernel hold SM resources the
// geometry stream's kernels and the running kernel's later waves need.  Off by default.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();
void note_launch();  // counts every kernel launch of the library (pn2_kernel_launches())

template <typename... P, typename... A>
inline cudaError_t launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  note_launch();
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

}  // namespace pn2
