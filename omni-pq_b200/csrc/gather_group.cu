// gather_points / group_points and their gradients for sm_100a -- replace gather_points_kernel,
// gather_points_grad_kernel (reference pointnet2/_ext_src/src/sampling_gpu.cu:13-62) and
// group_points_kernel, group_points_grad_kernel (group_points_gpu.cu:13-80).
//
// The reference launches one block per cloud (grid = B, or (B,C) for gather), so a single 40k-point
// cloud uses one SM.  Here the output index space is flattened over the whole chip: one thread owns
// one output column p (a sampled point j, or a (j,k) neighbour slot) and walks a strip of channels,
// so the int32 index is read once per strip, stores are coalesced along p, and the random reads hit
// one (b,c) row of N floats at a time (L1/L2 resident).  Gradients use fp32 red.global.add exactly
// like the reference's atomicAdd (order unspecified there as well).
#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kThreads = 256;
constexpr int kStrip = 8;  // channels per thread: amortises the index load, keeps 8 loads in flight

// out[b,c,p] = points[b,c,idx[b,p]] for p < cols (cols = m for gather, npoints*nsample for group).
__global__ void __launch_bounds__(kThreads)
column_gather_kernel(int c, int n, int cols, const float *__restrict__ points, const int *__restrict__ idx,
                     float *__restrict__ out) {
  pdl_prologue();
  const int b = blockIdx.z;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= cols) return;
  const int c0 = blockIdx.y * kStrip;
  const int a = idx[static_cast<size_t>(b) * cols + p];
  const float *src = points + (static_cast<size_t>(b) * c + c0) * n + a;
  float *dst = out + (static_cast<size_t>(b) * c + c0) * cols + p;
  const int cn = min(kStrip, c - c0);
  float v[kStrip];
#pragma unroll
  for (int l = 0; l < kStrip; ++l)
    if (l < cn) v[l] = __ldg(src + static_cast<size_t>(l) * n);
#pragma unroll
  for (int l = 0; l < kStrip; ++l)
    if (l < cn) dst[static_cast<size_t>(l) * cols] = v[l];
}

// grad_points[b,c,idx[b,p]] += grad_out[b,c,p]
__global__ void __launch_bounds__(kThreads)
column_scatter_add_kernel(int c, int n, int cols, const float *__restrict__ grad_out, const int *__restrict__ idx,
                          float *__restrict__ grad_points) {
  pdl_prologue();
  const int b = blockIdx.z;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= cols) return;
  const int c0 = blockIdx.y * kStrip;
  const int a = idx[static_cast<size_t>(b) * cols + p];
  const float *src = grad_out + (static_cast<size_t>(b) * c + c0) * cols + p;
  float *dst = grad_points + (static_cast<size_t>(b) * c + c0) * n + a;
  const int cn = min(kStrip, c - c0);
  float v[kStrip];
#pragma unroll
  for (int l = 0; l < kStrip; ++l)
    if (l < cn) v[l] = __ldg(src + static_cast<size_t>(l) * cols);
#pragma unroll
  for (int l = 0; l < kStrip; ++l)
    if (l < cn) atomicAdd(dst + static_cast<size_t>(l) * n, v[l]);
}

int launch_columns(bool scatter, const char *what, int b, int c, int n, long long cols, const float *src, const int *idx,
                   float *dst, void *stream) {
  PN2_REQUIRE(b >= 0 && c >= 0 && n > 0 && cols >= 0, "%s: bad extents b=%d c=%d n=%d cols=%lld", what, b, c, n, cols);
  PN2_REQUIRE(cols <= 0x7fffffffLL, "%s: npoints*nsample=%lld overflows int32", what, cols);
  if (b == 0 || c == 0 || cols == 0) return PN2_OK;
  PN2_REQUIRE(src && idx && dst, "%s: null pointer", what);
  PN2_REQUIRE(b <= 65535 && (c + kStrip - 1) / kStrip <= 65535, "%s: b=%d or c=%d exceeds the grid limits", what, b, c);
  dim3 grid(static_cast<unsigned>((cols + kThreads - 1) / kThreads), (c + kStrip - 1) / kStrip, b);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (scatter)
    pn2::launch(column_scatter_add_kernel, dim3(grid), dim3(kThreads), 0, s, c, n, static_cast<int>(cols), src, idx, dst);
  else
    pn2::launch(column_gather_kernel, dim3(grid), dim3(kThreads), 0, s, c, n, static_cast<int>(cols), src, idx, dst);
  return check_launch(what);
}

}  // namespace
}  // namespace pn2

PN2_EXPORT int pn2_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out,
                                 void *stream) {
  return pn2::launch_columns(false, "pn2_gather_points", b, c, n, m, points, idx, out, stream);
}

PN2_EXPORT int pn2_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                      float *grad_points, void *stream) {
  return pn2::launch_columns(true, "pn2_gather_points_grad", b, c, n, m, grad_out, idx, grad_points, stream);
}

PN2_EXPORT int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                                float *out, void *stream) {
  PN2_REQUIRE(npoints >= 0 && nsample >= 0, "pn2_group_points: bad extents npoints=%d nsample=%d", npoints, nsample);
  return pn2::launch_columns(false, "pn2_group_points", b, c, n, static_cast<long long>(npoints) * nsample, points, idx,
                             out, stream);
}

PN2_EXPORT int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                     const int *idx, float *grad_points, void *stream) {
  PN2_REQUIRE(npoints >= 0 && nsample >= 0, "pn2_group_points_grad: bad extents npoints=%d nsample=%d", npoints,
              nsample);
  return pn2::launch_columns(true, "pn2_group_points_grad", b, c, n, static_cast<long long>(npoints) * nsample, grad_out,
                             idx, grad_points, stream);
}
