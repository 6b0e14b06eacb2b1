// three_nn / three_interpolate / three_interpolate_grad for sm_100a -- replace three_nn_kernel,
// three_interpolate_kernel and three_interpolate_grad_kernel (reference
// pointnet2/_ext_src/src/interpolate_gpu.cu:14-73, 77-116, 121-159; each launched with grid = B).
//
// three_nn: one thread per unknown point, the known set streamed through a shared-memory tile that
// the whole CTA loads with coalesced reads; grid = (unknown tiles, clouds) so one cloud fills the
// chip.  The reference keeps its three best distances in doubles initialised to 1e40 and compares
// the fp32 distance against them with strict '<' (:32-54); fp32 registers initialised to +inf give
// the same decisions ((float)1e40 == +inf is also what the reference finally stores) without
// touching the fp64 pipe.  d = fmaf(dz,dz, fmaf(dx,dx, dy*dy)) with dx = unknown - known.
//
// three_interpolate: one thread per output column j walks a strip of channels, so idx/weight are read
// once per strip; out = fmaf(p3,w3, fmaf(p1,w1, p2*w2)) (the contraction nvcc emits for :103-104).
#include <math_constants.h>

#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kNnThreads = 128;
constexpr int kNnTile = 1024;  // known points per shared-memory tile (12 KB)

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2_out, int *__restrict__ idx_out) {
  pdl_prologue();
  __shared__ float tile[kNnTile * 3];
  const int b = blockIdx.y;
  unknown += static_cast<size_t>(b) * n * 3;
  known += static_cast<size_t>(b) * m * 3;
  const int j = blockIdx.x * kNnThreads + threadIdx.x;
  const bool live = j < n;
  const int jj = live ? j : n - 1;
  const float ux = unknown[jj * 3 + 0], uy = unknown[jj * 3 + 1], uz = unknown[jj * 3 + 2];
  float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int base = 0; base < m; base += kNnTile) {
    const int tn = min(kNnTile, m - base);
    __syncthreads();
    for (int i = threadIdx.x; i < tn * 3; i += kNnThreads) tile[i] = known[base * 3 + i];
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < tn; ++k) {
      const float d = dist2(ux, uy, uz, tile[k * 3 + 0], tile[k * 3 + 1], tile[k * 3 + 2]);
      if (d < b3) {  // interpolate_gpu.cu:39-53 as a sorted insert; strict '<' keeps the earlier k on ties
        const int kk = base + k;
        if (d < b1) {
          b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk;
        } else if (d < b2) {
          b3 = b2; i3 = i2; b2 = d; i2 = kk;
        } else {
          b3 = d; i3 = kk;
        }
      }
    }
  }
  if (live) {
    float *dd = dist2_out + (static_cast<size_t>(b) * n + j) * 3;
    int *ii = idx_out + (static_cast<size_t>(b) * n + j) * 3;
    dd[0] = b1; dd[1] = b2; dd[2] = b3;
    ii[0] = i1; ii[1] = i2; ii[2] = i3;
  }
}

constexpr int kIpThreads = 256;
constexpr int kIpStrip = 8;

__global__ void __launch_bounds__(kIpThreads)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points, const int *__restrict__ idx,
                         const float *__restrict__ weight, float *__restrict__ out) {
  pdl_prologue();
  const int b = blockIdx.z;
  const int j = blockIdx.x * kIpThreads + threadIdx.x;
  if (j >= n) return;
  const int c0 = blockIdx.y * kIpStrip;
  const int *ix = idx + (static_cast<size_t>(b) * n + j) * 3;
  const float *w = weight + (static_cast<size_t>(b) * n + j) * 3;
  const int a1 = ix[0], a2 = ix[1], a3 = ix[2];
  const float w1 = w[0], w2 = w[1], w3 = w[2];
  const int cn = min(kIpStrip, c - c0);
#pragma unroll
  for (int l = 0; l < kIpStrip; ++l) {
    if (l >= cn) break;
    const float *p = points + (static_cast<size_t>(b) * c + c0 + l) * m;
    const float t = __fmaf_rn(__ldg(p + a3), w3, __fmaf_rn(__ldg(p + a1), w1, __fmul_rn(__ldg(p + a2), w2)));
    out[(static_cast<size_t>(b) * c + c0 + l) * n + j] = t;
  }
}

__global__ void __launch_bounds__(kIpThreads)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out, const int *__restrict__ idx,
                              const float *__restrict__ weight, float *__restrict__ grad_points) {
  pdl_prologue();
  const int b = blockIdx.z;
  const int j = blockIdx.x * kIpThreads + threadIdx.x;
  if (j >= n) return;
  const int c0 = blockIdx.y * kIpStrip;
  const int *ix = idx + (static_cast<size_t>(b) * n + j) * 3;
  const float *w = weight + (static_cast<size_t>(b) * n + j) * 3;
  const int a1 = ix[0], a2 = ix[1], a3 = ix[2];
  const float w1 = w[0], w2 = w[1], w3 = w[2];
  const int cn = min(kIpStrip, c - c0);
#pragma unroll
  for (int l = 0; l < kIpStrip; ++l) {
    if (l >= cn) break;
    const float g = __ldg(grad_out + (static_cast<size_t>(b) * c + c0 + l) * n + j);
    float *gp = grad_points + (static_cast<size_t>(b) * c + c0 + l) * m;
    atomicAdd(gp + a1, __fmul_rn(g, w1));  // interpolate_gpu.cu:144-146
    atomicAdd(gp + a2, __fmul_rn(g, w2));
    atomicAdd(gp + a3, __fmul_rn(g, w3));
  }
}

}  // namespace
}  // namespace pn2

PN2_EXPORT int pn2_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                            void *stream) {
  using namespace pn2;
  PN2_REQUIRE(b >= 0 && n >= 0 && m >= 0, "pn2_three_nn: bad extents b=%d n=%d m=%d", b, n, m);
  if (b == 0 || n == 0) return PN2_OK;
  PN2_REQUIRE(unknown && dist2 && idx && (known || m == 0), "pn2_three_nn: null pointer");
  PN2_REQUIRE(b <= 65535, "pn2_three_nn: b=%d exceeds the grid limit", b);
  dim3 grid((n + kNnThreads - 1) / kNnThreads, b);
  pn2::launch(three_nn_kernel, dim3(grid), dim3(kNnThreads), 0, static_cast<cudaStream_t>(stream), n, m, unknown, known, dist2, idx);
  return check_launch("pn2_three_nn");
}

PN2_EXPORT int pn2_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                     const float *weight, float *out, void *stream) {
  using namespace pn2;
  PN2_REQUIRE(b >= 0 && c >= 0 && m > 0 && n >= 0, "pn2_three_interpolate: bad extents b=%d c=%d m=%d n=%d", b, c, m, n);
  if (b == 0 || c == 0 || n == 0) return PN2_OK;
  PN2_REQUIRE(points && idx && weight && out, "pn2_three_interpolate: null pointer");
  PN2_REQUIRE(b <= 65535 && (c + kIpStrip - 1) / kIpStrip <= 65535, "pn2_three_interpolate: grid limits exceeded");
  dim3 grid((n + kIpThreads - 1) / kIpThreads, (c + kIpStrip - 1) / kIpStrip, b);
  pn2::launch(three_interpolate_kernel, dim3(grid), dim3(kIpThreads), 0, static_cast<cudaStream_t>(stream), c, m, n, points, idx, weight, out);
  return check_launch("pn2_three_interpolate");
}

PN2_EXPORT int pn2_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                          const float *weight, float *grad_points, void *stream) {
  using namespace pn2;
  PN2_REQUIRE(b >= 0 && c >= 0 && m > 0 && n >= 0, "pn2_three_interpolate_grad: bad extents b=%d c=%d n=%d m=%d", b, c, n,
              m);
  if (b == 0 || c == 0 || n == 0) return PN2_OK;
  PN2_REQUIRE(grad_out && idx && weight && grad_points, "pn2_three_interpolate_grad: null pointer");
  PN2_REQUIRE(b <= 65535 && (c + kIpStrip - 1) / kIpStrip <= 65535, "pn2_three_interpolate_grad: grid limits exceeded");
  dim3 grid((n + kIpThreads - 1) / kIpThreads, (c + kIpStrip - 1) / kIpStrip, b);
  pn2::launch(three_interpolate_grad_kernel, dim3(grid), dim3(kIpThreads), 0, static_cast<cudaStream_t>(stream), c, n, m, grad_out, idx, weight,
                                                                                          grad_points);
  return check_launch("pn2_three_interpolate_grad");
}
