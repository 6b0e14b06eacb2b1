// Ball query for sm_100a -- replaces query_ball_point_kernel (reference
// pointnet2/_ext_src/src/ball_query_gpu.cu:14-59: one block per cloud, one THREAD per centre
// scanning all n points sequentially from global memory).
//
// Here one WARP owns kCentres centres and scans the cloud 32 points at a time from a shared-memory
// tile: every lane tests one point against each centre, a ballot gives the in-ball lanes in index
// order and a prefix popcount places them, so the result is exactly "the first nsample indices in
// ascending order, padded with the first hit, zeros if none" (ball_query_gpu.cu:32-46) with
// d2 = fmaf(dz,dz, fmaf(dx,dx, dy*dy)), dx = centre - point, compared strictly against the fp32
// product radius*radius (:27,36-38).  The grid is (centre groups, clouds) so B=1 still fills the
// chip, xyz is staged once per CTA tile with coalesced loads, and rows leave through a per-warp
// staging buffer as coalesced 128-byte stores.
#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kBqWarps = 4;
constexpr int kBqCentres = 2;   // centres per warp
constexpr int kBqTile = 2048;   // points per shared-memory tile (24 KB)
constexpr int kBqMaxSample = 128;

__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx) {
  __shared__ float tile[kBqTile * 3];
  __shared__ int stage[kBqWarps][kBqCentres][kBqMaxSample];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int batch = blockIdx.y;
  xyz += static_cast<size_t>(batch) * n * 3;
  new_xyz += static_cast<size_t>(batch) * m * 3;
  idx += static_cast<size_t>(batch) * m * nsample;
  const float radius2 = __fmul_rn(radius, radius);
  const int j0 = (blockIdx.x * kBqWarps + warp) * kBqCentres;

  float cx[kBqCentres], cy[kBqCentres], cz[kBqCentres];
  int cnt[kBqCentres], first[kBqCentres];
#pragma unroll
  for (int c = 0; c < kBqCentres; ++c) {
    const int j = min(j0 + c, m - 1);
    cx[c] = new_xyz[j * 3 + 0];
    cy[c] = new_xyz[j * 3 + 1];
    cz[c] = new_xyz[j * 3 + 2];
    cnt[c] = (j0 + c < m) ? 0 : nsample;  // out-of-range centres are "full" from the start
    first[c] = 0;
  }
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int base = 0; base < n; base += kBqTile) {
    const int tn = min(kBqTile, n - base);
    __syncthreads();
    for (int i = threadIdx.x; i < tn * 3; i += kBqWarps * 32) tile[i] = xyz[base * 3 + i];
    __syncthreads();
    bool open = false;
#pragma unroll
    for (int c = 0; c < kBqCentres; ++c) open |= cnt[c] < nsample;
    if (!open) continue;  // warp-uniform: every centre of this warp is full (ball_query_gpu.cu:32)
    for (int t = 0; t < tn; t += 32) {
      const int k = t + lane;
      const bool in_range = k < tn;
      const float x = in_range ? tile[k * 3 + 0] : 0.f;
      const float y = in_range ? tile[k * 3 + 1] : 0.f;
      const float z = in_range ? tile[k * 3 + 2] : 0.f;
#pragma unroll
      for (int c = 0; c < kBqCentres; ++c) {
        const float d2 = dist2(cx[c], cy[c], cz[c], x, y, z);
        const bool hit = in_range && d2 < radius2;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal != 0u && cnt[c] < nsample) {
          if (cnt[c] == 0) first[c] = base + t + __ffs(bal) - 1;
          const int pos = cnt[c] + __popc(bal & lt_mask);
          if (hit && pos < nsample) stage[warp][c][pos] = base + k;
          cnt[c] = min(nsample, cnt[c] + __popc(bal));
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < kBqCentres; ++c) {
    const int j = j0 + c;
    if (j >= m) break;
    for (int s = lane; s < nsample; s += 32)
      idx[static_cast<size_t>(j) * nsample + s] = s < cnt[c] ? stage[warp][c][s] : first[c];
  }
}

}  // namespace
}  // namespace pn2

PN2_EXPORT int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              const float *xyz, int *idx, void *stream) {
  using namespace pn2;
  PN2_REQUIRE(b >= 0 && n > 0 && m >= 0 && nsample > 0, "pn2_ball_query: bad extents b=%d n=%d m=%d nsample=%d", b, n, m,
              nsample);
  if (nsample > kBqMaxSample) {
    set_error("pn2_ball_query: nsample=%d exceeds the supported maximum %d", nsample, kBqMaxSample);
    return PN2_ERR_UNSUPPORTED;
  }
  if (b == 0 || m == 0) return PN2_OK;
  PN2_REQUIRE(new_xyz && xyz && idx, "pn2_ball_query: null pointer");
  const int per_block = kBqWarps * kBqCentres;
  dim3 grid((m + per_block - 1) / per_block, b);
  ball_query_kernel<<<grid, kBqWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(n, m, radius, nsample, new_xyz, xyz,
                                                                                  idx);
  return check_launch("pn2_ball_query");
}
