// Ball query for sm_100a -- replaces query_ball_point_kernel (reference
// pointnet2/_ext_src/src/ball_query_gpu.cu:14-59: one block per cloud, one THREAD per centre
// scanning all n points sequentially from global memory).
//
// Semantics kept bit-exact: for every centre the first `nsample` point indices, in ascending index
// order, with d2 < radius*radius (d2 = fmaf(dz,dz, fmaf(dx,dx, dy*dy)), dx = centre - point, fp32
// product radius*radius, strict '<'; ball_query_gpu.cu:27,36-38), remaining slots = the first hit,
// all zeros if the ball is empty (:39-46 over a zero-filled output).
//
// Parallelisation: a CTA of 8 warps owns 8 centres.  The cloud is cut into 8 index-contiguous
// segments, one per warp; a warp tests 32 points per step against all 8 centres (8 independent
// distance chains per lane), a ballot + prefix popcount appends the in-ball lanes of a step in index
// order to the (centre, segment) list in shared memory.  Because the segments are index-ordered,
// concatenating the 8 lists of a centre gives exactly the reference's scan order; the merge step
// writes the first `nsample` entries (coalesced 128-byte rows) and pads with the first hit.
// A warp stops recording for a centre once its own list holds `nsample` entries (later entries
// could never be among the first `nsample`).  grid = (centre groups, clouds): one 40k-point cloud
// with 2048 centres runs 2048 warps across the chip instead of one block.
#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kBqWarps = 8;      // = segments per cloud
constexpr int kBqCentres = 8;    // centres per CTA
constexpr int kBqMaxSample = 128;

__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx) {
  pdl_prologue();
  extern __shared__ int stage[];  // [kBqCentres][kBqWarps][nsample]
  __shared__ int counts[kBqCentres][kBqWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int batch = blockIdx.y;
  xyz += static_cast<size_t>(batch) * n * 3;
  new_xyz += static_cast<size_t>(batch) * m * 3;
  idx += static_cast<size_t>(batch) * m * nsample;
  const float radius2 = __fmul_rn(radius, radius);
  const int j0 = blockIdx.x * kBqCentres;

  float cx[kBqCentres], cy[kBqCentres], cz[kBqCentres];
  int cnt[kBqCentres];
#pragma unroll
  for (int c = 0; c < kBqCentres; ++c) {
    const int j = min(j0 + c, m - 1);
    cx[c] = __ldg(new_xyz + j * 3 + 0);
    cy[c] = __ldg(new_xyz + j * 3 + 1);
    cz[c] = __ldg(new_xyz + j * 3 + 2);
    cnt[c] = (j0 + c < m) ? 0 : nsample;  // out-of-range centres are "full" from the start
  }
  const unsigned lt_mask = (1u << lane) - 1u;
  const int seg = ((n + kBqWarps - 1) / kBqWarps + 31) / 32 * 32;  // segment length, multiple of 32
  const int begin = warp * seg, end = min(n, begin + seg);

  for (int t = begin; t < end; t += 32) {
    const int k = t + lane;
    const bool in_range = k < end;
    const int kk = in_range ? k : begin;
    const float x = __ldg(xyz + kk * 3 + 0), y = __ldg(xyz + kk * 3 + 1), z = __ldg(xyz + kk * 3 + 2);
    bool open = false;
#pragma unroll
    for (int c = 0; c < kBqCentres; ++c) {
      const float d2 = dist2(cx[c], cy[c], cz[c], x, y, z);
      const bool hit = in_range && d2 < radius2;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal != 0u && cnt[c] < nsample) {
        const int pos = cnt[c] + __popc(bal & lt_mask);
        if (hit && pos < nsample) stage[(c * kBqWarps + warp) * nsample + pos] = k;
        cnt[c] = min(nsample, cnt[c] + __popc(bal));
      }
      open |= cnt[c] < nsample;
    }
    if (!open) break;  // warp-uniform: every centre's list of this segment is full
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kBqCentres; ++c) counts[c][warp] = (j0 + c < m) ? cnt[c] : 0;
  }
  __syncthreads();

  // merge: output slot s of centre c = s-th entry of the concatenated segment lists, else the first hit
  for (int o = threadIdx.x; o < kBqCentres * nsample; o += kBqWarps * 32) {
    const int c = o / nsample, s = o - c * nsample;
    const int j = j0 + c;
    if (j >= m) continue;
    int acc = 0, value = 0, first = 0;
    bool found = false, have_first = false;
#pragma unroll
    for (int g = 0; g < kBqWarps; ++g) {
      const int cg = counts[c][g];
      if (!have_first && cg > 0) { first = stage[(c * kBqWarps + g) * nsample]; have_first = true; }
      if (!found && s < acc + cg) { value = stage[(c * kBqWarps + g) * nsample + (s - acc)]; found = true; }
      acc += cg;
    }
    idx[static_cast<size_t>(j) * nsample + s] = found ? value : first;  // empty ball: first == 0
  }
}

// Any nsample (the staged kernel above keeps 8 x 8 x nsample indices in shared memory and is used up to
// kBqMaxSample): one warp per centre scans the whole cloud 32 points per step and appends the in-ball lanes straight
// to the output row in index order (ballot + prefix popcount), then pads with the first hit.  Same results as the
// reference's sequential scan (ball_query_gpu.cu:27-46), which accepts every nsample.
__global__ void __launch_bounds__(256)
ball_query_warp_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz,
                       const float *__restrict__ xyz, int *__restrict__ idx) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int batch = blockIdx.y;
  if (j >= m) return;
  xyz += static_cast<size_t>(batch) * n * 3;
  new_xyz += (static_cast<size_t>(batch) * m + j) * 3;
  int *row = idx + (static_cast<size_t>(batch) * m + j) * nsample;
  const float radius2 = __fmul_rn(radius, radius);
  const float cx = __ldg(new_xyz + 0), cy = __ldg(new_xyz + 1), cz = __ldg(new_xyz + 2);
  const unsigned lt_mask = (1u << lane) - 1u;
  int cnt = 0, first = 0;
  for (int t = 0; t < n && cnt < nsample; t += 32) {
    const int k = t + lane;
    const bool in_range = k < n;
    const int kk = in_range ? k : 0;
    const float d2 = dist2(cx, cy, cz, __ldg(xyz + kk * 3 + 0), __ldg(xyz + kk * 3 + 1), __ldg(xyz + kk * 3 + 2));
    const bool hit = in_range && d2 < radius2;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (bal == 0u) continue;
    if (cnt == 0) first = t + __ffs(bal) - 1;
    const int pos = cnt + __popc(bal & lt_mask);
    if (hit && pos < nsample) row[pos] = k;
    cnt = min(nsample, cnt + __popc(bal));
  }
  for (int s = cnt + lane; s < nsample; s += 32) row[s] = first;  // empty ball: first == 0 (zero-filled in the reference)
}

}  // namespace
}  // namespace pn2

PN2_EXPORT int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              const float *xyz, int *idx, void *stream) {
  using namespace pn2;
  PN2_REQUIRE(b >= 0 && n > 0 && m >= 0 && nsample > 0, "pn2_ball_query: bad extents b=%d n=%d m=%d nsample=%d", b, n, m,
              nsample);
  if (b == 0 || m == 0) return PN2_OK;
  PN2_REQUIRE(new_xyz && xyz && idx, "pn2_ball_query: null pointer");
  PN2_REQUIRE(b <= 65535, "pn2_ball_query: b=%d exceeds the grid limit", b);
  if (nsample > kBqMaxSample) {  // staging buffer would exceed 32 KB: warp-per-centre kernel, any nsample
    dim3 wgrid((m + 7) / 8, b);
    pn2::launch(ball_query_warp_kernel, dim3(wgrid), dim3(256), 0, static_cast<cudaStream_t>(stream), n, m, radius, nsample, new_xyz,
                xyz, idx);
    return check_launch("pn2_ball_query(warp)");
  }
  dim3 grid((m + kBqCentres - 1) / kBqCentres, b);
  const size_t smem = sizeof(int) * kBqCentres * kBqWarps * static_cast<size_t>(nsample);  // <= 32 KB
  pn2::launch(ball_query_kernel, dim3(grid), dim3(kBqWarps * 32), smem, static_cast<cudaStream_t>(stream), n, m, radius, nsample, new_xyz, xyz,
                                                                                     idx);
  return check_launch("pn2_ball_query");
}
