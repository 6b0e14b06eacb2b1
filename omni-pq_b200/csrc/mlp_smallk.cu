// First-layer kernels for tiny input widths (K = kp <= 16): the backbone's SA1 layer 0 is a [131072 x 8] x [8 x 128]
// contraction (3 colour channels + pad, 3 local coordinates + pad).  On the 128x128x32 tensor-core tiles that is one
// nearly empty k-block per tile and, for the weight gradient, a 128 x 8 output spread over 148 position slices:
// 58 us forward / 125 us weight gradient at 0.4-1.1 TB/s (profiles/r1_bench.json).  Both are plain streaming
// problems -- 8 FMAs per output element forward, 64 MB of output; 134 MB of dY/dz input for the weight gradient --
// so they run here as fp32 FFMA kernels with coalesced 128-bit rows:
//   forward : CTA = one 128-row statistics tile, lane = 4 output channels, warp = 16 rows; every lane evaluates the
//             row source itself (the loads are warp-uniform, i.e. one transaction); BatchNorm partial sums per tile
//             in the layout the tensor-core epilogue writes ([tile][2][np], fixed-order cross-warp combine).
//   wgrad   : CTA = a slice of positions, lane = 4 dY channels, acc[4][kp] per thread, fixed-order cross-warp
//             combine into ws[slice][np][kp]; the existing split reduction (+ channel permutation) finishes it.
// Sources: A = PLAIN / GATHER (first layers), dY = DY (the layer above handed down dz; a single-layer MLP, whose dY
// is DYPOOL, stays on the tensor-core path).
// Measured (profiles/r1_bench_smallk.json): correct (68/68 GPU tests) but no faster -- forward 60 vs 56 us, weight
// gradient 102 vs 113 us once the split reduction is parallel; every lane re-evaluating the gather context (index,
// coordinates, three divisions) per row costs what the tensor-core tile wastes.  Opt-in with PN2_SMALLK=1.
#include <stdlib.h>

#include "mlp_rows.cuh"

namespace pn2 {
namespace {

constexpr int SK_THREADS = 256, SK_WARPS = SK_THREADS / 32;
constexpr int SK_KMAX = 16;

template <int AKIND, int KP>
__global__ void __launch_bounds__(SK_THREADS)
smallk_forward_kernel(const __grid_constant__ GemmArgs g) {
  pdl_prologue();
  __shared__ float red[2][SK_WARPS][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x, row0 = tile * 128;
  for (int c0 = 0; c0 < g.N; c0 += 128) {  // 128 output channels per pass
    const int c = c0 + lane * 4;
    const bool col_ok = c < g.N;
    float w[4][KP];  // B rows = output channels, K-major (wp [np][kp])
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k4 = 0; k4 < KP; k4 += 4) {
        const float4 v = col_ok ? ldg4(g.B.x + static_cast<size_t>(c + j) * g.B.ld + k4) : zero4();
        w[j][k4 + 0] = v.x; w[j][k4 + 1] = v.y; w[j][k4 + 2] = v.z; w[j][k4 + 3] = v.w;
      }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int r = 0; r < 128 / SK_WARPS; ++r) {
      const int row = row0 + warp * (128 / SK_WARPS) + r;
      const RowCtx ctx = row_ctx<AKIND>(g.A, row < g.M ? row : 0x7fffffff);
      float a[KP];
#pragma unroll
      for (int k4 = 0; k4 < KP; k4 += 4) {
        const float4 v = load4<AKIND>(g.A, ctx, k4);
        a[k4 + 0] = v.x; a[k4 + 1] = v.y; a[k4 + 2] = v.z; a[k4 + 3] = v.w;
      }
      float y[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) acc = fmaf(a[k], w[j][k], acc);
        y[j] = acc;
      }
      if (row < g.M && col_ok) {
        *reinterpret_cast<float4 *>(g.out + static_cast<size_t>(row) * g.ldo + c) = make_float4(y[0], y[1], y[2], y[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { s1[j] += y[j]; s2[j] += y[j] * y[j]; }
      }
    }
    if (g.stats != nullptr) {
      __syncthreads();  // the previous pass has read red[]
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        red[0][warp][lane * 4 + j] = s1[j];
        red[1][warp][lane * 4 + j] = s2[j];
      }
      __syncthreads();
      if (threadIdx.x < 128 && c0 + threadIdx.x < g.stats_ld) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int wv = 0; wv < SK_WARPS; ++wv) { t1 += red[0][wv][threadIdx.x]; t2 += red[1][wv][threadIdx.x]; }
        float *dst = g.stats + static_cast<size_t>(tile) * 2 * g.stats_ld + c0 + threadIdx.x;
        dst[0] = t1;
        dst[g.stats_ld] = t2;
      }
    }
  }
}

// dW partial of one position slice: ws[blockIdx.x][n][k] = sum over the slice's rows of dy[row][n] * a[row][k]
template <int BKIND, int KP>
__global__ void __launch_bounds__(SK_THREADS)
smallk_wgrad_kernel(const __grid_constant__ GemmArgs g) {
  pdl_prologue();
  extern __shared__ float sred[];  // [SK_WARPS][128][KP]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r_begin = blockIdx.x * g.k_per_split, r_end = min(g.K, r_begin + g.k_per_split);  // g.K = positions
  const pn2_rows &D = g.A;  // dY source (PN2_ROWS_DY): dy = c0*dz + c1 + c2*y
  for (int c0 = 0; c0 < g.M; c0 += 128) {  // g.M = np
    const int c = c0 + lane * 4;
    const bool col_ok = c < D.cols;
    const float4 ka = col_ok ? ldg4(D.c0 + c) : zero4(), kb = col_ok ? ldg4(D.c1 + c) : zero4(),
                 kc = col_ok ? ldg4(D.c2 + c) : zero4();
    float acc[4][KP];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < KP; ++k) acc[j][k] = 0.f;
#pragma unroll 4
    for (int row = r_begin + warp; row < r_end; row += SK_WARPS) {
      float4 dy = zero4();
      if (col_ok) {
        const float4 y = ldg4(D.x + static_cast<size_t>(row) * D.ld + c), dz = ldg4(D.dz + static_cast<size_t>(row) * D.ld + c);
        dy = make_float4(fmaf(kc.x, y.x, fmaf(ka.x, dz.x, kb.x)), fmaf(kc.y, y.y, fmaf(ka.y, dz.y, kb.y)),
                         fmaf(kc.z, y.z, fmaf(ka.z, dz.z, kb.z)), fmaf(kc.w, y.w, fmaf(ka.w, dz.w, kb.w)));
      }
      const RowCtx ctx = row_ctx<BKIND>(g.B, row);
      float a[KP];
#pragma unroll
      for (int k4 = 0; k4 < KP; k4 += 4) {
        const float4 v = load4<BKIND>(g.B, ctx, k4);
        a[k4 + 0] = v.x; a[k4 + 1] = v.y; a[k4 + 2] = v.z; a[k4 + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        acc[0][k] = fmaf(dy.x, a[k], acc[0][k]);
        acc[1][k] = fmaf(dy.y, a[k], acc[1][k]);
        acc[2][k] = fmaf(dy.z, a[k], acc[2][k]);
        acc[3][k] = fmaf(dy.w, a[k], acc[3][k]);
      }
    }
    __syncthreads();  // the previous pass has read sred[]
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < KP; ++k) sred[(warp * 128 + lane * 4 + j) * KP + k] = acc[j][k];
    __syncthreads();
    for (int e = threadIdx.x; e < 128 * KP; e += SK_THREADS) {  // e = (channel in pass) * KP + k
      const int n = c0 + e / KP, k = e % KP;
      if (n >= g.M || k >= g.N) continue;  // g.N = kp
      float t = 0.f;
#pragma unroll
      for (int wv = 0; wv < SK_WARPS; ++wv) t += sred[wv * 128 * KP + e];
      g.out[static_cast<size_t>(blockIdx.x) * g.out_split_stride + static_cast<size_t>(n) * g.ldo + k] = t;
    }
  }
}

bool smallk_on() {
  static const bool on = [] {
    const char *e = getenv("PN2_SMALLK");
    return e != nullptr && e[0] == '1';  // opt-in, see the header comment
  }();
  return on;
}

template <int KIND, int KP>
int launch_fwd(const GemmArgs &g, cudaStream_t s) {
  pn2::launch(smallk_forward_kernel<KIND, KP>, dim3((g.M + 127) / 128), dim3(SK_THREADS), 0, s, g);
  return check_launch("smallk_forward_kernel");
}

template <int KIND, int KP>
int launch_wg(const GemmArgs &g, int splits, cudaStream_t s) {
  auto kernel = smallk_wgrad_kernel<KIND, KP>;
  const size_t smem = sizeof(float) * SK_WARPS * 128 * KP;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured_dev = dev;
  }
  pn2::launch(kernel, dim3(splits), dim3(SK_THREADS), smem, s, g);
  return check_launch("smallk_wgrad_kernel");
}

}  // namespace

bool smallk_eligible(int akind, int kp, int np) {
  return smallk_on() && gemm_tc_enabled() && kp >= 4 && kp <= SK_KMAX && (kp % 4) == 0 && np <= 256 &&
         (akind == PN2_ROWS_PLAIN || akind == PN2_ROWS_GATHER);
}

// position slices of the small-K weight gradient: ~4 slices per SM, at least 64 rows each
int smallk_wgrad_splits(int rows) {
  long long s = 4ll * sm_count();
  const long long cap = (rows + 63) / 64;
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return static_cast<int>(s);
}

// g as pn2_mlp_forward builds it for the tensor-core path (B = wp [np][kp] K-major, stats per 128-row tile)
int smallk_forward_launch(const void *gemm_args, cudaStream_t s) {
  const GemmArgs &g = *static_cast<const GemmArgs *>(gemm_args);
  const int kp = (g.K + 3) / 4 * 4;
#define PN2_SK_F(KIND)                                        \
  if (g.A.kind == KIND) {                                     \
    if (kp <= 4) return launch_fwd<KIND, 4>(g, s);            \
    if (kp <= 8) return launch_fwd<KIND, 8>(g, s);            \
    if (kp <= 12) return launch_fwd<KIND, 12>(g, s);          \
    return launch_fwd<KIND, 16>(g, s);                        \
  }
  PN2_SK_F(PN2_ROWS_PLAIN)
  PN2_SK_F(PN2_ROWS_GATHER)
#undef PN2_SK_F
  return PN2_TC_UNSUPPORTED;
}

// g as pn2_mlp_wgrad builds it (A = dY source, B = activation source, M = np, N = kp, K = positions), with
// k_per_split / out_split_stride set for `splits` slices
int smallk_wgrad_launch(const void *gemm_args, int splits, cudaStream_t s) {
  const GemmArgs &g = *static_cast<const GemmArgs *>(gemm_args);
  if (g.A.kind != PN2_ROWS_DY) return PN2_TC_UNSUPPORTED;
  const int kp = g.N;
#define PN2_SK_W(KIND)                                        \
  if (g.B.kind == KIND) {                                     \
    if (kp <= 4) return launch_wg<KIND, 4>(g, splits, s);     \
    if (kp <= 8) return launch_wg<KIND, 8>(g, splits, s);     \
    if (kp <= 12) return launch_wg<KIND, 12>(g, splits, s);   \
    return launch_wg<KIND, 16>(g, splits, s);                 \
  }
  PN2_SK_W(PN2_ROWS_PLAIN)
  PN2_SK_W(PN2_ROWS_GATHER)
#undef PN2_SK_W
  return PN2_TC_UNSUPPORTED;
}

}  // namespace pn2
