// Shared-MLP contractions for sm_100a: forward, data-gradient and weight-gradient GEMMs of the 1x1
// convolutions in pytorch_utils.SharedMLP (reference pointnet2/pytorch_utils.py:11-36,67-120; the
// reference runs them as cuDNN convolutions with separate BatchNorm / ReLU / max_pool2d / grouping
// kernels around them, pointnet2_modules.py:233-267, pointnet2_utils.py:334-359).
//
// One templated fp32 kernel, C[M x N] = sum_k A(m,k) * B(k,n), whose operands are *row sources*
// (include/pn2_b200.h `pn2_rows`): the element-wise work that surrounds the contraction in the
// reference -- neighbourhood grouping with centre subtraction, BatchNorm+ReLU of the previous layer,
// the BatchNorm / ReLU / max-pool backward -- is applied in registers while a tile is staged into
// shared memory, and the epilogue produces BatchNorm partial statistics (or scatter-adds into the
// gathered inputs), so none of those intermediates makes its own trip through HBM.
//
// Arithmetic: plain fp32 FFMA with a fixed k-ascending accumulation order (deterministic, and the
// 1e-5 relative parity bar of the float paths rules out bf16 / single-pass TF32).  Tiling: 128x128x16
// (256 threads, 8x8 register micro-tile, double-buffered shared memory, 128-bit loads everywhere) or
// 64x64x16 when the problem would leave SMs idle; weight-gradient launches split the position
// dimension across CTAs and reduce the partial tiles in a fixed order.
#include <stdlib.h>

#include "mlp_rows.cuh"

namespace pn2 {
namespace {

constexpr int BK = 16;

template <int BM, int BN, int AKIND, bool ATRANS, int BKIND, int EPI>
__global__ void __launch_bounds__((BM / 8) * (BN / 8), (BM == 128 ? 2 : 6))
gemm_kernel(const __grid_constant__ GemmArgs g) {
  pdl_prologue();
  constexpr int TX = BN / 8, TY = BM / 8, THREADS = TX * TY;
  constexpr int A_LD = BM * BK / 4 / THREADS;  // float4 loads per thread per k-tile
  constexpr int B_LD = BN * BK / 4 / THREADS;
  constexpr int TPR = THREADS / BM;            // ATRANS: threads per A row
  static_assert(!ATRANS || (THREADS % BM == 0 && (4 % TPR) == 0 && A_LD == 4 / TPR), "A loader mapping");
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);

  // ATRANS: this thread always loads the same A row (a position) -> resolve its context once
  RowCtx actx;
  if (ATRANS) actx = row_ctx<AKIND>(g.A, m0 + tid % BM);

  float4 pa[A_LD], pb[B_LD];
  auto fetch = [&](int k0) {
    if (ATRANS) {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) pa[i] = load4<AKIND>(g.A, actx, k0 + ((tid / BM) * A_LD + i) * 4);
    } else {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int e = tid + i * THREADS, r = e / (BM / 4), c4 = (e % (BM / 4)) * 4;
        const int row = k0 + r;
        const RowCtx c = row_ctx<AKIND>(g.A, row < k_end ? row : 0x7fffffff);
        pa[i] = load4<AKIND>(g.A, c, m0 + c4);
      }
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int e = tid + i * THREADS, r = e / (BN / 4), c4 = (e % (BN / 4)) * 4;
      const int row = k0 + r;
      const RowCtx c = row_ctx<BKIND>(g.B, row < k_end ? row : 0x7fffffff);
      pb[i] = load4<BKIND>(g.B, c, n0 + c4);
    }
  };
  auto stage = [&](int buf) {
    if (ATRANS) {
      const int r = tid % BM;
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int kq = ((tid / BM) * A_LD + i) * 4;
        As[buf][kq + 0][r] = pa[i].x;
        As[buf][kq + 1][r] = pa[i].y;
        As[buf][kq + 2][r] = pa[i].z;
        As[buf][kq + 3][r] = pa[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        const int e = tid + i * THREADS, r = e / (BM / 4), c4 = (e % (BM / 4)) * 4;
        *reinterpret_cast<float4 *>(&As[buf][r][c4]) = pa[i];
      }
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int e = tid + i * THREADS, r = e / (BN / 4), c4 = (e % (BN / 4)) * 4;
      *reinterpret_cast<float4 *>(&Bs[buf][r][c4]) = pb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int buf = 0;
  if (k_begin < k_end) {
    fetch(k_begin);
    stage(0);
  }
  __syncthreads();
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][BM / 2 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][BN / 2 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) stage(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // ---- epilogue ----
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
    if (row >= g.M) continue;
    RowCtx gctx;
    if (EPI == EPI_SCATTER) gctx = row_ctx<PN2_ROWS_GATHER>(g.G, row);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = n0 + (h == 0 ? tx * 4 : BN / 2 + tx * 4);
      if (col >= g.N) continue;
      float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      if (EPI == EPI_STORE || EPI == EPI_STORE_STATS) {
        float *dst = g.out + blockIdx.z * g.out_split_stride + static_cast<size_t>(row) * g.ldo + col;
        *reinterpret_cast<float4 *>(dst) = v;
        if (EPI == EPI_STORE_STATS) {
          s1[h * 4 + 0] += v.x; s1[h * 4 + 1] += v.y; s1[h * 4 + 2] += v.z; s1[h * 4 + 3] += v.w;
          s2[h * 4 + 0] = fmaf(v.x, v.x, s2[h * 4 + 0]); s2[h * 4 + 1] = fmaf(v.y, v.y, s2[h * 4 + 1]);
          s2[h * 4 + 2] = fmaf(v.z, v.z, s2[h * 4 + 2]); s2[h * 4 + 3] = fmaf(v.w, v.w, s2[h * 4 + 3]);
        }
      } else if (EPI == EPI_DGRAD_MASK) {
        const float4 y = ldg4(g.prev_y + static_cast<size_t>(row) * g.ld_prev + col);
        const float4 sc = ldg4(g.prev_scale + col), sh = ldg4(g.prev_shift + col);
        v.x = fmaf(y.x, sc.x, sh.x) > 0.f ? v.x : 0.f;
        v.y = fmaf(y.y, sc.y, sh.y) > 0.f ? v.y : 0.f;
        v.z = fmaf(y.z, sc.z, sh.z) > 0.f ? v.z : 0.f;
        v.w = fmaf(y.w, sc.w, sh.w) > 0.f ? v.w : 0.f;
        *reinterpret_cast<float4 *>(g.out + static_cast<size_t>(row) * g.ldo + col) = v;
        s1[h * 4 + 0] += v.x; s1[h * 4 + 1] += v.y; s1[h * 4 + 2] += v.z; s1[h * 4 + 3] += v.w;
        s2[h * 4 + 0] = fmaf(v.x, y.x, s2[h * 4 + 0]); s2[h * 4 + 1] = fmaf(v.y, y.y, s2[h * 4 + 1]);
        s2[h * 4 + 2] = fmaf(v.z, y.z, s2[h * 4 + 2]); s2[h * 4 + 3] = fmaf(v.w, y.w, s2[h * 4 + 3]);
      } else {  // EPI_SCATTER: transpose of the gather
        if (col < g.G.feat_cols) {
          if (g.dfeat) {
            float *dst = g.dfeat + gctx.goff * g.ldf + col;
            atomicAdd(dst + 0, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
          }
        } else if (g.dxyz && g.G.use_xyz) {
          const float gx = __fdiv_rn(v.x, g.G.inv_scale), gy = __fdiv_rn(v.y, g.G.inv_scale),
                      gz = __fdiv_rn(v.z, g.G.inv_scale);
          float *dn = g.dxyz + gctx.goff * 3;
          atomicAdd(dn + 0, gx); atomicAdd(dn + 1, gy); atomicAdd(dn + 2, gz);
          const int centre = row / g.G.nsample, cloud = row / (g.G.npoint * g.G.nsample);
          float *dc = g.dxyz + (static_cast<size_t>(cloud) * g.G.n_src + __ldg(g.centre_src + centre)) * 3;
          atomicAdd(dc + 0, -gx); atomicAdd(dc + 1, -gy); atomicAdd(dc + 2, -gz);
        }
      }
    }
  }

  if (EPI == EPI_STORE_STATS || EPI == EPI_DGRAD_MASK) {
    if (g.stats == nullptr) return;
    // column sums over the tile's rows: registers -> shared (one row of partials per ty) -> BN threads
    float *red = &As[0][0][0];  // 2*BK*BM floats >= TY*BN floats; second quantity goes to Bs
    float *red2 = &Bs[0][0][0];
    static_assert(2 * BK * BM >= TY * BN && 2 * BK * BN >= TY * BN, "reduction scratch");
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = (h == 0 ? tx * 4 : BN / 2 + tx * 4) + q;
        red[ty * BN + c] = s1[h * 4 + q];
        red2[ty * BN + c] = s2[h * 4 + q];
      }
    __syncthreads();
    for (int c = tid; c < BN; c += THREADS) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int t = 0; t < TY; ++t) {
        a += red[t * BN + c];
        b += red2[t * BN + c];
      }
      if (n0 + c < g.stats_ld) {
        float *dst = g.stats + static_cast<size_t>(blockIdx.x) * 2 * g.stats_ld + n0 + c;
        dst[0] = a;
        dst[g.stats_ld] = b;
      }
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------
inline bool big_tile(long long m, long long n) {
  return ((m + 127) / 128) * ((n + 127) / 128) >= 96;  // else 64x64 tiles to keep the 148 SMs busy
}

template <int AKIND, bool ATRANS, int BKIND, int EPI>
int launch_gemm(const GemmArgs &g, int splits, cudaStream_t stream, const char *what) {
  // while the tcgen05 path is enabled the statistics buffers are sized for 128-row tiles (pn2_mlp_tiles), so a
  // call that falls back to this kernel (K beyond the tensor-core kernel's coefficient staging) keeps that tiling
  if (big_tile(g.M, g.N) || (gemm_tc_enabled() && g.stats != nullptr)) {
    dim3 grid((g.M + 127) / 128, (g.N + 127) / 128, splits);
    pn2::launch(gemm_kernel<128, 128, AKIND, ATRANS, BKIND, EPI>, dim3(grid), dim3(256), 0, stream, g);
  } else {
    dim3 grid((g.M + 63) / 64, (g.N + 63) / 64, splits);
    pn2::launch(gemm_kernel<64, 64, AKIND, ATRANS, BKIND, EPI>, dim3(grid), dim3(64), 0, stream, g);
  }
  return check_launch(what);
}

int check_rows(const pn2_rows *r, const char *what) {
  PN2_REQUIRE(r != nullptr, "%s: null row source", what);
  PN2_REQUIRE(r->rows >= 0 && r->cols >= 0 && (r->cols % 4) == 0 && (r->ld % 4) == 0, "%s: rows=%d cols=%d ld=%d must be >= 0 and multiples of 4",
              what, r->rows, r->cols, r->ld);
  switch (r->kind) {
    case PN2_ROWS_PLAIN: PN2_REQUIRE(r->x, "%s: null x", what); break;
    case PN2_ROWS_BNRELU: PN2_REQUIRE(r->x && r->c0 && r->c1, "%s: null x/scale/shift", what); break;
    case PN2_ROWS_GATHER:
      PN2_REQUIRE(r->idx && r->n_src > 0 && r->npoint > 0 && r->nsample > 0, "%s: bad gather geometry", what);
      PN2_REQUIRE((r->feat_cols == 0 || r->x) && (r->feat_cols % 4) == 0, "%s: bad gather features", what);
      PN2_REQUIRE(!r->use_xyz || (r->xyz && r->centres && r->inv_scale != 0.f), "%s: gather needs xyz/centres", what);
      PN2_REQUIRE(r->cols == r->feat_cols + (r->use_xyz ? 4 : 0), "%s: gather cols=%d != feat_cols+4", what, r->cols);
      break;
    case PN2_ROWS_DY: PN2_REQUIRE(r->x && r->dz && r->c0 && r->c1 && r->c2, "%s: null dy inputs", what); break;
    case PN2_ROWS_DYPOOL:
      PN2_REQUIRE(r->x && r->dz && r->arg && r->c0 && r->c1 && r->c2 && r->group > 0 && r->group <= 256,
                  "%s: bad pooled dy inputs", what);
      break;
    default: PN2_REQUIRE(false, "%s: unknown row source kind %d", what, r->kind);
  }
  return PN2_OK;
}

pn2_rows plain_rows(const float *x, int rows, int cols, int ld) {
  pn2_rows r = {};
  r.kind = PN2_ROWS_PLAIN;
  r.rows = rows; r.cols = cols; r.ld = ld; r.x = x;
  return r;
}

// ---- weight preparation / wgrad reduction -------------------------------------------------------------
// k' (permuted, padded input channel) -> original input channel, or -1 for padding
__device__ __forceinline__ int orig_cin(int kq, int cin, int xyz_first, int feat_pad) {
  if (!xyz_first) return kq < cin ? kq : -1;
  if (kq < feat_pad) return kq < cin - 3 ? kq + 3 : -1;
  const int d = kq - feat_pad;
  return d < 3 ? d : -1;
}

// Tensor-core image of a K-major operand matrix B [rows][cols] (mlp_gemm_tc.cu, gemm_tc_async_kernel): for every
// (R-row tile, 32-column k-block) one stage = [hi R x 32 | lo R x 32] floats, each half in the UMMA K-major
// 128-byte-swizzle layout (row r at (r/8)*1024 + (r%8)*128 bytes, 16-byte chunk c XOR (r%8)), zero padded.
// R = 256 when the matrix has a multiple of 256 rows (the kernel then runs 128 x 256 tiles), else 128.
inline int img_tile_rows(int rows) { return (gemm_tc_wide_enabled() && rows % 256 == 0) ? 256 : 128; }
inline long long img_floats(int rows, int cols) {  // (the same for either tile height)
  return static_cast<long long>((rows + 127) / 128) * ((cols + 31) / 32) * (2 * 128 * 32);
}
// e enumerates (tile, k-block, row-in-tile, col-in-block) of a matrix imaged with R-row tiles
__device__ __forceinline__ void img_store(float *img, int R, long long e, float v) {
  const int c = static_cast<int>(e & 31), r = static_cast<int>((e >> 5) % R);
  const long long stage = e / (32 * R);
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  const float hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));  // rounded, not left to the tensor core's truncation
  const float lo = __uint_as_float(l);
  const int off = (r >> 3) * 256 + (r & 7) * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3));  // floats inside a half
  float *st = img + stage * (2 * R * 32);
  st[off] = hi;
  st[R * 32 + off] = lo;
}

__global__ void prep_weights_kernel(int cout, int cin, int xyz_first, int feat_pad, int kp, int np,
                                    const float *__restrict__ w, float *__restrict__ wt, float *__restrict__ wp,
                                    long long img_t, long long img_p, int rows_t, int rows_p) {
  pdl_prologue();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<long long>(kp) * np) {
    {  // wt[k][n]
      const int k = static_cast<int>(i / np), n = static_cast<int>(i % np);
      const int c = orig_cin(k, cin, xyz_first, feat_pad);
      wt[i] = (n < cout && c >= 0) ? w[static_cast<size_t>(n) * cin + c] : 0.f;
    }
    {  // wp[n][k]
      const int n = static_cast<int>(i / kp), k = static_cast<int>(i % kp);
      const int c = orig_cin(k, cin, xyz_first, feat_pad);
      wp[i] = (n < cout && c >= 0) ? w[static_cast<size_t>(n) * cin + c] : 0.f;
    }
  }
  if (i < img_t) {  // image of wt: operand rows = k (kp of them), operand columns = n
    const int R = rows_t, nkb = (np + 31) / 32;
    const long long stage = i / (32 * R);
    const int k = static_cast<int>(stage / nkb) * R + static_cast<int>((i >> 5) % R);
    const int n = static_cast<int>(stage % nkb) * 32 + static_cast<int>(i & 31);
    const int c = k < kp ? orig_cin(k, cin, xyz_first, feat_pad) : -1;
    img_store(wt + static_cast<size_t>(kp) * np, R, i, (n < cout && c >= 0) ? w[static_cast<size_t>(n) * cin + c] : 0.f);
  }
  if (i < img_p) {  // image of wp: operand rows = n, operand columns = k
    const int R = rows_p, nkb = (kp + 31) / 32;
    const long long stage = i / (32 * R);
    const int n = static_cast<int>(stage / nkb) * R + static_cast<int>((i >> 5) % R);
    const int k = static_cast<int>(stage % nkb) * 32 + static_cast<int>(i & 31);
    const int c = k < kp ? orig_cin(k, cin, xyz_first, feat_pad) : -1;
    img_store(wp + static_cast<size_t>(kp) * np, R, i, (n < cout && c >= 0) ? w[static_cast<size_t>(n) * cin + c] : 0.f);
  }
}

__global__ void wgrad_reduce_kernel(int cout, int cin, int xyz_first, int feat_pad, int kp, int np, int splits,
                                    const float *__restrict__ ws, float *__restrict__ dw) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * kp) return;
  const int n = i / kp, k = i % kp;
  const int c = orig_cin(k, cin, xyz_first, feat_pad);
  if (c < 0) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[(static_cast<size_t>(z) * np + n) * kp + k];
  dw[static_cast<size_t>(n) * cin + c] = s;
}

// Same reduction with the slices spread over the warps of a block: lane = element (consecutive k, coalesced), warp w
// sums slices w, w + nw, ... and the nw partials are combined in a fixed order through shared memory.  A thread per
// element walking hundreds of 4 KB-strided slices one after the other was latency-bound (the 128 x 8 first-layer
// gradient: ~100 us of a 125 us call).  Deterministic: the summation order depends on (splits, nw) only.
__global__ void __launch_bounds__(512)
wgrad_reduce_wide_kernel(int cout, int cin, int xyz_first, int feat_pad, int kp, int np, int splits,
                         const float *__restrict__ ws, float *__restrict__ dw) {
  pdl_prologue();
  __shared__ float part[16][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int i = blockIdx.x * 32 + lane;  // element of the [cout][kp] gradient
  const bool live = i < cout * kp;
  const int n = live ? i / kp : 0, k = live ? i % kp : 0;
  float s = 0.f;
  if (live)
    for (int z = warp; z < splits; z += nw) s += ws[(static_cast<size_t>(z) * np + n) * kp + k];
  part[warp][lane] = s;
  __syncthreads();
  if (warp != 0 || !live) return;
  const int c = orig_cin(k, cin, xyz_first, feat_pad);
  if (c < 0) return;
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += part[w][lane];
  dw[static_cast<size_t>(n) * cin + c] = t;
}

// fold: the 4-column xyz block of a gathered source rides in the last feature tile (wgrad_fold_ok) -- one n-tile fewer
int wgrad_splits(int rows, int np, int kp, bool fold = false) {
  if (gemm_tc_enabled()) {  // tcgen05 kernel: 128x128 tiles, 1 CTA/SM, position slices of >= 4 k-blocks of 32
    const long long t = ((np + 127) / 128) * ((kp + 127) / 128 - (fold ? 1 : 0));
    // tiles * splits <= SMs: one wave (rounding up gave e.g. 6 x 25 = 150 CTAs on 148 SMs, i.e. a 2-CTA second wave).
    // Leaving 16 SMs to the geometry stream here as the persistent GEMMs do made no difference (3.079 vs 3.076 ms per step).
    long long s = static_cast<long long>(sm_count()) / t;
    const long long cap = (rows + 127) / 128;
    if (s > cap) s = cap;
    if (s < 1) s = 1;
    if (s > 512) s = 512;
    return static_cast<int>(s);
  }
  const long long tiles = big_tile(np, kp) ? ((np + 127) / 128) * ((kp + 127) / 128) : ((np + 63) / 64) * ((kp + 63) / 64);
  const int per_sm = big_tile(np, kp) ? 2 : 6;
  long long s = (static_cast<long long>(sm_count()) * per_sm + tiles - 1) / tiles;
  const long long max_s = (rows + 4 * BK - 1) / (4 * BK);  // at least 4 k-tiles per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  return static_cast<int>(s);
}

}  // namespace
}  // namespace pn2

using namespace pn2;

PN2_EXPORT long long pn2_mlp_weight_floats(int rows, int cols) {
  return static_cast<long long>(rows) * cols + img_floats(rows, cols);
}

PN2_EXPORT int pn2_mlp_prep_weights(int cout, int cin, int xyz_first, int feat_pad, int kp, int np, const float *w,
                                    float *wt, float *wp, void *stream) {
  PN2_REQUIRE(cout > 0 && cin > 0 && kp >= 4 && np >= cout && (kp % 4) == 0 && (np % 4) == 0 && w && wt && wp,
              "pn2_mlp_prep_weights: bad arguments cout=%d cin=%d kp=%d np=%d", cout, cin, kp, np);
  PN2_REQUIRE(xyz_first ? (cin >= 3 && feat_pad >= cin - 3 && kp == feat_pad + 4) : kp >= cin,
              "pn2_mlp_prep_weights: inconsistent padding cin=%d feat_pad=%d kp=%d", cin, feat_pad, kp);
  const long long img_t = img_floats(kp, np) / 2, img_p = img_floats(np, kp) / 2;  // one thread per (hi, lo) pair
  long long total = static_cast<long long>(kp) * np;
  if (img_t > total) total = img_t;
  if (img_p > total) total = img_p;
  pn2::launch(prep_weights_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      cout, cin, xyz_first, feat_pad, kp, np, w, wt, wp, img_t, img_p, img_tile_rows(kp), img_tile_rows(np));
  return check_launch("pn2_mlp_prep_weights");
}

PN2_EXPORT int pn2_mlp_tiles(int rows, int ncols) {
  if (gemm_tc_enabled()) return (rows + 127) / 128;  // the tcgen05 kernel always tiles rows by 128
  return big_tile(rows, ncols) ? (rows + 127) / 128 : (rows + 63) / 64;
}

PN2_EXPORT int pn2_mlp_forward(const pn2_rows *a, int kp, int np, const float *wt, const float *wp, float *y,
                               int ldy, float *stats, int *tiles, void *stream_) {
  if (int rc = check_rows(a, "pn2_mlp_forward")) return rc;
  PN2_REQUIRE(a->cols == kp && (np % 4) == 0 && np > 0 && wt && y && ldy >= np && (ldy % 4) == 0,
              "pn2_mlp_forward: operand widths disagree (a.cols=%d kp=%d np=%d ldy=%d)", a->cols, kp, np, ldy);
  PN2_REQUIRE(a->kind == PN2_ROWS_PLAIN || a->kind == PN2_ROWS_BNRELU || a->kind == PN2_ROWS_GATHER,
              "pn2_mlp_forward: unsupported row source %d", a->kind);
  if (tiles) *tiles = pn2_mlp_tiles(a->rows, np);
  if (a->rows == 0) return PN2_OK;
  GemmArgs g = {};
  g.A = *a;
  g.B = plain_rows(wt, kp, np, np);
  g.M = a->rows; g.N = np; g.K = kp;
  g.k_per_split = (kp + BK - 1) / BK * BK;
  g.out = y; g.ldo = ldy;
  g.stats = stats; g.stats_ld = np;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (wp != nullptr && gemm_tc_enabled()) {  // tensor-core path: B = W as [np][kp], K-major
    GemmArgs t = g;
    t.B = plain_rows(wp, np, kp, kp);
    t.b_img = wp + static_cast<size_t>(np) * kp;
    t.b_img_kblocks = (kp + 31) / 32;
    t.b_tile_rows = img_tile_rows(np);
    const int rc = gemm_tc_launch(a->kind, EPI_STORE_STATS, &t, s);
    if (rc != PN2_TC_UNSUPPORTED) return rc;
  }
  switch (a->kind) {
    case PN2_ROWS_PLAIN: return launch_gemm<PN2_ROWS_PLAIN, true, PN2_ROWS_PLAIN, EPI_STORE_STATS>(g, 1, s, "pn2_mlp_forward");
    case PN2_ROWS_BNRELU: return launch_gemm<PN2_ROWS_BNRELU, true, PN2_ROWS_PLAIN, EPI_STORE_STATS>(g, 1, s, "pn2_mlp_forward");
    default: return launch_gemm<PN2_ROWS_GATHER, true, PN2_ROWS_PLAIN, EPI_STORE_STATS>(g, 1, s, "pn2_mlp_forward");
  }
}

PN2_EXPORT int pn2_mlp_dgrad(int mode, const pn2_rows *dy, int ncols, const float *wp, int ldw, const float *wt,
                             float *out, int ldo,
                             const float *prev_y, int ld_prev, const float *prev_scale, const float *prev_shift,
                             float *stats, int *tiles, const pn2_rows *gather, float *dfeat, int ldf, float *dxyz,
                             const int *centre_src, void *stream_) {
  if (int rc = check_rows(dy, "pn2_mlp_dgrad")) return rc;
  PN2_REQUIRE(dy->kind == PN2_ROWS_DY || dy->kind == PN2_ROWS_DYPOOL, "pn2_mlp_dgrad: dy must be a DY / DYPOOL source");
  PN2_REQUIRE(wp && ncols > 0 && (ncols % 4) == 0 && ldw >= ncols && (ldw % 4) == 0, "pn2_mlp_dgrad: bad weight operand");
  if (tiles) *tiles = pn2_mlp_tiles(dy->rows, ncols);
  if (dy->rows == 0) return PN2_OK;
  GemmArgs g = {};
  g.A = *dy;
  g.B = plain_rows(wp, dy->cols, ncols, ldw);
  g.M = dy->rows; g.N = ncols; g.K = dy->cols;
  g.k_per_split = (g.K + BK - 1) / BK * BK;
  g.out = out; g.ldo = ldo;
  g.stats = stats; g.stats_ld = ncols;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const bool pool = dy->kind == PN2_ROWS_DYPOOL;
  const bool tc = wt != nullptr && gemm_tc_enabled() && mode != PN2_DGRAD_SCATTER;
  GemmArgs t;
  if (mode == PN2_DGRAD_MASK) {
    PN2_REQUIRE(out && prev_y && prev_scale && prev_shift && ldo >= ncols && ld_prev >= ncols, "pn2_mlp_dgrad: MASK needs out/prev_*");
    g.prev_y = prev_y; g.prev_scale = prev_scale; g.prev_shift = prev_shift; g.ld_prev = ld_prev;
    if (tc) {  // B = W^T as [ncols][dy.cols], K-major: that is wt
      t = g;
      t.B = plain_rows(wt, ncols, dy->cols, dy->cols);
      t.b_img = wt + static_cast<size_t>(ncols) * dy->cols;
      t.b_img_kblocks = (dy->cols + 31) / 32;
      t.b_tile_rows = img_tile_rows(ncols);
      const int rc = gemm_tc_launch(dy->kind, EPI_DGRAD_MASK, &t, s);
      if (rc != PN2_TC_UNSUPPORTED) return rc;
    }
    return pool ? launch_gemm<PN2_ROWS_DYPOOL, true, PN2_ROWS_PLAIN, EPI_DGRAD_MASK>(g, 1, s, "pn2_mlp_dgrad")
                : launch_gemm<PN2_ROWS_DY, true, PN2_ROWS_PLAIN, EPI_DGRAD_MASK>(g, 1, s, "pn2_mlp_dgrad");
  }
  if (mode == PN2_DGRAD_STORE) {
    PN2_REQUIRE(out && ldo >= ncols, "pn2_mlp_dgrad: STORE needs out");
    if (tc) {
      t = g;
      t.B = plain_rows(wt, ncols, dy->cols, dy->cols);
      t.b_img = wt + static_cast<size_t>(ncols) * dy->cols;
      t.b_img_kblocks = (dy->cols + 31) / 32;
      t.b_tile_rows = img_tile_rows(ncols);
      const int rc = gemm_tc_launch(dy->kind, EPI_STORE, &t, s);
      if (rc != PN2_TC_UNSUPPORTED) return rc;
    }
    return pool ? launch_gemm<PN2_ROWS_DYPOOL, true, PN2_ROWS_PLAIN, EPI_STORE>(g, 1, s, "pn2_mlp_dgrad")
                : launch_gemm<PN2_ROWS_DY, true, PN2_ROWS_PLAIN, EPI_STORE>(g, 1, s, "pn2_mlp_dgrad");
  }
  PN2_REQUIRE(mode == PN2_DGRAD_SCATTER, "pn2_mlp_dgrad: unknown mode %d", mode);
  if (int rc = check_rows(gather, "pn2_mlp_dgrad(gather)")) return rc;
  PN2_REQUIRE(gather->kind == PN2_ROWS_GATHER && gather->rows == dy->rows && gather->cols == ncols,
              "pn2_mlp_dgrad: gather source does not match (rows %d vs %d, cols %d vs %d)", gather->rows, dy->rows,
              gather->cols, ncols);
  PN2_REQUIRE((dfeat == nullptr || ldf >= gather->feat_cols) && (dxyz == nullptr || centre_src != nullptr),
              "pn2_mlp_dgrad: SCATTER targets inconsistent");
  g.G = *gather; g.dfeat = dfeat; g.ldf = ldf; g.dxyz = dxyz; g.centre_src = centre_src;
  // without a coordinate gradient the xyz block (the last 4 columns) is never consumed: do not compute it
  // (for 256 / 512 feature channels that is a whole extra 128-column tile)
  if (dxyz == nullptr && gather->feat_cols > 0) g.N = gather->feat_cols;
  if (dfeat == nullptr && dxyz == nullptr) return PN2_OK;
  if (wt != nullptr && gemm_tc_enabled()) {
    GemmArgs t2 = g;
    t2.B = plain_rows(wt, g.N, dy->cols, dy->cols);
    t2.b_img = wt + static_cast<size_t>(ncols) * dy->cols;  // the image follows the full [ncols][dy.cols] matrix
    t2.b_img_kblocks = (dy->cols + 31) / 32;
    t2.b_tile_rows = img_tile_rows(ncols);
    const int rc = gemm_tc_launch(dy->kind, EPI_SCATTER, &t2, s);
    if (rc != PN2_TC_UNSUPPORTED) return rc;
  }
  return pool ? launch_gemm<PN2_ROWS_DYPOOL, true, PN2_ROWS_PLAIN, EPI_SCATTER>(g, 1, s, "pn2_mlp_dgrad")
              : launch_gemm<PN2_ROWS_DY, true, PN2_ROWS_PLAIN, EPI_SCATTER>(g, 1, s, "pn2_mlp_dgrad");
}

PN2_EXPORT long long pn2_mlp_wgrad_workspace(int rows, int np, int kp) {
  const int a = wgrad_splits(rows, np, kp), b = kp > 128 && kp % 128 == 4 ? wgrad_splits(rows, np, kp, true) : 0;
  return static_cast<long long>(a > b ? a : b) * np * kp;
}

PN2_EXPORT int pn2_mlp_wgrad(const pn2_rows *dy, const pn2_rows *a, int cout, int cin, int xyz_first, int feat_pad,
                             float *ws, float *dw, void *stream_) {
  if (int rc = check_rows(dy, "pn2_mlp_wgrad(dy)")) return rc;
  if (int rc = check_rows(a, "pn2_mlp_wgrad(a)")) return rc;
  PN2_REQUIRE(dy->kind == PN2_ROWS_DY || dy->kind == PN2_ROWS_DYPOOL, "pn2_mlp_wgrad: dy must be a DY / DYPOOL source");
  PN2_REQUIRE(a->kind == PN2_ROWS_PLAIN || a->kind == PN2_ROWS_BNRELU || a->kind == PN2_ROWS_GATHER,
              "pn2_mlp_wgrad: unsupported activation source %d", a->kind);
  PN2_REQUIRE(dy->rows == a->rows && ws && dw && cout <= dy->cols && cin <= a->cols, "pn2_mlp_wgrad: shapes disagree");
  const int np = dy->cols, kp = a->cols, rows = dy->rows;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  GemmArgs g = {};
  g.A = *dy; g.B = *a;
  g.M = np; g.N = kp; g.K = rows;
  g.wg_fold = wgrad_fold_ok(&g) ? 1 : 0;
  int splits = wgrad_splits(rows, np, kp, g.wg_fold != 0);
  g.k_per_split = ((rows + splits - 1) / splits + 31) / 32 * 32;  // multiple of both kernels' k-block
  g.out = ws; g.ldo = kp; g.out_split_stride = static_cast<long long>(np) * kp;
  int rc = PN2_TC_UNSUPPORTED;
  if (rows > 0 && rc == PN2_TC_UNSUPPORTED && gemm_tc_enabled()) {
    rc = gemm_tc_wgrad_launch(&g, splits, s);
    if (rc != PN2_TC_UNSUPPORTED && rc != PN2_OK) return rc;
  }
  if (rows > 0 && rc == PN2_TC_UNSUPPORTED) {
#define PN2_WGRAD(DK, AK) launch_gemm<DK, false, AK, EPI_STORE>(g, splits, s, "pn2_mlp_wgrad")
    const bool pool = dy->kind == PN2_ROWS_DYPOOL;
    switch (a->kind) {
      case PN2_ROWS_PLAIN: rc = pool ? PN2_WGRAD(PN2_ROWS_DYPOOL, PN2_ROWS_PLAIN) : PN2_WGRAD(PN2_ROWS_DY, PN2_ROWS_PLAIN); break;
      case PN2_ROWS_BNRELU: rc = pool ? PN2_WGRAD(PN2_ROWS_DYPOOL, PN2_ROWS_BNRELU) : PN2_WGRAD(PN2_ROWS_DY, PN2_ROWS_BNRELU); break;
      default: rc = pool ? PN2_WGRAD(PN2_ROWS_DYPOOL, PN2_ROWS_GATHER) : PN2_WGRAD(PN2_ROWS_DY, PN2_ROWS_GATHER); break;
    }
#undef PN2_WGRAD
    if (rc) return rc;
  }
  const int total = cout * kp;
  static const bool wide = [] {
    const char *e = getenv("PN2_WGRAD_REDUCE_WIDE");
    return e == nullptr || e[0] != '0';
  }();
  const int nsl = rows > 0 ? splits : 0;
  if (wide && nsl >= 64) {  // few, large slices (big outputs): the thread-per-element kernel is faster (measured)
    int nw = 1;
    while (nw < 16 && nw < nsl) nw *= 2;
    pn2::launch(wgrad_reduce_wide_kernel, dim3((total + 31) / 32), dim3(32 * nw), 0, s, cout, cin, xyz_first, feat_pad, kp, np, nsl,
                ws, dw);
  } else {
    pn2::launch(wgrad_reduce_kernel, dim3((total + 255) / 256), dim3(256), 0, s, cout, cin, xyz_first, feat_pad, kp, np, nsl, ws, dw);
  }
  return check_launch("pn2_mlp_wgrad(reduce)");
}
