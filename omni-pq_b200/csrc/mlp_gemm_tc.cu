// Shared-MLP contraction on the 5th-generation tensor cores (tcgen05, sm_100a): same operands, row sources
// and epilogues as the FFMA kernel in mlp_gemm.cu, for the forward and data-gradient GEMMs
// (C[M x N] = A[M x K] * B[N x K]^T with both operands K-major).
//
// Precision: the parity bar of the float paths is 1e-5 relative, which a single TF32 pass (10-bit mantissa)
// cannot meet.  Every fp32 operand is split while it is staged into  v = hi + lo  (hi = nearest tf32,
// lo = v - hi exactly) and each k-step issues three kind::tf32 MMAs, hi*hi + hi*lo + lo*hi, accumulated in
// fp32 in tensor memory; the dropped lo*lo term is below 2^-22 relative per product.
//
// Structure (one CTA = one 128 x 128 output tile, 256 threads, 1 CTA/SM):
//   * all 8 warps are producers: they evaluate the row source (gather / BN+ReLU / BN-backward ...) for a
//     128 x 32 k-block of A and of B, split it and store hi/lo tiles in the canonical K-major 128-byte-swizzle
//     layout the UMMA descriptor expects (conflict-free 128-bit stores), 3-stage ring;
//   * fence.proxy.async + one barrier per k-block, then ONE thread issues 12 tcgen05.mma (4 k-steps x 3 split
//     terms) and commits them to the stage's "empty" mbarrier, so the tensor pipe works on block i while the
//     producers stage block i+1, i+2;
//   * epilogue: accumulator tile read back with tcgen05.ld (each thread one row, 32 columns at a time),
//     stored with 128-byte row segments; BatchNorm partial column sums by a warp butterfly transpose-reduce.
// The operands are produced by threads rather than TMA because every A element needs an element-wise
// transform (that fusion is the point of the kernel); the weights are small and L2 resident.
#include "mlp_rows.cuh"
#include "pn2_sm100.cuh"

namespace pn2 {
namespace {

using namespace sm100;

constexpr int TM = 128, TN = 128, TK = 32;          // CTA tile; TK fp32 = one 128-byte swizzle row
constexpr int TC_THREADS = 256;
constexpr int TC_STAGES = 3;
constexpr int TILE_BYTES = TM * TK * 4;             // 16 KB per operand half
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // A_hi, A_lo, B_hi, B_lo
constexpr int TC_SMEM = TC_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers, scratch*/;
constexpr uint32_t TMEM_COLS = 128;

enum { TC_EPI_STORE = 0, TC_EPI_STORE_STATS = 1, TC_EPI_DGRAD_MASK = 2 };

// column totals over the 32 lanes of a warp: afterwards lane l holds the sum of v[l] over all lanes
__device__ __forceinline__ float warp_column_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float send = upper ? v[j] : v[j + o];
      const float keep = upper ? v[j + o] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

template <int AKIND, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *empty_bar = reinterpret_cast<uint64_t *>(tiles + TC_STAGES * STAGE_BYTES);  // [TC_STAGES]
  uint64_t *done_bar = empty_bar + TC_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
  __shared__ float red[2][8][32];  // per-warp column partials for the statistics epilogues

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(&empty_bar[s], 1);
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  // producer mapping: thread -> one tile row (A: a position, B: an output channel) and 4 of its 8 chunks
  const int prow = tid & (TM - 1);
  const int chunk0 = (tid >> 7) * 4;
  const RowCtx actx = row_ctx<AKIND>(g.A, m0 + prow);
  const RowCtx bctx = row_ctx<PN2_ROWS_PLAIN>(g.B, n0 + prow);
  uint32_t off[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) off[i] = sw128_offset(prow, chunk0 + i);

  const uint32_t idesc = idesc_tf32(TM, TN);
  const int num_kb = (g.K + TK - 1) / TK;

  for (int kb = 0; kb < num_kb; ++kb) {
    const int s = kb % TC_STAGES;
    if (kb >= TC_STAGES) mbar_wait(&empty_bar[s], ((kb / TC_STAGES) - 1) & 1);  // MMAs of block kb-STAGES done
    unsigned char *st = tiles + s * STAGE_BYTES;
    const int k0 = kb * TK;
    float4 va[4], vb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      va[i] = load4<AKIND>(g.A, actx, k0 + (chunk0 + i) * 4);
      vb[i] = load4<PN2_ROWS_PLAIN>(g.B, bctx, k0 + (chunk0 + i) * 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 hi, lo;
      split_tf32(va[i].x, hi.x, lo.x); split_tf32(va[i].y, hi.y, lo.y);
      split_tf32(va[i].z, hi.z, lo.z); split_tf32(va[i].w, hi.w, lo.w);
      *reinterpret_cast<float4 *>(st + 0 * TILE_BYTES + off[i]) = hi;
      *reinterpret_cast<float4 *>(st + 1 * TILE_BYTES + off[i]) = lo;
      split_tf32(vb[i].x, hi.x, lo.x); split_tf32(vb[i].y, hi.y, lo.y);
      split_tf32(vb[i].z, hi.z, lo.z); split_tf32(vb[i].w, hi.w, lo.w);
      *reinterpret_cast<float4 *>(st + 2 * TILE_BYTES + off[i]) = hi;
      *reinterpret_cast<float4 *>(st + 3 * TILE_BYTES + off[i]) = lo;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
      const uint32_t base = smem_addr(st);
      const uint64_t a_hi = smem_desc_sw128(base), a_lo = smem_desc_sw128(base + TILE_BYTES);
      const uint64_t b_hi = smem_desc_sw128(base + 2 * TILE_BYTES), b_lo = smem_desc_sw128(base + 3 * TILE_BYTES);
#pragma unroll
      for (int ks = 0; ks < TK / 8; ++ks) {
        const uint64_t adv = static_cast<uint64_t>(ks * 32 >> 4);  // +32 bytes per k-step inside the swizzle row
        mma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, kb > 0 || ks > 0);
        mma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, true);
        mma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, true);
      }
      mma_commit(&empty_bar[s]);
      if (kb == num_kb - 1) mma_commit(done_bar);
    }
  }
  mbar_wait(done_bar, 0);
  tc_fence_after_sync();

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (rows), warps 0-3 columns 0-63, warps 4-7 columns 64-127
  const int row = m0 + (warp & 3) * 32 + lane;
  const int cbase = (warp >> 2) * 64;
  const bool row_ok = row < g.M;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c_local = cbase + cc * 32;
    const int col = n0 + c_local;
    float v[32];
    tmem_ld32(tmem_d + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(c_local), v);
    float q[32];  // second statistic operand
    if (EPI == TC_EPI_DGRAD_MASK) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 y = zero4(), sc = zero4(), sh = zero4();
        if (row_ok && col + j < g.N) {
          y = ldg4(g.prev_y + static_cast<size_t>(row) * g.ld_prev + col + j);
          sc = ldg4(g.prev_scale + col + j);
          sh = ldg4(g.prev_shift + col + j);
        }
        v[j + 0] = fmaf(y.x, sc.x, sh.x) > 0.f ? v[j + 0] : 0.f;
        v[j + 1] = fmaf(y.y, sc.y, sh.y) > 0.f ? v[j + 1] : 0.f;
        v[j + 2] = fmaf(y.z, sc.z, sh.z) > 0.f ? v[j + 2] : 0.f;
        v[j + 3] = fmaf(y.w, sc.w, sh.w) > 0.f ? v[j + 3] : 0.f;
        q[j + 0] = v[j + 0] * y.x; q[j + 1] = v[j + 1] * y.y; q[j + 2] = v[j + 2] * y.z; q[j + 3] = v[j + 3] * y.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) q[j] = v[j] * v[j];
    }
    if (row_ok) {
      float *dst = g.out + static_cast<size_t>(row) * g.ldo + col;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (col + j < g.N) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (EPI != TC_EPI_STORE && g.stats != nullptr) {
      // rows beyond M hold exact zeros (their A rows were zero), so they do not disturb the sums
      const float s1 = warp_column_sum(v, lane);
      const float s2 = warp_column_sum(q, lane);
      red[0][warp][lane] = s1;
      red[1][warp][lane] = s2;
      __syncthreads();
      if (tid < 64) {  // tid -> (half h = tid/32 selects warps 4h..4h+3, column lane)
        const int h = tid >> 5, l = tid & 31;
        const int c = n0 + h * 64 + cc * 32 + l;
        if (c < g.stats_ld) {
          const float a = red[0][4 * h][l] + red[0][4 * h + 1][l] + red[0][4 * h + 2][l] + red[0][4 * h + 3][l];
          const float b = red[1][4 * h][l] + red[1][4 * h + 1][l] + red[1][4 * h + 2][l] + red[1][4 * h + 3][l];
          float *dst = g.stats + static_cast<size_t>(blockIdx.x) * 2 * g.stats_ld + c;
          dst[0] = a;
          dst[g.stats_ld] = b;
        }
      }
      __syncthreads();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_d);
}

template <int AKIND, int EPI>
int launch_tc(const GemmArgs &g, cudaStream_t stream) {
  auto kernel = gemm_tc_kernel<AKIND, EPI>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    configured_dev = dev;
  }
  dim3 grid((g.M + TM - 1) / TM, (g.N + TN - 1) / TN);
  kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(g);
  return check_launch("gemm_tc_kernel");
}

}  // namespace

bool gemm_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("PN2_TC");
    on = (e == nullptr || e[0] != '0') ? 1 : 0;  // default on; PN2_TC=0 selects the FFMA kernel everywhere
  }
  return on == 1;
}

// epi uses mlp_gemm.cu's numbering: 0 store, 1 store+stats, 2 dgrad mask.  Tiles follow mlp_gemm.cu's
// 128-row tiling, so the statistics buffer sized by pn2_mlp_tiles() fits.
int gemm_tc_launch(int akind, int epi, const void *gemm_args, cudaStream_t stream) {
  const GemmArgs &g = *static_cast<const GemmArgs *>(gemm_args);
  if (!gemm_tc_enabled() || g.B.kind != PN2_ROWS_PLAIN || epi > 2) return PN2_TC_UNSUPPORTED;
#define PN2_TC_CASE(AK, EP) \
  if (akind == AK && epi == EP) return launch_tc<AK, EP>(g, stream);
  PN2_TC_CASE(PN2_ROWS_PLAIN, TC_EPI_STORE_STATS)
  PN2_TC_CASE(PN2_ROWS_BNRELU, TC_EPI_STORE_STATS)
  PN2_TC_CASE(PN2_ROWS_GATHER, TC_EPI_STORE_STATS)
  PN2_TC_CASE(PN2_ROWS_DY, TC_EPI_DGRAD_MASK)
  PN2_TC_CASE(PN2_ROWS_DYPOOL, TC_EPI_DGRAD_MASK)
  PN2_TC_CASE(PN2_ROWS_DY, TC_EPI_STORE)
  PN2_TC_CASE(PN2_ROWS_DYPOOL, TC_EPI_STORE)
#undef PN2_TC_CASE
  return PN2_TC_UNSUPPORTED;
}

}  // namespace pn2
