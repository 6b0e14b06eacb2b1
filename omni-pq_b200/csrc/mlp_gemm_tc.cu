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
#include <atomic>

#include "mlp_rows.cuh"
#include "pn2_sm100.cuh"

namespace pn2 {
namespace {

using namespace sm100;

constexpr int TM = 128, TN = 128, TK = 32;          // CTA tile; TK fp32 = one 128-byte swizzle row
constexpr int TC_THREADS = 512;                    // 16 producer / epilogue warps
constexpr int TC_CTA_THREADS = TC_THREADS + 32;    // + one warp whose lane 0 only issues the MMAs
constexpr int TC_ROWS_PER_THREAD = TM * 8 / TC_THREADS;  // 16-byte chunks of an operand k-block per thread (2)
constexpr int TC_STAGES = 3;
constexpr int TILE_BYTES = TM * TK * 4;             // 16 KB per operand half
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // A_hi, A_lo, B_hi, B_lo
constexpr int TC_KMAX = 1152;                        // largest K of the non-transposed form (coefficient staging)
constexpr int TC_COEF_FLOATS = 3 * TC_KMAX + 2 * 128; // A: up to 3 vectors over K (or over 128 tile channels); B: 2 x 128
constexpr int TC_SMEM = TC_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers, scratch*/ + TC_COEF_FLOATS * 4;
constexpr uint32_t TMEM_COLS = 128;

enum { TC_EPI_STORE = 0, TC_EPI_STORE_STATS = 1, TC_EPI_DGRAD_MASK = 2, TC_EPI_SCATTER = 3 };


// ---- two-phase row sources: issue the global loads of the NEXT k-block (fetch), run the element-wise
// transform of the CURRENT one (apply) with per-channel coefficients taken from shared memory, so a warp
// never stalls on a load it has just issued -------------------------------------------------------------
struct Raw {
  float4 x;   // matrix value (PLAIN / BNRELU / GATHER features / DY*: pre-BN y)
  float4 d;   // DY: dz; DYPOOL: pooled gz
  uchar4 a;   // DYPOOL: arg-max slot
};

template <int KIND>
__device__ __forceinline__ Raw fetch_raw(const pn2_rows &s, const RowCtx &c, int c4) {
  Raw r;
  r.x = zero4(); r.d = zero4(); r.a = make_uchar4(0, 0, 0, 0);
  if (!c.valid || c4 >= s.cols) return r;
  if (KIND == PN2_ROWS_GATHER) {
    if (c4 < s.feat_cols) r.x = ldg4(s.x + c.off + c4);
    return r;
  }
  r.x = ldg4(s.x + c.off + c4);
  if (KIND == PN2_ROWS_DY) r.d = ldg4(s.dz + c.off + c4);
  if (KIND == PN2_ROWS_DYPOOL) {
    r.d = ldg4(s.dz + c.goff + c4);
    r.a = __ldg(reinterpret_cast<const uchar4 *>(s.arg + c.goff + c4));
  }
  return r;
}

// coef: shared-memory copies of the source's per-channel vectors c0,c1,c2, indexed by (c4 - coef_base)
template <int KIND>
__device__ __forceinline__ float4 apply_raw(const pn2_rows &s, const RowCtx &c, int c4, const Raw &r,
                                            const float *coef, int coef_ld, int coef_base) {
  if (!c.valid || c4 >= s.cols) return zero4();
  if (KIND == PN2_ROWS_PLAIN) return r.x;
  if (KIND == PN2_ROWS_GATHER) return c4 < s.feat_cols ? r.x : make_float4(c.gx, c.gy, c.gz, 0.f);
  const int ci = c4 - coef_base;
  const float4 k0 = *reinterpret_cast<const float4 *>(coef + ci);
  const float4 k1 = *reinterpret_cast<const float4 *>(coef + coef_ld + ci);
  if (KIND == PN2_ROWS_BNRELU)
    return make_float4(relu_nan(fmaf(r.x.x, k0.x, k1.x)), relu_nan(fmaf(r.x.y, k0.y, k1.y)),
                       relu_nan(fmaf(r.x.z, k0.z, k1.z)), relu_nan(fmaf(r.x.w, k0.w, k1.w)));
  const float4 k2 = *reinterpret_cast<const float4 *>(coef + 2 * coef_ld + ci);
  float4 dz = r.d;
  if (KIND == PN2_ROWS_DYPOOL)
    dz = make_float4(r.a.x == c.slot ? dz.x : 0.f, r.a.y == c.slot ? dz.y : 0.f, r.a.z == c.slot ? dz.z : 0.f,
                     r.a.w == c.slot ? dz.w : 0.f);
  return make_float4(fmaf(k2.x, r.x.x, fmaf(k0.x, dz.x, k1.x)), fmaf(k2.y, r.x.y, fmaf(k0.y, dz.y, k1.y)),
                     fmaf(k2.z, r.x.z, fmaf(k0.z, dz.z, k1.z)), fmaf(k2.w, r.x.w, fmaf(k0.w, dz.w, k1.w)));
}

template <int KIND>
__device__ __forceinline__ void stage_coef(const pn2_rows &s, float *coef, int coef_ld, int base, int count, int tid,
                                           int nthreads = TC_THREADS) {
  if (KIND == PN2_ROWS_PLAIN || KIND == PN2_ROWS_GATHER) return;
  for (int i = tid; i < count; i += nthreads) {
    const int c = base + i;
    const bool ok = c < s.cols;
    coef[i] = ok ? __ldg(s.c0 + c) : 0.f;
    coef[coef_ld + i] = ok ? __ldg(s.c1 + c) : 0.f;
    if (KIND == PN2_ROWS_DY || KIND == PN2_ROWS_DYPOOL) coef[2 * coef_ld + i] = ok ? __ldg(s.c2 + c) : 0.f;
  }
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
// phase stamp of CTA `cta` (thread 0 only; no-op unless a trace buffer is installed)
__device__ __forceinline__ void trace_stamp(const GemmArgs &g, int cta, int slot, unsigned long long v) {
  if (g.trace != nullptr && threadIdx.x == 0 && cta < g.trace_cap) g.trace[static_cast<size_t>(cta) * 6 + slot] = v;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// mbarrier wait that traps instead of hanging the GPU if a transaction count was ever wrong
__device__ __forceinline__ void mbar_wait_guarded(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (!done && (spins & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// barrier over the 512 producer / epilogue threads only (the MMA warp does not take part)
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// ---- epilogue (shared by the kernels below) ------------------------------------------------------------------
// Drains the 128 x 128 fp32 accumulator at `tmem_d` of the tile at (m0, n0); `scratch` = the (now idle) operand
// stages, `red` = per-warp column partials.  m_tile indexes the per-row-tile statistics, split the wgrad slice.
// DGRAD_MASK reads the previous layer's pre-activations next to every output element: issue those loads before
// the accumulator is complete, so their latency hides behind the tail of the MMAs and the TMEM drain.
template <int EPI>
__device__ __forceinline__ void tc_epilogue_prefetch(const GemmArgs &g, int m0, int n0, float4 (&yv)[8]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = n0 + (warp >> 2) * 32 + (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (warp & 3) * 32 + (lane >> 3) + 4 * i;
    yv[i] = zero4();
    if (EPI == TC_EPI_DGRAD_MASK && row < g.M && col < g.N) yv[i] = ldg4(g.prev_y + static_cast<size_t>(row) * g.ld_prev + col);
  }
}

template <int EPI>
__device__ __forceinline__ void tc_epilogue(const GemmArgs &g, uint32_t tmem_d, unsigned char *tiles,
                                            float (&red)[2][TC_THREADS / 32][32], int m0, int n0, int m_tile, int split,
                                            bool have_acc, const float4 (&yv)[8]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 16 warps, one 32 x 32 chunk each: warp w reads TMEM lanes 32*(w%4)..+31 (tile rows), columns 32*(w/4)..+31.
  // The chunk goes through a per-warp shared-memory tile (row stride 36 floats, conflict-free 128-bit accesses
  // both ways) so that global accesses are row-contiguous: lane -> (row rs + 4*i, 4 columns 4*cq..).
  float *wt_tile = reinterpret_cast<float *>(tiles) + warp * (32 * 36);  // the operand stages are free now
  const int rbase = m0 + (warp & 3) * 32;
  const int cq = lane & 7, rs = lane >> 3;
  const int c_local = (warp >> 2) * 32;
  const int col = n0 + c_local + cq * 4;  // this lane's 4 columns
  {
    float v[32];
    if (have_acc) {
      tmem_ld32(tmem_d + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(c_local), v);
    } else {  // empty position slice of a split weight gradient: the accumulator was never written
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4 *>(wt_tile + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  __syncwarp();
  const bool col_ok = col < g.N;
  float4 sc = zero4(), sh = zero4();
  if (EPI == TC_EPI_DGRAD_MASK && col_ok) {
    sc = ldg4(g.prev_scale + col);
    sh = ldg4(g.prev_shift + col);
  }
  float4 s1 = zero4(), s2 = zero4();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = rs + 4 * i;
    const int row = rbase + rr;
    float4 v = *reinterpret_cast<const float4 *>(wt_tile + rr * 36 + cq * 4);
    if (row >= g.M || !col_ok) continue;
    if (EPI == TC_EPI_SCATTER) {  // transpose of the gather: scatter-add into the neighbour's feature row / xyz
      const int cloud = row / (g.G.npoint * g.G.nsample);
      const size_t src = static_cast<size_t>(cloud) * g.G.n_src + __ldg(g.G.idx + row);
      const int fc = g.G.feat_cols;
      if (col < fc) {
        if (g.dfeat) {
          float *dst = g.dfeat + src * g.ldf + col;
          atomicAdd(dst + 0, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
        }
      } else if (col == fc && g.dxyz && g.G.use_xyz) {
        const float gx = __fdiv_rn(v.x, g.G.inv_scale), gy = __fdiv_rn(v.y, g.G.inv_scale),
                    gz = __fdiv_rn(v.z, g.G.inv_scale);
        float *dn = g.dxyz + src * 3;
        atomicAdd(dn + 0, gx); atomicAdd(dn + 1, gy); atomicAdd(dn + 2, gz);
        const int centre = row / g.G.nsample;
        float *dc = g.dxyz + (static_cast<size_t>(cloud) * g.G.n_src + __ldg(g.centre_src + centre)) * 3;
        atomicAdd(dc + 0, -gx); atomicAdd(dc + 1, -gy); atomicAdd(dc + 2, -gz);
      }
      continue;
    }
    float4 qv;
    if (EPI == TC_EPI_DGRAD_MASK) {
      const float4 y = yv[i];
      v.x = fmaf(y.x, sc.x, sh.x) > 0.f ? v.x : 0.f;
      v.y = fmaf(y.y, sc.y, sh.y) > 0.f ? v.y : 0.f;
      v.z = fmaf(y.z, sc.z, sh.z) > 0.f ? v.z : 0.f;
      v.w = fmaf(y.w, sc.w, sh.w) > 0.f ? v.w : 0.f;
      qv = make_float4(v.x * y.x, v.y * y.y, v.z * y.z, v.w * y.w);
    } else {
      qv = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
    }
    *reinterpret_cast<float4 *>(g.out + split * g.out_split_stride + static_cast<size_t>(row) * g.ldo + col) = v;
    s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
    s2.x += qv.x; s2.y += qv.y; s2.z += qv.z; s2.w += qv.w;
  }
  if ((EPI == TC_EPI_STORE_STATS || EPI == TC_EPI_DGRAD_MASK) && g.stats != nullptr) {
    // column totals of this warp's 32 rows: combine the 4 row groups (lanes l, l^8, l^16, l^24)
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
      s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
      s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
      s2.z += __shfl_xor_sync(0xffffffffu, s2.z, o); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, o);
    }
    if (lane < 8) {
      *reinterpret_cast<float4 *>(&red[0][warp][cq * 4]) = s1;
      *reinterpret_cast<float4 *>(&red[1][warp][cq * 4]) = s2;
    }
    producers_sync();
    if (tid < 128) {  // tid -> (column group h = tid/32: warps 4h..4h+3 hold its four row blocks, column l)
      const int h = tid >> 5, l = tid & 31;
      const int c = n0 + h * 32 + l;
      if (c < g.stats_ld) {
        const float a = red[0][4 * h][l] + red[0][4 * h + 1][l] + red[0][4 * h + 2][l] + red[0][4 * h + 3][l];
        const float b = red[1][4 * h][l] + red[1][4 * h + 1][l] + red[1][4 * h + 2][l] + red[1][4 * h + 3][l];
        float *dst = g.stats + static_cast<size_t>(m_tile) * 2 * g.stats_ld + c;
        dst[0] = a;
        dst[g.stats_ld] = b;
      }
    }
  }

}

// TRANS = false: A rows are tile rows (positions), B rows are output channels, both K-contiguous in memory.
// TRANS = true (weight gradient): K runs over POSITIONS; A = dY source and B = activation source are both
// position rows with channels contiguous.  They are transposed while they are staged into the same K-major
// tiles: a lane holds 4 channels of one position and writes them with four 4-byte stores whose order is
// rotated per lane, so that the 32 lanes of a store (8 channel quads x 4 consecutive positions) hit 32
// different banks.  blockIdx.z selects a slice of positions and the partial tile goes to
// out + blockIdx.z * out_split_stride.
// Global access pattern (both forms): 8 consecutive lanes cover one 128-byte row segment, a warp-wide
// 128-bit access touches 4 lines instead of 32.
template <int AKIND, int BKIND, bool TRANS, int EPI, bool ROT>
__global__ void __launch_bounds__(TC_CTA_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
  pdl_prologue();
  extern __shared__ unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *empty_bar = reinterpret_cast<uint64_t *>(tiles + TC_STAGES * STAGE_BYTES);  // [TC_STAGES] MMAs of the stage done
  uint64_t *full_bar = empty_bar + TC_STAGES;                                           // [TC_STAGES] stage written (16 warps)
  uint64_t *done_bar = full_bar + TC_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
  float *coef_a = reinterpret_cast<float *>(tiles + TC_STAGES * STAGE_BYTES + 256);  // [3][coef_ld_a]
  float *coef_b = coef_a + 3 * TC_KMAX;                                              // [2][128]
  __shared__ float red[2][TC_THREADS / 32][32];  // per-warp column partials for the statistics epilogues

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int cta_id = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  trace_stamp(g, cta_id, 0, smid());
  trace_stamp(g, cta_id, 1, globaltimer_ns());

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&empty_bar[s], 1);
      mbar_init(&full_bar[s], TC_THREADS / 32);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  const bool producer = warp < TC_THREADS / 32;  // warp 16: lane 0 issues the MMAs, nothing else

  const uint32_t idesc = idesc_tf32(TM, TN, TRANS);
  const int k_begin = TRANS ? blockIdx.z * g.k_per_split : 0;
  const int k_end = TRANS ? min(g.K, k_begin + g.k_per_split) : g.K;
  const int num_kb = k_end > k_begin ? (k_end - k_begin + TK - 1) / TK : 0;

  const int coef_ld_a = TRANS ? 128 : TC_KMAX;
  const int coef_base_a = TRANS ? m0 : 0, coef_base_b = TRANS ? n0 : 0;

  // ---- producer mapping --------------------------------------------------------------------------------
  // plain form : chunk = tid % 8 (16 bytes of the 128-byte k-row), rows rsub + 64*i (i < 2) of the A / B tile
  // transposed : lane = 16-byte channel quad of the tile's 128 channels (a warp reads one position's 512
  //              contiguous bytes), positions warp + 16*i (i < 2) of the k-block, stored MN-major:
  //              offset(k, q) = (q/8)*4096 + (k/4)*512 + (k%4)*128 + (((q%8)/2 ^ k%4) * 32) + (q%2)*16
  constexpr int R = TC_ROWS_PER_THREAD;
  const int chunk = tid & 7, rsub = tid >> 3;
  uint32_t off[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    if (!TRANS) {
      off[i] = sw128_offset(rsub + 64 * i, chunk);
    } else {
      const int k = warp + 16 * i;
      off[i] = static_cast<uint32_t>((lane >> 3) * 4096 + (k >> 2) * 512 + (k & 3) * 128 + ((((lane & 7) >> 1) ^ (k & 3)) << 5) +
                                     (lane & 1) * 16);
    }
  }

  RowCtx ca[R], cb[R];
  auto make_ctx = [&](int kb, RowCtx *xa, RowCtx *xb) {  // transposed form: this thread's positions in k-block kb
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int p = k_begin + kb * TK + warp + 16 * i;
      const int r = (kb < num_kb && p < k_end) ? p : 0x7fffffff;
      xa[i] = row_ctx<AKIND>(g.A, r);
      xb[i] = row_ctx<BKIND>(g.B, r);
    }
  };
  RowCtx na[R], nb[R];  // transposed form: contexts of the NEXT k-block, resolved one block ahead of its loads so
                        // that a gather's index load and the row loads that depend on it sit in different iterations
  auto col_a = [&](int kb) { return TRANS ? m0 + lane * 4 : k_begin + kb * TK + chunk * 4; };
  auto col_b = [&](int kb) { return TRANS ? n0 + lane * 4 : k_begin + kb * TK + chunk * 4; };
  Raw da[2][R], db[2][R];  // two prefetch register sets, addressed with compile-time indices (loop unrolled by 2)
  if (producer) {
    if (TRANS) {
      make_ctx(0, ca, cb);
      make_ctx(1, na, nb);
    } else {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        ca[i] = row_ctx<AKIND>(g.A, m0 + rsub + 64 * i);
        cb[i] = row_ctx<BKIND>(g.B, n0 + rsub + 64 * i);
      }
    }
    if (num_kb > 0) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        da[0][i] = fetch_raw<AKIND>(g.A, ca[i], col_a(0));
        db[0][i] = fetch_raw<BKIND>(g.B, cb[i], col_b(0));
      }
    }
    // per-channel coefficient vectors of the sources -> shared memory (channels = k for the plain form, the
    // tile's 128 output rows / columns for the transposed form); staged while the first k-block's loads fly
    stage_coef<AKIND>(g.A, coef_a, coef_ld_a, coef_base_a, TRANS ? 128 : min(g.K, TC_KMAX), tid);
    if (TRANS) stage_coef<BKIND>(g.B, coef_b, 128, coef_base_b, 128, tid);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  trace_stamp(g, cta_id, 2, globaltimer_ns());

  // ---- MMA issue (lane 0 of warp 16, which does nothing else): wait until all 16 producer warps have written
  // the stage, issue its 12 MMAs and commit them to the stage's "empty" barrier.  No block barrier, and the
  // issuing thread is not a producer: a thread that stages operands and issues MMAs holds its warp back
  // (a __syncthreads per k-block serialised staging and MMA issue: 1.3-1.9 us per k-block against 0.4 us of MMA
  // time, profiles/c3_gemm_trace.txt).
  auto issue_mmas = [&](int kb) {
    const int s = kb % TC_STAGES;
    mbar_wait_guarded(&full_bar[s], (kb / TC_STAGES) & 1);
    tc_fence_after_sync();
    const uint32_t base = smem_addr(tiles + s * STAGE_BYTES);
    uint64_t a_hi, a_lo, b_hi, b_lo, step;
    if (TRANS) {
      a_hi = smem_desc_mn_sw128_32b(base, 4096, 512); a_lo = smem_desc_mn_sw128_32b(base + TILE_BYTES, 4096, 512);
      b_hi = smem_desc_mn_sw128_32b(base + 2 * TILE_BYTES, 4096, 512);
      b_lo = smem_desc_mn_sw128_32b(base + 3 * TILE_BYTES, 4096, 512);
      step = 1024 >> 4;  // 8 k-rows per MMA = two groups of 4 rows
    } else {
      a_hi = smem_desc_sw128(base); a_lo = smem_desc_sw128(base + TILE_BYTES);
      b_hi = smem_desc_sw128(base + 2 * TILE_BYTES); b_lo = smem_desc_sw128(base + 3 * TILE_BYTES);
      step = 32 >> 4;    // +32 bytes per k-step inside the 128-byte swizzle row
    }
#pragma unroll
    for (int ks = 0; ks < TK / 8; ++ks) {
      const uint64_t adv = step * ks;
      mma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, kb > 0 || ks > 0);
      mma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, true);
      mma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, true);
    }
    mma_commit(&empty_bar[s]);
    if (kb == num_kb - 1) mma_commit(done_bar);
  };
  if (!producer) {
    if (lane == 0)
      for (int kb = 0; kb < num_kb; ++kb) issue_mmas(kb);
  } else {
    // ---- producers: no block-wide barrier inside the loop; a stage is handed over with one mbarrier arrival
    // per warp and reclaimed when the MMAs that read it have completed
    // ROT = false: the two prefetch register sets are addressed with compile-time indices (loop unrolled by 2);
    // ROT = true: one loop body, the sets are rotated with register copies (fewer live registers, but the copy
    // waits for the load it has just issued).  PN2_TC_ROT selects; both are kept until one is measured better
    // on every shape.
    constexpr int NU = ROT ? 1 : 2;
    for (int kb0 = 0; kb0 < num_kb; kb0 += NU) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int kb = kb0 + u;
        if (kb >= num_kb) break;
        const int s = kb % TC_STAGES;
        // 1. put the next k-block's loads in flight (into the other register set: no copies, a move out of a
        //    register with a load in flight would stall until the data is back)
        if (kb + 1 < num_kb) {
#pragma unroll
          for (int i = 0; i < R; ++i) {
            da[u ^ 1][i] = fetch_raw<AKIND>(g.A, TRANS ? na[i] : ca[i], col_a(kb + 1));
            db[u ^ 1][i] = fetch_raw<BKIND>(g.B, TRANS ? nb[i] : cb[i], col_b(kb + 1));
          }
        }
        // 2. the stage is free once the MMAs that read it (block kb - STAGES) have completed
        if (kb >= TC_STAGES) mbar_wait_guarded(&empty_bar[s], ((kb / TC_STAGES) - 1) & 1);
        unsigned char *st = tiles + s * STAGE_BYTES;
        // 3. transform + split + store the current block (128-bit stores, conflict-free in both layouts)
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float4 va = apply_raw<AKIND>(g.A, ca[i], col_a(kb), da[u][i], coef_a, coef_ld_a, coef_base_a);
          const float4 vb = apply_raw<BKIND>(g.B, cb[i], col_b(kb), db[u][i], coef_b, 128, coef_base_b);
          float4 hi, lo;
          split_tf32(va.x, hi.x, lo.x); split_tf32(va.y, hi.y, lo.y);
          split_tf32(va.z, hi.z, lo.z); split_tf32(va.w, hi.w, lo.w);
          *reinterpret_cast<float4 *>(st + 0 * TILE_BYTES + off[i]) = hi;
          *reinterpret_cast<float4 *>(st + 1 * TILE_BYTES + off[i]) = lo;
          split_tf32(vb.x, hi.x, lo.x); split_tf32(vb.y, hi.y, lo.y);
          split_tf32(vb.z, hi.z, lo.z); split_tf32(vb.w, hi.w, lo.w);
          *reinterpret_cast<float4 *>(st + 2 * TILE_BYTES + off[i]) = hi;
          *reinterpret_cast<float4 *>(st + 3 * TILE_BYTES + off[i]) = lo;
        }
        fence_proxy_async_smem();  // this thread's generic-proxy stores -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        // 4. contexts are computed values (cheap to move); resolve the block after the next one
        if (TRANS) {
#pragma unroll
          for (int i = 0; i < R; ++i) { ca[i] = na[i]; cb[i] = nb[i]; }
          make_ctx(kb + 2, na, nb);
        }
        if (ROT) {
#pragma unroll
          for (int i = 0; i < R; ++i) { da[0][i] = da[1][i]; db[0][i] = db[1][i]; }
        }
      }
    }
    float4 yv[8];
    tc_epilogue_prefetch<EPI>(g, m0, n0, yv);
    if (num_kb > 0) mbar_wait_guarded(done_bar, 0);
    tc_fence_after_sync();
    trace_stamp(g, cta_id, 3, globaltimer_ns());
    tc_epilogue<EPI>(g, tmem_d, tiles, red, m0, n0, blockIdx.x, blockIdx.z, num_kb > 0, yv);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_d);
  trace_stamp(g, cta_id, 4, globaltimer_ns());
  trace_stamp(g, cta_id, 5, static_cast<unsigned long long>(num_kb));
}

// ---- forward / data-gradient kernel with a bulk-copied weight operand ---------------------------------------
// ncu of gemm_tc_kernel on the backbone's shapes (profiles/c1_ncu_gemm.json): the top stall is long_scoreboard
// (global-load latency) at 16 resident warps -- one k-block of A and B in flight per thread is ~16 KB per SM,
// far below what HBM needs.  For the non-transposed GEMMs B is the layer's weight matrix, identical for every
// row tile, so its hi/lo split and 128-byte-swizzle layout are computed ONCE by pn2_mlp_prep_weights into an
// image [n-tile][k-block][hi 16 KB | lo 16 KB]; here one thread fetches each 32 KB stage with a single
// cp.async.bulk (complete_tx on a "full" mbarrier), two k-blocks ahead, into a 4-stage ring.  That removes half
// of the staging instructions and the B registers, which pays for a second k-block of A prefetch per thread
// (register sets: current, +1, +2, +3).  A: 2-stage ring written by all 16 warps as before.
// Tiles are numbered with the n-tile fastest so that the CTAs sharing a row tile run together and the second
// one reads A from L2.
constexpr int BK_A_STAGES = 3, BK_B_STAGES = 3;
constexpr int BK_A_BYTES = 2 * TILE_BYTES, BK_B_BYTES = 2 * TILE_BYTES;  // hi + lo
constexpr int BK_RING = BK_A_STAGES * BK_A_BYTES + BK_B_STAGES * BK_B_BYTES;
constexpr int BK_SMEM = BK_RING + 1024 /*align*/ + 256 /*barriers*/ + TC_COEF_FLOATS * 4;

template <int AKIND, int EPI>
__global__ void __launch_bounds__(TC_CTA_THREADS, 1)
gemm_tc_bulk_kernel(const __grid_constant__ GemmArgs g) {
  pdl_prologue();
  extern __shared__ unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *ring_a = tiles, *ring_b = tiles + BK_A_STAGES * BK_A_BYTES;
  uint64_t *empty_bar = reinterpret_cast<uint64_t *>(tiles + BK_RING);  // [2]: MMAs of k-block kb done (A stage kb%2, B stage kb%4)
  uint64_t *full_a = empty_bar + BK_A_STAGES;                           // [2]: A stage written (one arrival per producer warp)
  uint64_t *full_b = full_a + BK_A_STAGES;                              // [4]: weight stage landed
  uint64_t *done_bar = full_b + BK_B_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
  float *coef_a = reinterpret_cast<float *>(tiles + BK_RING + 256);     // [3][TC_KMAX]
  __shared__ float red[2][TC_THREADS / 32][32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp < TC_THREADS / 32;  // warp 16: lane 0 issues the MMAs (see gemm_tc_kernel)
  const int ntn = (g.N + TN - 1) / TN;
  const int m_tile = blockIdx.x / ntn, n_tile = blockIdx.x - m_tile * ntn;
  const int m0 = m_tile * TM, n0 = n_tile * TN;
  const int num_kb = (g.K + TK - 1) / TK;
  const float *b_src = g.b_img + (static_cast<size_t>(n_tile) * g.b_img_kblocks) * (BK_B_BYTES / 4);
  trace_stamp(g, blockIdx.x, 0, smid());
  trace_stamp(g, blockIdx.x, 1, globaltimer_ns());

  if (tid == 0) {
    for (int s = 0; s < BK_A_STAGES; ++s) {
      mbar_init(&empty_bar[s], 1);
      mbar_init(&full_a[s], TC_THREADS / 32);
    }
    for (int s = 0; s < BK_B_STAGES; ++s) mbar_init(&full_b[s], 1);
    mbar_init(done_bar, 1);
    mbar_fence_init();
    fence_proxy_async_smem();  // the initialised barriers must be visible to the async proxy (bulk-copy complete_tx)
    for (int kb = 0; kb < BK_B_STAGES && kb < num_kb; ++kb) {
      mbar_expect_tx(&full_b[kb], BK_B_BYTES);
      bulk_g2s(ring_b + kb * BK_B_BYTES, b_src + static_cast<size_t>(kb) * (BK_B_BYTES / 4), BK_B_BYTES, &full_b[kb]);
    }
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);

  const uint32_t idesc = idesc_tf32(TM, TN, false);
  constexpr int R = TC_ROWS_PER_THREAD;
  const int chunk = tid & 7, rsub = tid >> 3;
  uint32_t off[R];
  RowCtx ca[R];
  // A prefetch: DEPTH k-blocks ahead of the one being staged (register sets rr[0] = current .. rr[DEPTH])
  // k-blocks of A in flight per thread beyond the current one (96 registers with 17 warps; DYPOOL chunks carry 9)
  constexpr int DEPTH = AKIND == PN2_ROWS_DYPOOL ? 1 : 2;
  Raw rr[DEPTH + 1][R];
  if (producer) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      off[i] = sw128_offset(rsub + 64 * i, chunk);
      ca[i] = row_ctx<AKIND>(g.A, m0 + rsub + 64 * i);
    }
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int i = 0; i < R; ++i) rr[d][i] = fetch_raw<AKIND>(g.A, ca[i], d < num_kb ? d * TK + chunk * 4 : 0x3fffffff);
    stage_coef<AKIND>(g.A, coef_a, TC_KMAX, 0, min(g.K, TC_KMAX), tid);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  trace_stamp(g, blockIdx.x, 2, globaltimer_ns());

  auto issue_mmas = [&](int kb) {  // lane 0 of the MMA warp
    const int sa = kb % BK_A_STAGES, sb = kb % BK_B_STAGES;
    mbar_wait_guarded(&full_a[sa], (kb / BK_A_STAGES) & 1);
    mbar_wait_guarded(&full_b[sb], (kb / BK_B_STAGES) & 1);
    tc_fence_after_sync();
    const uint32_t abase = smem_addr(ring_a + sa * BK_A_BYTES), bbase = smem_addr(ring_b + sb * BK_B_BYTES);
    const uint64_t a_hi = smem_desc_sw128(abase), a_lo = smem_desc_sw128(abase + TILE_BYTES);
    const uint64_t b_hi = smem_desc_sw128(bbase), b_lo = smem_desc_sw128(bbase + TILE_BYTES);
    if (g.debug != 1 || kb == 0) {
#pragma unroll
      for (int ks = 0; ks < TK / 8; ++ks) {
        const uint64_t adv = static_cast<uint64_t>(2 * ks);  // +32 bytes per k-step inside the 128-byte swizzle row
        mma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, kb > 0 || ks > 0);
        mma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, true);
        mma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, true);
      }
    }
    mma_commit(&empty_bar[sa]);
    if (kb == num_kb - 1) mma_commit(done_bar);
  };
  if (!producer) {
    if (lane == 0)
      for (int kb = 0; kb < num_kb; ++kb) issue_mmas(kb);
  } else {
    // The prefetch register sets are addressed with compile-time indices (loop unrolled by the number of
    // sets): rotating them with register copies would make every iteration wait for the loads it has just
    // issued -- a move out of a register with a load in flight stalls until the data is back.
    constexpr int NSET = DEPTH + 1;
    for (int kb0 = 0; kb0 < num_kb; kb0 += NSET) {
#pragma unroll
      for (int u = 0; u < NSET; ++u) {
        const int kb = kb0 + u;
        if (kb >= num_kb) break;
        const int sa = kb % BK_A_STAGES;
        // 1. A loads of k-block kb + DEPTH in flight
#pragma unroll
        for (int i = 0; i < R; ++i)
          rr[(u + DEPTH) % NSET][i] =
              fetch_raw<AKIND>(g.A, ca[i], kb + DEPTH < num_kb ? (kb + DEPTH) * TK + chunk * 4 : 0x3fffffff);
        // 2. MMAs of k-block kb - 3 done: A stage sa is free, and so is the weight stage that k-block read, which
        //    thread 0 refills with k-block kb (the first three were requested in the prologue)
        if (kb >= BK_A_STAGES) {
          mbar_wait_guarded(&empty_bar[sa], ((kb / BK_A_STAGES) - 1) & 1);
          if (tid == 0) {
            mbar_expect_tx(&full_b[sa], BK_B_BYTES);
            bulk_g2s(ring_b + sa * BK_B_BYTES, b_src + static_cast<size_t>(kb) * (BK_B_BYTES / 4), BK_B_BYTES, &full_b[sa]);
          }
        }
        // 3. transform + split + store A, hand the stage to the MMA thread (one arrival per warp)
        unsigned char *st = ring_a + sa * BK_A_BYTES;
        if (g.debug != 2) {
#pragma unroll
          for (int i = 0; i < R; ++i) {
            const float4 va = apply_raw<AKIND>(g.A, ca[i], kb * TK + chunk * 4, rr[u][i], coef_a, TC_KMAX, 0);
            float4 hi, lo;
            split_tf32(va.x, hi.x, lo.x); split_tf32(va.y, hi.y, lo.y);
            split_tf32(va.z, hi.z, lo.z); split_tf32(va.w, hi.w, lo.w);
            *reinterpret_cast<float4 *>(st + off[i]) = hi;
            *reinterpret_cast<float4 *>(st + TILE_BYTES + off[i]) = lo;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[sa]);
      }
    }
    float4 yv[8];
    tc_epilogue_prefetch<EPI>(g, m0, n0, yv);
    if (num_kb > 0) mbar_wait_guarded(done_bar, 0);
    tc_fence_after_sync();
    trace_stamp(g, blockIdx.x, 3, globaltimer_ns());
    tc_epilogue<EPI>(g, tmem_d, tiles, red, m0, n0, m_tile, 0, num_kb > 0, yv);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_d);
  trace_stamp(g, blockIdx.x, 4, globaltimer_ns());
  trace_stamp(g, blockIdx.x, 5, static_cast<unsigned long long>(num_kb));
}

// ---- forward / data-gradient kernel with the A operand in tensor memory -------------------------------------
// Measured on gemm_tc_bulk_kernel (profiles/c5_gemm_floor.txt): a k-block costs ~1.0 us against 0.40 us of MMA
// time, and the same loop with the MMAs removed still needs 0.7 us -- the producer -> MMA-thread -> producer
// hand-over of a shared-memory A stage (proxy fence, barrier round trip) is ~1.4 us long and a 2-stage ring
// hides only half of it; the 192 KB of shared memory are full, so the ring cannot get deeper there.  The A
// operand is produced in REGISTERS anyway (row source + hi/lo split), so this kernel writes it straight into
// tensor memory with tcgen05.st (thread = tile row = TMEM lane, 8 consecutive k columns per warp quarter) and
// the MMAs take A from TMEM (tcgen05.mma [d], [a], b-desc): no shared-memory traffic and no proxy fence for
// A, 6 A stages in the 384 TMEM columns next to the accumulator, and all of shared memory for a 6-stage ring
// of the bulk-copied weight image.  Shared-memory traffic per k-block drops from 160 KB to 80 KB.
// All 16 warps are producers + epilogue; thread 0 additionally loads the weight stages and issues the MMAs.
constexpr int TS_A_STAGES = 6, TS_B_STAGES = 6;
constexpr int TS_THREADS = TC_THREADS;
constexpr uint32_t TS_TMEM_COLS = 512, TS_A_COL0 = 128, TS_A_STAGE_COLS = 64;  // per stage: 32 hi + 32 lo columns
constexpr int TS_RING = TS_B_STAGES * BK_B_BYTES;
constexpr int TS_SMEM = TS_RING + 1024 /*align*/ + 256 /*barriers*/ + TC_COEF_FLOATS * 4;

template <int AKIND, int EPI>
__global__ void __launch_bounds__(TS_THREADS, 1)
gemm_tc_ts_kernel(const __grid_constant__ GemmArgs g) {
  pdl_prologue();
  extern __shared__ unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *ring_b = tiles;
  uint64_t *empty_a = reinterpret_cast<uint64_t *>(tiles + TS_RING);  // [6] MMAs that read A stage s have completed
  uint64_t *full_a = empty_a + TS_A_STAGES;                           // [6] A stage written (one arrival per producer warp)
  uint64_t *empty_b = full_a + TS_A_STAGES;                           // [6] MMAs that read weight stage s have completed
  uint64_t *full_b = empty_b + TS_B_STAGES;                           // [6] weight stage landed
  uint64_t *done_bar = full_b + TS_B_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
  float *coef_a = reinterpret_cast<float *>(tiles + TS_RING + 256);   // [3][TC_KMAX]
  __shared__ float red[2][TC_THREADS / 32][32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr bool producer = true;
  const int ntn = (g.N + TN - 1) / TN;
  const int m_tile = blockIdx.x / ntn, n_tile = blockIdx.x - m_tile * ntn;
  const int m0 = m_tile * TM, n0 = n_tile * TN;
  const int num_kb = (g.K + TK - 1) / TK;
  const float *b_src = g.b_img + (static_cast<size_t>(n_tile) * g.b_img_kblocks) * (BK_B_BYTES / 4);
  trace_stamp(g, blockIdx.x, 0, smid());
  trace_stamp(g, blockIdx.x, 1, globaltimer_ns());

  if (tid == 0) {
    for (int s = 0; s < TS_A_STAGES; ++s) {
      mbar_init(&empty_a[s], 1);
      mbar_init(&full_a[s], TC_THREADS / 32);
    }
    for (int s = 0; s < TS_B_STAGES; ++s) {
      mbar_init(&empty_b[s], 1);
      mbar_init(&full_b[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
    fence_proxy_async_smem();  // the initialised barriers must be visible to the async proxy (bulk-copy complete_tx)
  }
  if (warp == 0) tmem_alloc<TS_TMEM_COLS>(tmem_slot);

  const uint32_t idesc = idesc_tf32(TM, TN, false);
  // producer mapping: thread = tile row 32*(warp%4) + lane (its TMEM lane), columns 8*(warp/4)..+7 of the k-block
  constexpr int R = 2;
  const int quarter = warp & 3, cgrp = (warp >> 2) & 3;
  RowCtx ca;
  constexpr int DEPTH = (AKIND == PN2_ROWS_DYPOOL || AKIND == PN2_ROWS_DY) ? 2 : 3;  // register budget: 96 per thread
  Raw rr[DEPTH + 1][R];
  auto col_of = [&](int kb, int j) { return kb < num_kb ? kb * TK + cgrp * 8 + 4 * j : 0x3fffffff; };
  if (producer) {
    ca = row_ctx<AKIND>(g.A, m0 + quarter * 32 + lane);
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int j = 0; j < R; ++j) rr[d][j] = fetch_raw<AKIND>(g.A, ca, col_of(d, j));
    stage_coef<AKIND>(g.A, coef_a, TC_KMAX, 0, min(g.K, TC_KMAX), tid);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  trace_stamp(g, blockIdx.x, 2, globaltimer_ns());

  // thread 0 also loads the weight stages (one 32 KB bulk copy per k-block, TS_B_LEAD blocks ahead) and issues the MMAs
  constexpr int TS_B_LEAD = TS_B_STAGES - 2;  // the stage being refilled was read two k-blocks ago: its MMAs are done
  auto load_b = [&](int kb) {
    const int s = kb % TS_B_STAGES;
    if (kb >= TS_B_STAGES) mbar_wait_guarded(&empty_b[s], ((kb / TS_B_STAGES) - 1) & 1);
    mbar_expect_tx(&full_b[s], BK_B_BYTES);
    bulk_g2s(ring_b + s * BK_B_BYTES, b_src + static_cast<size_t>(kb) * (BK_B_BYTES / 4), BK_B_BYTES, &full_b[s]);
  };
  auto issue_mmas = [&](int kb) {
    const int sa = kb % TS_A_STAGES, sb = kb % TS_B_STAGES;
    mbar_wait_guarded(&full_a[sa], (kb / TS_A_STAGES) & 1);
    mbar_wait_guarded(&full_b[sb], (kb / TS_B_STAGES) & 1);
    tc_fence_after_sync();
    const uint32_t a_hi = tmem_d + TS_A_COL0 + sa * TS_A_STAGE_COLS, a_lo = a_hi + 32;
    const uint32_t bbase = smem_addr(ring_b + sb * BK_B_BYTES);
    const uint64_t b_hi = smem_desc_sw128(bbase), b_lo = smem_desc_sw128(bbase + TILE_BYTES);
#pragma unroll
    for (int ks = 0; ks < TK / 8; ++ks) {
      const uint64_t adv = static_cast<uint64_t>(2 * ks);  // +32 bytes per k-step inside the 128-byte swizzle row
      mma_tf32_ts(tmem_d, a_hi + 8 * ks, b_hi + adv, idesc, kb > 0 || ks > 0);
      mma_tf32_ts(tmem_d, a_hi + 8 * ks, b_lo + adv, idesc, true);
      mma_tf32_ts(tmem_d, a_lo + 8 * ks, b_hi + adv, idesc, true);
    }
    mma_commit(&empty_a[sa]);
    mma_commit(&empty_b[sb]);
    if (kb == num_kb - 1) mma_commit(done_bar);
  };
  if (tid == 0)
    for (int kb = 0; kb < TS_B_LEAD && kb < num_kb; ++kb) load_b(kb);
  {
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + TS_A_COL0 + cgrp * 8;
    constexpr int NSET = DEPTH + 1;  // compile-time register-set indices, see gemm_tc_bulk_kernel
    for (int kb0 = 0; kb0 < num_kb; kb0 += NSET) {
#pragma unroll
      for (int u = 0; u < NSET; ++u) {
        const int kb = kb0 + u;
        if (kb >= num_kb) break;
        const int sa = kb % TS_A_STAGES;
#pragma unroll
        for (int j = 0; j < R; ++j) rr[(u + DEPTH) % NSET][j] = fetch_raw<AKIND>(g.A, ca, col_of(kb + DEPTH, j));
        if (tid == 0 && kb + TS_B_LEAD < num_kb) load_b(kb + TS_B_LEAD);
        if (kb >= TS_A_STAGES) {
          mbar_wait_guarded(&empty_a[sa], ((kb / TS_A_STAGES) - 1) & 1);
          tc_fence_after_sync();
        }
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const float4 v = apply_raw<AKIND>(g.A, ca, kb * TK + cgrp * 8 + 4 * j, rr[u][j], coef_a, TC_KMAX, 0);
          split_tf32(v.x, hi[4 * j + 0], lo[4 * j + 0]); split_tf32(v.y, hi[4 * j + 1], lo[4 * j + 1]);
          split_tf32(v.z, hi[4 * j + 2], lo[4 * j + 2]); split_tf32(v.w, hi[4 * j + 3], lo[4 * j + 3]);
        }
        const uint32_t dst = lane_base + sa * TS_A_STAGE_COLS;
        tmem_st8(dst, hi);
        tmem_st8(dst + 32, lo);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[sa]);
        if (tid == 0) issue_mmas(kb);
      }
    }
    float4 yv[8];
    tc_epilogue_prefetch<EPI>(g, m0, n0, yv);
    if (num_kb > 0) mbar_wait_guarded(done_bar, 0);
    tc_fence_after_sync();
    trace_stamp(g, blockIdx.x, 3, globaltimer_ns());
    tc_epilogue<EPI>(g, tmem_d, tiles, red, m0, n0, m_tile, 0, num_kb > 0, yv);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TS_TMEM_COLS>(tmem_d);
  trace_stamp(g, blockIdx.x, 4, globaltimer_ns());
  trace_stamp(g, blockIdx.x, 5, static_cast<unsigned long long>(num_kb));
}

template <int AKIND, int EPI>
int launch_tc_ts(const GemmArgs &g, cudaStream_t stream) {
  auto kernel = gemm_tc_ts_kernel<AKIND, EPI>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM);
    configured_dev = dev;
  }
  const unsigned grid = static_cast<unsigned>((g.M + TM - 1) / TM) * static_cast<unsigned>((g.N + TN - 1) / TN);
  GemmArgs a = g;
  gemm_trace_target(&a.trace, &a.trace_cap);
  pn2::launch(kernel, dim3(grid), dim3(TS_THREADS), TS_SMEM, stream, a);
  return check_launch("gemm_tc_ts_kernel");
}

// ---- persistent form of gemm_tc_bulk_kernel -------------------------------------------------------------------
// The per-CTA trace (profiles/r1_c8_gemm_trace.txt) shows 1.5 us of prologue (barrier init, TMEM allocation,
// coefficient staging) and 0.9 us of CTA turnaround per 128x128 tile next to a 4-10 us main loop.  Here one CTA per
// SM walks over the tiles (n-tile fastest, so neighbouring CTAs share A rows through L2): the prologue is paid
// once, barriers / TMEM / coefficients live for the whole kernel, ring stages and barrier phases are indexed by a
// running k-block counter.  A: 2-stage ring, weights: 4-stage ring, requested two k-blocks ahead inside a tile.
// The epilogue scratch aliases the rings, hence one 512-thread barrier per tile after the epilogue.
// Tiles are handed out dynamically (thread 0 draws tickets from a global counter one tile ahead and publishes
// them through shared memory + an mbarrier for the MMA thread): the geometry stream's FPS cluster occupies 16 SMs
// for most of a forward pass, and with a static tile list the CTAs that start late became a tail.  The counter
// resets itself: the CTA that draws the last of the ntiles + gridDim.x tickets knows nobody will draw again.
__device__ int g_tile_counters[1024];
constexpr int PB_A_STAGES = 2, PB_B_STAGES = 4;
constexpr int PB_RING = PB_A_STAGES * BK_A_BYTES + PB_B_STAGES * BK_B_BYTES;
constexpr int PB_SMEM = PB_RING + 1024 /*align*/ + 256 /*barriers*/ + TC_COEF_FLOATS * 4;

template <int AKIND, int EPI>
__global__ void __launch_bounds__(TC_CTA_THREADS, 1)
gemm_tc_pbulk_kernel(const __grid_constant__ GemmArgs g) {
  pdl_prologue();
  extern __shared__ unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *ring_a = tiles, *ring_b = tiles + PB_A_STAGES * BK_A_BYTES;
  uint64_t *empty_bar = reinterpret_cast<uint64_t *>(tiles + PB_RING);  // [2] MMAs of the k-block done (A stage, B stage)
  uint64_t *full_a = empty_bar + PB_A_STAGES;                           // [2] A stage written (one arrival per producer warp)
  uint64_t *full_b = full_a + PB_A_STAGES;                              // [4] weight stage landed
  uint64_t *done_bar = full_b + PB_B_STAGES;                            // accumulator of the tile complete
  uint64_t *tile_bar = done_bar + 1;                                    // next tile ticket published
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tile_bar + 1);
  float *coef_a = reinterpret_cast<float *>(tiles + PB_RING + 256);     // [3][TC_KMAX]
  __shared__ float red[2][TC_THREADS / 32][32];
  __shared__ int ticket[2];  // ticket[ti & 1] = tile of this CTA's ti-th iteration (>= ntiles: stop)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp < TC_THREADS / 32;  // warp 16: lane 0 issues the MMAs
  const int ntn = (g.N + TN - 1) / TN;
  const int ntiles = ((g.M + TM - 1) / TM) * ntn;
  const int num_kb = (g.K + TK - 1) / TK;
  trace_stamp(g, blockIdx.x, 0, smid());
  trace_stamp(g, blockIdx.x, 1, globaltimer_ns());

  if (tid == 0) {
    for (int s = 0; s < PB_A_STAGES; ++s) {
      mbar_init(&empty_bar[s], 1);
      mbar_init(&full_a[s], TC_THREADS / 32);
    }
    for (int s = 0; s < PB_B_STAGES; ++s) mbar_init(&full_b[s], 1);
    mbar_init(done_bar, 1);
    mbar_init(tile_bar, 1);
    mbar_fence_init();
    fence_proxy_async_smem();  // the initialised barriers must be visible to the async proxy (bulk-copy complete_tx)
    ticket[0] = atomicAdd(g.tile_counter, 1);
  }
  const int last_ticket = ntiles + static_cast<int>(gridDim.x) - 1;
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (producer) stage_coef<AKIND>(g.A, coef_a, TC_KMAX, 0, min(g.K, TC_KMAX), tid);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = idesc_tf32(TM, TN, false);
  trace_stamp(g, blockIdx.x, 2, globaltimer_ns());

  if (!producer) {
    if (lane == 0) {  // ---- MMA thread: the k-blocks of all of this CTA's tiles form one stream `it`
      int it = 0;
      for (int ti = 0;; ++ti) {
        if (ti > 0) mbar_wait_guarded(tile_bar, (ti - 1) & 1);
        if (ticket[ti & 1] >= ntiles) break;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int sa = it & (PB_A_STAGES - 1), sb = it & (PB_B_STAGES - 1);
          mbar_wait_guarded(&full_a[sa], (it >> 1) & 1);  // the first block of a tile arrives only after every
          mbar_wait_guarded(&full_b[sb], (it >> 2) & 1);  // warp has drained the previous accumulator
          tc_fence_after_sync();
          const uint32_t abase = smem_addr(ring_a + sa * BK_A_BYTES), bbase = smem_addr(ring_b + sb * BK_B_BYTES);
          const uint64_t a_hi = smem_desc_sw128(abase), a_lo = smem_desc_sw128(abase + TILE_BYTES);
          const uint64_t b_hi = smem_desc_sw128(bbase), b_lo = smem_desc_sw128(bbase + TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < TK / 8; ++ks) {
            const uint64_t adv = static_cast<uint64_t>(2 * ks);  // +32 bytes per k-step inside the 128-byte swizzle row
            mma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, kb > 0 || ks > 0);
            mma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, true);
            mma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, true);
          }
          mma_commit(&empty_bar[sa]);
          if (kb == num_kb - 1) mma_commit(done_bar);
        }
      }
    }
  } else {
    constexpr int R = TC_ROWS_PER_THREAD;
    constexpr int DEPTH = AKIND == PN2_ROWS_DYPOOL ? 1 : 2;  // k-blocks of A in flight beyond the current one
    const int chunk = tid & 7, rsub = tid >> 3;
    uint32_t off[R];
#pragma unroll
    for (int i = 0; i < R; ++i) off[i] = sw128_offset(rsub + 64 * i, chunk);
    int it = 0;
    for (int ti = 0;; ++ti) {
      const int tile = ticket[ti & 1];
      if (tile >= ntiles) {
        if (tid == 0 && tile == last_ticket) *g.tile_counter = 0;  // every CTA has drawn its last ticket
        break;
      }
      if (tid == 0) {  // draw the next tile now; everybody reads it after this tile's closing barrier
        ticket[(ti + 1) & 1] = atomicAdd(g.tile_counter, 1);
        mbar_arrive(tile_bar);
      }
      const int m_tile = tile / ntn, n_tile = tile - m_tile * ntn;
      const int m0 = m_tile * TM, n0 = n_tile * TN;
      const float *b_src = g.b_img + (static_cast<size_t>(n_tile) * g.b_img_kblocks) * (BK_B_BYTES / 4);
      auto load_b = [&](int kb, int stream) {  // thread 0: weight stage of k-block kb (stream index `stream`)
        const int s = stream & (PB_B_STAGES - 1);
        mbar_expect_tx(&full_b[s], BK_B_BYTES);
        bulk_g2s(ring_b + s * BK_B_BYTES, b_src + static_cast<size_t>(kb) * (BK_B_BYTES / 4), BK_B_BYTES, &full_b[s]);
      };
      if (tid == 0)
        for (int kb = 0; kb < 2 && kb < num_kb; ++kb) load_b(kb, it + kb);  // every earlier MMA has completed (done_bar)
      RowCtx ca[R];
      Raw rr[DEPTH + 1][R];
#pragma unroll
      for (int i = 0; i < R; ++i) ca[i] = row_ctx<AKIND>(g.A, m0 + rsub + 64 * i);
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int i = 0; i < R; ++i) rr[d][i] = fetch_raw<AKIND>(g.A, ca[i], d < num_kb ? d * TK + chunk * 4 : 0x3fffffff);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int sa = it & (PB_A_STAGES - 1);
#pragma unroll
        for (int i = 0; i < R; ++i)
          rr[DEPTH][i] = fetch_raw<AKIND>(g.A, ca[i], kb + DEPTH < num_kb ? (kb + DEPTH) * TK + chunk * 4 : 0x3fffffff);
        // MMAs of stream block it - 2 done: A stage sa and weight stage (it + 2) % 4 are free
        if (it >= PB_A_STAGES) mbar_wait_guarded(&empty_bar[sa], ((it >> 1) - 1) & 1);
        if (tid == 0 && kb + 2 < num_kb) load_b(kb + 2, it + 2);
        unsigned char *st = ring_a + sa * BK_A_BYTES;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float4 va = apply_raw<AKIND>(g.A, ca[i], kb * TK + chunk * 4, rr[0][i], coef_a, TC_KMAX, 0);
          float4 hi, lo;
          split_tf32(va.x, hi.x, lo.x); split_tf32(va.y, hi.y, lo.y);
          split_tf32(va.z, hi.z, lo.z); split_tf32(va.w, hi.w, lo.w);
          *reinterpret_cast<float4 *>(st + off[i]) = hi;
          *reinterpret_cast<float4 *>(st + TILE_BYTES + off[i]) = lo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[sa]);
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
          for (int i = 0; i < R; ++i) rr[d][i] = rr[d + 1][i];
      }
      float4 yv[8];
      tc_epilogue_prefetch<EPI>(g, m0, n0, yv);
      if (num_kb > 0) mbar_wait_guarded(done_bar, ti & 1);
      tc_fence_after_sync();
      tc_epilogue<EPI>(g, tmem_d, tiles, red, m0, n0, m_tile, 0, num_kb > 0, yv);
      tc_fence_before_sync();   // this warp's TMEM reads are complete before the next tile's first MMA can be issued
      producers_sync();         // the epilogue scratch aliases both rings
    }
  }
  trace_stamp(g, blockIdx.x, 3, globaltimer_ns());

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_d);
  trace_stamp(g, blockIdx.x, 4, globaltimer_ns());
  trace_stamp(g, blockIdx.x, 5, static_cast<unsigned long long>(num_kb));
}

template <int AKIND, int EPI>
int launch_tc_pbulk(const GemmArgs &g, cudaStream_t stream) {
  auto kernel = gemm_tc_pbulk_kernel<AKIND, EPI>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM);
    configured_dev = dev;
  }
  const int ntiles = ((g.M + TM - 1) / TM) * ((g.N + TN - 1) / TN);
  // PN2_TC_PERSISTENT_SPARE SMs are left to the geometry stream (FPS / ball query of the next level run underneath
  // the MLP; a persistent grid on every SM would make them wait for a whole GEMM)
  static const int spare = [] {
    const char *e = getenv("PN2_TC_PERSISTENT_SPARE");
    return e ? atoi(e) : 8;  // measured: 0 -> 4.11, 8 -> 3.92, 16 -> 3.93 ms per step (profiles/r1_bench_*spare*)
  }();
  const int sms = sm_count() - spare > 1 ? sm_count() - spare : 1;
  const int grid = ntiles < sms ? ntiles : sms;
  GemmArgs a = g;
  gemm_trace_target(&a.trace, &a.trace_cap);
  // one counter slot per launch in flight (a captured graph keeps replaying the slot it was captured with)
  static thread_local int *counters = nullptr;
  static thread_local int counters_dev = -1;
  static std::atomic<unsigned> next_slot{0};  // process-wide: launches from different threads / streams never share a slot
  if (counters_dev != dev) {
    void *p = nullptr;
    if (cudaGetSymbolAddress(&p, g_tile_counters) != cudaSuccess) return check_launch("gemm_tc_pbulk_kernel(counters)");
    counters = static_cast<int *>(p);
    counters_dev = dev;
  }
  a.tile_counter = counters + (next_slot.fetch_add(1, std::memory_order_relaxed) & 1023u);
  pn2::launch(kernel, dim3(grid), dim3(TC_CTA_THREADS), PB_SMEM, stream, a);
  return check_launch("gemm_tc_pbulk_kernel");
}

template <int AKIND, int EPI>
int launch_tc_bulk(const GemmArgs &g, cudaStream_t stream) {
  auto kernel = gemm_tc_bulk_kernel<AKIND, EPI>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM);
    configured_dev = dev;
  }
  const unsigned grid = static_cast<unsigned>((g.M + TM - 1) / TM) * static_cast<unsigned>((g.N + TN - 1) / TN);
  GemmArgs a = g;
  gemm_trace_target(&a.trace, &a.trace_cap);
  static const int debug = [] {
    const char *e = getenv("PN2_TC_DEBUG");
    return e ? atoi(e) : 0;
  }();
  a.debug = debug;
  pn2::launch(kernel, dim3(grid), dim3(TC_CTA_THREADS), BK_SMEM, stream, a);
  return check_launch("gemm_tc_bulk_kernel");
}

template <int AKIND, int BKIND, bool TRANS, int EPI, bool ROT>
int launch_tc_rot(const GemmArgs &g, int splits, cudaStream_t stream) {
  auto kernel = gemm_tc_kernel<AKIND, BKIND, TRANS, EPI, ROT>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    configured_dev = dev;
  }
  dim3 grid((g.M + TM - 1) / TM, (g.N + TN - 1) / TN, splits);
  GemmArgs a = g;
  gemm_trace_target(&a.trace, &a.trace_cap);
  pn2::launch(kernel, dim3(grid), dim3(TC_CTA_THREADS), TC_SMEM, stream, a);
  return check_launch("gemm_tc_kernel");
}

template <int AKIND, int BKIND, bool TRANS, int EPI>
int launch_tc(const GemmArgs &g, int splits, cudaStream_t stream) {
  // measured (profiles/c8_*): the same per-k-block time on plain shapes, but the statically indexed sets spill
  // under the 96-register cap of a 17-warp CTA for the DYPOOL / GATHER sources (wgrad 1.31 vs 1.09 ms per step)
  static const bool rot = [] {
    const char *e = getenv("PN2_TC_ROT");
    return e == nullptr || e[0] != '0';
  }();
  return rot ? launch_tc_rot<AKIND, BKIND, TRANS, EPI, true>(g, splits, stream)
             : launch_tc_rot<AKIND, BKIND, TRANS, EPI, false>(g, splits, stream);
}

}  // namespace

// development aid: per-CTA phase timestamps of the next tensor-core GEMM launches (tools/gemm_trace.py)
static unsigned long long *g_trace_buf = nullptr;
static int g_trace_cap = 0;
void gemm_trace_target(unsigned long long **buf, int *cap) {
  *buf = g_trace_buf;
  *cap = g_trace_cap;
}

bool gemm_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("PN2_TC");
    on = (e == nullptr || e[0] != '0') ? 1 : 0;  // default on; PN2_TC=0 selects the FFMA kernel everywhere
  }
  return on == 1;
}

// epi uses mlp_gemm.cu's numbering: 0 store, 1 store+stats, 2 dgrad mask, 3 scatter.  Row tiles are always 128
// rows, which is what pn2_mlp_tiles() reports while this path is enabled.
int gemm_tc_launch(int akind, int epi, const void *gemm_args, cudaStream_t stream) {
  const GemmArgs &g = *static_cast<const GemmArgs *>(gemm_args);
  if (!gemm_tc_enabled() || g.B.kind != PN2_ROWS_PLAIN || g.K > TC_KMAX) return PN2_TC_UNSUPPORTED;
  // PN2_TC_BULK=0 keeps every operand thread-staged (gemm_tc_kernel) for A/B measurements
  static const bool bulk_on = [] {
    const char *e = getenv("PN2_TC_BULK");
    return e == nullptr || e[0] != '0';
  }();
  const bool use_bulk = bulk_on && g.b_img != nullptr;
  // PN2_TC_TS=1 selects the A-in-tensor-memory variant (correct, but measured slower: profiles/c7_gemm_trace_ts.txt)
  static const bool ts_on = [] {
    const char *e = getenv("PN2_TC_TS");
    return e != nullptr && e[0] == '1';
  }();
  // PN2_TC_PERSISTENT=0 launches one CTA per tile (gemm_tc_bulk_kernel) instead of the persistent form
  static const bool persistent = [] {
    const char *e = getenv("PN2_TC_PERSISTENT");
    return e == nullptr || e[0] != '0';
  }();
#define PN2_TC_CASE(AK, EP)                                                                     \
  if (akind == AK && epi == EP)                                                                 \
    return !use_bulk ? launch_tc<AK, PN2_ROWS_PLAIN, false, EP>(g, 1, stream)                   \
           : ts_on   ? launch_tc_ts<AK, EP>(g, stream)                                          \
           : persistent ? launch_tc_pbulk<AK, EP>(g, stream) : launch_tc_bulk<AK, EP>(g, stream);
  PN2_TC_CASE(PN2_ROWS_PLAIN, TC_EPI_STORE_STATS)
  PN2_TC_CASE(PN2_ROWS_BNRELU, TC_EPI_STORE_STATS)
  PN2_TC_CASE(PN2_ROWS_GATHER, TC_EPI_STORE_STATS)
  PN2_TC_CASE(PN2_ROWS_DY, TC_EPI_DGRAD_MASK)
  PN2_TC_CASE(PN2_ROWS_DYPOOL, TC_EPI_DGRAD_MASK)
  PN2_TC_CASE(PN2_ROWS_DY, TC_EPI_STORE)
  PN2_TC_CASE(PN2_ROWS_DYPOOL, TC_EPI_STORE)
  PN2_TC_CASE(PN2_ROWS_DY, TC_EPI_SCATTER)
  PN2_TC_CASE(PN2_ROWS_DYPOOL, TC_EPI_SCATTER)
#undef PN2_TC_CASE
  return PN2_TC_UNSUPPORTED;
}

// weight gradient: g.A = dY source, g.B = activation source, g.M = np, g.N = kp, g.K = positions,
// g.k_per_split a multiple of 32, partial tiles to g.out + z * g.out_split_stride
int gemm_tc_wgrad_launch(const void *gemm_args, int splits, cudaStream_t stream) {
  const GemmArgs &g = *static_cast<const GemmArgs *>(gemm_args);
  if (!gemm_tc_enabled()) return PN2_TC_UNSUPPORTED;
#define PN2_TC_W(AK, BK) \
  if (g.A.kind == AK && g.B.kind == BK) return launch_tc<AK, BK, true, TC_EPI_STORE>(g, splits, stream);
  PN2_TC_W(PN2_ROWS_DY, PN2_ROWS_PLAIN)
  PN2_TC_W(PN2_ROWS_DY, PN2_ROWS_BNRELU)
  PN2_TC_W(PN2_ROWS_DY, PN2_ROWS_GATHER)
  PN2_TC_W(PN2_ROWS_DYPOOL, PN2_ROWS_PLAIN)
  PN2_TC_W(PN2_ROWS_DYPOOL, PN2_ROWS_BNRELU)
  PN2_TC_W(PN2_ROWS_DYPOOL, PN2_ROWS_GATHER)
#undef PN2_TC_W
  return PN2_TC_UNSUPPORTED;
}

}  // namespace pn2

PN2_EXPORT int pn2_debug_gemm_trace(unsigned long long *device_buf, int ctas) {
  pn2::g_trace_buf = device_buf;
  pn2::g_trace_cap = device_buf ? ctas : 0;
  return PN2_OK;
}
