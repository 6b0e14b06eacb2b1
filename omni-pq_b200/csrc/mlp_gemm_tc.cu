// Shared-MLP contraction on the 5th-generation tensor cores (tcgen05, sm_100a): same operands, row sources and
// epilogues as the FFMA kernel in mlp_gemm.cu (C[M x N] = A[M x K] * B[N x K]^T).
//
// Precision: the parity bar of the float paths is 1e-5 relative, which a single TF32 pass (10-bit mantissa) cannot
// meet.  Every fp32 operand is split while it is staged into  v = hi + lo  (hi = nearest tf32 of v, lo = nearest tf32
// of v - hi) and each 8-wide k-step issues three kind::tf32 MMAs, hi*hi + hi*lo + lo*hi, accumulated in fp32 in tensor
// memory; the dropped lo*lo term is below 2^-22 relative per product.  Measured against float64 the result is
// 1e-6 ... 7e-6 relative over the backbone's layer stacks (fp32 FMA: 3e-7): the tensor core adds into its fp32
// accumulator with truncation, 3 x K/8 times per output (tests/arbiter.py has the numbers).
//
// Two kernels:
//   * gemm_tc_async_kernel  -- forward and data-gradient GEMMs (A = a row source with its element-wise transform,
//     B = the layer's weights): persistent, one CTA per SM, 16 producer warps + a converged MMA warp + a weight-loader
//     warp; activations travel  global --cp.async--> raw ring in shared memory --transform, split--> TENSOR MEMORY
//     (tcgen05.st), weights as a pre-split image by cp.async.bulk; MMAs in TS form (A from tensor memory).
//   * wgrad_tc_async_kernel -- weight-gradient GEMMs (K = positions): both operands are row sources; raw ring by
//     cp.async, transposed on the way out of it with thread = channel (dY -> tensor memory, activation -> a K-major
//     shared-memory tile), one CTA per (tile, position slice).
// Both are bounded by INSTRUCTION ISSUE in the 16 producer warps (a k-block is ~250-350 instructions per warp, 4 issue
// slots per SM and cycle), which is why their loops avoid cvt.rna.tf32 (expanded to ~7 instructions by ptxas),
// generic-address loads and guarded spin loops on the hot path.
// Both share the epilogue (tcgen05.ld -> per-warp shared-memory transpose -> row-contiguous 128-bit stores, BatchNorm
// partial column sums, ReLU-mask / scatter-add variants).
#include <atomic>

#include "mlp_rows.cuh"
#include "pn2_sm100.cuh"

namespace pn2 {
namespace {

using namespace sm100;

constexpr int TM = 128, TN = 128, TK = 32;          // CTA tile; TK fp32 = one 128-byte swizzle row
constexpr int TC_THREADS = 512;                    // 16 producer / epilogue warps
constexpr int TC_CTA_THREADS = TC_THREADS + 32;    // + one warp whose lane 0 only issues the MMAs
constexpr int TILE_BYTES = TM * TK * 4;             // 16 KB per operand half
constexpr int TC_KMAX = 1152;                        // largest K of the forward / data-gradient form

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
// per-tile timeline of the persistent async kernel (development aid, pn2_debug_gemm_trace2): 8 stamps for each of the
// first 8 tiles of every CTA
__device__ unsigned long long *g_trace2 = nullptr;
__device__ int g_trace2_ctas = 0;
__device__ __forceinline__ void tile_stamp(int ti, int k) {
  if (g_trace2 != nullptr && ti < 8 && static_cast<int>(blockIdx.x) < g_trace2_ctas)
    g_trace2[(static_cast<size_t>(blockIdx.x) * 8 + ti) * 8 + k] = globaltimer_ns();
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
// mbarrier wait that traps instead of hanging the GPU if a transaction count was ever wrong.  The hot path -- the phase
// is already complete -- is one try_wait and a branch: the producers' loops are instruction-issue bound, and the clock
// bookkeeping of a guarded spin loop cost ~14 instructions per wait even when it never spun.
__device__ __forceinline__ bool mbar_try_wait(uint32_t a, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(a), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_guarded(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  if (mbar_try_wait(a, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(a, parity))
    if (clock64() - t0 > 4000000000ll) __trap();
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

// Measured per 128 x 128 half (round 2, stamps inside this function, profiles/r2_epilogue_trace.txt): tensor-memory load
// 0.1 us, shared-memory transpose 0.2-0.4, the store loop 1.1-1.3, shuffles + barrier + statistics 0.4-0.8.  The store
// loop is the SM's write path -- 64 KB at ~32 B/clk -- not the chip's: starting the CTAs of a launch in four phases so that
// their epilogues do not coincide changed nothing (93.2 vs 89.4 us on 131072 x 128 -> 256).  Hiding it needs stores that
// drain under the NEXT tile's main loop.  Tried: the tile staged row-major in the (idle) weight ring and written out by
// the weight-loader warp with 512-byte cp.async.bulk shared -> global stores while the producers go on -- correct
// (73 / 73 GPU tests) but much slower, 145 vs 85 us on 131072 x 128 -> 256: the bulk-store path drains a 64 KB block
// more slowly than 16 warps of 128-bit stores, and the next tile's weights wait for it (they share the memory).
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

// ---- forward / data-gradient kernels with a bulk-copied weight operand ---------------------------------------
// For the non-transposed GEMMs B is the layer's weight matrix, identical for every row tile, so its hi/lo split and
// 128-byte-swizzle layout are computed ONCE by pn2_mlp_prep_weights into an image [n-tile][k-block][hi 16 KB | lo 16 KB];
// one thread fetches each 32 KB stage with a single cp.async.bulk (complete_tx on a "full" mbarrier).  Tiles are
// numbered with the n-tile fastest so that the CTAs sharing a row tile run together and the second one reads A from L2.
constexpr int BK_A_BYTES = 2 * TILE_BYTES, BK_B_BYTES = 2 * TILE_BYTES;  // hi + lo

// Dynamic tile scheduler of the persistent kernel: thread 0 of a CTA draws tickets from a global counter (the geometry
// stream's FPS cluster occupies 16 SMs for most of a forward pass; with a static tile list the CTAs that start late
// became a tail).  The counter resets itself: the CTA that makes the last draw of the grid knows nobody draws again.
__device__ int g_tile_counters[1024];

// PN2_TC_PERSISTENT_SPARE SMs are left to the geometry stream (FPS / ball query of the next level run underneath
// the MLP; a persistent grid on every SM would make them wait for a whole GEMM)
inline int persistent_spare() {
  static const int spare = [] {
    const char *e = getenv("PN2_TC_PERSISTENT_SPARE");
    // round 1 (geometry inside the step): 0 -> 4.11, 8 -> 3.92, 16 -> 3.93 ms per step.  Round 2, with the next batch's
    // level-1 FPS (one 16-CTA cluster) prefetched underneath the step: 8 -> 3.24, 16 -> 3.09, 24 -> 3.27 ms
    // (profiles/r2_bench_spare*.json): the cluster needs 16 SMs free at the same time
    return e ? atoi(e) : 16;
  }();
  return spare;
}

// one self-resetting ticket counter per launch in flight (a captured graph keeps replaying the slot it was captured with)
int *tile_counter_slot(int dev) {
  static thread_local int *counters = nullptr;
  static thread_local int counters_dev = -1;
  static std::atomic<unsigned> next_slot{0};  // process-wide: launches from different threads / streams never share a slot
  if (counters_dev != dev) {
    void *p = nullptr;
    if (cudaGetSymbolAddress(&p, g_tile_counters) != cudaSuccess) return nullptr;
    counters = static_cast<int *>(p);
    counters_dev = dev;
  }
  return counters + (next_slot.fetch_add(1, std::memory_order_relaxed) & 1023u);
}

// ---- persistent forward / dgrad kernel: async raw ring -> tensor-memory A operand -----------------------------------
// Round-2 measurements (tools/mma_floor.cu, profiles/r2_mma_floor.txt): twelve kind::tf32 128x128x8 MMAs -- one 32-wide
// k-block of the 3xTF32 scheme -- take 768 cycles (0.39 us) in SS *and* TS form, with or without concurrent
// shared-memory store traffic, so the 1.0 / 1.25 us per k-block of round 1's kernels were NOT operand bandwidth: it
// was the producers.  Their global loads were prefetched through registers, and the depth that survives register
// rotation / the 96-register cap is about one k-block, i.e. one memory latency (~1 us) per k-block.
//
// Here the loads do not pass through registers:
//   * every producer thread issues its 16-byte pieces of the next k-blocks with cp.async (LDGSTS) into a RAW ring in
//     shared memory (5 stages of 16 KB for plain / BatchNorm / gather sources, 3 stages of 32-36 KB for the
//     BatchNorm-backward sources that read y and dz), coalesced (8 lanes = one 128-byte row segment), completion
//     signalled per stage on an mbarrier (cp.async.mbarrier.arrive.noinc) -- 64-96 KB in flight per SM, independent of
//     the register allocator; the ring runs ACROSS tile boundaries (tickets are drawn five tiles ahead), so the first
//     k-blocks of the next tile are in flight during the epilogue of the current one;
//   * the transform (gather / BN+ReLU / BN backward + pool routing) and the hi/lo split read the raw stage with
//     thread = tile row (16-byte chunks XOR-swizzled by row: conflict-free) and write the operand straight into TENSOR
//     MEMORY (tcgen05.st, thread = TMEM lane); the MMAs take A from TMEM (TS form).  The A operand therefore needs no
//     shared memory at all, which is what pays for the raw ring; 6 (128-wide tile) or 4 (256-wide) A stages live in
//     the tensor-memory columns next to the accumulator;
//   * weights: the pre-split image in 32 KB slots (128 output rows x one k-block, hi | lo), 3-4 slot ring; a 256-wide
//     tile consumes two slots per k-block.
// 16 producer warps + one converged MMA warp (issues under elect.sync) + one weight-loader warp (lane 0 issues the bulk
// copies, so that waiting for a free weight slot never stalls a producer), one CTA per SM, five static tile tickets,
// then dynamic ones.
constexpr int AS_CTA_THREADS = TC_THREADS + 64;
constexpr uint32_t AS_TMEM_COLS = 512, AS_A_STAGE_COLS = 64;  // accumulator first, then the A stages: 32 hi + 32 lo columns each
constexpr int AS_ARG_BYTES = TM * TK;                                             // uint8 arg-max slots of a k-block

// TNW = tile width: 128, or 256 for weight matrices with a multiple of 256 rows -- A is then staged once per 256 output
// columns and a k-block's 12 MMAs run 1536 cycles, above the producers' ~1650-cycle chain instead of far below it.
template <int AKIND, int TNW>
struct AsCfg {
  static constexpr bool kDz = AKIND == PN2_ROWS_DY || AKIND == PN2_ROWS_DYPOOL;
  static constexpr bool kArg = AKIND == PN2_ROWS_DYPOOL;
  static constexpr bool kCoef = !(AKIND == PN2_ROWS_PLAIN || AKIND == PN2_ROWS_GATHER);
  static constexpr int kRawBytes = TILE_BYTES * (kDz ? 2 : 1) + (kArg ? AS_ARG_BYTES : 0);
  // weight ring: slots of one k-block x 128 output rows (hi 16 KB | lo 16 KB).  A 256-wide tile consumes two slots per
  // k-block, each with its own barriers and its own six N = 128 MMA pairs: with whole 64 KB stages the ring was two deep
  // and a k-block waited ~1.1 us for its weights where the MMAs need 0.78
  static constexpr int kBSlot = 2 * TN * TK * 4;
  static constexpr int kH = TNW / TN;                                           // slots per k-block
  // (data-gradient sources: a raw stage is 32-36 KB, so 3 + 3; a wide tile with 2 raw stages + 4 slots staged a k-block
  // in 1.8 us -- one copy in flight -- against 1.0 us here: 83.9 -> 65.7 us on 32768 x 512 -> 256)
  static constexpr int kNR = kDz ? 3 : 5;                                       // raw stages
  static constexpr int kNB = kDz ? 3 : 4;                                       // weight slots
  static constexpr int kNA = (static_cast<int>(AS_TMEM_COLS) - TNW) / static_cast<int>(AS_A_STAGE_COLS);  // A stages: 6 / 4
  static constexpr int kCoefK = kCoef ? 640 : 0;         // largest K with per-channel coefficient vectors (else FFMA kernel)
  static constexpr int kRing = kNB * kBSlot + kNR * kRawBytes;
  static constexpr int kSmem = kRing + 1024 /*align*/ + 512 /*barriers, tickets*/ + 3 * kCoefK * 4;
  static_assert(kSmem + 4096 /*static: statistics partials*/ <= 232448, "shared memory budget");
  static_assert(kNB * kBSlot >= (TC_THREADS / 32) * 32 * 36 * 4, "epilogue scratch aliases the weight ring");
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once every cp.async issued so far by this thread has landed (the count is part of the init value)
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ float4 lds4(const unsigned char *p) { return *reinterpret_cast<const float4 *>(p); }
// shared-window loads with an explicit 32-bit address: base register + immediate, no generic-address arithmetic per load
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return static_cast<int>(v);
}

// apply_raw with the per-channel coefficient vectors read through the shared window (coef = 32-bit address of c0;
// c1, c2 follow at coef_ld floats)
template <int KIND>
__device__ __forceinline__ float4 apply_raw_s(const pn2_rows &s, const RowCtx &c, int c4, const Raw &r, uint32_t coef, int coef_ld) {
  if (!c.valid || c4 >= s.cols) return zero4();
  if (KIND == PN2_ROWS_PLAIN) return r.x;
  if (KIND == PN2_ROWS_GATHER) return c4 < s.feat_cols ? r.x : make_float4(c.gx, c.gy, c.gz, 0.f);
  const float4 k0 = lds_v4(coef + c4 * 4), k1 = lds_v4(coef + (coef_ld + c4) * 4);
  if (KIND == PN2_ROWS_BNRELU)
    return make_float4(relu_nan(fmaf(r.x.x, k0.x, k1.x)), relu_nan(fmaf(r.x.y, k0.y, k1.y)),
                       relu_nan(fmaf(r.x.z, k0.z, k1.z)), relu_nan(fmaf(r.x.w, k0.w, k1.w)));
  const float4 k2 = lds_v4(coef + (2 * coef_ld + c4) * 4);
  float4 dz = r.d;
  if (KIND == PN2_ROWS_DYPOOL)
    dz = make_float4(r.a.x == c.slot ? dz.x : 0.f, r.a.y == c.slot ? dz.y : 0.f, r.a.z == c.slot ? dz.z : 0.f,
                     r.a.w == c.slot ? dz.w : 0.f);
  return make_float4(fmaf(k2.x, r.x.x, fmaf(k0.x, dz.x, k1.x)), fmaf(k2.y, r.x.y, fmaf(k0.y, dz.y, k1.y)),
                     fmaf(k2.z, r.x.z, fmaf(k0.z, dz.z, k1.z)), fmaf(k2.w, r.x.w, fmaf(k0.w, dz.w, k1.w)));
}

// what the ISSUING thread needs of a row: where it lives (element offsets, resolved once per tile)
struct IssueCtx {
  bool valid;
  const float *px, *pd;  // the row in x; DY / DYPOOL: the row in dz (pooled row for DYPOOL)
};
template <int AKIND>
__device__ __forceinline__ IssueCtx issue_ctx(const pn2_rows &s, int row) {
  IssueCtx c;
  c.valid = row < s.rows;
  c.px = s.x; c.pd = s.dz;
  if (!c.valid) return c;
  if (AKIND == PN2_ROWS_GATHER) {
    const int cloud = row / (s.npoint * s.nsample);
    c.px = s.x + (static_cast<size_t>(cloud) * s.n_src + __ldg(s.idx + row)) * s.ld;
  } else {
    c.px = s.x + static_cast<size_t>(row) * s.ld;
    if (AKIND == PN2_ROWS_DY) c.pd = s.dz + static_cast<size_t>(row) * s.ld;
    if (AKIND == PN2_ROWS_DYPOOL) c.pd = s.dz + static_cast<size_t>(row / s.group) * s.ld;
  }
  return c;
}
// what the TRANSFORMING thread needs of its row: validity, pool slot, local coordinates of a gathered neighbour
struct ConsCtx {
  bool valid;
  int slot;
  float gx, gy, gz;
};
template <int AKIND>
__device__ __forceinline__ ConsCtx cons_ctx(const pn2_rows &s, int row, bool want_xyz) {
  ConsCtx c;
  c.valid = row < s.rows;
  c.slot = 0; c.gx = c.gy = c.gz = 0.f;
  if (!c.valid) return c;
  if (AKIND == PN2_ROWS_DYPOOL) c.slot = row % s.group;
  if (AKIND == PN2_ROWS_GATHER && want_xyz && s.use_xyz) {
    const int cloud = row / (s.npoint * s.nsample), centre = row / s.nsample;
    const size_t src = static_cast<size_t>(cloud) * s.n_src + __ldg(s.idx + row);
    const float *p = s.xyz + src * 3, *q = s.centres + static_cast<size_t>(centre) * 3;
    c.gx = __fdiv_rn(__fsub_rn(__ldg(p + 0), __ldg(q + 0)), s.inv_scale);  // pointnet2_utils.py:350-352
    c.gy = __fdiv_rn(__fsub_rn(__ldg(p + 1), __ldg(q + 1)), s.inv_scale);
    c.gz = __fdiv_rn(__fsub_rn(__ldg(p + 2), __ldg(q + 2)), s.inv_scale);
  }
  return c;
}

template <int AKIND, int EPI, int TNW>
__global__ void __launch_bounds__(AS_CTA_THREADS, 1)  // 96 registers: more (__maxnreg__ 104 / 112) does not launch with 18 warps
gemm_tc_async_kernel(const __grid_constant__ GemmArgs g) {
  using Cfg = AsCfg<AKIND, TNW>;
  constexpr int NR = Cfg::kNR, NB = Cfg::kNB, NA = Cfg::kNA;
  constexpr int BSL = Cfg::kBSlot, H = Cfg::kH;
  constexpr uint32_t AS_A_COL0 = TNW;
  pdl_prologue();
  extern __shared__ unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *ring_b = tiles, *ring_r = tiles + NB * BSL;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + Cfg::kRing);
  uint64_t *full_b = bars, *empty_b = full_b + NB;            // weight stage landed / its MMAs done
  uint64_t *raw_full = empty_b + NB, *raw_empty = raw_full + NR;  // raw stage landed (512 async arrivals) / read (16 warps)
  uint64_t *full_a = raw_empty + NR, *empty_a = full_a + NA;  // A stage in tensor memory written (16 warps) / its MMAs done
  uint64_t *done_bar = empty_a + NA, *tile_bar = done_bar + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tile_bar + 1);
  int *ticket = reinterpret_cast<int *>(tmem_slot + 1);       // [8] ticket[ti & 7] = tile of this CTA's ti-th iteration
  float *coef_a = reinterpret_cast<float *>(tiles + Cfg::kRing + 512);  // [3][kCoefK]
  __shared__ float red[2][TC_THREADS / 32][32];
  static_assert((2 * NB + 2 * NR + 2 * NA + 2) * 8 + 4 + 32 <= 512, "barrier block");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp < TC_THREADS / 32;
  const int ntn = (g.N + TNW - 1) / TNW;
  const int ntiles = ((g.M + TM - 1) / TM) * ntn;
  const int num_kb = (g.K + TK - 1) / TK;
  // Tile tickets: the first five of every CTA are STATIC (blockIdx + i * grid -- drawn dynamically up front, the CTAs that
  // start first would take five tiles each and leave late starters with none), the later ones come from the global
  // counter (values 5 * grid + k).  Every CTA draws once per tile it processes, so exactly `ntiles` dynamic draws are
  // made per launch and the one that returns ntiles - 1 resets the counter for the next launch that uses this slot.
  const int grid_n = static_cast<int>(gridDim.x);
  trace_stamp(g, blockIdx.x, 0, smid());
  trace_stamp(g, blockIdx.x, 1, globaltimer_ns());

  if (tid == 0) {
    for (int s = 0; s < NB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < NR; ++s) { mbar_init(&raw_full[s], TC_THREADS); mbar_init(&raw_empty[s], TC_THREADS / 32); }
    for (int s = 0; s < NA; ++s) { mbar_init(&full_a[s], TC_THREADS / 32); mbar_init(&empty_a[s], 1); }
    mbar_init(done_bar, 1);
    mbar_init(tile_bar, 1);
    mbar_fence_init();
    fence_proxy_async_smem();  // the initialised barriers must be visible to the async proxy (bulk-copy complete_tx)
    for (int i = 0; i < 5; ++i) ticket[i] = static_cast<int>(blockIdx.x) + i * grid_n;
  }
  if (warp == 0) tmem_alloc<AS_TMEM_COLS>(tmem_slot);
  if (producer && Cfg::kCoef) stage_coef<AKIND>(g.A, coef_a, Cfg::kCoefK, 0, min(g.K, Cfg::kCoefK), tid);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = idesc_tf32(TM, TN);
  trace_stamp(g, blockIdx.x, 2, globaltimer_ns());

  if (!producer) {
    // tile_bar arrival #ti is made by thread 0 when the producers START tile ti: the epilogue of tile ti-1 (whose
    // scratch aliases the weight ring) is over and ticket[ti & 3] has long been written.  Both service threads below
    // wait for it once per tile; thread 0 cannot get two arrivals ahead of them (arrival #ti+1 needs the MMAs of tile ti).
    if (warp == TC_THREADS / 32) {  // ---- MMA warp, converged: the k-blocks of all of this CTA's tiles form one stream `it`
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);  // warp-uniform for the compiler too
      int it = 0;
      int qb = 0, pb = 0;  // weight-slot cursor: slot, phase parity of its current use
      for (int ti = 0;; ++ti) {
        mbar_wait_guarded(tile_bar, ti & 1);
        if (__shfl_sync(0xffffffffu, ticket[ti & 7], 0) >= ntiles) break;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int sa = it % NA;
          mbar_wait_guarded(&full_a[sa], (it / NA) & 1);
          if (kb == 0 && lane == 0) tile_stamp(ti, 5);
          const uint32_t a_hi = tmem_u + AS_A_COL0 + sa * AS_A_STAGE_COLS, a_lo = a_hi + 32;
#pragma unroll
          for (int h = 0; h < H; ++h) {
            mbar_wait_guarded(&full_b[qb], pb);
            tc_fence_after_sync();
            if (kb == 0 && h == 0 && lane == 0) tile_stamp(ti, 6);
            const uint32_t bbase = smem_addr(ring_b + qb * BSL);
            const uint64_t b_hi = smem_desc_sw128(bbase), b_lo = smem_desc_sw128(bbase + BSL / 2);
            const uint32_t acc = tmem_u + h * TN;
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < TK / 8; ++ks) {
                const uint64_t adv = static_cast<uint64_t>(2 * ks);  // +32 bytes per k-step inside the 128-byte swizzle row
                mma_tf32_ts(acc, a_hi + 8 * ks, b_hi + adv, idesc, kb > 0 || ks > 0);
                mma_tf32_ts(acc, a_hi + 8 * ks, b_lo + adv, idesc, true);
                mma_tf32_ts(acc, a_lo + 8 * ks, b_hi + adv, idesc, true);
              }
              mma_commit(&empty_b[qb]);
              if (h == H - 1) {
                mma_commit(&empty_a[sa]);
                if (kb == num_kb - 1) mma_commit(done_bar);
              }
            }
            __syncwarp();
            if (++qb == NB) { qb = 0; pb ^= 1; }
          }
        }
      }
    } else if (warp == TC_THREADS / 32 + 1 && lane == 0) {  // ---- weight loader: two 16 KB bulk copies (hi, lo) per slot
      // image tile = g.b_tile_rows (128 or 256) weight rows x 32 k: [hi rows x 128 B | lo rows x 128 B]; the 128-row
      // half `hh` of a 256-row tile is a contiguous 16 KB piece of each
      const int R = g.b_tile_rows, per = R / TN;
      const size_t stage_f = static_cast<size_t>(2) * R * TK;  // floats per image k-block
      int qb = 0, pb = 0;
      for (int ti = 0;; ++ti) {
        mbar_wait_guarded(tile_bar, ti & 1);
        const int tile = ticket[ti & 7];
        if (tile >= ntiles) break;
        const int n_tile = tile % ntn;
        for (int kb = 0; kb < num_kb; ++kb) {
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const int n128 = n_tile * H + h, T = n128 / per, hh = n128 - T * per;
            const float *src = g.b_img + (static_cast<size_t>(T) * g.b_img_kblocks + kb) * stage_f + static_cast<size_t>(hh) * TN * TK;
            mbar_wait_guarded(&empty_b[qb], pb ^ 1);  // the MMAs that read this slot are done (first use: passes)
            if (kb == 0 && h == 0) tile_stamp(ti, 7);
            mbar_expect_tx(&full_b[qb], BSL);
            bulk_g2s(ring_b + qb * BSL, src, BSL / 2, &full_b[qb]);
            bulk_g2s(ring_b + qb * BSL + BSL / 2, src + static_cast<size_t>(R) * TK, BSL / 2, &full_b[qb]);
            if (++qb == NB) { qb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else {
    // ---- producers.  What a k-block costs here is neither bandwidth nor the tensor pipe but INSTRUCTION ISSUE: 16 warps x
    // ~300 instructions per k-block (ring top-up, three barrier waits, shared-memory reads, transform, split,
    // tcgen05.st + wait, fence + arrive) over 4 issue slots per cycle is the measured 0.6-0.8 us, although a thread only
    // moves 8 values (loop bodies counted in the SASS).  Hence: ring positions and phase bits
    // carried incrementally, row pointers resolved once per tile, shared-window loads with immediates, one try_wait on
    // the hot path of every barrier, tf32 rounding on the bit pattern.  Rejected variants: TWO k-blocks per iteration (half
    // the barrier round trips per k-block; 63.9 vs 49.1 us on 32768 x 256 -> 256: the MMA warp then receives its stages
    // in bursts) and four warp groups owning every fourth k-block (2.6 us per group and k-block).
    // issue mapping: 16-byte chunk `chunk` of rows rsub, rsub + 64 (8 consecutive lanes = one 128-byte row segment)
    const int chunk = tid & 7, rsub = tid >> 3;
    // transform mapping: tile row 32*(warp%4) + lane (= this thread's TMEM lane), k columns 8*(warp/4) .. +7
    const int quarter = warp & 3, cgrp = warp >> 2, crow = quarter * 32 + lane;
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + AS_A_COL0 + cgrp * 8;
    const bool xyz_mine = AKIND == PN2_ROWS_GATHER && ((g.A.feat_cols % TK) / 8) == cgrp;
    const void *dummy = g.b_img;  // a valid address for zero-filling copies (src size 0 reads nothing)
    // per-thread constants of the two mappings
    const uint32_t ring_r_s = smem_addr(ring_r), coef_s = smem_addr(coef_a);
    uint32_t ioff[2];   // issue: byte offset of (row rsub + 64 i, chunk) inside a raw x / dz tile
#pragma unroll
    for (int i = 0; i < 2; ++i) ioff[i] = (rsub + 64 * i) * 128 + ((chunk ^ ((rsub + 64 * i) & 7)) << 4);
    uint32_t coff[2];   // transform: byte offset of (row crow, chunk 2 cgrp + j)
#pragma unroll
    for (int j = 0; j < 2; ++j) coff[j] = crow * 128 + (((2 * cgrp + j) ^ (crow & 7)) << 4);

    // ring cursors, carried incrementally: slot index and phase parity of the slot's CURRENT use.  A first-time wait for
    // "the previous use of this slot is over" asks for the phase before the barrier's first one, which reads as complete.
    int c_sr = 0, c_pr = 0;               // consume side of the raw ring
    int c_sa = 0, c_pa = 0;               // A stages in tensor memory
    int i_sr = 0, i_pr = 0;               // issue side of the raw ring
    int il = 0, it = 0;                   // stream indices: next k-block to issue / to transform
    int lt = 0, lkb = 0;                  // the load cursor's tile sequence number and k-block inside that tile
    IssueCtx lc[2], nlc[2];               // issue contexts of the current tile / of the tile after it
    ConsCtx cc, ncc;
    bool have_next = false;
    {
      const int t0 = ticket[0];
      if (t0 < ntiles) {
        const int m0 = (t0 / ntn) * TM;
        lc[0] = issue_ctx<AKIND>(g.A, m0 + rsub);
        lc[1] = issue_ctx<AKIND>(g.A, m0 + rsub + 64);
        cc = cons_ctx<AKIND>(g.A, m0 + crow, xyz_mine);
      }
    }
    const int a_cols = AKIND == PN2_ROWS_GATHER ? g.A.feat_cols : g.A.cols;
    auto issue = [&](const IssueCtx (&c)[2], int kb) {  // this thread's copies of k-block kb into raw stage i_sr
      mbar_wait_guarded(&raw_empty[i_sr], i_pr ^ 1);
      const uint32_t raw = ring_r_s + i_sr * Cfg::kRawBytes;
      const int c4 = kb * TK + chunk * 4;
      const bool col_ok = c4 < a_cols;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool ok = c[i].valid && col_ok;
        const int nb = ok ? 16 : 0;
        cp_async16(raw + ioff[i], ok ? static_cast<const void *>(c[i].px + c4) : dummy, nb);
        if (Cfg::kDz) {
          cp_async16(raw + TILE_BYTES + ioff[i], ok ? static_cast<const void *>(c[i].pd + c4) : dummy, nb);
          if (Cfg::kArg)
            cp_async4(raw + 2 * TILE_BYTES + (chunk * TM + rsub + 64 * i) * 4,
                      ok ? static_cast<const void *>(g.A.arg + (c[i].pd - g.A.dz) + c4) : dummy, ok ? 4 : 0);
        }
      }
      cp_async_arrive(&raw_full[i_sr]);
      if (++i_sr == NR) { i_sr = 0; i_pr ^= 1; }
      ++il;
      if (++lkb == num_kb) { lkb = 0; ++lt; }
    };

    for (int ti = 0;; ++ti) {
      const int tile = ticket[ti & 7];
      if (tile >= ntiles) {
        if (tid == 0) mbar_arrive(tile_bar);  // releases the MMA / weight-loader warps, which then see the end ticket too
        break;
      }
      const int next_tile = ticket[(ti + 1) & 7];
      if (tid == 0) {
        mbar_arrive(tile_bar);  // arrival #ti: tile ti has started
        tile_stamp(ti, 0);
      }
      const int m_tile = tile / ntn, n_tile = tile - m_tile * ntn;
      const int m0 = m_tile * TM, n0 = n_tile * TNW;
      // contexts of the NEXT tile, resolved one tile ahead so that their (dependent) loads are long back when needed
      have_next = next_tile < ntiles;
      if (have_next) {
        const int nm0 = (next_tile / ntn) * TM;
        nlc[0] = issue_ctx<AKIND>(g.A, nm0 + rsub);
        nlc[1] = issue_ctx<AKIND>(g.A, nm0 + rsub + 64);
        ncc = cons_ctx<AKIND>(g.A, nm0 + crow, xyz_mine);
      }
      RowCtx rc;
      rc.valid = cc.valid; rc.off = 0; rc.goff = 0; rc.slot = cc.slot; rc.gx = cc.gx; rc.gy = cc.gy; rc.gz = cc.gz;
      for (int kb = 0; kb < num_kb; ++kb) {
        // 1. keep the raw ring full: up to NR - 1 k-blocks beyond this one, into the next tile if need be
        while (il - it < NR - 1) {
          if (lt == ti) issue(lc, lkb);
          else if (lt == ti + 1 && have_next) issue(nlc, lkb);
          else break;
        }
        // 2. the raw stage has landed (all 512 threads' copies): read this thread's pieces, release the stage
        mbar_wait_guarded(&raw_full[c_sr], c_pr);
        const uint32_t raw = ring_r_s + c_sr * Cfg::kRawBytes;
        Raw rw[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          rw[j].x = lds_v4(raw + coff[j]);
          rw[j].d = zero4(); rw[j].a = make_uchar4(0, 0, 0, 0);
          if (Cfg::kDz) rw[j].d = lds_v4(raw + TILE_BYTES + coff[j]);
          if (Cfg::kArg) {
            const uint32_t a = lds_u32(raw + 2 * TILE_BYTES + ((2 * cgrp + j) * TM + crow) * 4);
            rw[j].a = make_uchar4(a & 0xff, (a >> 8) & 0xff, (a >> 16) & 0xff, a >> 24);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[c_sr]);  // the stage may be refilled once all 16 warps have read it
        if (++c_sr == NR) { c_sr = 0; c_pr ^= 1; }
        // 3. transform + split
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float4 v = apply_raw_s<AKIND>(g.A, rc, kb * TK + cgrp * 8 + 4 * j, rw[j], coef_s, Cfg::kCoefK);
          split_tf32(v.x, hi[4 * j + 0], lo[4 * j + 0]); split_tf32(v.y, hi[4 * j + 1], lo[4 * j + 1]);
          split_tf32(v.z, hi[4 * j + 2], lo[4 * j + 2]); split_tf32(v.w, hi[4 * j + 3], lo[4 * j + 3]);
        }
        // 4. straight into tensor memory: the A stage is free once the MMAs that read it have completed
        mbar_wait_guarded(&empty_a[c_sa], c_pa ^ 1);
        tc_fence_after_sync();
        tmem_st8(lane_base + c_sa * AS_A_STAGE_COLS, hi);
        tmem_st8(lane_base + c_sa * AS_A_STAGE_COLS + 32, lo);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[c_sa]);
        if (++c_sa == NA) { c_sa = 0; c_pa ^= 1; }
        ++it;
        if (tid == 0 && kb == 0) tile_stamp(ti, 1);
      }
      if (tid == 0) {
        tile_stamp(ti, 2);
        // draw the ticket five tiles ahead -- here, where this thread would otherwise only wait for the accumulator (the
        // atomic's ~1 us round trip at the START of a tile delayed warp 0 and with it everybody, also when only its
        // issue sits there and the result is first read here: 60.5 vs 55.6 us on 131072 x 128 -> 128); everybody reads
        // it after the closing barrier below
        const int k = atomicAdd(g.tile_counter, 1);
        if (k == ntiles - 1) *g.tile_counter = 0;  // the last draw of the launch: nobody draws again
        ticket[(ti + 5) & 7] = 5 * grid_n + k;
      }
      float4 yv[8];
      tc_epilogue_prefetch<EPI>(g, m0, n0, yv);
      mbar_wait_guarded(done_bar, ti & 1);
      tc_fence_after_sync();
      if (tid == 0) tile_stamp(ti, 3);
#pragma unroll
      for (int h = 0; h < TNW / TN; ++h) {  // a 256-wide tile is drained as two 128-column halves
        if (h > 0) tc_epilogue_prefetch<EPI>(g, m0, n0 + h * TN, yv);
        tc_epilogue<EPI>(g, tmem_d + h * TN, tiles, red, m0, n0 + h * TN, m_tile, 0, true, yv);  // scratch = the (idle) weight ring
        if (h + 1 < TNW / TN) producers_sync();  // the statistics partials of this half have been read
      }
      tc_fence_before_sync();   // this warp's TMEM reads are complete before the next tile's first MMA can be issued
      producers_sync();
      if (tid == 0) tile_stamp(ti, 4);
      cc = ncc;
      lc[0] = nlc[0];
      lc[1] = nlc[1];
    }
  }
  trace_stamp(g, blockIdx.x, 3, globaltimer_ns());

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<AS_TMEM_COLS>(tmem_d);
  trace_stamp(g, blockIdx.x, 4, globaltimer_ns());
  trace_stamp(g, blockIdx.x, 5, static_cast<unsigned long long>(num_kb));
}

template <int AKIND, int EPI, int TNW>
int launch_tc_async_w(const GemmArgs &g, cudaStream_t stream) {
  using Cfg = AsCfg<AKIND, TNW>;
  auto kernel = gemm_tc_async_kernel<AKIND, EPI, TNW>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    configured_dev = dev;
  }
  const int ntiles = ((g.M + TM - 1) / TM) * ((g.N + TNW - 1) / TNW);
  const int sms = sm_count() - persistent_spare() > 1 ? sm_count() - persistent_spare() : 1;
  const int grid = ntiles < sms ? ntiles : sms;
  GemmArgs a = g;
  gemm_trace_target(&a.trace, &a.trace_cap);
  a.tile_counter = tile_counter_slot(dev);
  if (a.tile_counter == nullptr) return check_launch("gemm_tc_async_kernel(counters)");
  pn2::launch(kernel, dim3(grid), dim3(AS_CTA_THREADS), Cfg::kSmem, stream, a);
  return check_launch("gemm_tc_async_kernel");
}

// 128 x 256 tiles when the weight image was built with 256-row tiles and the cost model favours them: a wide tile stages
// A once for 256 columns (0.8 us per k-block, MMA bound, against 2 x 0.6 us) but drains two accumulator halves, so with
// few tiles (the deep levels: 4 ... 64 row tiles) twice as many narrow tiles on twice as many SMs finish earlier.
// Per-tile constants from profiles/r2_tile_trace.txt.
template <int AKIND, int EPI>
int launch_tc_async(const GemmArgs &g, cudaStream_t stream) {
  if (g.b_tile_rows != 128 && g.b_tile_rows != 256) return PN2_TC_UNSUPPORTED;
  if (g.b_tile_rows == 256 && (g.N % 256) == 0) {
    const int sms = sm_count() - persistent_spare() > 1 ? sm_count() - persistent_spare() : 1;
    const int wide = ((g.M + TM - 1) / TM) * (g.N / 256), kb = (g.K + TK - 1) / TK;
    const float cost_w = static_cast<float>((wide + sms - 1) / sms) * (0.8f * kb + 5.0f);
    const float cost_n = static_cast<float>((2 * wide + sms - 1) / sms) * (0.6f * kb + 3.3f);
    if (cost_w <= cost_n) return launch_tc_async_w<AKIND, EPI, 256>(g, stream);
  }
  return launch_tc_async_w<AKIND, EPI, 128>(g, stream);
}

// ---- weight-gradient kernel: async raw ring, thread = channel ---------------------------------------------------------
// dW[cout tile 128][cin tile 128] over a slice of positions (blockIdx.z).  K runs over POSITIONS, and both operands are row
// sources with the channels contiguous, so the operand tiles are the TRANSPOSE of what arrives from memory.
// Round 1's kernel did that transposition with MN-major tiles staged through registers, one k-block of prefetch deep:
// 2.8 us per 32-position k-block where the 48 KB it reads cost ~1 us of the SM's share of HBM.  Here:
//   * raw ring (3 stages): per k-block the 32 positions x 128 channels of  y | dz | activation  (+ the pooled arg-max
//     bytes, + the neighbour / centre coordinates of a gathered source) are copied by cp.async exactly as they lie in
//     memory -- a warp copies one position's 512 contiguous bytes -- with completion on an mbarrier; two k-blocks
//     (96 KB) are in flight per SM beyond the one being transformed;
//   * transform with thread = CHANNEL (32 * (warp % 4) + lane) for 8 positions (8 * (warp / 4) ..): the 4-byte reads of
//     a warp are 32 consecutive words of one position row (conflict-free without a swizzle), the per-channel BatchNorm
//     coefficients live in registers, and the thread's 8 values are 8 consecutive K elements of ITS row of the operand:
//     dY goes straight into tensor memory (tcgen05.st, lane = channel), the activation into a K-major 128-byte-swizzle
//     tile with four 16-byte stores -- the same operand forms as the forward kernel, no MN-major descriptors;
//   * the MMA warp (converged, elect.sync) issues the 12 TS-form MMAs of a k-block; 2 operand stages.
struct IncDiv {  // quotient of a position that advances by a fixed step per k-block, without dividing
  int q, rem, d, dq, dr;
  __device__ __forceinline__ void init(int r, int d_, int step) {
    d = d_ > 0 ? d_ : 1;
    q = r / d; rem = r - q * d;
    dq = step / d; dr = step - dq * d;
  }
  __device__ __forceinline__ void advance() {
    q += dq; rem += dr;
    if (rem >= d) { rem -= d; ++q; }
  }
  __device__ __forceinline__ int q_at(int delta) const {  // quotient `delta` (small) positions further on
    int qq = q, rr = rem + delta;
    while (rr >= d) { rr -= d; ++qq; }
    return qq;
  }
};

template <int AKIND, int BKIND>
struct WgCfg {
  static constexpr bool kArg = AKIND == PN2_ROWS_DYPOOL;
  static constexpr bool kXyz = BKIND == PN2_ROWS_GATHER;
  static constexpr int kOffB = 2 * TILE_BYTES;                     // raw stage: y | dz | activation | arg | xyz
  static constexpr int kOffArg = 3 * TILE_BYTES;
  static constexpr int kOffXyz = kOffArg + (kArg ? TK * TM : 0);
  static constexpr int kRawBytes = kOffXyz + (kXyz ? 1024 : 0);    // [6][32] floats: neighbour xyz, centre xyz
  static constexpr int kNR = 3, kNS = 2;
  // A gathered source whose feature width is a multiple of 128 puts its 4-column xyz block into an n-tile of its own, which
  // cost a full share of the CTAs for 3 useful channels (kp = 260: 6 tiles instead of 4).  Folded: the LAST feature tile
  // runs its MMAs with N = 144 and carries the xyz channels as operand rows 128..130 (rows 131..143 stay zero).
  static constexpr bool kFoldable = kXyz && !kArg;
  static constexpr int kOpTile = kFoldable ? 18 * 1024 : TILE_BYTES;  // one half (hi or lo) of an activation operand stage
  static constexpr int kACol0 = kFoldable ? 160 : TN;                // first dY column in tensor memory (accumulator below)
  static constexpr uint32_t kTmemCols = kFoldable ? 512 : 256;       // accumulator + 2 dY stages of 32 hi + 32 lo columns
  static constexpr int kOps = kNS * 2 * kOpTile;                     // activation operand stages: hi | lo
  static constexpr int kRing = kOps + kNR * kRawBytes;
  static constexpr int kSmem = kRing + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kSmem <= 232448, "shared memory budget");
  static_assert(kRing >= (TC_THREADS / 32) * 32 * 36 * 4 + 4096, "epilogue scratch aliases the rings");
};
template <int AKIND, int BKIND>
__global__ void __launch_bounds__(TC_CTA_THREADS, 1)
wgrad_tc_async_kernel(const __grid_constant__ GemmArgs g) {
  using Cfg = WgCfg<AKIND, BKIND>;
  constexpr int NR = Cfg::kNR, NS = Cfg::kNS;
  pdl_prologue();
  extern __shared__ unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *ops = tiles, *ring_r = tiles + Cfg::kOps;
  uint64_t *full = reinterpret_cast<uint64_t *>(tiles + Cfg::kRing);  // operand stage written (16 warps)
  uint64_t *empty = full + NS;                                        // its MMAs done
  uint64_t *raw_full = empty + NS, *raw_empty = raw_full + NR;        // raw stage landed (512 async arrivals) / read (16 warps)
  uint64_t *done_bar = raw_empty + NR;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
  static_assert((2 * NS + 2 * NR + 1) * 8 + 4 <= 256, "barrier block");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp < TC_THREADS / 32;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const bool fold = Cfg::kFoldable && g.wg_fold && n0 + TN == g.B.feat_cols;  // this tile carries the xyz block as well
  const int cta_id = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int num_kb = k_end > k_begin ? (k_end - k_begin + TK - 1) / TK : 0;
  trace_stamp(g, cta_id, 0, smid());
  trace_stamp(g, cta_id, 1, globaltimer_ns());

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], TC_THREADS / 32); mbar_init(&empty[s], 1); }
    for (int s = 0; s < NR; ++s) { mbar_init(&raw_full[s], TC_THREADS); mbar_init(&raw_empty[s], TC_THREADS / 32); }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  if (fold && producer)  // operand rows 128..143 of every stage half: zero, rows 128..130 are rewritten per k-block
    for (int i = tid; i < NS * 2 * 128; i += TC_THREADS)
      *reinterpret_cast<float4 *>(ops + (i >> 7) * Cfg::kOpTile + 16 * 1024 + (i & 127) * 16) = zero4();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = fold ? idesc_tf32(TM, TN + 16) : idesc_tf32(TM, TN);
  trace_stamp(g, cta_id, 2, globaltimer_ns());

  if (!producer) {  // ---- MMA warp, converged
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % NS;
      mbar_wait_guarded(&full[s], (kb / NS) & 1);
      tc_fence_after_sync();
      const uint32_t a_hi = tmem_u + Cfg::kACol0 + s * 64, a_lo = a_hi + 32;
      const uint32_t bbase = smem_addr(ops + s * 2 * Cfg::kOpTile);
      const uint64_t b_hi = smem_desc_sw128(bbase), b_lo = smem_desc_sw128(bbase + Cfg::kOpTile);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
          const uint64_t adv = static_cast<uint64_t>(2 * ks);
          mma_tf32_ts(tmem_u, a_hi + 8 * ks, b_hi + adv, idesc, kb > 0 || ks > 0);
          mma_tf32_ts(tmem_u, a_hi + 8 * ks, b_lo + adv, idesc, true);
          mma_tf32_ts(tmem_u, a_lo + 8 * ks, b_hi + adv, idesc, true);
        }
        mma_commit(&empty[s]);
        if (kb == num_kb - 1) mma_commit(done_bar);
      }
      __syncwarp();
    }
  } else {
    // transform mapping: channel ch of both 128-channel tiles (= this thread's TMEM lane), positions 8 pg .. 8 pg + 7
    const int quarter = warp & 3, pg = warp >> 2, ch = quarter * 32 + lane;
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + Cfg::kACol0 + pg * 8;
    uint32_t boff[2];  // activation operand: 16-byte chunks 2 pg, 2 pg + 1 of row ch
#pragma unroll
    for (int j = 0; j < 2; ++j) boff[j] = sw128_offset(ch, 2 * pg + j);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f;  // per-channel coefficients of the two sources
    if (m0 + ch < g.A.cols) {
      a0 = __ldg(g.A.c0 + m0 + ch); a1 = __ldg(g.A.c1 + m0 + ch); a2 = __ldg(g.A.c2 + m0 + ch);
    }
    if (BKIND == PN2_ROWS_BNRELU && n0 + ch < g.B.cols) {
      b0 = __ldg(g.B.c0 + n0 + ch); b1 = __ldg(g.B.c1 + n0 + ch);
    }
    const int xd = n0 + ch - g.B.feat_cols;  // gathered source: 0..2 = this thread produces a local coordinate
    const bool xyz_mine = Cfg::kXyz && g.B.use_xyz && xd >= 0 && xd < 3;
    const bool xyz_tile = Cfg::kXyz && g.B.use_xyz && ((g.B.feat_cols >= n0 && g.B.feat_cols < n0 + TN) || fold);
    // issue mapping: 16-byte chunk `lane` (channels 4 lane ..) of positions warp, warp + 16
    const void *dummy = g.A.x;  // a valid address for zero-filling copies (src size 0 reads nothing)
    const uint32_t ring_r_s = smem_addr(ring_r);
    const bool col_a = m0 + 4 * lane < g.A.cols;
    const bool col_b = n0 + 4 * lane < (BKIND == PN2_ROWS_GATHER ? g.B.feat_cols : g.B.cols);
    int i_sr = 0, i_pr = 0, il = 0;
    int c_sr = 0, c_pr = 0;
    // positions advance by 32 per k-block: centre (r / nsample), cloud (r / (npoint nsample)) and pooled row (r / group)
    // are carried incrementally -- three integer divisions per position and k-block were a third of this loop
    IncDiv d_centre, d_cloud, d_group;
    d_centre.init(k_begin + warp, BKIND == PN2_ROWS_GATHER ? g.B.nsample : 1, TK);
    d_cloud.init(k_begin + warp, BKIND == PN2_ROWS_GATHER ? g.B.npoint * g.B.nsample : 1, TK);
    d_group.init(k_begin + warp, AKIND == PN2_ROWS_DYPOOL ? g.A.group : 1, TK);
    // this thread's 16 bytes of position k_begin + warp in each source (the second copy is 16 rows further on), advanced by
    // 32 rows per k-block
    const size_t lda = static_cast<size_t>(g.A.ld), ldb = static_cast<size_t>(g.B.ld);
    const size_t row0 = static_cast<size_t>(k_begin + warp);
    const float *pax = g.A.x + row0 * lda + m0 + 4 * lane;
    const float *pad = g.A.dz + (AKIND == PN2_ROWS_DYPOOL ? 0 : row0 * lda) + m0 + 4 * lane;
    const float *pbx = g.B.x + (BKIND == PN2_ROWS_GATHER ? 0 : row0 * ldb) + n0 + 4 * lane;
    int r0 = k_begin + warp;
    int nidx[2] = {0, 0};
    auto load_idx = [&](int kb) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = k_begin + kb * TK + warp + 16 * i;
        nidx[i] = (BKIND == PN2_ROWS_GATHER && kb < num_kb && r < k_end) ? __ldg(g.B.idx + r) : 0;
      }
    };
    load_idx(0);
    auto issue = [&](int kb) {  // this thread's copies of k-block kb into raw stage i_sr
      const int idx_now[2] = {nidx[0], nidx[1]};
      load_idx(kb + 1);  // consumed by the next call: the index load and the copies that depend on it never meet
      mbar_wait_guarded(&raw_empty[i_sr], i_pr ^ 1);  // (first use: the phase before the barrier's first reads as complete)
      const uint32_t raw = ring_r_s + i_sr * Cfg::kRawBytes;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int p = warp + 16 * i;
        const bool rv = r0 + 16 * i < k_end;
        const uint32_t dst = raw + p * 512 + lane * 16;
        const bool oka = rv && col_a;
        const float *ax = pax + (i ? 16 * lda : 0);
        const float *ad = AKIND == PN2_ROWS_DYPOOL ? pad + static_cast<size_t>(i ? d_group.q_at(16) : d_group.q) * lda
                                                   : pad + (i ? 16 * lda : 0);
        cp_async16(dst, oka ? static_cast<const void *>(ax) : dummy, oka ? 16 : 0);
        cp_async16(dst + TILE_BYTES, oka ? static_cast<const void *>(ad) : dummy, oka ? 16 : 0);
        if (Cfg::kArg) {
          const unsigned char *aa = g.A.arg + (ad - g.A.dz);
          cp_async4(raw + Cfg::kOffArg + p * TM + lane * 4, oka ? static_cast<const void *>(aa) : dummy, oka ? 4 : 0);
        }
        size_t src = 0;
        const float *bx = pbx + (i ? 16 * ldb : 0);
        if (BKIND == PN2_ROWS_GATHER) {
          src = static_cast<size_t>(i ? d_cloud.q_at(16) : d_cloud.q) * g.B.n_src + idx_now[i];
          bx = pbx + src * ldb;
        }
        const bool okb = rv && col_b;
        cp_async16(dst + Cfg::kOffB, okb ? static_cast<const void *>(bx) : dummy, okb ? 16 : 0);
        if (Cfg::kXyz && xyz_tile && lane < 6) {
          const float *q = lane < 3 ? g.B.xyz + src * 3 + lane
                                    : g.B.centres + static_cast<size_t>(i ? d_centre.q_at(16) : d_centre.q) * 3 + (lane - 3);
          cp_async4(raw + Cfg::kOffXyz + (lane * 32 + p) * 4, rv ? static_cast<const void *>(q) : dummy, rv ? 4 : 0);
        }
      }
      r0 += TK;
      pax += TK * lda;
      if (AKIND != PN2_ROWS_DYPOOL) pad += TK * lda;
      if (BKIND != PN2_ROWS_GATHER) pbx += TK * ldb;
      cp_async_arrive(&raw_full[i_sr]);
      if (++i_sr == NR) { i_sr = 0; i_pr ^= 1; }
      ++il;
      if (BKIND == PN2_ROWS_GATHER) { d_centre.advance(); d_cloud.advance(); }
      if (AKIND == PN2_ROWS_DYPOOL) d_group.advance();
    };

    int slot_pg = AKIND == PN2_ROWS_DYPOOL ? (k_begin + pg * 8) % g.A.group : 0;  // pool slot of this thread's first position
    const int slot_step = AKIND == PN2_ROWS_DYPOOL ? TK % g.A.group : 0;
    // the three local-coordinate rows of a gathered source are produced by warps 0..2, lane = position (ONE correctly
    // rounded division per thread; eight per thread in the channel mapping made those warps the tail of every k-block)
    const bool xyz_warp = xyz_tile && warp < 3;
    const uint32_t xoff = sw128_offset(fold ? TN + warp : g.B.feat_cols - n0 + warp, lane >> 2) + (lane & 3) * 4;
    // An operand stage is SIGNALLED one iteration late: tcgen05.wait::st and the async-proxy fence then find their stores
    // long complete instead of stalling all 16 warps (which run in lock-step) at the end of every k-block.
    auto signal = [&](int s) {
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    };
    for (int kb = 0; kb < num_kb; ++kb) {
      // 1. keep the raw ring full: the block about to be read + NR - 1 behind it
      while (il < num_kb && il - kb < NR) issue(il);
      // 2. this k-block has landed (all 512 threads' copies): read this thread's channel of 8 positions, release the stage
      mbar_wait_guarded(&raw_full[c_sr], c_pr);
      const uint32_t rbase = ring_r_s + c_sr * Cfg::kRawBytes;
      const uint32_t rmine = rbase + pg * 8 * 512 + ch * 4;  // this thread's channel of its first position
      const int nvalid = k_end - (k_begin + kb * TK + pg * 8);  // < 8 only in the last k-block of a slice
      float y[8], dz[8], xb[8];
      int slot = slot_pg;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[j] = lds_f32(rmine + j * 512);
        dz[j] = lds_f32(rmine + TILE_BYTES + j * 512);
        xb[j] = lds_f32(rmine + Cfg::kOffB + j * 512);
        if (Cfg::kArg) {
          const int a = lds_u8(rbase + Cfg::kOffArg + pg * 8 * TM + ch + j * TM);
          dz[j] = a == slot ? dz[j] : 0.f;
          if (++slot == g.A.group) slot = 0;
        }
      }
      if (Cfg::kArg) {
        slot_pg += slot_step;
        if (slot_pg >= g.A.group) slot_pg -= g.A.group;
      }
      float xv = 0.f;
      if (Cfg::kXyz && xyz_warp) {  // pointnet2_utils.py:350-352: grouped_xyz -= new_xyz; grouped_xyz /= radius
        const uint32_t rx = rbase + Cfg::kOffXyz + (warp * 32 + lane) * 4;
        xv = __fdiv_rn(__fsub_rn(lds_f32(rx), lds_f32(rx + 3 * 32 * 4)), g.B.inv_scale);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[c_sr]);
      if (++c_sr == NR) { c_sr = 0; c_pr ^= 1; }
      // 3. hand the PREVIOUS k-block's operand stage to the MMA warp; this one's stage is free once the MMAs that read
      //    it (k-block kb - NS) have completed
      const int s = kb % NS;
      if (kb > 0) signal((kb - 1) % NS);
      mbar_wait_guarded(&empty[s], ((kb / NS) & 1) ^ 1);
      tc_fence_after_sync();
      // 4. dY = c0 dz + c1 + c2 y  -> tensor memory (positions past the slice contribute nothing: dY = 0 there)
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf(a2, y[j], fmaf(a0, dz[j], a1));
      if (nvalid < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = j < nvalid ? y[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) split_tf32(y[j], hi[j], lo[j]);
      tmem_st8(lane_base + s * 64, hi);
      tmem_st8(lane_base + s * 64 + 32, lo);
      // 5. activation -> K-major operand tile
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = xb[j];
        if (BKIND == PN2_ROWS_BNRELU) v = relu_nan(fmaf(v, b0, b1));
        split_tf32(v, hi[j], lo[j]);
      }
      unsigned char *st = ops + s * 2 * Cfg::kOpTile;
      if (!xyz_mine) {  // (a local-coordinate row belongs to its warp below)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          *reinterpret_cast<float4 *>(st + boff[j]) = make_float4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          *reinterpret_cast<float4 *>(st + Cfg::kOpTile + boff[j]) = make_float4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
      }
      if (Cfg::kXyz && xyz_warp) {
        float h, l;
        split_tf32(xv, h, l);
        *reinterpret_cast<float *>(st + xoff) = h;
        *reinterpret_cast<float *>(st + Cfg::kOpTile + xoff) = l;
      }
    }
    if (num_kb > 0) signal((num_kb - 1) % NS);
    float4 yv[8];
    tc_epilogue_prefetch<TC_EPI_STORE>(g, m0, n0, yv);
    if (num_kb > 0) mbar_wait_guarded(done_bar, 0);
    tc_fence_after_sync();
    trace_stamp(g, cta_id, 3, globaltimer_ns());
    if (Cfg::kFoldable && fold && pg == 0 && num_kb > 0) {  // accumulator columns 128..130: the xyz block's gradient
      float v[32];
      tmem_ld32(tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + TN, v);
      if (m0 + ch < g.M)
        *reinterpret_cast<float4 *>(g.out + blockIdx.z * g.out_split_stride + static_cast<size_t>(m0 + ch) * g.ldo + n0 + TN) =
            make_float4(v[0], v[1], v[2], v[3]);
    } else if (Cfg::kFoldable && fold && pg == 0 && m0 + ch < g.M) {
      *reinterpret_cast<float4 *>(g.out + blockIdx.z * g.out_split_stride + static_cast<size_t>(m0 + ch) * g.ldo + n0 + TN) = zero4();
    }
    float(&red)[2][TC_THREADS / 32][32] =
        *reinterpret_cast<float(*)[2][TC_THREADS / 32][32]>(tiles + (TC_THREADS / 32) * 32 * 36 * 4);  // unused by EPI_STORE
    tc_epilogue<TC_EPI_STORE>(g, tmem_d, tiles, red, m0, n0, blockIdx.x, blockIdx.z, num_kb > 0, yv);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::kTmemCols>(tmem_d);
  trace_stamp(g, cta_id, 4, globaltimer_ns());
  trace_stamp(g, cta_id, 5, static_cast<unsigned long long>(num_kb));
}

template <int AKIND, int BKIND>
int launch_wgrad_async(const GemmArgs &g, int splits, cudaStream_t stream) {
  using Cfg = WgCfg<AKIND, BKIND>;
  auto kernel = wgrad_tc_async_kernel<AKIND, BKIND>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    configured_dev = dev;
  }
  dim3 grid((g.M + TM - 1) / TM, g.wg_fold ? g.B.feat_cols / TN : (g.N + TN - 1) / TN, splits);
  if (g.wg_fold && !Cfg::kFoldable) return PN2_TC_UNSUPPORTED;  // the caller asks wgrad_fold_ok() first
  GemmArgs a = g;
  gemm_trace_target(&a.trace, &a.trace_cap);
  pn2::launch(kernel, grid, dim3(TC_CTA_THREADS), Cfg::kSmem, stream, a);
  return check_launch("wgrad_tc_async_kernel");
}

}  // namespace

// development aid: per-CTA phase timestamps of the next tensor-core GEMM launches (tools/gemm_trace.py)
static unsigned long long *g_trace_buf = nullptr;
static int g_trace_cap = 0;
void gemm_trace_target(unsigned long long **buf, int *cap) {
  *buf = g_trace_buf;
  *cap = g_trace_cap;
}

// xyz block of a gathered source folded into the last feature tile of the weight-gradient kernel (WgCfg::kFoldable)
bool wgrad_fold_ok(const void *gemm_args) {
  static const bool on = [] {
    const char *e = getenv("PN2_WGRAD_FOLD");
    return e == nullptr || e[0] != '0';
  }();
  const GemmArgs &g = *static_cast<const GemmArgs *>(gemm_args);
  return on && gemm_tc_enabled() && g.A.kind == PN2_ROWS_DY && g.B.kind == PN2_ROWS_GATHER && g.B.use_xyz && g.B.feat_cols >= TN &&
         g.B.feat_cols % TN == 0 && g.B.cols == g.B.feat_cols + 4;
}

static bool small_k_ffma() {
  static const bool on = [] {
    const char *e = getenv("PN2_TC_SMALLK_FFMA");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

bool gemm_tc_wide_enabled() {
  static const bool on = [] {
    const char *e = getenv("PN2_TC_WIDE");
    return e == nullptr || e[0] != '0';
  }();
  return on;
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
  if (akind != PN2_ROWS_PLAIN && akind != PN2_ROWS_GATHER && g.K > 640) return PN2_TC_UNSUPPORTED;  // coefficient staging
  if (g.b_img == nullptr) return PN2_TC_UNSUPPORTED;  // the weight operand comes as the pre-split image only
  // K <= 16 (the first layer of a level without input features: xyz only): one k-block per tile, i.e. a launch that is
  // all tile start-up and epilogue; the FFMA kernel writes the same output at memory speed
  if (g.K <= 16 && epi == TC_EPI_STORE_STATS && small_k_ffma()) return PN2_TC_UNSUPPORTED;
#define PN2_TC_CASE(AK, EP) \
  if (akind == AK && epi == EP) return launch_tc_async<AK, EP>(g, stream);
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
  if (g.A.kind == AK && g.B.kind == BK) return launch_wgrad_async<AK, BK>(g, splits, stream);
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

PN2_EXPORT int pn2_debug_gemm_trace2(unsigned long long *device_buf, int ctas) {
  int n = device_buf ? ctas : 0;
  if (cudaMemcpyToSymbol(pn2::g_trace2, &device_buf, sizeof(device_buf)) != cudaSuccess) return pn2::check_launch("pn2_debug_gemm_trace2");
  if (cudaMemcpyToSymbol(pn2::g_trace2_ctas, &n, sizeof(n)) != cudaSuccess) return pn2::check_launch("pn2_debug_gemm_trace2");
  return PN2_OK;
}

PN2_EXPORT int pn2_debug_gemm_trace(unsigned long long *device_buf, int ctas) {
  pn2::g_trace_buf = device_buf;
  pn2::g_trace_cap = device_buf ? ctas : 0;
  return PN2_OK;
}
