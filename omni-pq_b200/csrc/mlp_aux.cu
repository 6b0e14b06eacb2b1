// Element-wise / reduction companions of the shared-MLP GEMMs (mlp_gemm.cu): layout changes between the
// reference's channel-major API tensors (B,C,N) and the position-major activations the GEMMs use,
// BatchNorm statistics -> folded scale/shift (forward) and -> backward coefficients, the fused
// BatchNorm+ReLU+max-pool over nsample (reference F.max_pool2d, pointnet2_modules.py:254-257) with its
// backward preparation, and the feature-propagation front end (three_nn + inverse-distance weights +
// three_interpolate, pointnet2_modules.py:393-401) as one gather-MAC kernel.
// All of these are HBM-bound streaming kernels: 128-bit accesses along the channel dimension, grids
// sized from the problem so a single cloud still covers the chip.
#include <math_constants.h>

#include "pn2_common.cuh"

namespace pn2 {
namespace {

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float relu_nan(float v) { return !(v <= 0.f) ? v : 0.f; }  // NaN-propagating ReLU

// ---- (B,C,N) <-> (B,N,ld) ----------------------------------------------------------------------------
// 32x32 tiles through shared memory, coalesced on both sides; columns c..ld-1 of the point-major side
// are written as zeros.
__global__ void __launch_bounds__(256)
to_point_major_kernel(int c, int n, int ld, int stride, const float *__restrict__ src, float *__restrict__ dst) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
  src += static_cast<size_t>(b) * c * n;
  dst += static_cast<size_t>(b) * n * stride;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int cc = c0 + r, nn = n0 + tx;
    tile[r][tx] = (cc < c && nn < n) ? src[static_cast<size_t>(cc) * n + nn] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int nn = n0 + r, cc = c0 + tx;
    if (nn < n && cc < ld) dst[static_cast<size_t>(nn) * stride + cc] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(256)
to_channel_major_kernel(int c, int n, int stride, const float *__restrict__ src, float *__restrict__ dst) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  src += static_cast<size_t>(b) * n * stride;
  dst += static_cast<size_t>(b) * c * n;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int nn = n0 + r, cc = c0 + tx;
    tile[r][tx] = (nn < n && cc < c) ? src[static_cast<size_t>(nn) * stride + cc] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int cc = c0 + r, nn = n0 + tx;
    if (cc < c && nn < n) dst[static_cast<size_t>(cc) * n + nn] = tile[tx][r];
  }
}

// ---- BatchNorm statistics ---------------------------------------------------------------------------
// stats[tiles][2][np] fp32 partials -> fp64 totals.  Block = 8 channels x 128 tile lanes: lane ty walks
// tiles ty, ty+128, ... (one 32-byte sector per tile row), the lane totals are combined in a fixed
// order through shared memory, so the result is deterministic.  Every thread with ty == 0 gets the totals.
constexpr int kStatCh = 8, kStatLanes = 128;  // 8 channels x 128 tile lanes per CTA: np/8 CTAs share the reduction

__device__ __forceinline__ void total_of(const float *stats, int tiles, int np, int ch, bool live, double &s1,
                                         double &s2) {
  __shared__ double red[2][kStatLanes][kStatCh];
  __shared__ double red8[2][8][kStatCh];
  const int tx = threadIdx.x % kStatCh, ty = threadIdx.x / kStatCh;
  double a = 0.0, b = 0.0;
  if (live && stats) {
    for (int t = ty; t < tiles; t += kStatLanes) {
      a += static_cast<double>(stats[(static_cast<size_t>(t) * 2 + 0) * np + ch]);
      b += static_cast<double>(stats[(static_cast<size_t>(t) * 2 + 1) * np + ch]);
    }
  }
  red[0][ty][tx] = a;
  red[1][ty][tx] = b;
  __syncthreads();
  if (ty < 8) {  // fixed-order two-level combine: 8 partial sums of 16 lanes each, then 8 -> 1
    double pa = 0.0, pb = 0.0;
#pragma unroll
    for (int l = 0; l < kStatLanes / 8; ++l) {
      pa += red[0][ty * (kStatLanes / 8) + l][tx];
      pb += red[1][ty * (kStatLanes / 8) + l][tx];
    }
    red8[0][ty][tx] = pa;
    red8[1][ty][tx] = pb;
  }
  __syncthreads();
  s1 = 0.0; s2 = 0.0;
  if (ty == 0) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      s1 += red8[0][l][tx];
      s2 += red8[1][l][tx];
    }
  }
}

__global__ void __launch_bounds__(kStatCh * kStatLanes)
bn_reduce_stats_kernel(int tiles, int c, int np, double count, const float *__restrict__ stats,
                       double *__restrict__ sums) {
  pdl_prologue();
  const int ch = blockIdx.x * kStatCh + threadIdx.x % kStatCh;
  double s1, s2;
  total_of(stats, tiles, np, ch, ch < c, s1, s2);
  if (blockIdx.x == 0 && threadIdx.x == 0) sums[2 * c] = count;  // rows behind these totals: reduced with them
  if (threadIdx.x / kStatCh != 0 || ch >= c) return;
  sums[ch] = s1;
  sums[c + ch] = s2;
}

// ---- SyncBatchNorm: statistics reduction + cross-rank exchange in ONE kernel, over NVLink peer memory -----------------
// (models/pq_transformer.py:194 converts every hot-path BatchNorm into a SyncBatchNorm, which torch
// serves with an all_gather + an all_reduce and their host-side bookkeeping per layer: 38 small NCCL collectives per
// step on the path.)  Every rank owns a buffer in symmetric memory (torch.distributed._symmetric_memory: the same
// allocation mapped into every peer's address space): [2 slots][2c+2 doubles] + one 32-bit signal word per sender.
//   1. each CTA reduces its 8 channels' per-tile partials to fp64 totals (as bn_reduce_stats_kernel) and stores them
//      into THIS rank's slot (epoch parity); CTA 0 adds the rank's row count;
//   2. the last CTA to finish (device-scope ticket) publishes the slot: one st.release.sys of the epoch number into
//      every peer's signal word for this rank -- the only "send" of the exchange;
//   3. every CTA waits until all peers' signal words have reached the epoch (ld.acquire.sys; traps after ~2 s instead of
//      hanging the GPU) and then LOADS the peers' totals of its 8 channels straight from their memory over NVLink,
//      adding them in rank order -- every rank computes bit-identical totals, no reduction tree, no second kernel.
// Output: sums[2c+1] = (sum, second sum, row count) over all ranks, consumed on the device by the finalize kernels.
// Slot reuse is safe with two slots: a rank can enter exchange e+2 only after every peer has signalled e+1, i.e. after
// the peers' kernels of exchange e (which read this rank's slot e) have completed.
__device__ __forceinline__ void st_release_sys_u32(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kStatCh * kStatLanes)
bn_sync_exchange_kernel(int tiles, int c, int np, double count, const float *__restrict__ stats,
                        const unsigned long long *__restrict__ peers /* [world] base address of every rank's buffer */,
                        int rank, int world, unsigned epoch, int slot_doubles, unsigned *__restrict__ cta_ticket,
                        double *__restrict__ sums) {
  pdl_prologue();
  const int ch = blockIdx.x * kStatCh + threadIdx.x % kStatCh;
  const bool writer = threadIdx.x / kStatCh == 0 && ch < c;
  double s1, s2;
  total_of(stats, tiles, np, ch, ch < c, s1, s2);
  double *mine = reinterpret_cast<double *>(peers[rank]) + static_cast<size_t>(epoch & 1u) * slot_doubles;
  if (writer) {
    mine[ch] = s1;
    mine[c + ch] = s2;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) mine[2 * c] = count;
  __threadfence_system();  // this thread's slot stores are visible system-wide before the ticket below
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(cta_ticket, 1u);
    if (done == gridDim.x - 1) {  // every CTA of this rank has stored and fenced its part of the slot: publish it
      *cta_ticket = 0;
      __threadfence_system();
      for (int r = 0; r < world; ++r) {
        unsigned *sig = reinterpret_cast<unsigned *>(reinterpret_cast<double *>(peers[r]) + 2 * slot_doubles) + rank;
        st_release_sys_u32(sig, epoch);
      }
    }
    const unsigned *my_sig = reinterpret_cast<const unsigned *>(reinterpret_cast<double *>(peers[rank]) + 2 * slot_doubles);
    long long t0 = 0;
    for (int r = 0; r < world; ++r) {
      unsigned spins = 0;
      while (static_cast<int>(ld_acquire_sys_u32(my_sig + r) - epoch) < 0) {  // wrap-safe "signal < epoch"
        if ((++spins & 1023u) == 0) {
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000ll) __trap();  // a peer never arrived: fail loudly instead of hanging the GPU
        }
      }
    }
  }
  __syncthreads();
  if (!writer && !(blockIdx.x == 0 && threadIdx.x == 0)) return;
  double t1 = 0.0, t2 = 0.0, n = 0.0;
  for (int r = 0; r < world; ++r) {  // fixed rank order: bit-identical totals on every rank
    const double *theirs = reinterpret_cast<const double *>(peers[r]) + static_cast<size_t>(epoch & 1u) * slot_doubles;
    if (writer) {
      t1 += ld_relaxed_sys_f64(theirs + ch);
      t2 += ld_relaxed_sys_f64(theirs + c + ch);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) n += ld_relaxed_sys_f64(theirs + 2 * c);
  }
  if (writer) {
    sums[ch] = t1;
    sums[c + ch] = t2;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) sums[2 * c] = n;
}

__global__ void bn_finalize_kernel(int training, int tiles, int c, int np, double count,
                                   const float *__restrict__ stats, const double *__restrict__ sums,
                                   const float *__restrict__ gamma, const float *__restrict__ beta,
                                   float *__restrict__ running_mean, float *__restrict__ running_var,
                                   long long *__restrict__ nbt, float momentum, float eps, float *__restrict__ scale,
                                   float *__restrict__ shift, float *__restrict__ mean_out,
                                   float *__restrict__ invstd_out) {
  pdl_prologue();
  const int ch = blockIdx.x * kStatCh + threadIdx.x % kStatCh;
  // num_batches_tracked += 1 rides along when the momentum is fixed (nobody reads the counter then); with
  // momentum=None every channel reads it, so the host launches a separate increment afterwards
  if (training && nbt && running_mean && momentum >= 0.f && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
  double s1 = 0.0, s2 = 0.0;
  if (training && !sums) total_of(stats, tiles, np, ch, ch < c, s1, s2);  // block-cooperative: before any return
  if (threadIdx.x / kStatCh != 0 || ch >= np) return;
  if (ch >= c) {  // zero padding: padded channels stay exactly 0 through BN+ReLU
    scale[ch] = 0.f; shift[ch] = 0.f; mean_out[ch] = 0.f; invstd_out[ch] = 0.f;
    return;
  }
  float mean, invstd;
  if (training) {
    if (sums) { s1 = sums[ch]; s2 = sums[c + ch]; count = sums[2 * c]; }  // all-reduced totals carry the global row count
    const double mu = s1 / count;
    double var = s2 / count - mu * mu;  // biased variance (what BatchNorm normalises with)
    if (var < 0.0) var = 0.0;
    mean = static_cast<float>(mu);
    invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean && running_var) {
      // momentum < 0: cumulative moving average with factor 1/num_batches_tracked (after increment)
      const double f = momentum >= 0.f ? static_cast<double>(momentum)
                                       : 1.0 / static_cast<double>((nbt ? *nbt : 0) + 1);
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[ch] = static_cast<float>((1.0 - f) * running_mean[ch] + f * mu);
      running_var[ch] = static_cast<float>((1.0 - f) * running_var[ch] + f * unbiased);
    }
  } else {
    mean = running_mean[ch];
    invstd = static_cast<float>(1.0 / sqrt(static_cast<double>(running_var[ch]) + static_cast<double>(eps)));
  }
  const float gmm = gamma ? gamma[ch] : 1.f, bt = beta ? beta[ch] : 0.f;
  const float sc = gmm * invstd;
  scale[ch] = sc;
  shift[ch] = fmaf(-mean, sc, bt);
  mean_out[ch] = mean;
  invstd_out[ch] = invstd;
}

__global__ void bn_bump_counter_kernel(long long *nbt) {
  pdl_prologue(); *nbt += 1; }

__global__ void bn_bwd_finalize_kernel(int training, int tiles, int c, int np, double count,
                                       const float *__restrict__ stats, const double *__restrict__ sums,
                                       const double *__restrict__ count_dev, const float *__restrict__ gamma,
                                       const float *__restrict__ mean, const float *__restrict__ invstd,
                                       float *__restrict__ ca, float *__restrict__ cb, float *__restrict__ cc,
                                       float *__restrict__ dgamma, float *__restrict__ dbeta) {
  pdl_prologue();
  const int ch = blockIdx.x * kStatCh + threadIdx.x % kStatCh;
  // this rank's totals of (dz, dz*y).  SyncBatchNorm: `sums` holds the totals over ALL ranks; they enter the
  // input-gradient coefficients only -- dgamma / dbeta stay rank-local exactly like torch's SyncBatchNorm
  // (batch_norm_backward_reduce returns local grad_weight / grad_bias; DDP averages them afterwards).
  double l_dz = 0.0, l_dzy = 0.0;
  if (stats) total_of(stats, tiles, np, ch, ch < c, l_dz, l_dzy);  // block-cooperative: before any return
  if (threadIdx.x / kStatCh != 0 || ch >= np) return;
  if (ch >= c) {
    ca[ch] = 0.f; cb[ch] = 0.f; cc[ch] = 0.f;
    return;
  }
  if (count_dev) count = *count_dev;
  double s_dz = l_dz, s_dzy = l_dzy;
  if (sums) { s_dz = sums[ch]; s_dzy = sums[c + ch]; }
  if (!stats) { l_dz = s_dz; l_dzy = s_dzy; }
  const double mu = mean[ch], r = invstd[ch], gmm = gamma ? gamma[ch] : 1.0;
  const double s_dzn = r * (s_dzy - mu * s_dz);  // sum dz * normalised y (over every rank's rows)
  if (dgamma) dgamma[ch] = static_cast<float>(r * (l_dzy - mu * l_dz));
  if (dbeta) dbeta[ch] = static_cast<float>(l_dz);
  const double s = gmm * r;
  if (training) {
    const double m1 = s_dz / count, m2 = s_dzn / count;
    const double kc = -s * m2 * r;
    ca[ch] = static_cast<float>(s);
    cc[ch] = static_cast<float>(kc);
    cb[ch] = static_cast<float>(-s * m1 - kc * mu);
  } else {
    ca[ch] = static_cast<float>(s);
    cb[ch] = 0.f;
    cc[ch] = 0.f;
  }
}

// ---- BN + ReLU + max over the nsample rows of a group -----------------------------------------------
// one thread per (group, 4 channels); consecutive threads -> consecutive channels (coalesced 128-bit rows)
__global__ void __launch_bounds__(256)
bn_relu_pool_kernel(int groups, int group, int ld, const float *__restrict__ y, const float *__restrict__ scale,
                    const float *__restrict__ shift, float *__restrict__ out_pm, unsigned char *__restrict__ arg) {
  pdl_prologue();
  const int q = ld / 4;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(groups) * q) return;
  const int g = static_cast<int>(t / q), c4 = static_cast<int>(t % q) * 4;
  const float4 sc = ldg4(scale + c4), sh = ldg4(shift + c4);
  const float *row = y + static_cast<size_t>(g) * group * ld + c4;
  float4 best = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  uchar4 bi = make_uchar4(0, 0, 0, 0);
  for (int s = 0; s < group; ++s) {
    const float4 v = ldg4(row + static_cast<size_t>(s) * ld);
    // ReLU and max both propagate NaN like torch.relu / F.max_pool2d (a diverged run must not be masked)
    const float a = relu_nan(fmaf(v.x, sc.x, sh.x)), b = relu_nan(fmaf(v.y, sc.y, sh.y));
    const float c = relu_nan(fmaf(v.z, sc.z, sh.z)), d = relu_nan(fmaf(v.w, sc.w, sh.w));
    if (a > best.x || a != a) { best.x = a; bi.x = s; }
    if (b > best.y || b != b) { best.y = b; bi.y = s; }
    if (c > best.z || c != c) { best.z = c; bi.z = s; }
    if (d > best.w || d != d) { best.w = d; bi.w = s; }
  }
  *reinterpret_cast<float4 *>(out_pm + static_cast<size_t>(g) * ld + c4) = best;
  if (arg) *reinterpret_cast<uchar4 *>(arg + static_cast<size_t>(g) * ld + c4) = bi;
}

// gz = gout * (out > 0) in place; partial sums over a strip of groups of gz and gz * y[arg]
constexpr int kPrepGroups = 8;  // groups per CTA strip

__global__ void __launch_bounds__(128)
pool_bwd_prep_kernel(int groups, int group, int ld, float *__restrict__ gz, const float *__restrict__ out_pm,
                     const unsigned char *__restrict__ arg, const float *__restrict__ y,
                     float *__restrict__ stats) {
  pdl_prologue();
  const int g0 = blockIdx.x * kPrepGroups;
  for (int c4 = threadIdx.x * 4; c4 < ld; c4 += blockDim.x * 4) {
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    for (int g = g0; g < min(groups, g0 + kPrepGroups); ++g) {
      const size_t o = static_cast<size_t>(g) * ld + c4;
      float4 v = *reinterpret_cast<const float4 *>(gz + o);
      const float4 z = ldg4(out_pm + o);
      v.x = z.x > 0.f ? v.x : 0.f; v.y = z.y > 0.f ? v.y : 0.f;
      v.z = z.z > 0.f ? v.z : 0.f; v.w = z.w > 0.f ? v.w : 0.f;
      *reinterpret_cast<float4 *>(gz + o) = v;
      uchar4 a = make_uchar4(0, 0, 0, 0);
      if (arg) a = __ldg(reinterpret_cast<const uchar4 *>(arg + o));
      const float *yr = y + static_cast<size_t>(g) * group * ld + c4;
      const float y0 = __ldg(yr + static_cast<size_t>(a.x) * ld + 0), y1 = __ldg(yr + static_cast<size_t>(a.y) * ld + 1);
      const float y2 = __ldg(yr + static_cast<size_t>(a.z) * ld + 2), y3 = __ldg(yr + static_cast<size_t>(a.w) * ld + 3);
      s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
      s2.x = fmaf(v.x, y0, s2.x); s2.y = fmaf(v.y, y1, s2.y); s2.z = fmaf(v.z, y2, s2.z); s2.w = fmaf(v.w, y3, s2.w);
    }
    float *dst = stats + static_cast<size_t>(blockIdx.x) * 2 * ld + c4;
    *reinterpret_cast<float4 *>(dst) = s1;
    *reinterpret_cast<float4 *>(dst + ld) = s2;
  }
}

// ---- feature propagation front end --------------------------------------------------------------------
// Phase 1 (per unknown point): 3-NN exactly like three_nn_kernel (interpolate_gpu.cu:14-64) + weights
// w_t = (1/(sqrt(d2_t)+1e-8)) / sum (pointnet2_modules.py:395-397).  Phase 2: the warp that owns the point
// streams the three known rows (point-major, 128-bit) and writes the blended row.
constexpr int kFpThreads = 128;
constexpr int kFpTile = 1024;

// PPB = unknown points per CTA: 128 (one per thread) for large clouds; 32 for the backbone's FP modules (512 / 1024
// unknown points), where phase 2 -- a warp blending its points' rows one after the other, ~1 us of dependent row
// loads each -- is the whole run time and 4 CTAs of 128 points left it at 60 us (profiles/r1_ncu_rest.json).
template <int PPB>
__global__ void __launch_bounds__(kFpThreads)
fp_interpolate_kernel(int n, int m, int c4n, int ld_known, const float *__restrict__ unknown,
                      const float *__restrict__ known, const float *__restrict__ known_pm, float *__restrict__ out,
                      int ldo, int *__restrict__ idx_out, float *__restrict__ w_out) {
  pdl_prologue();
  __shared__ float tile[kFpTile * 3];
  __shared__ int s_idx[kFpThreads][3];
  __shared__ float s_w[kFpThreads][3];
  const int b = blockIdx.y;
  unknown += static_cast<size_t>(b) * n * 3;
  known += static_cast<size_t>(b) * m * 3;
  const int j = blockIdx.x * PPB + threadIdx.x;
  const bool live = j < n && threadIdx.x < PPB;
  const int jj = live ? j : n - 1;
  const float ux = unknown[jj * 3 + 0], uy = unknown[jj * 3 + 1], uz = unknown[jj * 3 + 2];
  float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int base = 0; base < m; base += kFpTile) {
    const int tn = min(kFpTile, m - base);
    __syncthreads();
    for (int i = threadIdx.x; i < tn * 3; i += kFpThreads) tile[i] = known[base * 3 + i];
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < (threadIdx.x < PPB ? tn : 0); ++k) {
      const float d = dist2(ux, uy, uz, tile[k * 3 + 0], tile[k * 3 + 1], tile[k * 3 + 2]);
      if (d < b3) {
        const int kk = base + k;
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = kk; }
        else { b3 = d; i3 = kk; }
      }
    }
  }
  // dist = sqrt(dist2); recip = 1/(dist + 1e-8); w = recip / sum(recip)   (left-to-right sum like torch.sum)
  const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
  const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
  const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
  const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
  const float w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm), w3 = __fdiv_rn(r3, norm);
  s_idx[threadIdx.x][0] = i1; s_idx[threadIdx.x][1] = i2; s_idx[threadIdx.x][2] = i3;
  s_w[threadIdx.x][0] = w1; s_w[threadIdx.x][1] = w2; s_w[threadIdx.x][2] = w3;
  if (live) {
    int *io = idx_out + (static_cast<size_t>(b) * n + j) * 3;
    float *wo = w_out + (static_cast<size_t>(b) * n + j) * 3;
    io[0] = i1; io[1] = i2; io[2] = i3;
    wo[0] = w1; wo[1] = w2; wo[2] = w3;
  }
  __syncthreads();
  // phase 2: each warp blends the rows of its PPB/4 points, one point at a time, lanes across channels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float *kp = known_pm + static_cast<size_t>(b) * m * ld_known;
  constexpr int PPW = PPB / (kFpThreads / 32);
  for (int p = 0; p < PPW; ++p) {
    const int t = warp * PPW + p, jp = blockIdx.x * PPB + t;
    if (jp >= n) break;
    const float *ra = kp + static_cast<size_t>(s_idx[t][0]) * ld_known;
    const float *rb = kp + static_cast<size_t>(s_idx[t][1]) * ld_known;
    const float *rc = kp + static_cast<size_t>(s_idx[t][2]) * ld_known;
    const float wa = s_w[t][0], wb = s_w[t][1], wc = s_w[t][2];
    float *o = out + (static_cast<size_t>(b) * n + jp) * ldo;
    for (int q = lane; q < c4n; q += 32) {
      const float4 a = ldg4(ra + q * 4), bb = ldg4(rb + q * 4), cc = ldg4(rc + q * 4);
      float4 r;  // interpolate_gpu.cu:103-104 contraction: p2*w2, then fma p1*w1, then fma p3*w3
      r.x = __fmaf_rn(cc.x, wc, __fmaf_rn(a.x, wa, __fmul_rn(bb.x, wb)));
      r.y = __fmaf_rn(cc.y, wc, __fmaf_rn(a.y, wa, __fmul_rn(bb.y, wb)));
      r.z = __fmaf_rn(cc.z, wc, __fmaf_rn(a.z, wa, __fmul_rn(bb.z, wb)));
      r.w = __fmaf_rn(cc.w, wc, __fmaf_rn(a.w, wa, __fmul_rn(bb.w, wb)));
      *reinterpret_cast<float4 *>(o + q * 4) = r;
    }
  }
}

// The backbone's FP modules have 512 / 1024 unknown points: a thread per point is 16 / 32 CTAs of one busy warp each
// (45 us per launch, on the critical path of the step).  Here a WARP owns a point: its lanes scan the known points
// 32 apart, each keeping its three nearest in index order, and the warp then extracts three times the smallest
// (distance, index) pair -- exactly the triple three_nn_kernel's sequential scan with strict `<` ends with (equal
// distances keep the lower index; slots nobody fills stay (inf, 0)) -- and blends the three rows, lanes across channels.
__global__ void __launch_bounds__(kFpThreads)
fp_interpolate_warp_kernel(int n, int m, int c4n, int ld_known, const float *__restrict__ unknown,
                           const float *__restrict__ known, const float *__restrict__ known_pm, float *__restrict__ out,
                           int ldo, int *__restrict__ idx_out, float *__restrict__ w_out) {
  pdl_prologue();
  __shared__ float tile[kFpTile * 3];
  const int b = blockIdx.y;
  unknown += static_cast<size_t>(b) * n * 3;
  known += static_cast<size_t>(b) * m * 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * (kFpThreads / 32) + warp;
  const bool live = j < n;
  const int jj = live ? j : n - 1;
  const float ux = unknown[jj * 3 + 0], uy = unknown[jj * 3 + 1], uz = unknown[jj * 3 + 2];
  float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int base = 0; base < m; base += kFpTile) {
    const int tn = min(kFpTile, m - base);
    __syncthreads();
    for (int i = threadIdx.x; i < tn * 3; i += kFpThreads) tile[i] = known[base * 3 + i];
    __syncthreads();
    for (int k = lane; k < tn; k += 32) {  // (stride of 3 floats between lanes: conflict-free)
      const float d = dist2(ux, uy, uz, tile[k * 3 + 0], tile[k * 3 + 1], tile[k * 3 + 2]);
      if (d < b3) {
        const int kk = base + k;
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = kk; }
        else { b3 = d; i3 = kk; }
      }
    }
  }
  // distances are >= 0 (or never stored: NaN fails `<`), so their bit patterns order like the values and
  // (distance bits << 32 | index) orders like (distance, index)
  float rd[3];
  int ri[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(b1)) << 32) | static_cast<unsigned>(i1);
    unsigned long long best = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    rd[t] = __uint_as_float(static_cast<unsigned>(best >> 32));
    ri[t] = static_cast<int>(best & 0xffffffffu);
    if (key == best) { b1 = b2; i1 = i2; b2 = b3; i2 = i3; b3 = CUDART_INF_F; i3 = 0; }  // (only fillers can tie across lanes)
  }
  if (!live) return;
  // dist = sqrt(dist2); recip = 1/(dist + 1e-8); w = recip / sum(recip)   (left-to-right sum like torch.sum)
  const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(rd[0]), 1e-8f));
  const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(rd[1]), 1e-8f));
  const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(rd[2]), 1e-8f));
  const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
  const float wa = __fdiv_rn(r1, norm), wb = __fdiv_rn(r2, norm), wc = __fdiv_rn(r3, norm);
  if (lane == 0) {
    int *io = idx_out + (static_cast<size_t>(b) * n + j) * 3;
    float *wo = w_out + (static_cast<size_t>(b) * n + j) * 3;
    io[0] = ri[0]; io[1] = ri[1]; io[2] = ri[2];
    wo[0] = wa; wo[1] = wb; wo[2] = wc;
  }
  const float *kp = known_pm + static_cast<size_t>(b) * m * ld_known;
  const float *ra = kp + static_cast<size_t>(ri[0]) * ld_known;
  const float *rb = kp + static_cast<size_t>(ri[1]) * ld_known;
  const float *rc = kp + static_cast<size_t>(ri[2]) * ld_known;
  float *o = out + (static_cast<size_t>(b) * n + j) * ldo;
  for (int q = lane; q < c4n; q += 32) {
    const float4 a = ldg4(ra + q * 4), bb = ldg4(rb + q * 4), cc = ldg4(rc + q * 4);
    float4 r;  // interpolate_gpu.cu:103-104 contraction: p2*w2, then fma p1*w1, then fma p3*w3
    r.x = __fmaf_rn(cc.x, wc, __fmaf_rn(a.x, wa, __fmul_rn(bb.x, wb)));
    r.y = __fmaf_rn(cc.y, wc, __fmaf_rn(a.y, wa, __fmul_rn(bb.y, wb)));
    r.z = __fmaf_rn(cc.z, wc, __fmaf_rn(a.z, wa, __fmul_rn(bb.z, wb)));
    r.w = __fmaf_rn(cc.w, wc, __fmaf_rn(a.w, wa, __fmul_rn(bb.w, wb)));
    *reinterpret_cast<float4 *>(o + q * 4) = r;
  }
}

// dknown_pm[idx_t][c] += w_t * dout[row][c]; one warp per unknown point, lanes across channels
__global__ void __launch_bounds__(256)
fp_interpolate_grad_kernel(int n, int m, int c4n, const float *__restrict__ dout, int ldo,
                           const int *__restrict__ idx, const float *__restrict__ weight,
                           float *__restrict__ dknown_pm, int ld_known) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (j >= n) return;
  const size_t row = static_cast<size_t>(b) * n + j;
  const int *ix = idx + row * 3;
  const float *w = weight + row * 3;
  float *base = dknown_pm + static_cast<size_t>(b) * m * ld_known;
  const float *g = dout + row * ldo;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    float *dst = base + static_cast<size_t>(ix[t]) * ld_known;
    const float wt = w[t];
    for (int q = lane; q < c4n; q += 32) {
      const float4 v = ldg4(g + q * 4);
      atomicAdd(dst + q * 4 + 0, __fmul_rn(v.x, wt));
      atomicAdd(dst + q * 4 + 1, __fmul_rn(v.y, wt));
      atomicAdd(dst + q * 4 + 2, __fmul_rn(v.z, wt));
      atomicAdd(dst + q * 4 + 3, __fmul_rn(v.w, wt));
    }
  }
}

}  // namespace
}  // namespace pn2

using namespace pn2;

static int transpose_launch(bool to_pm, int b, int c, int n, int ld, int stride, const float *src, float *dst,
                            void *stream, const char *what) {
  PN2_REQUIRE(b >= 0 && c >= 0 && n >= 0 && ld >= c && stride >= ld, "%s: bad extents b=%d c=%d n=%d ld=%d stride=%d", what, b,
              c, n, ld, stride);
  const int cc = to_pm ? ld : c;
  if (b == 0 || n == 0 || cc == 0) return PN2_OK;
  PN2_REQUIRE(src && dst && b <= 65535, "%s: null pointer or batch too large", what);
  dim3 grid((n + 31) / 32, (cc + 31) / 32, b);
  PN2_REQUIRE(grid.y <= 65535, "%s: too many channels", what);
  if (to_pm)
    pn2::launch(to_point_major_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), c, n, ld, stride, src, dst);
  else
    pn2::launch(to_channel_major_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), c, n, stride, src, dst);
  return check_launch(what);
}

PN2_EXPORT int pn2_to_point_major(int b, int c, int n, int ld, int stride, const float *src, float *dst, void *stream) {
  return transpose_launch(true, b, c, n, ld, stride, src, dst, stream, "pn2_to_point_major");
}
PN2_EXPORT int pn2_to_channel_major(int b, int c, int n, int stride, const float *src, float *dst, void *stream) {
  return transpose_launch(false, b, c, n, c, stride, src, dst, stream, "pn2_to_channel_major");
}

PN2_EXPORT int pn2_bn_reduce_stats(int tiles, int c, int np, double count, const float *stats, double *sums,
                                   void *stream) {
  PN2_REQUIRE(tiles >= 0 && c > 0 && np >= c && stats && sums, "pn2_bn_reduce_stats: bad arguments");
  pn2::launch(bn_reduce_stats_kernel, dim3((c + kStatCh - 1) / kStatCh), dim3(kStatCh * kStatLanes), 0, static_cast<cudaStream_t>(stream), tiles, c, np, count, stats, sums);
  return check_launch("pn2_bn_reduce_stats");
}

PN2_EXPORT int pn2_bn_sync_exchange(int tiles, int c, int np, double count, const float *stats,
                                    const unsigned long long *peers, int rank, int world, unsigned epoch, int slot_doubles,
                                    unsigned *cta_ticket, double *sums, void *stream) {
  PN2_REQUIRE(tiles >= 0 && c > 0 && np >= c && stats && peers && cta_ticket && sums, "pn2_bn_sync_exchange: bad arguments");
  PN2_REQUIRE(world >= 1 && rank >= 0 && rank < world && slot_doubles >= 2 * c + 2 && epoch != 0,
              "pn2_bn_sync_exchange: bad exchange geometry (world=%d rank=%d slot=%d c=%d epoch=%u)", world, rank, slot_doubles,
              c, epoch);
  const int grid = (c + kStatCh - 1) / kStatCh;
  PN2_REQUIRE(grid <= sm_count(), "pn2_bn_sync_exchange: %d CTAs must be co-resident (they wait for each other's peers)", grid);
  pn2::launch(bn_sync_exchange_kernel, dim3(grid), dim3(kStatCh * kStatLanes), 0, static_cast<cudaStream_t>(stream), tiles, c, np,
              count, stats, peers, rank, world, epoch, slot_doubles, cta_ticket, sums);
  return check_launch("pn2_bn_sync_exchange");
}

PN2_EXPORT int pn2_bn_finalize(int training, int tiles, int c, int np, double count, const float *stats,
                               const double *sums, const float *gamma, const float *beta, float *running_mean,
                               float *running_var, long long *num_batches_tracked, float momentum, float eps,
                               float *scale, float *shift, float *mean, float *invstd, void *stream_) {
  PN2_REQUIRE(c > 0 && np >= c && (np % 4) == 0 && scale && shift && mean && invstd, "pn2_bn_finalize: bad arguments");
  PN2_REQUIRE(training ? ((stats && count > 0.0) || sums) : (running_mean && running_var),
              "pn2_bn_finalize: %s", training ? "training needs statistics and a positive count" : "eval needs running statistics");
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  pn2::launch(bn_finalize_kernel, dim3((np + kStatCh - 1) / kStatCh), dim3(kStatCh * kStatLanes), 0, s, training, tiles, c, np, count, stats, sums, gamma, beta,
                                                       running_mean, running_var, num_batches_tracked, momentum, eps,
                                                       scale, shift, mean, invstd);
  if (int rc = check_launch("pn2_bn_finalize")) return rc;
  if (training && num_batches_tracked && running_mean && momentum < 0.f) {
    pn2::launch(bn_bump_counter_kernel, dim3(1), dim3(1), 0, s, num_batches_tracked);
    return check_launch("pn2_bn_finalize(counter)");
  }
  return PN2_OK;
}

PN2_EXPORT int pn2_bn_relu_pool(int groups, int group, int c, int ld, const float *y, const float *scale,
                                const float *shift, float *out_pm, unsigned char *arg, void *stream) {
  PN2_REQUIRE(groups >= 0 && group > 0 && group <= 256 && c > 0 && ld >= c && (ld % 4) == 0, "pn2_bn_relu_pool: bad extents groups=%d group=%d c=%d ld=%d",
              groups, group, c, ld);
  if (groups == 0) return PN2_OK;
  PN2_REQUIRE(y && scale && shift && out_pm, "pn2_bn_relu_pool: null pointer");
  const long long total = static_cast<long long>(groups) * (ld / 4);
  pn2::launch(bn_relu_pool_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      groups, group, ld, y, scale, shift, out_pm, arg);
  return check_launch("pn2_bn_relu_pool");
}

PN2_EXPORT int pn2_pool_bwd_tiles(int groups) { return (groups + kPrepGroups - 1) / kPrepGroups; }

PN2_EXPORT int pn2_pool_bwd_prep(int groups, int group, int c, int ld, float *gz, const float *out_pm,
                                 const unsigned char *arg, const float *y, float *stats, int *tiles, void *stream) {
  PN2_REQUIRE(groups >= 0 && group > 0 && c > 0 && ld >= c && (ld % 4) == 0, "pn2_pool_bwd_prep: bad extents");
  if (tiles) *tiles = pn2_pool_bwd_tiles(groups);
  if (groups == 0) return PN2_OK;
  PN2_REQUIRE(gz && out_pm && y && stats && (arg || group == 1), "pn2_pool_bwd_prep: null pointer");
  pn2::launch(pool_bwd_prep_kernel, dim3(pn2_pool_bwd_tiles(groups)), dim3(128), 0, static_cast<cudaStream_t>(stream), groups, group, ld, gz,
                                                                                                out_pm, arg, y, stats);
  return check_launch("pn2_pool_bwd_prep");
}

PN2_EXPORT int pn2_bn_bwd_finalize(int training, int tiles, int c, int np, double count, const float *stats,
                                   const double *sums, const double *count_dev, const float *gamma, const float *mean,
                                   const float *invstd, float *ca, float *cb, float *cc, float *dgamma, float *dbeta,
                                   void *stream) {
  PN2_REQUIRE(c > 0 && np >= c && (stats || sums) && mean && invstd && ca && cb && cc && (count > 0.0 || count_dev),
              "pn2_bn_bwd_finalize: bad arguments");
  pn2::launch(bn_bwd_finalize_kernel, dim3((np + kStatCh - 1) / kStatCh), dim3(kStatCh * kStatLanes), 0, static_cast<cudaStream_t>(stream), 
      training, tiles, c, np, count, stats, sums, count_dev, gamma, mean, invstd, ca, cb, cc, dgamma, dbeta);
  return check_launch("pn2_bn_bwd_finalize");
}

PN2_EXPORT int pn2_fp_interpolate(int b, int n, int m, int c, int ld_known, const float *unknown, const float *known,
                                  const float *known_pm, float *out, int ldo, int *idx, float *weight, void *stream) {
  PN2_REQUIRE(b >= 0 && n >= 0 && m > 0 && c > 0 && (c % 4) == 0 && ld_known >= c && ldo >= c && (ld_known % 4) == 0 && (ldo % 4) == 0,
              "pn2_fp_interpolate: bad extents b=%d n=%d m=%d c=%d ld_known=%d ldo=%d", b, n, m, c, ld_known, ldo);
  if (b == 0 || n == 0) return PN2_OK;
  PN2_REQUIRE(unknown && known && known_pm && out && idx && weight && b <= 65535, "pn2_fp_interpolate: null pointer");
  static const bool warp_per_point = [] {
    const char *e = getenv("PN2_FP_WARP");
    return e == nullptr || e[0] != '0';
  }();
  if (static_cast<long long>(b) * n <= 16384 && warp_per_point) {
    dim3 grid((n + kFpThreads / 32 - 1) / (kFpThreads / 32), b);
    pn2::launch(fp_interpolate_warp_kernel, dim3(grid), dim3(kFpThreads), 0, static_cast<cudaStream_t>(stream), n, m, c / 4, ld_known, unknown,
                known, known_pm, out, ldo, idx, weight);
  } else if (static_cast<long long>(b) * n <= 16384) {
    dim3 grid((n + 31) / 32, b);
    pn2::launch(fp_interpolate_kernel<32>, dim3(grid), dim3(kFpThreads), 0, static_cast<cudaStream_t>(stream), n, m, c / 4, ld_known, unknown,
                                                                                        known, known_pm, out, ldo, idx, weight);
  } else {
    dim3 grid((n + kFpThreads - 1) / kFpThreads, b);
    pn2::launch(fp_interpolate_kernel<kFpThreads>, dim3(grid), dim3(kFpThreads), 0, static_cast<cudaStream_t>(stream), 
        n, m, c / 4, ld_known, unknown, known, known_pm, out, ldo, idx, weight);
  }
  return check_launch("pn2_fp_interpolate");
}

PN2_EXPORT int pn2_fp_interpolate_grad(int b, int n, int m, int c, const float *dout, int ldo, const int *idx,
                                       const float *weight, float *dknown_pm, int ld_known, void *stream) {
  PN2_REQUIRE(b >= 0 && n >= 0 && m > 0 && c > 0 && (c % 4) == 0 && ld_known >= c && ldo >= c, "pn2_fp_interpolate_grad: bad extents");
  if (b == 0 || n == 0) return PN2_OK;
  PN2_REQUIRE(dout && idx && weight && dknown_pm && b <= 65535, "pn2_fp_interpolate_grad: null pointer");
  dim3 grid((n + 7) / 8, b);
  pn2::launch(fp_interpolate_grad_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), n, m, c / 4, dout, ldo, idx, weight,
                                                                                dknown_pm, ld_known);
  return check_launch("pn2_fp_interpolate_grad");
}
