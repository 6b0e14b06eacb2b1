// Furthest point sampling for sm_100a -- replaces furthest_point_sampling_kernel
// (reference pointnet2/_ext_src/src/sampling_gpu.cu:74-234, one 512-thread block per cloud that
// re-streams xyz and a global temp[] array from L2 on every one of the m-1 rounds and reduces
// through a 9-step __syncthreads tree).
//
// FPS is a chain of m-1 dependent arg-max rounds; its compulsory HBM traffic is 12*N + 4*m bytes,
// so it is latency-bound, not bandwidth-bound.  Design (see DESIGN.md "FPS"):
//   * one thread-block CLUSTER (1..16 CTAs of 512 threads) per cloud; every point lives in
//     registers (x, y, z, running min-distance) for the whole kernel -> xyz is read from HBM once
//     and the reference's temp[] scratch never exists;
//   * per exchange round: register update -> warp arg-max with redux.sync -> the warps' candidates AND a bound on
//     every other point through shared memory (one bar.sync) -> each CTA's two best candidates + its bound pushed to
//     every CTA of the cluster with st.async (DSMEM store that completes a transaction on the receiver's mbarrier)
//     -> every warp resolves as many picks as the bound proves exact (fps_multipick_kernel below, ~8 per round);
//     candidates carry their coordinates so the next round starts without another memory round trip;
//   * buffers are double-buffered on the round parity, which makes one barrier per level enough.
//
// Bit-exactness with the reference (index output):
//   * d = fmaf(dz,dz, fmaf(dx,dx, dy*dy)) with dx = x_k - x_old (pn2_common.cuh sq3), running
//     min with fminf, skip of points with (double)|p|^2 <= 1e-3 (sampling_gpu.cu:105-106);
//   * ties on the max are broken like the reference's tree: thread t = k mod bs scans k ascending
//     with strict '>' and the tree keeps the lower slot, i.e. the winner minimises
//     (bitrev(k mod bs), k div bs) -- encoded below as `rank`; bs = pn2_ref_block_size(n);
//   * if no point is a candidate (all skipped) the reference yields index 0.
#include <limits.h>
#include <math_constants.h>
#include <stdlib.h>

#include <type_traits>

#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kFpsThreads = 512;  // resident kernel: 16 warps = 4 per SM sub-partition (latency hiding for the chain)
constexpr int kFpsMaxCluster = 16;
constexpr int kStreamThreads = 512;  // streaming fallback for clouds beyond register capacity
constexpr int kFpsMaxPts = 16;       // register-resident points per thread (largest instantiation)

template <int NW>
struct alignas(16) FpsSmem {
  uint4 wkey[2][NW];              // per-warp candidate {dist bits, rank, x bits, y bits}
  uint4 ckey[2][kFpsMaxCluster];  // per-CTA candidates received from the whole cluster
  float wz[2][NW];
  float cz[2][kFpsMaxCluster];
  unsigned long long bar[2];  // one mbarrier per parity buffer
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// DSMEM stores that complete `bytes` on the receiving CTA's mbarrier.
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b,
                                            uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::
                   "r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t rbar, uint32_t a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
               "r"(a), "r"(rbar)
               : "memory");
}
constexpr uint32_t kCandBytes = 20;  // v4.b32 + b32 per CTA candidate

// Tie-break order of the reference tree: smaller rank wins.  rank = (bitrev(k mod bs), k div bs).
__device__ __forceinline__ uint32_t rank_of(int k, int bs_log2) {
  const uint32_t low = static_cast<uint32_t>(k) & ((1u << bs_log2) - 1u);
  return (__brev(low) >> 1) | (static_cast<uint32_t>(k) >> bs_log2);
}
__device__ __forceinline__ int index_of_rank(uint32_t r, int bs_log2) {
  const uint32_t qmask = (1u << (31 - bs_log2)) - 1u;
  const uint32_t low = __brev((r & ~qmask) << 1);
  return static_cast<int>(((r & qmask) << bs_log2) | low);
}

struct Cand {
  int d;       // float bits of the min-distance (>= 0) or of the -1 "no candidate" sentinel (< 0)
  uint32_t r;  // tie-break rank
  float x, y, z;
};

// Block- and cluster-level arg-max of the per-warp candidates already stored in S.wkey[p]/wz[p].
// Every thread returns the same winner.
template <int NW>
__device__ __forceinline__ Cand reduce_candidates(FpsSmem<NW> &S, int p, int j, int cs, uint32_t my_cta,
                                                  int warp, int lane) {
  __syncthreads();
  Cand c;
  {  // every warp reduces the NW per-warp candidates with redux (NW <= 32)
    int ed = INT_MIN;
    uint32_t er = 0xffffffffu;
    if (lane < NW) {
      const uint4 e = S.wkey[p][lane];
      ed = static_cast<int>(e.x);
      er = e.y;
    }
    const int gmax = __reduce_max_sync(0xffffffffu, ed);
    const uint32_t gr = __reduce_min_sync(0xffffffffu, ed == gmax ? er : 0xffffffffu);
    const int src = __ffs(__ballot_sync(0xffffffffu, lane < NW && ed == gmax && er == gr)) - 1;
    const uint4 e = S.wkey[p][src];
    c.d = static_cast<int>(e.x);
    c.r = e.y;
    c.x = __uint_as_float(e.z);
    c.y = __uint_as_float(e.w);
    c.z = S.wz[p][src];
  }
  if (cs > 1) {
    if (warp == 0 && lane < cs) {
      const uint32_t rbar = mapa_u32(smem_u32(&S.bar[p]), lane);
      st_async_v4(mapa_u32(smem_u32(&S.ckey[p][my_cta]), lane), rbar, static_cast<uint32_t>(c.d), c.r,
                  __float_as_uint(c.x), __float_as_uint(c.y));
      st_async_b32(mapa_u32(smem_u32(&S.cz[p][my_cta]), lane), rbar, __float_as_uint(c.z));
    }
    mbar_wait(&S.bar[p], ((j - 1) >> 1) & 1);
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&S.bar[p], cs * kCandBytes);  // re-arm for round j+2
    int ed = INT_MIN;
    uint32_t er = 0xffffffffu;
    if (lane < cs) {
      const uint4 e = S.ckey[p][lane];
      ed = static_cast<int>(e.x);
      er = e.y;
    }
    const int gmax = __reduce_max_sync(0xffffffffu, ed);
    const uint32_t gr = __reduce_min_sync(0xffffffffu, ed == gmax ? er : 0xffffffffu);
    const int src = __ffs(__ballot_sync(0xffffffffu, ed == gmax && er == gr)) - 1;
    const uint4 e = S.ckey[p][src];
    c.d = static_cast<int>(e.x);
    c.r = e.y;
    c.x = __uint_as_float(e.z);
    c.y = __uint_as_float(e.w);
    c.z = S.cz[p][src];
  }
  return c;
}

template <int NW>
__device__ __forceinline__ void setup_cluster(FpsSmem<NW> &S, int cs) {
  if (cs > 1) {
    if (threadIdx.x == 0) {
      mbar_init(&S.bar[0], 1);
      mbar_init(&S.bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_arrive_expect_tx(&S.bar[0], cs * kCandBytes);
      mbar_arrive_expect_tx(&S.bar[1], cs * kCandBytes);
    }
    cluster_sync_all();
  }
}

// ---- register-resident kernel ------------------------------------------------------------------------------
// Thread g = cta*512 + tid owns points k = g + i*T, T = cs*512.  T is a multiple of the reference block size
// bs (a power of two <= 512), so all points of one thread share k mod bs and their tie-break rank grows
// with i: inside a thread "strict > while scanning i upwards" IS the reference order.
//
// A round is a chain of dependent steps, so candidates carry the point index k itself and the tie-break rank
// (bit-reversed slot, stripe) is only evaluated when two candidates of a level have EXACTLY the same distance
// -- a warp-uniform, rarely taken branch.  The common path per level is one redux.max, one ballot and one
// popc; no rank arithmetic and no index decoding sit on the critical path.
struct FpsCand {
  int d;  // float bits of the min-distance (>= 0), or of the -1 "no candidate" sentinel (< 0)
  int k;  // point index
  float x, y, z;
};

// arg-max over the lanes' (d, k): returns the winning lane (ties -> smallest reference rank)
__device__ __forceinline__ int warp_argmax_lane(int d, int k, bool active, int bs_log2) {
  const int dm = __reduce_max_sync(0xffffffffu, active ? d : INT_MIN);
  const unsigned eq = __ballot_sync(0xffffffffu, active && d == dm);
  if (__popc(eq) == 1) return __ffs(eq) - 1;
  const uint32_t r = (active && d == dm) ? rank_of(k, bs_log2) : 0xffffffffu;  // exact tie: reference order
  const uint32_t rm = __reduce_min_sync(0xffffffffu, r);
  return __ffs(__ballot_sync(0xffffffffu, r == rm)) - 1;
}

// ---- register-resident kernel, several picks per exchange round ----------------------------------------------
// One block + cluster exchange per sampled point costs ~1300 cycles (the round's first kernel, 0.68 us per sample at 40k
// points; removed).  Most of those exchanges are avoidable without changing a single output bit:
//
//   After a round, let C be a set of candidate points whose current min-distances are known to every warp
//   (with coordinates), and let `bmax` bound the min-distance of every point NOT in C.  Min-distances only
//   decrease, so `bmax` stays a bound while further points are picked.  The next sample is the arg-max over
//   all points; if, after updating the candidates with the picks made so far, the best candidate's distance
//   is strictly above `bmax`, no outsider can beat or tie it, and it IS the reference's next pick (ties
//   between candidates are resolved with the reference rank).  Otherwise the round ends and the next
//   exchange recomputes exact candidates; its first pick is always the exact global arg-max.
//
// Candidates: per warp the exact arg-max (total order: distance, then reference rank) plus the value of the
// best OTHER point of the warp (thread-level runner-up / other lanes' maxima) as that warp's bound.  With a
// cluster, each CTA forwards its two best warp candidates and max(third candidate, warp bounds) to every
// CTA: 32 candidates, one per lane.  Every warp then resolves picks redundantly (identical arithmetic:
// dist2(candidate, pick) is the same expression its owner thread evaluates in the next round), and the
// next round applies all picks of the previous one to the register-resident points.
// On a 40k-point room scan this makes ~8 picks per exchange (2047 rounds -> ~250); on an FPS-ordered input
// (levels 2-4) the interleaved point ownership below keeps it at ~10.
template <int NW>
struct alignas(16) FpsSmem4 {
  uint4 wcand[2][NW];                  // per-warp candidate {key, k, x bits, y bits}
  uint4 ccand[2][2 * kFpsMaxCluster];  // [c]: best of CTA c, [16 + c]: its runner-up
  float wz[2][NW];
  int wbound[2][NW];                   // best key among the warp's points other than its candidate
  float cz[2][2 * kFpsMaxCluster];
  int cbound[2][kFpsMaxCluster];       // bound over the CTA's points other than its two candidates
  unsigned long long bar[2];
};
constexpr uint32_t kCand2Bytes = 44;   // best: v4 + z + bound, runner-up: v4 + z

// order-preserving int key of a min-distance: >= 0 -> float bits, "never a candidate" (-1) -> -1
__device__ __forceinline__ int fps_key(float v) { return v < 0.f ? -1 : __float_as_int(v); }

template <int PTS>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_multipick_kernel(int n, int m, int cs, int bs_log2, const float *__restrict__ xyz,
                     int *__restrict__ idxs, float *__restrict__ new_xyz, const int *__restrict__ identity_flag) {
  pdl_prologue();
  constexpr int NW = kFpsThreads / 32;
  if (identity_flag != nullptr && identity_flag[blockIdx.x / cs] == 0) {
    // the prefix-order check passed for this cloud: the samples are 0 .. m-1 (every CTA of the cluster takes this
    // branch, so no cluster barrier is left waiting)
    const int batch_ = blockIdx.x / cs;
    const int r = static_cast<int>(cs > 1 ? cluster_ctarank() : 0u) * kFpsThreads + threadIdx.x;
    for (int j = r; j < m; j += cs * kFpsThreads) {
      idxs[static_cast<size_t>(batch_) * m + j] = j;
      if (new_xyz) {
        const float *p = xyz + (static_cast<size_t>(batch_) * n + j) * 3;
        float *o = new_xyz + (static_cast<size_t>(batch_) * m + j) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
      }
    }
    return;
  }
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FpsSmem4<NW> &S = *reinterpret_cast<FpsSmem4<NW> *>(smem_raw);
  float *sx = reinterpret_cast<float *>(smem_raw + sizeof(FpsSmem4<NW>));
  float *sy = sx + PTS * kFpsThreads;
  float *sz = sy + PTS * kFpsThreads;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t my_cta = cs > 1 ? cluster_ctarank() : 0u;
  const int batch = blockIdx.x / cs;
  xyz += static_cast<size_t>(batch) * n * 3;
  idxs += static_cast<size_t>(batch) * m;
  if (new_xyz) new_xyz += static_cast<size_t>(batch) * m * 3;
  const int T = cs * kFpsThreads;
  // Point ownership: consecutive indices go to different CTAs, then to different warps, so that an input whose
  // order is itself an FPS order (every level after the first samples the previous level's centres) spreads
  // its next picks k, k+1, ... over all candidate slots instead of one warp.  k = g + i*T as before, so a
  // thread's points share k mod bs and scan order inside a thread is the reference tie-break order.
  const int g = (lane * NW + warp) * cs + static_cast<int>(my_cta);
  const bool writer = my_cta == 0 && tid == 0;

  float px[PTS], py[PTS], pz[PTS], pt[PTS];
#pragma unroll
  for (int i = 0; i < PTS; ++i) {
    const int k = g + i * T;
    float x = 0.f, y = 0.f, z = 0.f;
    bool valid = false;
    if (k < n) {
      x = xyz[k * 3 + 0];
      y = xyz[k * 3 + 1];
      z = xyz[k * 3 + 2];
      valid = !(static_cast<double>(sq3(x, y, z)) <= 1e-3);  // sampling_gpu.cu:105-106
    }
    px[i] = x; py[i] = y; pz[i] = z;
    pt[i] = valid ? 1e10f : -1.0f;  // -1: never a candidate, fminf keeps it at -1
    sx[i * kFpsThreads + tid] = x;
    sy[i * kFpsThreads + tid] = y;
    sz[i * kFpsThreads + tid] = z;
  }
  const float x0 = xyz[0], y0 = xyz[1], z0 = xyz[2];
  if (writer) {
    idxs[0] = 0;
    if (new_xyz) { new_xyz[0] = x0; new_xyz[1] = y0; new_xyz[2] = z0; }
  }
  if (cs > 1) {
    if (tid == 0) {
      mbar_init(&S.bar[0], 1);
      mbar_init(&S.bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_arrive_expect_tx(&S.bar[0], cs * kCand2Bytes);
      mbar_arrive_expect_tx(&S.bar[1], cs * kCand2Bytes);
    }
    cluster_sync_all();
  }
#pragma unroll
  for (int i = 0; i < PTS; ++i) pt[i] = fminf(dist2(px[i], py[i], pz[i], x0, y0, z0), pt[i]);  // sample 0

  int j = 1;  // next output slot
  for (int round = 0; j < m; ++round) {
    const int p = round & 1;
    // ---- the thread's best / runner-up (every pick so far is already applied to pt[]) ----
    float best = -2.0f, second = -2.0f;
    int ib = 0;
#pragma unroll
    for (int i = 0; i < PTS; ++i) {
      const float t = pt[i];
      if (t > best) { second = best; best = t; ib = i; }  // strict '>': lowest i (lowest rank in the thread) wins ties
      else second = fmaxf(second, t);
    }
    // ---- warp level: exact arg-max + bound over the warp's other points ----
    const int kb = fps_key(best), ks = fps_key(second);
    const int kmine = g + ib * T;
    {
      const int src = warp_argmax_lane(kb, kmine, true, bs_log2);
      const int wb = __reduce_max_sync(FULL, lane == src ? ks : kb);
      if (lane == src) {
        const int slot = ib * kFpsThreads + tid;
        S.wcand[p][warp] = make_uint4(static_cast<uint32_t>(kb), static_cast<uint32_t>(kmine), __float_as_uint(sx[slot]),
                                      __float_as_uint(sy[slot]));
        S.wz[p][warp] = sz[slot];
        S.wbound[p][warp] = wb;
      }
    }
    __syncthreads();
    // ---- candidates of this round, one per lane ----
    int cd = INT_MIN, ck = 0, bmax;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    bool valid;
    if (cs == 1) {
      valid = lane < NW;
      int wb = INT_MIN;
      if (valid) {
        const uint4 e = S.wcand[p][lane];
        cd = static_cast<int>(e.x); ck = static_cast<int>(e.y);
        cx = __uint_as_float(e.z); cy = __uint_as_float(e.w); cz = S.wz[p][lane];
        wb = S.wbound[p][lane];
      }
      bmax = __reduce_max_sync(FULL, wb);
    } else {
      if (warp < 2) {  // warp 0 forwards the CTA's best warp candidate (+ bound), warp 1 the runner-up
        int d = INT_MIN, k = 0, wb = INT_MIN;
        if (lane < NW) {
          const uint4 e = S.wcand[p][lane];
          d = static_cast<int>(e.x); k = static_cast<int>(e.y);
          wb = S.wbound[p][lane];
        }
        const int src1 = warp_argmax_lane(d, k, lane < NW, bs_log2);
        const bool rest = lane < NW && lane != src1;
        const int m2 = __reduce_max_sync(FULL, rest ? d : INT_MIN);
        const int src2 = __ffs(__ballot_sync(FULL, rest && d == m2)) - 1;
        if (warp == 0) {
          const int third = __reduce_max_sync(FULL, (rest && lane != src2) ? d : INT_MIN);
          const int cb = max(third, __reduce_max_sync(FULL, wb));
          if (lane < cs) {
            const uint4 w = S.wcand[p][src1];
            const uint32_t rbar = mapa_u32(smem_u32(&S.bar[p]), lane);
            st_async_v4(mapa_u32(smem_u32(&S.ccand[p][my_cta]), lane), rbar, w.x, w.y, w.z, w.w);
            st_async_b32(mapa_u32(smem_u32(&S.cz[p][my_cta]), lane), rbar, __float_as_uint(S.wz[p][src1]));
            st_async_b32(mapa_u32(smem_u32(&S.cbound[p][my_cta]), lane), rbar, static_cast<uint32_t>(cb));
          }
        } else if (lane < cs) {
          const uint4 w = S.wcand[p][src2];
          const uint32_t rbar = mapa_u32(smem_u32(&S.bar[p]), lane);
          st_async_v4(mapa_u32(smem_u32(&S.ccand[p][kFpsMaxCluster + my_cta]), lane), rbar, w.x, w.y, w.z, w.w);
          st_async_b32(mapa_u32(smem_u32(&S.cz[p][kFpsMaxCluster + my_cta]), lane), rbar, __float_as_uint(S.wz[p][src2]));
        }
      }
      mbar_wait(&S.bar[p], (round >> 1) & 1);
      if (tid == 0) mbar_arrive_expect_tx(&S.bar[p], cs * kCand2Bytes);  // re-arm for round + 2
      valid = (lane & (kFpsMaxCluster - 1)) < cs;
      if (valid) {
        const uint4 e = S.ccand[p][lane];
        cd = static_cast<int>(e.x); ck = static_cast<int>(e.y);
        cx = __uint_as_float(e.z); cy = __uint_as_float(e.w); cz = S.cz[p][lane];
      }
      bmax = __reduce_max_sync(FULL, lane < cs ? S.cbound[p][lane] : INT_MIN);
    }
    // ---- resolve as many picks as the bound allows (every warp, identical result).  Each accepted pick is
    // applied to the thread's resident points right away: that work is independent of the candidate chain, so
    // the warps of an SM overlap one another's redux / shuffle latency with it. ----
    int npick = 0;
    while (true) {
      const int src = warp_argmax_lane(cd, ck, valid, bs_log2);
      const int dbest = __shfl_sync(FULL, cd, src);
      if (npick == 0) {
        if (dbest < 0) {  // no point is a candidate: the reference yields index 0 from here on
          if (writer)
            for (int t = j; t < m; ++t) {
              idxs[t] = 0;
              if (new_xyz) { new_xyz[t * 3 + 0] = x0; new_xyz[t * 3 + 1] = y0; new_xyz[t * 3 + 2] = z0; }
            }
          npick = m - j;
          break;
        }
      } else if (dbest <= bmax) {
        break;  // an outsider may beat or tie it: exact candidates are needed
      }
      const int kq = __shfl_sync(FULL, ck, src);
      const float xq = __shfl_sync(FULL, cx, src), yq = __shfl_sync(FULL, cy, src), zq = __shfl_sync(FULL, cz, src);
      if (writer) {
        idxs[j + npick] = kq;
        if (new_xyz) { new_xyz[(j + npick) * 3 + 0] = xq; new_xyz[(j + npick) * 3 + 1] = yq; new_xyz[(j + npick) * 3 + 2] = zq; }
      }
      ++npick;
      if (valid && cd >= 0) cd = __float_as_int(fminf(dist2(cx, cy, cz, xq, yq, zq), __int_as_float(cd)));
#pragma unroll
      for (int i = 0; i < PTS; ++i) pt[i] = fminf(dist2(px[i], py[i], pz[i], xq, yq, zq), pt[i]);
      if (j + npick >= m) break;
    }
    j += npick;
  }
  if (cs > 1) cluster_sync_all();  // keep every CTA's shared memory alive until all DSMEM stores landed
}

// ---- streaming fallback (n beyond register capacity): same exchange, points from L2 ---------------
__global__ void __launch_bounds__(kStreamThreads, 1)
fps_streaming_kernel(int n, int m, int cs, int bs_log2, const float *__restrict__ xyz,
                     float *__restrict__ temp, int *__restrict__ idxs, float *__restrict__ new_xyz) {
  pdl_prologue();
  constexpr int NW = kStreamThreads / 32;
  __shared__ FpsSmem<NW> S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t my_cta = cs > 1 ? cluster_ctarank() : 0u;
  const int batch = blockIdx.x / cs;
  xyz += static_cast<size_t>(batch) * n * 3;
  temp += static_cast<size_t>(batch) * n;
  idxs += static_cast<size_t>(batch) * m;
  if (new_xyz) new_xyz += static_cast<size_t>(batch) * m * 3;
  const int T = cs * kStreamThreads;
  const int g = static_cast<int>(my_cta) * kStreamThreads + tid;

  for (int k = g; k < n; k += T) {
    const bool valid = !(static_cast<double>(sq3(xyz[k * 3], xyz[k * 3 + 1], xyz[k * 3 + 2])) <= 1e-3);
    temp[k] = valid ? 1e10f : -1.0f;
  }
  const float x0 = xyz[0], y0 = xyz[1], z0 = xyz[2];
  float x1 = x0, y1 = y0, z1 = z0;
  if (g == 0) {
    idxs[0] = 0;
    if (new_xyz) { new_xyz[0] = x0; new_xyz[1] = y0; new_xyz[2] = z0; }
  }
  setup_cluster(S, cs);

  for (int j = 1; j < m; ++j) {
    const int p = j & 1;
    float best = -2.0f;
    uint32_t brank = 0xffffffffu;
    for (int k = g; k < n; k += T) {
      const float d = dist2(xyz[k * 3], xyz[k * 3 + 1], xyz[k * 3 + 2], x1, y1, z1);
      const float t = fminf(d, temp[k]);
      temp[k] = t;
      if (t >= best) {
        const uint32_t r = rank_of(k, bs_log2);
        if (t > best || r < brank) { best = t; brank = r; }
      }
    }
    const int wmax = __reduce_max_sync(0xffffffffu, __float_as_int(best));
    const uint32_t myrank = (__float_as_int(best) == wmax) ? brank : 0xffffffffu;
    const uint32_t wrank = __reduce_min_sync(0xffffffffu, myrank);
    const uint32_t win = __ballot_sync(0xffffffffu, __float_as_int(best) == wmax && myrank == wrank);
    if (lane == __ffs(win) - 1) {
      float wx = 0.f, wy = 0.f, wzv = 0.f;
      if (wmax >= 0) {
        const int k = index_of_rank(wrank, bs_log2);
        wx = xyz[k * 3]; wy = xyz[k * 3 + 1]; wzv = xyz[k * 3 + 2];
      }
      S.wkey[p][warp] = make_uint4(static_cast<uint32_t>(wmax), wrank, __float_as_uint(wx), __float_as_uint(wy));
      S.wz[p][warp] = wzv;
    }
    const Cand c = reduce_candidates(S, p, j, cs, my_cta, warp, lane);
    int k = 0;
    if (c.d < 0) { x1 = x0; y1 = y0; z1 = z0; }
    else { k = index_of_rank(c.r, bs_log2); x1 = c.x; y1 = c.y; z1 = c.z; }
    if (g == 0) {
      idxs[j] = k;
      if (new_xyz) { new_xyz[j * 3 + 0] = x1; new_xyz[j * 3 + 1] = y1; new_xyz[j * 3 + 2] = z1; }
    }
  }
  if (cs > 1) cluster_sync_all();
}

// ---- prefix-order shortcut --------------------------------------------------------------------------------
// Every set-abstraction level after the first samples the PREVIOUS level's centres, i.e. a cloud that already is in
// FPS order; its own FPS then returns 0, 1, 2, ..., m-1 (the arg-max over the whole cloud at step j was point j, and
// it still is over the subset) unless an exact tie is broken differently by the subset's index ranks.  Verifying
// that claim is embarrassingly parallel, where producing it is a chain of m-1 dependent arg-max rounds:
//   D_j(k) = min_{i<j} dist2(p_k, p_i)   (fminf chain from 1e10; skipped points stay at -1)
//   identity holds at step j  <=>  point j is a candidate and no other candidate k has D_j(k) > D_j(j), or
//                                  D_j(k) == D_j(j) with a smaller reference rank.
// fps_prefix_diag_kernel computes diag[j] = D_j(j) (thread per j), fps_prefix_check_kernel walks j = 1..m-1 for every
// point k (thread per k, the prefix coordinates and diag[] in shared memory) and raises flag[cloud] on the first
// violation; the FPS kernel launched afterwards writes the identity and returns at once when the flag is still 0,
// otherwise it computes the samples as usual.  2048 -> 1024: ~20 us instead of 350 us; an unordered cloud fails at
// j = 1 and pays ~10 us.  Only attempted for n <= kPrefixMaxN.
constexpr int kPrefixMaxN = 8192;
constexpr int kPrefixThreads = 128;
constexpr int kPrefixPad = 8;  // steps per unrolled trip of the check kernel = padding rows behind the prefix

__global__ void __launch_bounds__(kPrefixThreads)
fps_prefix_diag_kernel(int n, int m, const float *__restrict__ xyz, float *__restrict__ diag, int *__restrict__ flag) {
  pdl_prologue();
  extern __shared__ __align__(16) float sp[];  // [m][3] prefix coordinates
  const int batch = blockIdx.y;
  xyz += static_cast<size_t>(batch) * n * 3;
  diag += static_cast<size_t>(batch) * m;
  if (blockIdx.x == 0 && threadIdx.x == 0) flag[batch] = 0;  // raised by the check kernel (stream order)
  for (int i = threadIdx.x; i < m * 3; i += kPrefixThreads) sp[i] = xyz[i];
  __syncthreads();
  const int j = blockIdx.x * kPrefixThreads + threadIdx.x;
  if (j >= m) return;
  const float x = sp[j * 3 + 0], y = sp[j * 3 + 1], z = sp[j * 3 + 2];
  float d = (static_cast<double>(sq3(x, y, z)) <= 1e-3) ? -1.0f : 1e10f;  // sampling_gpu.cu:105-106
#pragma unroll 8
  for (int i = 0; i < j; ++i) d = fminf(dist2(x, y, z, sp[i * 3 + 0], sp[i * 3 + 1], sp[i * 3 + 2]), d);
  diag[j] = d;
}

__global__ void __launch_bounds__(kPrefixThreads)
fps_prefix_check_kernel(int n, int m, int bs_log2, const float *__restrict__ xyz, const float *__restrict__ diag,
                        int *__restrict__ flag) {
  pdl_prologue();
  extern __shared__ __align__(16) float sp[];  // [m][4]: x, y, z of point j, diag[j]
  const int batch = blockIdx.y;
  xyz += static_cast<size_t>(batch) * n * 3;
  diag += static_cast<size_t>(batch) * m;
  for (int i = threadIdx.x; i < m + kPrefixPad; i += kPrefixThreads) {
    const bool in = i < m;  // padding rows: a step nobody can lose (diag = +inf), so the loop below needs no tail
    sp[i * 4 + 0] = in ? xyz[i * 3 + 0] : 0.f;
    sp[i * 4 + 1] = in ? xyz[i * 3 + 1] : 0.f;
    sp[i * 4 + 2] = in ? xyz[i * 3 + 2] : 0.f;
    sp[i * 4 + 3] = in ? diag[i] : CUDART_INF_F;
  }
  __syncthreads();
  const int k = blockIdx.x * kPrefixThreads + threadIdx.x;
  if (k >= n) return;
  const float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
  float d = (static_cast<double>(sq3(x, y, z)) <= 1e-3) ? -1.0f : 1e10f;
  const uint32_t rk = rank_of(k, bs_log2);
  bool bad = false;
  for (int j0 = 1; j0 < m; j0 += kPrefixPad) {  // kPrefixPad steps per trip: their shared-memory loads overlap
#pragma unroll
    for (int u = 0; u < kPrefixPad; ++u) {
      const int j = j0 + u;  // j >= m only touches the padding rows (after the last real step d is not used again)
      const float4 q = *reinterpret_cast<const float4 *>(sp + (j - 1) * 4);
      d = fminf(dist2(x, y, z, q.x, q.y, q.z), d);
      const float dj = sp[j * 4 + 3];
      if (k == j) bad |= dj < 0.f;                                              // a skipped point is never sampled
      else if (d >= 0.f) bad |= d > dj || (d == dj && rk < rank_of(j, bs_log2));  // somebody else wins step j
    }
    if ((((j0 - 1) / kPrefixPad) & 7) == 0 && (bad || *reinterpret_cast<volatile int *>(flag + batch) != 0)) break;
  }
  if (bad) flag[batch] = 1;
}

// Function attributes are per device: true if `slot` already names the current device, else records it.
inline bool configured_on(int &slot) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (slot == dev) return true;
  slot = dev;
  return false;
}

template <typename K>
int launch_cluster(K kernel, int grid, int block, size_t smem, int cs, cudaStream_t stream, void **args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  note_launch();
  cudaError_t e = cudaLaunchKernelExC(&cfg, reinterpret_cast<const void *>(kernel), args);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("pn2_furthest_point_sampling: launch (cluster %d) failed: %s", cs, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return PN2_OK;
}

template <int PTS>
int launch_multipick(int b, int n, int m, int cs, int bs_log2, const float *xyz, int *idxs, float *new_xyz,
                     const int *identity_flag, cudaStream_t stream) {
  auto kernel = fps_multipick_kernel<PTS>;
  const size_t smem = sizeof(FpsSmem4<kFpsThreads / 32>) + size_t(3) * PTS * kFpsThreads * sizeof(float);
  static thread_local int configured_dev = -1;
  if (!configured_on(configured_dev)) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  }
  void *args[] = {&n, &m, &cs, &bs_log2, &xyz, &idxs, &new_xyz, &identity_flag};
  return launch_cluster(kernel, b * cs, kFpsThreads, smem, cs, stream, args);
}

template <int PTS>
int launch_resident(int b, int n, int m, int cs, int bs_log2, const float *xyz, int *idxs, float *new_xyz,
                    const int *identity_flag, cudaStream_t stream) {
  return launch_multipick<PTS>(b, n, m, cs, bs_log2, xyz, idxs, new_xyz, identity_flag, stream);
}

constexpr int kPtsOptions[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};

int pick_pts(int need) {
  for (int v : kPtsOptions)
    if (v >= need) return v;
  return 0;
}

}  // namespace
}  // namespace pn2

PN2_EXPORT int pn2_fps_resident_capacity(void) { return pn2::kFpsMaxCluster * pn2::kFpsThreads * pn2::kFpsMaxPts; }

PN2_EXPORT int pn2_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idxs,
                                           float *new_xyz, void *stream_) {
  using namespace pn2;
  PN2_REQUIRE(b >= 0 && n > 0, "pn2_furthest_point_sampling: need b >= 0 and n > 0 (b=%d n=%d)", b, n);
  if (b == 0 || m <= 0) return PN2_OK;  // sampling_gpu.cu:78
  PN2_REQUIRE(xyz && idxs, "pn2_furthest_point_sampling: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int bs = pn2_ref_block_size(n);
  int bs_log2 = 0;
  while ((1 << (bs_log2 + 1)) <= bs) ++bs_log2;

  // Cluster size: smallest that keeps <= 5 points per thread (the per-round critical path is the register
  // update); halved while the batch would not fit the chip in one wave and registers allow.
  int cs = 1;
  while (cs < kFpsMaxCluster && (n + cs * kFpsThreads - 1) / (cs * kFpsThreads) > 5) cs *= 2;
  const int sms = sm_count();
  while (cs > 1 && b * cs > sms && (n + (cs / 2) * kFpsThreads - 1) / ((cs / 2) * kFpsThreads) <= kFpsMaxPts) cs /= 2;
  const int need = (n + cs * kFpsThreads - 1) / (cs * kFpsThreads);
  const int pts = pick_pts(need);
  // prefix-order shortcut (needs the scratch: diag [b][m] floats, then one flag per cloud)
  const int *identity_flag = nullptr;
  static const bool prefix_on = [] {
    const char *e = getenv("PN2_FPS_PREFIX");
    return e == nullptr || e[0] != '0';
  }();
  if (prefix_on && pts != 0 && temp != nullptr && n <= kPrefixMaxN && m >= 2 && m < n &&
      b <= 65535) {
    float *diag = temp;
    int *flag = reinterpret_cast<int *>(temp + static_cast<size_t>(b) * m);
    static thread_local int configured_dev = -1;
    if (!configured_on(configured_dev)) {
      cudaFuncSetAttribute(fps_prefix_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPrefixMaxN * 12);
      cudaFuncSetAttribute(fps_prefix_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (kPrefixMaxN + kPrefixPad) * 16);
    }
    pn2::launch(fps_prefix_diag_kernel, dim3(dim3((m + kPrefixThreads - 1) / kPrefixThreads, b)), dim3(kPrefixThreads), static_cast<size_t>(m) * 12, stream, n, m, xyz, diag, flag);
    pn2::launch(fps_prefix_check_kernel, dim3(dim3((n + kPrefixThreads - 1) / kPrefixThreads, b)), dim3(kPrefixThreads), static_cast<size_t>(m + kPrefixPad) * 16, stream, n, m, bs_log2, xyz, diag, flag);
    if (int rc = check_launch("pn2_furthest_point_sampling(prefix check)")) return rc;
    identity_flag = flag;
  }
  if (pts == 0) {
    PN2_REQUIRE(temp != nullptr, "pn2_furthest_point_sampling: n=%d exceeds the resident capacity %d, temp scratch required",
                n, pn2_fps_resident_capacity());
    cs = kFpsMaxCluster;
    static thread_local int configured_dev = -1;
    if (!configured_on(configured_dev))
      cudaFuncSetAttribute(fps_streaming_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    void *args[] = {&n, &m, &cs, &bs_log2, &xyz, &temp, &idxs, &new_xyz};
    return launch_cluster(fps_streaming_kernel, b * cs, kStreamThreads, 0, cs, stream, args);
  }
  switch (pts) {
#define PN2_FPS_CASE(P) \
  case P:               \
    return launch_resident<P>(b, n, m, cs, bs_log2, xyz, idxs, new_xyz, identity_flag, stream);
    PN2_FPS_CASE(1)
    PN2_FPS_CASE(2)
    PN2_FPS_CASE(3)
    PN2_FPS_CASE(4)
    PN2_FPS_CASE(5)
    PN2_FPS_CASE(6)
    PN2_FPS_CASE(8)
    PN2_FPS_CASE(10)
    PN2_FPS_CASE(12)
    PN2_FPS_CASE(16)
#undef PN2_FPS_CASE
  }
  set_error("pn2_furthest_point_sampling: internal dispatch error (pts=%d)", pts);
  return PN2_ERR_UNSUPPORTED;
}
