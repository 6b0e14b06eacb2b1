/* CPU emulation of fps_multipick_kernel (omni-pq_b200/csrc/fps.cu): the same thread / warp / CTA mapping, the
 * same candidate + bound construction and the same pick-resolution loop, in plain C.  Test infrastructure only
 * (tests/test_fps_multipick_cpu.py compares it with the oracle's one-pick-per-round restatement of the
 * reference kernel): it pins the exactness argument of the multi-pick scheme -- ties, skipped points, ragged
 * sizes, every cluster size -- on machines without a GPU.  Build: gcc -O2 -ffp-contract=off -shared -fPIC. */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define THREADS 512
#define NW 16
#define MAXCS 16

static float sq3(float a, float b, float c) { return fmaf(c, c, fmaf(a, a, b * b)); }
static float dist2(float ax, float ay, float az, float bx, float by, float bz) { return sq3(ax - bx, ay - by, az - bz); }
static int fkey(float v) { int k; if (v < 0.f) return -1; memcpy(&k, &v, 4); return k; }
static float keyf(int k) { float v; memcpy(&v, &k, 4); return v; }
static uint32_t brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
static uint32_t rank_of(int k, int bs_log2) {
  const uint32_t low = (uint32_t)k & ((1u << bs_log2) - 1u);
  return (brev(low) >> 1) | ((uint32_t)k >> bs_log2);
}
/* warp_argmax_lane: max key over active lanes, exact ties -> smallest reference rank */
static int argmax_lane(const int *d, const int *k, const int *active, int bs_log2) {
  int dm = INT_MIN, cnt = 0, first = -1;
  for (int l = 0; l < 32; ++l) if (active[l] && d[l] > dm) dm = d[l];
  for (int l = 0; l < 32; ++l) if (active[l] && d[l] == dm) { if (first < 0) first = l; ++cnt; }
  if (cnt == 1) return first;
  uint32_t rm = 0xffffffffu;
  for (int l = 0; l < 32; ++l) { const uint32_t r = (active[l] && d[l] == dm) ? rank_of(k[l], bs_log2) : 0xffffffffu; if (r < rm) rm = r; }
  for (int l = 0; l < 32; ++l) { const uint32_t r = (active[l] && d[l] == dm) ? rank_of(k[l], bs_log2) : 0xffffffffu; if (r == rm) return l; }
  return 0;
}

typedef struct { int d, k; float x, y, z; } Cand;

/* xyz: n x 3, idx: m ints.  cs = cluster size (1..16), bs = reference block size (power of two <= 512).
 * Returns the number of exchange rounds, or -1 on bad arguments. */
int fps_multipick_emul(int n, int m, int cs, int bs, const float *xyz, int *idx) {
  if (n <= 0 || m <= 0 || cs < 1 || cs > MAXCS) return -1;
  int bs_log2 = 0;
  while ((1 << (bs_log2 + 1)) <= bs) ++bs_log2;
  const int T = cs * THREADS;
  const int PTS = (n + T - 1) / T;
  float *pt = (float *)malloc(sizeof(float) * (size_t)T * PTS);
  for (int g = 0; g < T; ++g)
    for (int i = 0; i < PTS; ++i) {
      const int k = g + i * T;
      int valid = 0;
      if (k < n) valid = !((double)sq3(xyz[k * 3], xyz[k * 3 + 1], xyz[k * 3 + 2]) <= 1e-3);
      pt[(size_t)g * PTS + i] = valid ? 1e10f : -1.0f;
    }
  const float x0 = xyz[0], y0 = xyz[1], z0 = xyz[2];
  idx[0] = 0;
  int j = 1, rounds = 0;
#define APPLY_PICK(XQ, YQ, ZQ)                                                                          \
  for (int g_ = 0; g_ < T; ++g_)                                                                        \
    for (int i_ = 0; i_ < PTS; ++i_) {                                                                  \
      const int k_ = g_ + i_ * T;                                                                       \
      const float px_ = k_ < n ? xyz[k_ * 3] : 0.f, py_ = k_ < n ? xyz[k_ * 3 + 1] : 0.f,              \
                  pz_ = k_ < n ? xyz[k_ * 3 + 2] : 0.f;                                                 \
      float *q_ = pt + (size_t)g_ * PTS + i_;                                                           \
      *q_ = fminf(dist2(px_, py_, pz_, (XQ), (YQ), (ZQ)), *q_);                                         \
    }
  APPLY_PICK(x0, y0, z0) /* sample 0 */
  Cand wc[MAXCS][NW];
  int wb[MAXCS][NW];
  while (j < m) {
    ++rounds;
    for (int cta = 0; cta < cs; ++cta)
      for (int w = 0; w < NW; ++w) {
        int kb[32], ks[32], km[32], act[32];
        for (int l = 0; l < 32; ++l) {
          const int g = (l * NW + w) * cs + cta; /* interleaved ownership, as in the kernel */
          float *p = pt + (size_t)g * PTS;
          float best = -2.0f, second = -2.0f;
          int ib = 0;
          for (int i = 0; i < PTS; ++i) {
            const float t = p[i];
            if (t > best) { second = best; best = t; ib = i; }
            else second = fmaxf(second, t);
          }
          kb[l] = fkey(best); ks[l] = fkey(second); km[l] = g + ib * T; act[l] = 1;
        }
        const int src = argmax_lane(kb, km, act, bs_log2);
        int bound = INT_MIN;
        for (int l = 0; l < 32; ++l) { const int v = (l == src) ? ks[l] : kb[l]; if (v > bound) bound = v; }
        const int k = km[src];
        wc[cta][w].d = kb[src]; wc[cta][w].k = k;
        wc[cta][w].x = k < n ? xyz[k * 3] : 0.f; wc[cta][w].y = k < n ? xyz[k * 3 + 1] : 0.f; wc[cta][w].z = k < n ? xyz[k * 3 + 2] : 0.f;
        wb[cta][w] = bound;
      }
    /* candidates of the round, one per lane */
    Cand c[32];
    int cd[32], ck[32], valid[32], bmax = INT_MIN;
    memset(c, 0, sizeof(c));
    for (int l = 0; l < 32; ++l) { valid[l] = 0; cd[l] = INT_MIN; ck[l] = 0; }
    if (cs == 1) {
      for (int l = 0; l < NW; ++l) { c[l] = wc[0][l]; valid[l] = 1; if (wb[0][l] > bmax) bmax = wb[0][l]; }
    } else {
      for (int cta = 0; cta < cs; ++cta) {
        int d[32], k[32], act[32];
        for (int l = 0; l < 32; ++l) { act[l] = l < NW; d[l] = l < NW ? wc[cta][l].d : INT_MIN; k[l] = l < NW ? wc[cta][l].k : 0; }
        const int src1 = argmax_lane(d, k, act, bs_log2);
        int m2 = INT_MIN, src2 = -1, third = INT_MIN, mwb = INT_MIN;
        for (int l = 0; l < NW; ++l) if (l != src1 && d[l] > m2) m2 = d[l];
        for (int l = 0; l < NW; ++l) if (l != src1 && d[l] == m2) { src2 = l; break; }
        for (int l = 0; l < NW; ++l) if (l != src1 && l != src2 && d[l] > third) third = d[l];
        for (int l = 0; l < NW; ++l) if (wb[cta][l] > mwb) mwb = wb[cta][l];
        const int cb = third > mwb ? third : mwb;
        c[cta] = wc[cta][src1]; valid[cta] = 1;
        c[MAXCS + cta] = wc[cta][src2]; valid[MAXCS + cta] = 1;
        if (cb > bmax) bmax = cb;
      }
    }
    for (int l = 0; l < 32; ++l) if (valid[l]) { cd[l] = c[l].d; ck[l] = c[l].k; }
    /* resolution */
    int npick = 0;
    for (;;) {
      const int src = argmax_lane(cd, ck, valid, bs_log2);
      const int dbest = cd[src];
      if (npick == 0) {
        if (dbest < 0) { for (int t = j; t < m; ++t) idx[t] = 0; npick = m - j; break; }
      } else if (dbest <= bmax) break;
      const float xq = c[src].x, yq = c[src].y, zq = c[src].z;
      idx[j + npick] = ck[src];
      ++npick;
      for (int l = 0; l < 32; ++l)
        if (valid[l] && cd[l] >= 0) cd[l] = fkey(fminf(dist2(c[l].x, c[l].y, c[l].z, xq, yq, zq), keyf(cd[l])));
      APPLY_PICK(xq, yq, zq) /* the kernel applies an accepted pick to the resident points at once */
      if (j + npick >= m) break;
    }
    j += npick;
  }
  free(pt);
  return rounds;
}

/* CPU restatement of the prefix-order check (fps_prefix_diag_kernel + fps_prefix_check_kernel): returns 0 when the
 * check passes (the kernel then answers 0..m-1), 1 when it raises the flag. */
int fps_prefix_verify_emul(int n, int m, int bs, const float *xyz) {
  int bs_log2 = 0;
  while ((1 << (bs_log2 + 1)) <= bs) ++bs_log2;
  float *diag = (float *)malloc(sizeof(float) * (size_t)m);
  for (int j = 0; j < m; ++j) {
    const float x = xyz[j * 3], y = xyz[j * 3 + 1], z = xyz[j * 3 + 2];
    float d = ((double)sq3(x, y, z) <= 1e-3) ? -1.0f : 1e10f;
    for (int i = 0; i < j; ++i) d = fminf(dist2(x, y, z, xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]), d);
    diag[j] = d;
  }
  int flag = 0;
  for (int k = 0; k < n && !flag; ++k) {
    const float x = xyz[k * 3], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
    float d = ((double)sq3(x, y, z) <= 1e-3) ? -1.0f : 1e10f;
    const uint32_t rk = rank_of(k, bs_log2);
    for (int j = 1; j < m; ++j) {
      d = fminf(dist2(x, y, z, xyz[(j - 1) * 3], xyz[(j - 1) * 3 + 1], xyz[(j - 1) * 3 + 2]), d);
      const float dj = diag[j];
      if (k == j) { if (dj < 0.f) { flag = 1; break; } }
      else if (d >= 0.f && (d > dj || (d == dj && rk < rank_of(j, bs_log2)))) { flag = 1; break; }
    }
  }
  free(diag);
  return flag;
}
