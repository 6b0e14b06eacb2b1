/*
 * pn2_oracle.c -- CPU restatement of the reference PointNet++ ops (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for the B200 kernels in omni-pq_b200/csrc.  It is NOT part of
 * the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  It restates, loop for loop, the nine CUDA kernels of the reference
 * (AIR-DISCOVER/Omni-PQ, the .cu files under pointnet2/_ext_src/src) in plain C.  The reference kernels only run
 * on a GPU ("CPU not supported", sampling.cpp:41,67,89), so the "threads" of each kernel are
 * emulated sequentially here in an order that produces the same result.
 *
 * Floating point: the integer outputs (FPS order, ball-query idx, 3-NN idx) depend on the exact
 * fp32 contraction nvcc emits for a*a + b*b + c*c, which is  fmaf(c,c, fmaf(a,a, b*b))
 * (read from the SASS of the sm_100a build).  Every such expression is
 * spelt with fmaf()/explicit float temporaries below and the file is compiled with
 * -ffp-contract=off so the compiler can neither add nor remove a fused multiply-add.
 *
 * Parity pinning: the reference ships no golden vectors for these ops, so
 * this oracle is pinned against the reference's own kernels compiled from /root/reference into
 * oracle/_ref (oracle/build_ref.py) and run on the B200 box (tests/test_gpu_ops.py::test_index_ops_match_reference_kernels_live), and against
 * the frozen fixtures in tests/golden/ that both agree on.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PN2O_API __attribute__((visibility("default")))

/* a*a + b*b + c*c exactly as the reference's SASS evaluates it (FMUL b*b; FFMA a,a; FFMA c,c). */
static inline float sq3(float a, float b, float c) {
  float t = b * b;
  t = fmaf(a, a, t);
  t = fmaf(c, c, t);
  return t;
}

/* cuda_utils.h:20-24  opt_n_threads: 2^floor(log2 n) clamped to [1, 512]. */
PN2O_API int pn2o_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

PN2O_API int pn2o_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * K1  furthest_point_sampling_kernel  (sampling_gpu.cu:74-178, __update :64-70,
 *     launch :180-231 with block = opt_n_threads(n); temp initialised to 1e10 by sampling.cpp:80-82).
 * dataset (b,n,3) f32, temp (b,n) f32 scratch (caller fills with 1e10f), idxs (b,m) i32 zero-init.
 * The block's shared-memory tree reduction is emulated literally with two arrays of bs slots.
 * ------------------------------------------------------------------------------------------- */
PN2O_API void pn2o_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                           int32_t *idxs) {
  if (m <= 0) return; /* sampling_gpu.cu:78 */
  const int bs = pn2o_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float *ds = dataset + (size_t)bi * n * 3;
    float *tp = temp + (size_t)bi * n;
    int32_t *out = idxs + (size_t)bi * m;
    float *dists = (float *)malloc(sizeof(float) * (size_t)bs);
    int *dists_i = (int *)malloc(sizeof(int) * (size_t)bs);
    int old = 0;
    out[0] = old; /* :90-91 */
    for (int j = 1; j < m; ++j) {
      const float x1 = ds[old * 3 + 0], y1 = ds[old * 3 + 1], z1 = ds[old * 3 + 2];
      for (int tid = 0; tid < bs; ++tid) { /* one CUDA thread each */
        int besti = 0;
        float best = -1.0f;
        for (int k = tid; k < n; k += bs) {
          const float x2 = ds[k * 3 + 0], y2 = ds[k * 3 + 1], z2 = ds[k * 3 + 2];
          const float mag = sq3(x2, y2, z2);
          if ((double)mag <= 1e-3) continue; /* :105-106, double compare */
          const float d = sq3(x2 - x1, y2 - y1, z2 - z1);
          const float d2 = fminf(d, tp[k]); /* :111 */
          tp[k] = d2;
          besti = d2 > best ? k : besti; /* :113-114 strict > keeps the smallest k */
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      /* :120-173  tree: slot t absorbs slot t+s, tie keeps slot t (__update: v2 > v1 ? i2 : i1) */
      for (int s = bs / 2; s >= 1; s >>= 1) {
        for (int tid = 0; tid < s; ++tid) {
          const float v1 = dists[tid], v2 = dists[tid + s];
          const int i1 = dists_i[tid], i2 = dists_i[tid + s];
          dists[tid] = v1 > v2 ? v1 : v2; /* max(v1,v2) */
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
    free(dists);
    free(dists_i);
  }
}

/* K2 gather_points_kernel (sampling_gpu.cu:13-25): out[b,c,j] = points[b,c,idx[b,j]]. */
PN2O_API void pn2o_gather_points(int b, int c, int n, int m, const float *points,
                                 const int32_t *idx, float *out) {
#pragma omp parallel for collapse(2)
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j) {
        const int a = idx[(size_t)i * m + j];
        out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + a];
      }
}

/* K3 gather_points_grad_kernel (sampling_gpu.cu:39-52): atomicAdd scatter; grad_points zero-init
 * by the caller (sampling.cpp:55-57).  Sequential j order here (the GPU order is unspecified). */
PN2O_API void pn2o_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                                      const int32_t *idx, float *grad_points) {
#pragma omp parallel for collapse(2)
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j) {
        const int a = idx[(size_t)i * m + j];
        grad_points[((size_t)i * c + l) * n + a] += grad_out[((size_t)i * c + l) * m + j];
      }
}

/* K4 query_ball_point_kernel (ball_query_gpu.cu:14-49): new_xyz (b,m,3), xyz (b,n,3),
 * idx (b,m,nsample) zero-init by the caller (ball_query.cpp:27-29). */
PN2O_API void pn2o_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              const float *xyz, int32_t *idx) {
  const float radius2 = radius * radius; /* :27 fp32 product */
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < m; ++j) {
      const float *px = xyz + (size_t)bi * n * 3;
      const float *c = new_xyz + ((size_t)bi * m + j) * 3;
      int32_t *row = idx + ((size_t)bi * m + j) * nsample;
      const float nx = c[0], ny = c[1], nz = c[2];
      for (int k = 0, cnt = 0; k < n && cnt < nsample; ++k) {
        const float x = px[k * 3 + 0], y = px[k * 3 + 1], z = px[k * 3 + 2];
        const float d2 = sq3(nx - x, ny - y, nz - z);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) row[l] = k;
          row[cnt] = k;
          ++cnt;
        }
      }
    }
}

/* K5 group_points_kernel (group_points_gpu.cu:13-33): out[b,c,j,k] = points[b,c,idx[b,j,k]]. */
PN2O_API void pn2o_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                                const int32_t *idx, float *out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *p = points + ((size_t)bi * c + l) * n;
      const int32_t *ix = idx + (size_t)bi * npoints * nsample;
      float *o = out + ((size_t)bi * c + l) * npoints * nsample;
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) o[(size_t)j * nsample + k] = p[ix[(size_t)j * nsample + k]];
    }
}

/* K6 group_points_grad_kernel (group_points_gpu.cu:48-69): atomicAdd scatter, zero-init dst. */
PN2O_API void pn2o_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                     const float *grad_out, const int32_t *idx,
                                     float *grad_points) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      float *gp = grad_points + ((size_t)bi * c + l) * n;
      const int32_t *ix = idx + (size_t)bi * npoints * nsample;
      const float *go = grad_out + ((size_t)bi * c + l) * npoints * nsample;
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) gp[ix[(size_t)j * nsample + k]] += go[(size_t)j * nsample + k];
    }
}

/* K7 three_nn_kernel (interpolate_gpu.cu:14-64): unknown (b,n,3), known (b,m,3) ->
 * dist2 (b,n,3) f32, idx (b,n,3) i32.  best1..3 are doubles initialised to 1e40 and compared
 * against the float d (promoted), strict '<' at each level. */
PN2O_API void pn2o_three_nn(int b, int n, int m, const float *unknown, const float *known,
                            float *dist2, int32_t *idx) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < n; ++j) {
      const float *u = unknown + ((size_t)bi * n + j) * 3;
      const float *kn = known + (size_t)bi * m * 3;
      const float ux = u[0], uy = u[1], uz = u[2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        const float x = kn[k * 3 + 0], y = kn[k * 3 + 1], z = kn[k * 3 + 2];
        const float d = sq3(ux - x, uy - y, uz - z);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      float *dd = dist2 + ((size_t)bi * n + j) * 3;
      int32_t *ii = idx + ((size_t)bi * n + j) * 3;
      dd[0] = (float)best1; dd[1] = (float)best2; dd[2] = (float)best3;
      ii[0] = besti1; ii[1] = besti2; ii[2] = besti3;
    }
}

/* K8 three_interpolate_kernel (interpolate_gpu.cu:77-106): points (b,c,m), idx/weight (b,n,3)
 * -> out (b,c,n).  p1*w1 + p2*w2 + p3*w3 contracts like sq3: FMUL p2*w2, FFMA p1*w1, FFMA p3*w3. */
PN2O_API void pn2o_three_interpolate(int b, int c, int m, int n, const float *points,
                                     const int32_t *idx, const float *weight, float *out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *p = points + ((size_t)bi * c + l) * m;
      const int32_t *ix = idx + (size_t)bi * n * 3;
      const float *w = weight + (size_t)bi * n * 3;
      float *o = out + ((size_t)bi * c + l) * n;
      for (int j = 0; j < n; ++j) {
        float t = p[ix[j * 3 + 1]] * w[j * 3 + 1];
        t = fmaf(p[ix[j * 3 + 0]], w[j * 3 + 0], t);
        t = fmaf(p[ix[j * 3 + 2]], w[j * 3 + 2], t);
        o[j] = t;
      }
    }
}

/* K9 three_interpolate_grad_kernel (interpolate_gpu.cu:121-148): grad_out (b,c,n) ->
 * grad_points (b,c,m) zero-init; three atomicAdds per (c,j). */
PN2O_API void pn2o_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                          const int32_t *idx, const float *weight,
                                          float *grad_points) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *go = grad_out + ((size_t)bi * c + l) * n;
      const int32_t *ix = idx + (size_t)bi * n * 3;
      const float *w = weight + (size_t)bi * n * 3;
      float *gp = grad_points + ((size_t)bi * c + l) * m;
      for (int j = 0; j < n; ++j) {
        gp[ix[j * 3 + 0]] += go[j] * w[j * 3 + 0];
        gp[ix[j * 3 + 1]] += go[j] * w[j * 3 + 1];
        gp[ix[j * 3 + 2]] += go[j] * w[j * 3 + 2];
      }
    }
}
