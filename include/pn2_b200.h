/*
 * pn2_b200.h -- C ABI of libpn2_b200.so: PointNet++ set-abstraction / feature-propagation ops for
 * NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's native layer.  The reference (AIR-DISCOVER/Omni-PQ)
 * exposes nine functions through a pybind11/ATen module `pointnet2._ext`
 * (pointnet2/_ext_src/src/bindings.cpp:11-24, prototypes in pointnet2/_ext_src/include/{sampling,
 * ball_query,group_points,interpolate}.h).  Each op-level entry point below replaces one of them,
 * with plain device pointers, extents and a stream instead of at::Tensor:
 *
 *   - every pointer is a DEVICE pointer to a contiguous buffer (float = fp32, int = int32), laid
 *     out exactly like the reference tensors (xyz (B,N,3) AoS, features channel-major (B,C,N));
 *   - the caller allocates outputs (the reference allocates them in its .cpp wrappers); where the
 *     reference relies on zero-initialisation (torch::zeros) this header says so per function;
 *   - `stream` is a cudaStream_t passed as void* (the reference launches on ATen's current stream,
 *     e.g. ball_query_gpu.cu:54); launches are asynchronous, nothing synchronises;
 *   - return value: 0 = success; > 0 = a cudaError_t from the launch; < 0 = PN2_ERR_* argument
 *     error.  Nothing ever calls exit() (the reference's CUDA_CHECK_ERRORS does, cuda_utils.h:35-44).
 *     pn2_last_error() returns a message for the calling thread's last failure.
 *
 * There is no CPU path: every function needs a CUDA device (the reference likewise raises
 * "CPU not supported", sampling.cpp:41).
 */
#ifndef PN2_B200_H
#define PN2_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PN2_OK 0
#define PN2_ERR_INVALID_ARG (-1)
#define PN2_ERR_UNSUPPORTED (-2)

/* Library/ABI version (major*100+minor) and last error text of the calling thread. */
int pn2_version(void);
const char *pn2_last_error(void);

/* Block size the reference would launch FPS with for n points: clamp(2^floor(log2 n), 1, 512)
 * evaluated with the reference's double-precision formula (cuda_utils.h:20-24).  It fixes the
 * tie-break order of furthest point sampling, so it is part of the contract. */
int pn2_ref_block_size(int n);

/* Number of kernels this library has launched in the calling process so far (every launch goes through one
 * helper); bench.py reports the per-step difference as `gpu_launches`. */
long long pn2_kernel_launches(void);

/* ---- K1  furthest_point_sampling(points, nsamples)   sampling.cpp:72-93, sampling_gpu.cu:74-234 --
 * xyz (b,n,3) -> idxs (b,m) int32.  Starts at index 0, skips points with |p|^2 <= 1e-3, breaks
 * ties exactly like the reference's 512-slot shared-memory tree.
 * temp: (b,n) fp32 scratch.  Required when n exceeds the register-resident capacity
 *       (pn2_fps_resident_capacity()); optional otherwise -- when given and n <= 8192 it also enables the
 *       prefix-order check (an input that already is in FPS order, as every level after the first is, is
 *       verified in parallel and answered with 0..m-1 instead of being sampled again; same result).
 *       Contents on return are unspecified (the reference leaves min-distances there; no caller reads them).
 * new_xyz: optional (b,m,3) output = xyz gathered at idxs (what gather_points does next in
 *       pointnet2_modules.py:238); pass NULL to skip. */
int pn2_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idxs,
                                float *new_xyz, void *stream);
int pn2_fps_resident_capacity(void);

/* ---- K2/K3  gather_points / gather_points_grad   sampling.cpp:22-71, sampling_gpu.cu:13-62 ------
 * out[b,c,j] = points[b,c,idx[b,j]];  grad_points[b,c,idx[b,j]] += grad_out[b,c,j].
 * grad_points (b,c,n) must be zero-filled by the caller (torch::zeros in sampling.cpp:55-57). */
int pn2_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out,
                      void *stream);
int pn2_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                           float *grad_points, void *stream);

/* ---- K4  ball_query(new_xyz, xyz, radius, nsample)   ball_query.cpp:16-40, ball_query_gpu.cu:14-59
 * new_xyz (b,m,3), xyz (b,n,3) -> idx (b,m,nsample) int32: the first nsample points (ascending
 * index) with d2 < radius*radius, remaining slots = the first hit, all 0 if no hit.  Every slot is
 * written (the reference needs a zero-filled idx; this implementation does not). */
int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                   const float *xyz, int *idx, void *stream);

/* ---- K5/K6  group_points / group_points_grad   group_points.cpp:19-67, group_points_gpu.cu:13-80 -
 * out[b,c,j,k] = points[b,c,idx[b,j,k]];  grad_points[b,c,idx[b,j,k]] += grad_out[b,c,j,k].
 * grad_points (b,c,n) must be zero-filled by the caller. */
int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, void *stream);
int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, void *stream);

/* ---- K7  three_nn(unknowns, knows)   interpolate.cpp:22-48, interpolate_gpu.cu:14-73 -------------
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) squared distances ascending, idx (b,n,3) int32;
 * ties keep the lower index; m < 3 leaves +inf / 0 in the unused slots. */
int pn2_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx, void *stream);

/* ---- K8/K9  three_interpolate / three_interpolate_grad   interpolate.cpp:50-107 -----------------
 * points (b,c,m), idx/weight (b,n,3) -> out (b,c,n);  grad_points (b,c,m) zero-filled by caller. */
int pn2_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream);
int pn2_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, void *stream);


/* =====================================================================================================
 * Fused shared-MLP path (replaces the cuDNN/ATen calls the reference makes through
 * pytorch_utils.SharedMLP = Conv2d 1x1 (no bias) -> BatchNorm2d -> ReLU, pytorch_utils.py:11-36,67-120,
 * and F.max_pool2d over nsample, pointnet2_modules.py:254-257, plus their autograd backward).
 *
 * Everything works on POSITION-MAJOR activations: a matrix [rows][ld] with one row per position
 * (position = (cloud, centre, neighbour slot) for set abstraction, (cloud, point) for feature
 * propagation) and channels contiguous; ld and every channel count are multiples of 4 (zero padded).
 * A `pn2_rows` describes where the rows of a GEMM operand come from and which element-wise transform
 * is applied while they are loaded -- that is how grouping, BatchNorm(+ReLU) and the BatchNorm /
 * max-pool backward are fused into the contraction instead of being separate passes over HBM.
 * =================================================================================================== */
#define PN2_ROWS_PLAIN 0   /* v = x[row][c]                                                            */
#define PN2_ROWS_BNRELU 1  /* v = max(0, x[row][c]*c0[c] + c1[c])            (BN folded to scale/shift) */
#define PN2_ROWS_GATHER 2  /* v = feat[src(row)][c] for c < feat_cols, then ((xyz[src]-centre)/inv_scale, 0):
                              QueryAndGroup (pointnet2_utils.py:334-359) without materialising the groups */
#define PN2_ROWS_DY 3      /* v = c0[c]*dz[row][c] + c1[c] + c2[c]*x[row][c]  (BatchNorm backward, x = pre-BN y) */
#define PN2_ROWS_DYPOOL 4  /* as DY with dz[row][c] = (arg[g][c] == row % group) ? dz[g][c] : 0, g = row/group
                              (max-pool backward routed through the saved arg-max slot)                 */

typedef struct pn2_rows {
  int kind;
  int rows;  /* positions */
  int cols;  /* channels this source yields (multiple of 4) */
  int ld;    /* row stride of x / dz / arg in elements (multiple of 4) */
  const float *x;
  const float *c0, *c1, *c2; /* per-channel coefficient vectors, zero padded to cols */
  const float *dz;
  const unsigned char *arg;
  int group;
  /* PN2_ROWS_GATHER only */
  const int *idx;       /* [rows] neighbour index inside its cloud (ball query output) */
  const float *xyz;     /* [B*n_src][3] */
  const float *centres; /* [B*npoint][3] */
  int n_src, npoint, nsample;
  int feat_cols; /* padded feature width, 0 = no features; x = point-major features [B*n_src][ld] */
  int use_xyz;
  float inv_scale; /* radius when normalize_xyz, else 1 */
} pn2_rows;

/* Weights (cout,cin) row-major -> zero padded, optionally column-permuted copies used by the GEMMs:
 * wt [kp][np] (k-major, for forward) and wp [np][kp] (for dgrad).  `xyz_first` != 0 moves the first three
 * input channels (the xyz channels QueryAndGroup puts first) behind the `feat_pad` feature channels so
 * that gathered feature rows stay 16-byte aligned: k' = [features 0..cin-4 | pad | x y z 0].
 * Each buffer additionally carries, right after the plain matrix, the tensor-core image of that matrix
 * (every 128-row x 32-column block -- 256-row blocks when the matrix has a multiple of 256 rows -- split into tf32
 * hi / lo halves, in the UMMA 128-byte-swizzle layout; the kernels read either block size) which the
 * forward / dgrad kernels fetch with one bulk copy per k-block.  Sizes, in floats:
 * wt: pn2_mlp_weight_floats(kp, np), wp: pn2_mlp_weight_floats(np, kp). */
long long pn2_mlp_weight_floats(int rows, int cols);
int pn2_mlp_prep_weights(int cout, int cin, int xyz_first, int feat_pad, int kp, int np, const float *w,
                         float *wt, float *wp, void *stream);

/* y[rows][ldy] = A . W^T  (A = `a` rows x kp; wt [kp][np] and wp [np][kp] are the two prepared copies of W:
 * the FFMA kernel reads wt, the tcgen05 kernel the K-major wp); also writes per-row-tile partial column sums
 * stats[tiles][2][np] (sum, sum of squares) for BatchNorm; returns the number of row tiles in *tiles.
 * `stats` may be NULL.  PN2_TC=0 in the environment selects the FFMA kernel everywhere. */
int pn2_mlp_forward(const pn2_rows *a, int kp, int np, const float *wt, const float *wp, float *y, int ldy,
                    float *stats, int *tiles, void *stream);
int pn2_mlp_tiles(int rows, int np); /* row tiles pn2_mlp_forward / pn2_mlp_dgrad use for this shape */

/* BatchNorm statistics -> folded scale/shift (+ saved mean / invstd, running-stat update).
 * training != 0: batch statistics from `stats` (tiles partials, `count` rows) or, if sums != NULL, from the fp64
 * totals sums[2*c + 1] = (sum[c], sum of squares[c], row count) -- SyncBatchNorm (torch/nn/modules/_functions.py
 * SyncBatchNorm.forward; enabled by the reference at models/pq_transformer.py:194): pn2_bn_reduce_stats writes
 * this rank's totals and row count, the caller all-reduces the 2c+1 doubles, and the row count is consumed ON THE
 * DEVICE (no host synchronisation).  training == 0: running statistics.  gamma/beta may be NULL (affine=False).
 * momentum < 0 means cumulative average. */
int pn2_bn_reduce_stats(int tiles, int c, int np, double count, const float *stats, double *sums, void *stream);
/* SyncBatchNorm exchange fused into the statistics kernel (replaces pn2_bn_reduce_stats + the caller's all-reduce):
 * `peers[r]` = device address, in THIS process, of rank r's exchange buffer (symmetric / peer-mapped memory;
 * 2 * slot_doubles doubles + `world` 32-bit signal words, zero-filled once, slot_doubles >= 2c+2); `epoch` = 1, 2, 3 ...
 * incremented by the caller for every exchange (identically on every rank); `cta_ticket` = a zeroed device word.
 * Writes sums[2c+1] = totals and row count over all ranks.  Every rank of the group must issue the same sequence of
 * exchanges; a peer that never arrives makes the kernel trap after ~2 s (CUDA error) rather than hang.  Not for use
 * inside CUDA-graph capture (the epoch is a launch argument). */
int pn2_bn_sync_exchange(int tiles, int c, int np, double count, const float *stats, const unsigned long long *peers,
                         int rank, int world, unsigned epoch, int slot_doubles, unsigned *cta_ticket, double *sums,
                         void *stream);
int pn2_bn_finalize(int training, int tiles, int c, int np, double count, const float *stats, const double *sums,
                    const float *gamma, const float *beta, float *running_mean, float *running_var,
                    long long *num_batches_tracked, float momentum, float eps, float *scale, float *shift,
                    float *mean, float *invstd, void *stream);

/* out_pm[g][c] = max_s relu(y[g*group+s][c]*scale[c]+shift[c]), arg[g][c] = first slot attaining it. */
int pn2_bn_relu_pool(int groups, int group, int c, int ld, const float *y, const float *scale, const float *shift,
                     float *out_pm, unsigned char *arg, void *stream);

/* (B,C,N) channel-major <-> point-major rows of `stride` floats per point: to_point_major writes columns
 * [0,ld) of every row (zeros beyond c), to_channel_major reads columns [0,c). */
int pn2_to_point_major(int b, int c, int n, int ld, int stride, const float *src, float *dst, void *stream);
int pn2_to_channel_major(int b, int c, int n, int stride, const float *src, float *dst, void *stream);

/* Backward of pool+ReLU: gz[g][c] = gout_pm[g][c] * (out_pm[g][c] > 0) in place, and partial sums over
 * groups of gz and gz*y[g*group+arg][c] -> stats[tiles][2][ld]; returns tiles. */
int pn2_pool_bwd_prep(int groups, int group, int c, int ld, float *gz, const float *out_pm, const unsigned char *arg,
                      const float *y, float *stats, int *tiles, void *stream);
int pn2_pool_bwd_tiles(int groups);

/* BatchNorm backward coefficients: dy = ca*dz + cb + cc*y, dgamma, dbeta (training: batch statistics;
 * eval: ca = gamma*invstd, cb = cc = 0).  Input: this rank's partial sums (sum dz, sum dz*y) in `stats`.
 * SyncBatchNorm: `sums` = the totals over all ranks (2c doubles, all-reduced by the caller) and `count_dev` = the
 * global row count on the device (element 2c of the forward's sums); they determine ca/cb/cc, while dgamma/dbeta
 * are computed from the LOCAL partial sums like torch's SyncBatchNorm does (DDP averages them afterwards). */
int pn2_bn_bwd_finalize(int training, int tiles, int c, int np, double count, const float *stats, const double *sums,
                        const double *count_dev, const float *gamma, const float *mean, const float *invstd, float *ca,
                        float *cb, float *cc, float *dgamma, float *dbeta, void *stream);

/* dX = dY . W with dY given by `dy` (kind DY / DYPOOL); wp [np of forward][ldw] (FFMA kernel) and
 * wt [kp][np of forward] (tcgen05 kernel, may be NULL) are the prepared copies of W.  Modes:
 *  PN2_DGRAD_MASK: dz_prev = dX * (prev_y*prev_scale+prev_shift > 0) stored to out[rows][ldo], partial
 *                  sums (dz_prev, dz_prev*prev_y) to stats;
 *  PN2_DGRAD_STORE: out = dX;
 *  PN2_DGRAD_SCATTER: dX scatter-added through `gather` (the forward's PN2_ROWS_GATHER source) into
 *                  dfeat [B*n_src][ldf] and, if dxyz != NULL, into dxyz [B*n_src][3] (both the neighbour's
 *                  +dX/inv_scale and the centre's -dX/inv_scale via centre_src[B*npoint]). */
#define PN2_DGRAD_MASK 0
#define PN2_DGRAD_STORE 1
#define PN2_DGRAD_SCATTER 2
int pn2_mlp_dgrad(int mode, const pn2_rows *dy, int ncols, const float *wp, int ldw, const float *wt, float *out, int ldo,
                  const float *prev_y, int ld_prev, const float *prev_scale, const float *prev_shift, float *stats,
                  int *tiles, const pn2_rows *gather, float *dfeat, int ldf, float *dxyz, const int *centre_src,
                  void *stream);

/* dW[cout][cin] = sum_p dY[p][:]^T A[p][:], dY from `dy`, A from `a` (PLAIN / BNRELU / GATHER), written
 * in the original (cout,cin) layout (undoing the xyz_first permutation).  ws: workspace of
 * pn2_mlp_wgrad_workspace() floats. */
long long pn2_mlp_wgrad_workspace(int rows, int np, int kp);
int pn2_mlp_wgrad(const pn2_rows *dy, const pn2_rows *a, int cout, int cin, int xyz_first, int feat_pad, float *ws,
                  float *dw, void *stream);

/* Feature propagation front end (three_nn + inverse-distance weights + three_interpolate in one
 * gather-MAC kernel, pointnet2_modules.py:393-401): writes rows [n][ld] = sum_t w_t * known_pm[idx_t][:]
 * into out (column offset applied by the caller), plus idx / weight for the backward. */
int pn2_fp_interpolate(int b, int n, int m, int c, int ld_known, const float *unknown, const float *known,
                       const float *known_pm, float *out, int ldo, int *idx, float *weight, void *stream);
/* dknown_pm[idx_t][c] += w_t * dout[row][c]  (dknown_pm zero-filled by the caller). */
int pn2_fp_interpolate_grad(int b, int n, int m, int c, const float *dout, int ldo, const int *idx,
                            const float *weight, float *dknown_pm, int ld_known, void *stream);

/* Development aid: while `device_buf` (6 x `ctas` uint64, device memory) is installed, every CTA of the tensor-core
 * GEMM kernels records [smid, t_start, t_prologue_done, t_mainloop_done, t_end, k_blocks] (globaltimer ns) for
 * tools/gemm_trace.py.  Pass NULL to switch it off (the default). */
int pn2_debug_gemm_trace(unsigned long long *device_buf, int ctas);
/* Same for the persistent async kernel, per tile: 8 x 8 uint64 per CTA = for each of a CTA's first 8 tiles
 * [tile start, first k-block staged, last k-block staged, accumulator complete, epilogue done, MMA thread: first
 * A stage seen, first weight stage seen, weight loader: first copy issued] (globaltimer ns). */
int pn2_debug_gemm_trace2(unsigned long long *device_buf, int ctas);

#ifdef __cplusplus
}
#endif
#endif /* PN2_B200_H */
