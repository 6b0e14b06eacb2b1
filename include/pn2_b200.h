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

/* ---- K1  furthest_point_sampling(points, nsamples)   sampling.cpp:72-93, sampling_gpu.cu:74-234 --
 * xyz (b,n,3) -> idxs (b,m) int32.  Starts at index 0, skips points with |p|^2 <= 1e-3, breaks
 * ties exactly like the reference's 512-slot shared-memory tree.
 * temp: (b,n) fp32 scratch, only touched when n exceeds the register-resident capacity
 *       (pn2_fps_resident_capacity()); may be NULL otherwise.  Contents on return are unspecified
 *       (the reference leaves min-distances there; no caller reads them).
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

#ifdef __cplusplus
}
#endif
#endif /* PN2_B200_H */
