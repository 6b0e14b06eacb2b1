// Row sources and argument block shared by the FFMA GEMM (mlp_gemm.cu) and the tcgen05 GEMM (mlp_gemm_tc.cu).
#pragma once
#include "pn2_common.cuh"

namespace pn2 {
namespace {

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// ReLU that propagates NaN like torch.relu does (fmaxf(NaN, 0) would return 0 and mask a diverged run)
__device__ __forceinline__ float relu_nan(float v) { return !(v <= 0.f) ? v : 0.f; }

// ---- row sources ------------------------------------------------------------------------------------
struct RowCtx {
  bool valid;
  size_t off;   // element offset of the row in x (PLAIN/BNRELU/DY*: row*ld; GATHER: src*ld)
  size_t goff;  // DYPOOL: g*ld; GATHER: source row (cloud*n_src + idx)
  int slot;     // DYPOOL: row % group
  float gx, gy, gz;  // GATHER: local coordinates
};

template <int KIND>
__device__ __forceinline__ RowCtx row_ctx(const pn2_rows &s, int row) {
  RowCtx c;
  c.valid = row < s.rows;
  c.off = 0; c.goff = 0; c.slot = 0; c.gx = c.gy = c.gz = 0.f;
  if (!c.valid) return c;
  if (KIND == PN2_ROWS_GATHER) {
    const int cloud = row / (s.npoint * s.nsample);
    const int centre = row / s.nsample;
    const size_t src = static_cast<size_t>(cloud) * s.n_src + __ldg(s.idx + row);
    c.off = src * s.ld;
    c.goff = src;
    if (s.use_xyz) {
      const float *p = s.xyz + src * 3, *q = s.centres + static_cast<size_t>(centre) * 3;
      // pointnet2_utils.py:350-352: grouped_xyz -= new_xyz; grouped_xyz /= radius
      c.gx = __fdiv_rn(__fsub_rn(__ldg(p + 0), __ldg(q + 0)), s.inv_scale);
      c.gy = __fdiv_rn(__fsub_rn(__ldg(p + 1), __ldg(q + 1)), s.inv_scale);
      c.gz = __fdiv_rn(__fsub_rn(__ldg(p + 2), __ldg(q + 2)), s.inv_scale);
    }
  } else {
    c.off = static_cast<size_t>(row) * s.ld;
    if (KIND == PN2_ROWS_DYPOOL) {
      const int g = row / s.group;
      c.goff = static_cast<size_t>(g) * s.ld;
      c.slot = row - g * s.group;
    }
  }
  return c;
}

template <int KIND>
__device__ __forceinline__ float4 load4(const pn2_rows &s, const RowCtx &c, int c4) {
  if (!c.valid || c4 >= s.cols) return zero4();
  if (KIND == PN2_ROWS_PLAIN) return ldg4(s.x + c.off + c4);
  if (KIND == PN2_ROWS_BNRELU) {
    const float4 v = ldg4(s.x + c.off + c4), a = ldg4(s.c0 + c4), b = ldg4(s.c1 + c4);
    return make_float4(relu_nan(fmaf(v.x, a.x, b.x)), relu_nan(fmaf(v.y, a.y, b.y)),
                       relu_nan(fmaf(v.z, a.z, b.z)), relu_nan(fmaf(v.w, a.w, b.w)));
  }
  if (KIND == PN2_ROWS_GATHER) {
    if (c4 < s.feat_cols) return ldg4(s.x + c.off + c4);
    return make_float4(c.gx, c.gy, c.gz, 0.f);  // c4 == feat_cols: the xyz block
  }
  // DY / DYPOOL: dy = c0*dz + c1 + c2*y
  const float4 y = ldg4(s.x + c.off + c4);
  const float4 ca = ldg4(s.c0 + c4), cb = ldg4(s.c1 + c4), cc = ldg4(s.c2 + c4);
  float4 dz;
  if (KIND == PN2_ROWS_DY) {
    dz = ldg4(s.dz + c.off + c4);
  } else {
    const float4 g = ldg4(s.dz + c.goff + c4);
    const uchar4 a = __ldg(reinterpret_cast<const uchar4 *>(s.arg + c.goff + c4));
    dz = make_float4(a.x == c.slot ? g.x : 0.f, a.y == c.slot ? g.y : 0.f, a.z == c.slot ? g.z : 0.f,
                     a.w == c.slot ? g.w : 0.f);
  }
  return make_float4(fmaf(cc.x, y.x, fmaf(ca.x, dz.x, cb.x)), fmaf(cc.y, y.y, fmaf(ca.y, dz.y, cb.y)),
                     fmaf(cc.z, y.z, fmaf(ca.z, dz.z, cb.z)), fmaf(cc.w, y.w, fmaf(ca.w, dz.w, cb.w)));
}

// ---- epilogues ----------------------------------------------------------------------------------------
enum { EPI_STORE = 0, EPI_STORE_STATS = 1, EPI_DGRAD_MASK = 2, EPI_SCATTER = 3 };

struct GemmArgs {
  pn2_rows A, B;
  int M, N, K;          // logical extents (all multiples of 4 where they index channels)
  int k_per_split;      // multiple of BK; blockIdx.z selects the split
  float *out;           // [M][ldo] (+ blockIdx.z * out_split_stride)
  int ldo;
  long long out_split_stride;
  float *stats;         // [gridDim.x][2][stats_ld]
  int stats_ld;
  // EPI_DGRAD_MASK
  const float *prev_y, *prev_scale, *prev_shift;
  int ld_prev;
  // EPI_SCATTER
  pn2_rows G;           // the forward's gather source
  float *dfeat;
  int ldf;
  float *dxyz;
  const int *centre_src;
  // tcgen05 forward / dgrad: pre-split, pre-swizzled image of B (pn2_mlp_prep_weights), or null
  const float *b_img;
  int b_img_kblocks;
  int b_tile_rows;    // rows per image tile: 256 when the weight matrix has a multiple of 256 rows, else 128
  // development aid (pn2_debug_gemm_trace): per-CTA phase timestamps [smid, start, prologue, main loop, end, k-blocks]
  unsigned long long *trace;
  int trace_cap;
  int *tile_counter;  // persistent GEMM: dynamic tile scheduler (self-resetting ticket counter)
  int wg_fold;        // weight gradient: the gathered source's xyz block rides in the last feature tile (wgrad_fold_ok)
};

}  // namespace

// tcgen05 path (mlp_gemm_tc.cu): returns PN2_TC_UNSUPPORTED when the shape / source / epilogue is not covered
constexpr int PN2_TC_UNSUPPORTED = -100;
int gemm_tc_launch(int akind, int epi, const void *gemm_args, cudaStream_t stream);
int gemm_tc_wgrad_launch(const void *gemm_args, int splits, cudaStream_t stream);
bool gemm_tc_enabled();
bool wgrad_fold_ok(const void *gemm_args);  // weight gradient of a gathered source: xyz block folded into the last feature tile
bool gemm_tc_wide_enabled();  // PN2_TC_WIDE=0: 128 x 128 tiles (and weight images) everywhere
void gemm_trace_target(unsigned long long **buf, int *cap);

}  // namespace pn2
