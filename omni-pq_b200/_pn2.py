"""ctypes binding of libpn2_b200.so (include/pn2_b200.h) -- the stand-in for the reference's pybind
module `pointnet2._ext` (pointnet2/_ext_src/src/bindings.cpp:11-24).

Each function below has the name, argument order, output allocation (zeros / int32) and error
behaviour of the reference wrapper it replaces (the `.cpp` files under pointnet2/_ext_src/src):
argument errors raise RuntimeError with the reference's wording ("... must be a contiguous tensor",
"... must be a CUDA tensor" -- there is no CPU path, exactly like "CPU not supported" in the
reference, sampling.cpp:41).  Unlike the reference, a failed kernel launch raises instead of calling
exit(-1) (cuda_utils.h:35-44).

torch is plumbing here: it owns device memory and the current stream; every computation is a kernel
of libpn2_b200.so launched on `torch.cuda.current_stream()`.  Importing this module without the built
library raises ImportError -- there is no fallback.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpn2_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python omni-pq_b200/build.py` "
        "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the PointNet++ ops.")

lib = ctypes.CDLL(LIB_PATH)
lib.pn2_last_error.restype = ctypes.c_char_p
lib.pn2_version.restype = ctypes.c_int
lib.pn2_kernel_launches.restype = ctypes.c_longlong


def kernel_launches():
    """Kernels launched by libpn2_b200.so in this process so far (counted inside the library)."""
    return int(lib.pn2_kernel_launches())


_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float

#: number of kernels launched through this binding (bench.py reports it as gpu_launches)
launch_count = 0


# ---- optional per-launch CUDA-event timing (bench.py's roofline pass) ----------------------------------
_prof = None          # list of (label, start_event, end_event, flops, bytes) while profiling
_work = (None, 0.0, 0.0)


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> [(label, milliseconds, algorithmic flops, algorithmic bytes)] (call after a synchronize)."""
    global _prof
    recs, _prof = _prof, None
    return [(lbl, s.elapsed_time(e), fl, by) for lbl, s, e, fl, by in recs]


def _annotate(label, flops=0.0, nbytes=0.0):
    """Describe the next launch for the profile: a label and its algorithmic work."""
    global _work
    if _prof is not None:
        _work = (label, float(flops), float(nbytes))


def _sig(name, *argtypes, kernel=True):
    fn = getattr(lib, name)
    fn.argtypes = list(argtypes)
    fn.restype = ctypes.c_int
    if not kernel:
        return fn

    def call(*args):
        global _work
        if _prof is None:
            return fn(*args)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        _prof.append((_work[0] or name, s, e, _work[1], _work[2]))
        _work = (None, 0.0, 0.0)
        return rc

    return call


def _check(rc):
    if rc != 0:
        msg = lib.pn2_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libpn2_b200: {msg} (code {rc})")


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return _vp(t.data_ptr())


# --- the reference's CHECK_* macros (pointnet2/_ext_src/include/utils.h:10-30) --------------------
def _want_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("CPU not supported: tensor must be a CUDA tensor")


def _want_float(t, name):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")


def _want_int(t, name):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")


_fps = _sig("pn2_furthest_point_sampling", _i, _i, _i, _vp, _vp, _vp, _vp, _vp)
_gather = _sig("pn2_gather_points", _i, _i, _i, _i, _vp, _vp, _vp, _vp)
_gather_grad = _sig("pn2_gather_points_grad", _i, _i, _i, _i, _vp, _vp, _vp, _vp)
_ball = _sig("pn2_ball_query", _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp)
_group = _sig("pn2_group_points", _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp)
_group_grad = _sig("pn2_group_points_grad", _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp)
_three_nn = _sig("pn2_three_nn", _i, _i, _i, _vp, _vp, _vp, _vp, _vp)
_interp = _sig("pn2_three_interpolate", _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp)
_interp_grad = _sig("pn2_three_interpolate_grad", _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp)
lib.pn2_ref_block_size.argtypes = [_i]
lib.pn2_ref_block_size.restype = _i
lib.pn2_fps_resident_capacity.restype = _i


def _launched(n=1):
    global launch_count
    launch_count += n


def furthest_point_sampling(points, nsamples, return_xyz=False):
    """sampling.cpp:72-93.  points (B,N,3) f32 -> (B,nsamples) int32.
    return_xyz=True also returns points gathered at the result, (B,nsamples,3), from the same kernel."""
    _want_float(points, "points")
    _want_cuda(points)
    b, n, _ = points.shape
    nsamples = int(nsamples)
    with torch.cuda.device(points.device):
        out = torch.zeros(b, nsamples, dtype=torch.int32, device=points.device)
        new_xyz = torch.empty(b, nsamples, 3, dtype=torch.float32, device=points.device) if return_xyz else None
        tmp = None
        if n > lib.pn2_fps_resident_capacity() or n <= 8192:  # streaming scratch / prefix-order check scratch
            tmp = torch.empty(b, n, dtype=torch.float32, device=points.device)
        _annotate(f"fps_kernel[{b}x{n}->{nsamples}]", nbytes=b * (12.0 * n + 16.0 * nsamples))
        _check(_fps(b, n, nsamples, _ptr(points), _ptr(tmp) if tmp is not None else None, _ptr(out),
                    _ptr(new_xyz) if return_xyz else None, _stream()))
    _launched()
    return (out, new_xyz) if return_xyz else out


def gather_points(points, idx):
    """sampling.cpp:22-46.  points (B,C,N), idx (B,m) -> (B,C,m)."""
    _want_float(points, "points")
    _want_int(idx, "idx")
    _want_cuda(points, idx)
    b, c, n = points.shape
    m = idx.shape[1]
    with torch.cuda.device(points.device):
        out = torch.empty(b, c, m, dtype=torch.float32, device=points.device)
        _check(_gather(b, c, n, m, _ptr(points), _ptr(idx), _ptr(out), _stream()))
    _launched()
    return out


def gather_points_grad(grad_out, idx, n):
    """sampling.cpp:48-71.  grad_out (B,C,m), idx (B,m) -> (B,C,n)."""
    _want_float(grad_out, "grad_out")
    _want_int(idx, "idx")
    _want_cuda(grad_out, idx)
    b, c, m = grad_out.shape
    with torch.cuda.device(grad_out.device):
        out = torch.zeros(b, c, int(n), dtype=torch.float32, device=grad_out.device)
        _check(_gather_grad(b, c, int(n), m, _ptr(grad_out), _ptr(idx), _ptr(out), _stream()))
    _launched()
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """ball_query.cpp:16-40 (note the ext-level argument order: centres first)."""
    _want_float(new_xyz, "new_xyz")
    _want_float(xyz, "xyz")
    _want_cuda(new_xyz, xyz)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    with torch.cuda.device(xyz.device):
        idx = torch.empty(b, m, int(nsample), dtype=torch.int32, device=xyz.device)
        _annotate(f"ball_query_kernel[{b}x{n}x{m}]", nbytes=b * (12.0 * (n + m) + 4.0 * m * int(nsample)))
        _check(_ball(b, n, m, float(radius), int(nsample), _ptr(new_xyz), _ptr(xyz), _ptr(idx), _stream()))
    _launched()
    return idx


def group_points(points, idx):
    """group_points.cpp:19-42.  points (B,C,N), idx (B,npoints,nsample) -> (B,C,npoints,nsample)."""
    _want_float(points, "points")
    _want_int(idx, "idx")
    _want_cuda(points, idx)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    with torch.cuda.device(points.device):
        out = torch.empty(b, c, npoints, nsample, dtype=torch.float32, device=points.device)
        _check(_group(b, c, n, npoints, nsample, _ptr(points), _ptr(idx), _ptr(out), _stream()))
    _launched()
    return out


def group_points_grad(grad_out, idx, n):
    """group_points.cpp:44-67.  grad_out (B,C,npoints,nsample) -> (B,C,n)."""
    _want_float(grad_out, "grad_out")
    _want_int(idx, "idx")
    _want_cuda(grad_out, idx)
    b, c, npoints, nsample = grad_out.shape
    with torch.cuda.device(grad_out.device):
        out = torch.zeros(b, c, int(n), dtype=torch.float32, device=grad_out.device)
        _check(_group_grad(b, c, int(n), npoints, nsample, _ptr(grad_out), _ptr(idx), _ptr(out), _stream()))
    _launched()
    return out


def three_nn(unknowns, knows):
    """interpolate.cpp:22-48.  -> [dist2 (B,n,3) f32, idx (B,n,3) int32] (squared distances)."""
    _want_float(unknowns, "unknowns")
    _want_float(knows, "knows")
    _want_cuda(unknowns, knows)
    b, n, _ = unknowns.shape
    m = knows.shape[1]
    with torch.cuda.device(unknowns.device):
        idx = torch.empty(b, n, 3, dtype=torch.int32, device=unknowns.device)
        dist2 = torch.empty(b, n, 3, dtype=torch.float32, device=unknowns.device)
        _check(_three_nn(b, n, m, _ptr(unknowns), _ptr(knows), _ptr(dist2), _ptr(idx), _stream()))
    _launched()
    return dist2, idx


def three_interpolate(points, idx, weight):
    """interpolate.cpp:50-78.  points (B,C,m), idx/weight (B,n,3) -> (B,C,n)."""
    _want_float(points, "points")
    _want_int(idx, "idx")
    _want_float(weight, "weight")
    _want_cuda(points, idx, weight)
    b, c, m = points.shape
    n = idx.shape[1]
    with torch.cuda.device(points.device):
        out = torch.empty(b, c, n, dtype=torch.float32, device=points.device)
        _check(_interp(b, c, m, n, _ptr(points), _ptr(idx), _ptr(weight), _ptr(out), _stream()))
    _launched()
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """interpolate.cpp:79-107.  grad_out (B,C,n) -> (B,C,m)."""
    _want_float(grad_out, "grad_out")
    _want_int(idx, "idx")
    _want_float(weight, "weight")
    _want_cuda(grad_out, idx, weight)
    b, c, n = grad_out.shape
    with torch.cuda.device(grad_out.device):
        out = torch.zeros(b, c, int(m), dtype=torch.float32, device=grad_out.device)
        _check(_interp_grad(b, c, n, int(m), _ptr(grad_out), _ptr(idx), _ptr(weight), _ptr(out), _stream()))
    _launched()
    return out


# ======================================================================================================
# Fused shared-MLP path (include/pn2_b200.h, second half).  Thin wrappers: torch allocates, the library
# computes.  Position-major activations are 2-D tensors [rows, ld] with ld % 4 == 0.
# ======================================================================================================
ROWS_PLAIN, ROWS_BNRELU, ROWS_GATHER, ROWS_DY, ROWS_DYPOOL = 0, 1, 2, 3, 4
DGRAD_MASK, DGRAD_STORE, DGRAD_SCATTER = 0, 1, 2


class Rows(ctypes.Structure):
    """Mirror of `pn2_rows`; `keep` holds the tensors whose pointers it carries."""
    _fields_ = [("kind", _i), ("rows", _i), ("cols", _i), ("ld", _i),
                ("x", _vp), ("c0", _vp), ("c1", _vp), ("c2", _vp), ("dz", _vp), ("arg", _vp),
                ("group", _i),
                ("idx", _vp), ("xyz", _vp), ("centres", _vp),
                ("n_src", _i), ("npoint", _i), ("nsample", _i), ("feat_cols", _i), ("use_xyz", _i),
                ("inv_scale", _f)]


def pad4(c):
    return (int(c) + 3) // 4 * 4


def _p(t):
    return None if t is None else t.data_ptr()


def rows_plain(x, rows, cols, ld):
    r = Rows(kind=ROWS_PLAIN, rows=rows, cols=cols, ld=ld, x=_p(x))
    r.keep = (x,)
    return r


def rows_bnrelu(y, rows, cols, ld, scale, shift):
    r = Rows(kind=ROWS_BNRELU, rows=rows, cols=cols, ld=ld, x=_p(y), c0=_p(scale), c1=_p(shift))
    r.keep = (y, scale, shift)
    return r


def rows_gather(feat_pm, ldf, feat_cols, idx, xyz, centres, n_src, npoint, nsample, use_xyz, inv_scale):
    rows = idx.numel()
    r = Rows(kind=ROWS_GATHER, rows=rows, cols=feat_cols + (4 if use_xyz else 0), ld=ldf, x=_p(feat_pm),
             idx=_p(idx), xyz=_p(xyz), centres=_p(centres), n_src=n_src, npoint=npoint, nsample=nsample,
             feat_cols=feat_cols, use_xyz=int(bool(use_xyz)), inv_scale=float(inv_scale))
    r.keep = (feat_pm, idx, xyz, centres)
    return r


def rows_dy(y, dz, rows, cols, ld, ca, cb, cc, arg=None, group=1):
    kind = ROWS_DY if arg is None else ROWS_DYPOOL
    r = Rows(kind=kind, rows=rows, cols=cols, ld=ld, x=_p(y), c0=_p(ca), c1=_p(cb), c2=_p(cc), dz=_p(dz),
             arg=_p(arg), group=group)
    r.keep = (y, dz, ca, cb, cc, arg)
    return r


_rp = ctypes.POINTER(Rows)
_ip_ = ctypes.POINTER(ctypes.c_int)
_d = ctypes.c_double
_prep = _sig("pn2_mlp_prep_weights", _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp)
_mlp_fwd = _sig("pn2_mlp_forward", _rp, _i, _i, _vp, _vp, _vp, _i, _vp, _ip_, _vp)
_mlp_tiles = _sig("pn2_mlp_tiles", _i, _i, kernel=False)
_bn_reduce = _sig("pn2_bn_reduce_stats", _i, _i, _i, _d, _vp, _vp, _vp)
_bn_sync = _sig("pn2_bn_sync_exchange", _i, _i, _i, _d, _vp, _vp, _i, _i, ctypes.c_uint, _i, _vp, _vp, _vp)
_bn_fin = _sig("pn2_bn_finalize", _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp)
_pool = _sig("pn2_bn_relu_pool", _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp)
_to_pm = _sig("pn2_to_point_major", _i, _i, _i, _i, _i, _vp, _vp, _vp)
_to_cm = _sig("pn2_to_channel_major", _i, _i, _i, _i, _vp, _vp, _vp)
_pool_prep = _sig("pn2_pool_bwd_prep", _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _ip_, _vp)
_pool_tiles = _sig("pn2_pool_bwd_tiles", _i, kernel=False)
_bn_bwd = _sig("pn2_bn_bwd_finalize", _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp)
_dgrad = _sig("pn2_mlp_dgrad", _i, _rp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _ip_, _rp, _vp, _i, _vp, _vp, _vp)
_wgrad = _sig("pn2_mlp_wgrad", _rp, _rp, _i, _i, _i, _i, _vp, _vp, _vp)
lib.pn2_mlp_weight_floats.argtypes = [_i, _i]
lib.pn2_mlp_weight_floats.restype = ctypes.c_longlong
lib.pn2_mlp_wgrad_workspace.argtypes = [_i, _i, _i]
lib.pn2_mlp_wgrad_workspace.restype = ctypes.c_longlong
_fp_interp = _sig("pn2_fp_interpolate", _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp)
_fp_interp_grad = _sig("pn2_fp_interpolate_grad", _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp)


_POISON = bool(os.environ.get("PN2_DEBUG_POISON"))  # debugging aid: NaN-fill every scratch/output buffer


def _f32(dev, *shape, zero=False):
    t = (torch.zeros if zero else torch.empty)(*shape, dtype=torch.float32, device=dev)
    if _POISON and not zero:
        t.fill_(float("nan"))
    return t


def mlp_prep_weights(w2d, xyz_first, feat_pad, kp, np_):
    """(cout,cin) weights -> (wt [kp,np], wp [np,kp]) padded / permuted copies."""
    cout, cin = w2d.shape
    # each buffer = plain matrix + its tensor-core image (pn2_b200.h); the returned tensors view the plain part
    wt = _f32(w2d.device, lib.pn2_mlp_weight_floats(kp, np_))[:kp * np_].view(kp, np_)
    wp = _f32(w2d.device, lib.pn2_mlp_weight_floats(np_, kp))[:kp * np_].view(np_, kp)
    _check(_prep(cout, cin, int(xyz_first), feat_pad, kp, np_, _ptr(w2d), _ptr(wt), _ptr(wp), _stream()))
    _launched()
    return wt, wp


def mlp_forward(rows, kp, np_, wt, wp=None, want_stats=True):
    dev = wt.device
    y = _f32(dev, rows.rows, np_)
    tiles = _mlp_tiles(rows.rows, np_)
    stats = _f32(dev, max(tiles, 1), 2, np_) if want_stats else None
    t = ctypes.c_int(0)
    _annotate(f"gemm_kernel<forward>[{rows.rows}x{kp}->{np_}]", flops=2.0 * rows.rows * kp * np_, nbytes=4.0 * rows.rows * (kp + np_))
    _check(_mlp_fwd(ctypes.byref(rows), kp, np_, _ptr(wt), _p(wp), _ptr(y), np_, _p(stats), ctypes.byref(t), _stream()))
    _launched()
    return y, stats, tiles


def bn_reduce_stats(stats, tiles, c, np_, count):
    """Per-tile fp32 partials -> this rank's fp64 totals [2c+1] = (sums[c], second sums[c], `count` rows): the
    buffer a SyncBatchNorm exchange all-reduces in one piece (the row count travels with the totals and is consumed
    on the device, so the exchange needs no host synchronisation)."""
    sums = torch.empty(2 * c + 1, dtype=torch.float64, device=stats.device)
    _check(_bn_reduce(tiles, c, np_, float(count), _ptr(stats), _ptr(sums), _stream()))
    _launched()
    return sums


def bn_sync_exchange(stats, tiles, c, np_, count, peers, rank, world, epoch, slot_doubles, ticket):
    """Statistics reduction + cross-rank exchange over peer memory in one kernel -> sums [2c+1] over all ranks."""
    sums = torch.empty(2 * c + 1, dtype=torch.float64, device=stats.device)
    _check(_bn_sync(tiles, c, np_, float(count), _ptr(stats), _ptr(peers), rank, world, epoch, slot_doubles, _ptr(ticket),
                    _ptr(sums), _stream()))
    _launched()
    return sums


def bn_finalize(training, tiles, c, np_, count, stats, sums, bn):
    """-> (scale, shift, mean, invstd), each [np_].  Updates bn's running statistics when training."""
    dev = bn.running_mean.device if bn.running_mean is not None else (stats if stats is not None else sums).device
    out = _f32(dev, 4, np_)
    track = training and bn.track_running_stats and bn.running_mean is not None
    mom = -1.0 if bn.momentum is None else float(bn.momentum)
    _check(_bn_fin(int(training), tiles, c, np_, float(count), _p(stats), _p(sums), _p(bn.weight), _p(bn.bias),
                   _p(bn.running_mean) if (track or not training) else None,
                   _p(bn.running_var) if (track or not training) else None,
                   _p(bn.num_batches_tracked) if track else None, mom, float(bn.eps),
                   _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), _stream()))
    _launched(2 if (track and bn.num_batches_tracked is not None and bn.momentum is None) else 1)
    return out[0], out[1], out[2], out[3]


def bn_relu_pool(y, groups, group, c, ld, scale, shift, want_arg=True):
    out_pm = _f32(y.device, groups, ld)
    arg = torch.empty(groups, ld, dtype=torch.uint8, device=y.device) if want_arg else None
    _annotate("bn_relu_pool_kernel", nbytes=4.0 * groups * ld * (group + 1.25))
    _check(_pool(groups, group, c, ld, _ptr(y), _ptr(scale), _ptr(shift), _ptr(out_pm), _p(arg), _stream()))
    _launched()
    return out_pm, arg


def to_point_major(src, ld=None, out=None, col0=0):
    """(B,C,N) -> [B*N, ld] (or into columns col0.. of `out`, whose row stride is out.shape[1])."""
    b, c, n = src.shape
    ld = pad4(c) if ld is None else ld
    if out is None:
        out = _f32(src.device, b * n, ld)
    stride = out.shape[1]
    _annotate("to_point_major_kernel", nbytes=8.0 * b * c * n)
    _check(_to_pm(b, c, n, ld, stride, _ptr(src), _vp(out.data_ptr() + 4 * col0), _stream()))
    _launched()
    return out


def to_channel_major(src_pm, b, c, n, col0=0):
    """[B*N, stride] (columns col0..col0+c) -> (B,C,N)."""
    out = _f32(src_pm.device, b, c, n)
    _annotate("to_channel_major_kernel", nbytes=8.0 * b * c * n)
    _check(_to_cm(b, c, n, src_pm.shape[1], _vp(src_pm.data_ptr() + 4 * col0), _ptr(out), _stream()))
    _launched()
    return out


def pool_bwd_prep(gz, out_pm, arg, y, groups, group, c, ld):
    tiles = _pool_tiles(groups)
    stats = _f32(gz.device, max(tiles, 1), 2, ld)
    t = ctypes.c_int(0)
    _annotate("pool_bwd_prep_kernel", nbytes=4.0 * groups * ld * 4.25)
    _check(_pool_prep(groups, group, c, ld, _ptr(gz), _ptr(out_pm), _p(arg), _ptr(y), _ptr(stats), ctypes.byref(t),
                      _stream()))
    _launched()
    return stats, tiles


def bn_bwd_finalize(training, tiles, c, np_, count, stats, sums, gamma, mean, invstd, count_dev=None,
                    dgamma_out=None, dbeta_out=None):
    """-> (ca, cb, cc [np_] coefficient vectors, dgamma [c], dbeta [c]).  `sums` / `count_dev`: SyncBatchNorm totals
    over all ranks and the global row count on the device; dgamma / dbeta always come from the local `stats`."""
    dev = mean.device
    co = _f32(dev, 3, np_)
    if dgamma_out is None or dbeta_out is None:
        dg = _f32(dev, 2, c)
        dgamma_out, dbeta_out = dg[0], dg[1]
    _check(_bn_bwd(int(training), tiles, c, np_, float(count), _p(stats), _p(sums), _p(count_dev), _p(gamma), _ptr(mean), _ptr(invstd),
                   _ptr(co[0]), _ptr(co[1]), _ptr(co[2]), _ptr(dgamma_out), _ptr(dbeta_out), _stream()))
    _launched()
    return co[0], co[1], co[2], dgamma_out, dbeta_out


def mlp_dgrad_mask(dy, ncols, wp, prev_y, prev_scale, prev_shift, wt=None):
    """dz_prev [rows, ncols] and its BatchNorm-backward partial sums."""
    dev = wp.device
    out = _f32(dev, dy.rows, ncols)
    tiles = _mlp_tiles(dy.rows, ncols)
    stats = _f32(dev, max(tiles, 1), 2, ncols)
    t = ctypes.c_int(0)
    _annotate(f"gemm_kernel<dgrad>[{dy.rows}x{dy.cols}->{ncols}]", flops=2.0 * dy.rows * dy.cols * ncols, nbytes=4.0 * dy.rows * (2 * dy.cols + 2 * ncols))
    _check(_dgrad(DGRAD_MASK, ctypes.byref(dy), ncols, _ptr(wp), wp.shape[1], _p(wt), _ptr(out), ncols, _ptr(prev_y),
                  prev_y.shape[1], _ptr(prev_scale), _ptr(prev_shift), _ptr(stats), ctypes.byref(t), None, None, 0,
                  None, None, _stream()))
    _launched()
    return out, stats, tiles


def mlp_dgrad_store(dy, ncols, wp, wt=None):
    out = _f32(wp.device, dy.rows, ncols)
    _annotate(f"gemm_kernel<dgrad>[{dy.rows}x{dy.cols}->{ncols}]", flops=2.0 * dy.rows * dy.cols * ncols, nbytes=4.0 * dy.rows * (2 * dy.cols + ncols))
    _check(_dgrad(DGRAD_STORE, ctypes.byref(dy), ncols, _ptr(wp), wp.shape[1], _p(wt), _ptr(out), ncols, None, 0, None, None,
                  None, None, None, None, 0, None, None, _stream()))
    _launched()
    return out


def mlp_dgrad_scatter(dy, ncols, wp, gather, dfeat, dxyz, centre_src, wt=None):
    _annotate(f"gemm_kernel<dgrad+scatter>[{dy.rows}x{dy.cols}->{ncols}]", flops=2.0 * dy.rows * dy.cols * ncols, nbytes=4.0 * dy.rows * (2 * dy.cols + ncols))
    _check(_dgrad(DGRAD_SCATTER, ctypes.byref(dy), ncols, _ptr(wp), wp.shape[1], _p(wt), None, 0, None, 0, None, None, None,
                  None, ctypes.byref(gather), _p(dfeat), dfeat.shape[1] if dfeat is not None else 0, _p(dxyz),
                  _p(centre_src), _stream()))
    _launched()


def mlp_wgrad(dy, a, cout, cin, xyz_first, feat_pad, dev, out=None):
    """-> dW (cout, cin); written into `out` (any tensor of cout*cin contiguous floats) when given."""
    ws = _f32(dev, max(1, lib.pn2_mlp_wgrad_workspace(dy.rows, dy.cols, a.cols)))
    dw = _f32(dev, cout, cin) if out is None else out
    _annotate(f"gemm_kernel<wgrad>[{dy.rows}:{dy.cols}x{a.cols}]", flops=2.0 * dy.rows * dy.cols * a.cols, nbytes=4.0 * dy.rows * (2 * dy.cols + a.cols))
    _check(_wgrad(ctypes.byref(dy), ctypes.byref(a), cout, cin, int(xyz_first), feat_pad, _ptr(ws), _ptr(dw), _stream()))
    _launched(2)
    return dw


def fp_interpolate(unknown, known, known_pm, c, out, ldo):
    """three_nn + weights + three_interpolate into columns [0,c) of `out` [B*n, ldo]; -> (idx, weight)."""
    b, n, _ = unknown.shape
    m = known.shape[1]
    idx = torch.empty(b, n, 3, dtype=torch.int32, device=unknown.device)
    w = _f32(unknown.device, b, n, 3)
    _annotate("fp_interpolate_kernel", nbytes=4.0 * b * (3.0 * (n + m) + n * (4.0 * c + 6)))
    _check(_fp_interp(b, n, m, c, known_pm.shape[1], _ptr(unknown), _ptr(known), _ptr(known_pm), _ptr(out), ldo,
                      _ptr(idx), _ptr(w), _stream()))
    _launched()
    return idx, w


def fp_interpolate_grad(dout, ldo, idx, weight, b, n, m, c, ld_known):
    dk = _f32(dout.device, b * m, ld_known, zero=True)
    _check(_fp_interp_grad(b, n, m, c, _ptr(dout), ldo, _ptr(idx), _ptr(weight), _ptr(dk), ld_known, _stream()))
    _launched()
    return dk
