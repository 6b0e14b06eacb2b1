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

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float

#: number of kernels launched through this binding (bench.py reports it as gpu_launches)
launch_count = 0


def _sig(name, *argtypes):
    fn = getattr(lib, name)
    fn.argtypes = list(argtypes)
    fn.restype = ctypes.c_int
    return fn


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
        if n > lib.pn2_fps_resident_capacity():
            tmp = torch.empty(b, n, dtype=torch.float32, device=points.device)
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
