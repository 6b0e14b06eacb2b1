"""GPU parity of the nine op kernels (through the C ABI) against the CPU oracle, the frozen golden
vectors and -- when it was built into oracle/_ref -- the reference's own CUDA kernels on the same box.
Index outputs: bit-exact.  Float gathers: bit-exact forward; scatter-add gradients to 1e-5 relative
(the reference's atomics are order-nondeterministic too)."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import _pn2
    return _pn2


@pytest.fixture(scope="module")
def oracle():
    from oracle import pn2_oracle as O
    return O


@pytest.fixture(scope="module")
def ref_ext():
    from oracle import build_ref
    try:
        return build_ref.load()
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference extension not loadable: {e}")


def _assert_rel(a, b, tol=1e-5):
    scale = b.abs().max().clamp_min(1e-30)
    assert float((a - b).abs().max() / scale) <= tol


@pytest.mark.parametrize("name", list(cases.CASES))
def test_index_ops_match_golden(name, ext, golden):
    outs, dig = cases.run_case(name, ext, "cuda")
    assert golden, "no golden files"
    for kind, blob in golden.items():
        assert bytes(blob[f"{name}/input_digest"]).decode() == dig
        for k, v in outs.items():
            assert np.array_equal(v.numpy(), blob[f"{name}/{k}"], equal_nan=True), f"{kind}:{name}/{k}"


@pytest.mark.parametrize("name", list(cases.CASES))
def test_index_ops_match_reference_kernels_live(name, ext, ref_ext):
    if ref_ext is None:
        pytest.skip("oracle/_ref/pn2_ref_ext.so not present")
    ours, _ = cases.run_case(name, ext, "cuda")
    theirs, _ = cases.run_case(name, ref_ext, "cuda")
    for k in ours:
        assert torch.equal(ours[k], theirs[k]), f"{name}/{k}"


@pytest.mark.parametrize("b,n,m", [(1, 1, 1), (2, 5, 5), (3, 31, 7), (2, 128, 128), (4, 129, 40), (2, 255, 33),
                                   (1, 2048, 1024), (8, 2048, 1024), (2, 4097, 300), (3, 8192, 512),
                                   (1, 20000, 700), (9, 40000, 64), (1, 60000, 256), (160, 512, 256)])
def test_fps_shapes_vs_oracle(b, n, m, ext, oracle):
    """Every register-resident instantiation / cluster size, batches above the SM count, ragged n."""
    rng = np.random.Generator(np.random.PCG64(1000 + n))
    xyz = torch.from_numpy(rng.random((b, n, 3)).astype(np.float32) * 4 - 1)
    if n >= 128:
        xyz[:, n // 2:n // 2 + n // 8] = xyz[:, :n // 8]  # exact duplicates -> ties
    want = oracle.ext.furthest_point_sampling(xyz, m)
    got, new_xyz = ext.furthest_point_sampling(xyz.cuda(), m, return_xyz=True)
    assert torch.equal(got.cpu(), want)
    gathered = torch.gather(xyz, 1, want.long()[..., None].expand(-1, -1, 3))
    assert torch.equal(new_xyz.cpu(), gathered), "fused centre gather"


def test_fps_streaming_fallback_vs_oracle(ext, oracle):
    n = ext.lib.pn2_fps_resident_capacity() + 1234
    rng = np.random.Generator(np.random.PCG64(7))
    xyz = torch.from_numpy(rng.standard_normal((2, n, 3)).astype(np.float32))
    want = oracle.ext.furthest_point_sampling(xyz, 96)
    got = ext.furthest_point_sampling(xyz.cuda(), 96)
    assert torch.equal(got.cpu(), want)


def test_fps_all_points_skipped_yields_zero(ext, oracle):
    xyz = torch.full((2, 300, 3), 0.001)
    want = oracle.ext.furthest_point_sampling(xyz, 10)
    assert (want == 0).all()
    assert torch.equal(ext.furthest_point_sampling(xyz.cuda(), 10).cpu(), want)


@pytest.mark.parametrize("n,m,r,ns", [(1, 1, 0.5, 1), (33, 5, 0.3, 4), (1000, 77, 0.2, 64), (5000, 300, 0.1, 128),
                                      (2049, 256, 0.4, 32), (40000, 100, 0.05, 16)])
def test_ball_query_shapes_vs_oracle(n, m, r, ns, ext, oracle):
    rng = np.random.Generator(np.random.PCG64(n))
    xyz = torch.from_numpy(rng.random((2, n, 3)).astype(np.float32))
    centres = torch.from_numpy(rng.random((2, m, 3)).astype(np.float32))
    centres[:, -1] += 5.0  # an empty ball
    want = oracle.ext.ball_query(centres, xyz, r, ns)
    got = ext.ball_query(centres.cuda(), xyz.cuda(), r, ns)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("n,m", [(1, 1), (10, 2), (100, 3), (1000, 257), (3000, 1025), (50000, 2048)])
def test_three_nn_shapes_vs_oracle(n, m, ext, oracle):
    rng = np.random.Generator(np.random.PCG64(n + m))
    unknown = torch.from_numpy(rng.random((2, n, 3)).astype(np.float32))
    known = torch.from_numpy(rng.random((2, m, 3)).astype(np.float32))
    if m >= 100:
        known[:, m // 2:m // 2 + 20] = known[:, :20]
    d2w, iw = oracle.ext.three_nn(unknown, known)
    d2, idx = ext.three_nn(unknown.cuda(), known.cuda())
    assert torch.equal(idx.cpu(), iw)
    assert torch.equal(d2.cpu(), d2w)


def test_gather_group_interpolate_forward_exact_and_grads(ext, oracle):
    rng = np.random.Generator(np.random.PCG64(5))
    b, c, n, m, ns = 2, 19, 500, 64, 9
    feats = torch.from_numpy(rng.standard_normal((b, c, n)).astype(np.float32))
    idx1 = torch.from_numpy(rng.integers(0, n, (b, m)).astype(np.int32))
    idx2 = torch.from_numpy(rng.integers(0, n, (b, m, ns)).astype(np.int32))
    idx3 = torch.from_numpy(rng.integers(0, n, (b, 300, 3)).astype(np.int32))
    w3 = torch.from_numpy(rng.random((b, 300, 3)).astype(np.float32))
    O = oracle.ext
    dev = lambda t: t.cuda()
    assert torch.equal(ext.gather_points(dev(feats), dev(idx1)).cpu(), O.gather_points(feats, idx1))
    assert torch.equal(ext.group_points(dev(feats), dev(idx2)).cpu(), O.group_points(feats, idx2))
    assert torch.equal(ext.three_interpolate(dev(feats), dev(idx3), dev(w3)).cpu(), O.three_interpolate(feats, idx3, w3))
    g1 = torch.from_numpy(rng.standard_normal((b, c, m)).astype(np.float32))
    g2 = torch.from_numpy(rng.standard_normal((b, c, m, ns)).astype(np.float32))
    g3 = torch.from_numpy(rng.standard_normal((b, c, 300)).astype(np.float32))
    _assert_rel(ext.gather_points_grad(dev(g1), dev(idx1), n).cpu(), O.gather_points_grad(g1, idx1, n))
    _assert_rel(ext.group_points_grad(dev(g2), dev(idx2), n).cpu(), O.group_points_grad(g2, idx2, n))
    _assert_rel(ext.three_interpolate_grad(dev(g3), dev(idx3), dev(w3), n).cpu(),
                O.three_interpolate_grad(g3, idx3, w3, n))


def test_reference_gradcheck_case(ext):
    """pointnet2_test.py:18-30, the reference's only test, against our three_interpolate."""
    import pointnet2_utils
    torch.manual_seed(0)
    feats = torch.randn(1, 2, 4, requires_grad=True).float().cuda()
    idx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32).cuda()
    weight = torch.tensor([[[1, 1, 1], [2, 2, 2]]], dtype=torch.float32).cuda()
    assert torch.autograd.gradcheck(lambda f: pointnet2_utils.three_interpolate(f, idx, weight), feats,
                                    atol=1e-1, rtol=1e-1, eps=1e-2)


def test_query_and_group_matches_oracle(ext, oracle):
    import pointnet2_utils as U
    xyz, feats = oracle.uniform_cloud(2, 1024, 3, seed=0)
    inds = oracle.ext.furthest_point_sampling(xyz, 128)
    new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    want, want_xyz, _ = oracle.query_and_group(0.2, 32, xyz, new_xyz, feats, use_xyz=True, normalize_xyz=True)
    got, got_xyz = U.QueryAndGroup(0.2, 32, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=True)(
        xyz.cuda(), new_xyz.cuda(), feats.cuda())
    _assert_rel(got.cpu(), want, 1e-6)
    _assert_rel(got_xyz.cpu(), want_xyz, 1e-6)
