"""CPU tests of the oracle itself: golden vectors, and independent restatements of the semantics."""
import numpy as np
import pytest
import torch

import cases
from oracle import pn2_oracle as O


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_matches_golden(name, golden):
    """The C oracle reproduces every frozen vector: the oracle-generated file always, the file generated
    by the reference's own CUDA kernels on the B200 box when it has been committed."""
    assert "oracle" in golden
    if name not in cases.SMALL and "ref" not in golden:
        pytest.skip("large cases are only re-run on CPU against the reference-generated goldens")
    outs, dig = cases.run_case(name, O.ext, "cpu")
    for kind, blob in golden.items():
        assert bytes(blob[f"{name}/input_digest"]).decode() == dig, "synthetic input generator drifted"
        for k, v in outs.items():
            ref = blob[f"{name}/{k}"]
            assert np.array_equal(v.numpy(), ref, equal_nan=True), f"{kind}:{name}/{k}"


def _bitrev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def _fps_rank_formula(xyz, m):
    """Independent FPS: numpy fp32 with the fmaf contraction emulated in float64 (exact for fp32
    products/sums of this size? no -- so use math.fma-free exact rational trick: float64 holds the
    exact product of two float32, and one rounding to float32 after the add equals fmaf)."""
    n = xyz.shape[0]
    bs = O.ext.opt_n_threads(n)
    bits = bs.bit_length() - 1
    x = xyz.astype(np.float32)

    def sq3(a, b, c):
        t = (b.astype(np.float32) * b.astype(np.float32)).astype(np.float32)
        t = (a.astype(np.float64) * a.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
        return (c.astype(np.float64) * c.astype(np.float64) + t.astype(np.float64)).astype(np.float32)

    mag = sq3(x[:, 0], x[:, 1], x[:, 2])
    live = ~(mag.astype(np.float64) <= 1e-3)
    rank = np.array([(_bitrev(k % bs, bits), k // bs) for k in range(n)])
    order = np.lexsort((rank[:, 1], rank[:, 0]))  # ascending tie-break preference
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    temp = np.full(n, 1e10, dtype=np.float32)
    out = np.zeros(m, dtype=np.int32)
    old = 0
    for j in range(1, m):
        d = sq3(x[:, 0] - x[old, 0], x[:, 1] - x[old, 1], x[:, 2] - x[old, 2])
        temp = np.where(live, np.minimum(d, temp), temp)
        if not live.any():
            old = 0
        else:
            best = temp[live].max()
            cand = np.nonzero(live & (temp == best))[0]
            old = int(cand[np.argmin(pos[cand])])
        out[j] = old
    return out


@pytest.mark.parametrize("name", ["c1_uniform", "c1_duplicates", "c1_near_origin", "c1_grid", "c1_small"])
def test_fps_literal_tree_equals_rank_formula(name):
    """The oracle emulates the reference's shared-memory tree literally; the closed-form tie-break
    (bit-reversed slot, then stripe) used by the B200 kernel must select the same points."""
    make, stages = cases.CASES[name]
    xyz = make()
    m = stages[0][1]
    got = O.ext.furthest_point_sampling(xyz, m).numpy()
    for b in range(xyz.shape[0]):
        want = _fps_rank_formula(xyz[b].numpy(), m)
        assert np.array_equal(got[b], want), f"cloud {b}"


def test_fps_skips_near_origin_points():
    xyz = cases.c1_near_origin()
    inds = O.ext.furthest_point_sampling(xyz, 256).numpy()
    mag = (xyz.double() ** 2).sum(-1).numpy()
    for b in range(2):
        sel = inds[b, 1:]
        assert (mag[b, sel] > 1e-3).all(), "a skipped point was selected after slot 0"
        assert inds[b, 0] == 0


def _sq3_np(a, b, c):
    """fp32 a*a+b*b+c*c with the reference contraction; float64 holds fp32 products exactly, so one
    rounding after the add equals fmaf."""
    t = (b * b).astype(np.float32)
    t = (a.astype(np.float64) * a.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
    return (c.astype(np.float64) * c.astype(np.float64) + t.astype(np.float64)).astype(np.float32)


def test_ball_query_semantics_small():
    """First nsample in-ball indices in ascending order, padded with the first hit, zeros if empty
    (ball_query_gpu.cu:32-46) -- against an independent numpy evaluation of the same fp32 arithmetic."""
    xyz, _ = O.uniform_cloud(1, 300, 3, seed=9)
    centres = xyz[:, :20].contiguous()
    ns = 12
    idx = O.ext.ball_query(centres, xyz, 0.25, ns).numpy()[0]
    x = xyz[0].numpy()
    r2 = np.float32(0.25) * np.float32(0.25)
    saw_partial = saw_full = False
    for j in range(20):
        d2 = _sq3_np(x[j, 0] - x[:, 0], x[j, 1] - x[:, 1], x[j, 2] - x[:, 2])
        hits = np.nonzero(d2 < r2)[0]
        want = np.full(ns, hits[0], dtype=np.int32)
        want[:min(ns, len(hits))] = hits[:ns]
        assert np.array_equal(idx[j], want), j
        saw_partial |= len(hits) < ns
        saw_full |= len(hits) >= ns
    assert saw_partial and saw_full
    far = (xyz[:, :4] + 10.0).contiguous()
    assert (O.ext.ball_query(far, xyz, 0.25, ns) == 0).all(), "empty balls stay zero"


def test_three_nn_matches_sort_and_ties_keep_lower_index():
    xyz, _ = O.uniform_cloud(1, 400, 3, seed=10)
    known = xyz[:, :50].contiguous()
    d2, idx = O.ext.three_nn(xyz, known)
    full = ((xyz[0, :, None, :].double() - known[0, None, :, :].double()) ** 2).sum(-1)
    top = torch.topk(full, 3, dim=1, largest=False).indices
    assert (idx[0, 50:].long() == top[50:]).float().mean() > 0.999
    assert torch.all(d2[0, :, 0] <= d2[0, :, 1]) and torch.all(d2[0, :, 1] <= d2[0, :, 2])
    dup = torch.cat([known, known], dim=1)  # every distance appears twice: ties -> lower index first
    _, idx2 = O.ext.three_nn(xyz, dup)
    assert torch.all(idx2[0, :, 0] < 50) and torch.all(idx2[0, :, 1] == idx2[0, :, 0] + 50)
    d2s, idxs = O.ext.three_nn(xyz, known[:, :2].contiguous())  # m < 3: +inf / 0 in the unused slot
    assert torch.isinf(d2s[0, :, 2]).all() and (idxs[0, :, 2] == 0).all()


def test_group_gather_interpolate_grads_are_transposes():
    """Scatter-add gradients are the exact transposes of the gathers (checked through autograd)."""
    torch.manual_seed(0)
    feats = torch.randn(2, 5, 40, requires_grad=True)
    idx = torch.randint(0, 40, (2, 7, 3), dtype=torch.int32)
    out = O.grouping_operation(feats, idx)
    assert torch.equal(out, torch.stack([feats[b][:, idx[b].long()] for b in range(2)]))
    g = torch.randn_like(out)
    out.backward(g)
    want = torch.zeros(2, 5, 40)
    for b in range(2):
        want[b].index_add_(1, idx[b].reshape(-1).long(), g[b].reshape(5, -1))
    assert torch.allclose(feats.grad, want, atol=1e-6)
    w = torch.rand(2, 7, 3)
    feats.grad = None
    o2 = O.three_interpolate(feats, idx, w)
    want2 = sum(torch.stack([feats[b][:, idx[b, :, t].long()] for b in range(2)]) * w[:, None, :, t] for t in range(3))
    assert torch.allclose(o2, want2, atol=1e-6)


def test_reference_gradcheck_case_on_oracle():
    """pointnet2_test.py:18-30 (the reference's only test): gradcheck of three_interpolate with its
    fixed idx / weight, same tolerances, here on the CPU oracle."""
    torch.manual_seed(0)
    feats = torch.randn(1, 2, 4, requires_grad=True).float()
    idx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32)
    weight = torch.tensor([[[1, 1, 1], [2, 2, 2]]], dtype=torch.float32)
    assert torch.autograd.gradcheck(lambda f: O.three_interpolate(f, idx, weight), feats, atol=1e-1, rtol=1e-1,
                                    eps=1e-2)
