"""The warp-per-point 3-NN of fp_interpolate_warp_kernel (omni-pq_b200/csrc/mlp_aux.cu), restated on the CPU: every lane
scans the known points 32 apart with three_nn_kernel's strict-`<` insertion (interpolate_gpu.cu:14-64), then the warp
extracts three times the smallest (distance bits << 32 | index) key.  The result must be the triple the reference's
sequential scan ends with -- under ties, NaN and infinite distances and with fewer than three known points."""
import struct

import numpy as np

INF = np.float32(np.inf)


def _insert(b, i, x, k):
    if x < b[2]:
        if x < b[0]:
            return [x, b[0], b[1]], [k, i[0], i[1]]
        if x < b[1]:
            return [b[0], x, b[1]], [i[0], k, i[1]]
        return [b[0], b[1], x], [i[0], i[1], k]
    return b, i


def sequential(d):
    b, i = [INF, INF, INF], [0, 0, 0]
    for k, x in enumerate(d):
        b, i = _insert(b, i, x, k)
    return b, i


def _key(x, k):
    return (struct.unpack("<I", struct.pack("<f", float(x)))[0] << 32) | k


def warp_merge(d):
    lanes = []
    for lane in range(32):
        b, i = [INF, INF, INF], [0, 0, 0]
        for k in range(lane, len(d), 32):
            b, i = _insert(b, i, d[k], k)
        lanes.append((b, i))
    out_b, out_i = [], []
    for _ in range(3):
        keys = [_key(b[0], i[0]) for b, i in lanes]
        best = min(keys)
        out_b.append(struct.unpack("<f", struct.pack("<I", best >> 32))[0])
        out_i.append(best & 0xFFFFFFFF)
        lanes = [([b[1], b[2], INF], [i[1], i[2], 0]) if key == best else (b, i) for key, (b, i) in zip(keys, lanes)]
    return out_b, out_i


def test_warp_merge_equals_sequential_scan():
    rng = np.random.default_rng(0)
    for trial in range(2000):
        m = int(rng.integers(1, 200))
        mode = trial % 4
        if mode == 0:
            d = rng.random(m).astype(np.float32)
        elif mode == 1:
            d = rng.integers(0, 4, m).astype(np.float32)                                          # heavy ties
        elif mode == 2:
            d = np.where(rng.random(m) < 0.2, np.nan, rng.integers(0, 6, m)).astype(np.float32)   # NaN distances
        else:
            d = np.where(rng.random(m) < 0.3, np.inf, rng.integers(0, 3, m)).astype(np.float32)   # overflowed distances
        b1, i1 = sequential(d)
        b2, i2 = warp_merge(d)
        assert list(i1) == list(i2), (d, i1, i2)
        assert all(np.float32(x) == np.float32(y) for x, y in zip(b1, b2)), (d, b1, b2)
