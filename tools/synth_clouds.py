"""Synthetic input clouds, numpy PCG64 so every implementation sees identical bits.

Neutral module: shared by bench.py, the parity tests and the oracle; it contains no implementation of the
hot path (input generation only)."""
import numpy as np
import torch


def uniform_cloud(b, n, c_feat=3, seed=0):
    """C1: xyz ~ U[0,1)^3, feats ~ N(0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xyz = rng.random((b, n, 3), dtype=np.float64).astype(np.float32)
    feats = rng.standard_normal((b, c_feat, n)).astype(np.float32)
    return torch.from_numpy(xyz), torch.from_numpy(feats)


def scannet_like_cloud(n=40000, seed=1234, c_feat=3, centred=False, yaw=False):
    """C2/C5: points on the walls/floor/ceiling of an axis-aligned room plus 5-20 furniture boxes,
    area-proportional sampling, 5 mm jitter, random point order.  Returns (n, 3+c_feat) float32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    W, L, H = rng.uniform(3, 8), rng.uniform(3, 8), rng.uniform(2.4, 3.0)
    boxes = [(0.0, 0.0, 0.0, W, L, H)]
    for _ in range(int(rng.integers(5, 21))):
        sx, sy, sz = rng.uniform(0.3, 2.0), rng.uniform(0.3, 2.0), rng.uniform(0.3, 1.5)
        ox, oy = rng.uniform(0, max(W - sx, 0.1)), rng.uniform(0, max(L - sy, 0.1))
        boxes.append((ox, oy, 0.0, sx, sy, sz))
    faces = []  # (origin, u, v, area)
    for (ox, oy, oz, sx, sy, sz) in boxes:
        o = np.array([ox, oy, oz])
        ex, ey, ez = np.array([sx, 0, 0]), np.array([0, sy, 0]), np.array([0, 0, sz])
        for (p, u, v) in [(o, ex, ey), (o + ez, ex, ey), (o, ex, ez), (o + ey, ex, ez), (o, ey, ez), (o + ex, ey, ez)]:
            faces.append((p, u, v, np.linalg.norm(np.cross(u, v))))
    areas = np.array([f[3] for f in faces])
    which = rng.choice(len(faces), size=n, p=areas / areas.sum())
    uv = rng.random((n, 2))
    P = np.stack([faces[w][0] + uv[i, 0] * faces[w][1] + uv[i, 1] * faces[w][2] for i, w in enumerate(which)])
    P = P + rng.normal(0.0, 0.005, size=P.shape)
    if centred:
        P = P - np.array([W / 2, L / 2, 0.0])
    if yaw:
        a = rng.uniform(0, 2 * np.pi)
        R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        P = P @ R.T
    P = P[rng.permutation(n)]
    col = rng.uniform(-0.5, 0.5, size=(n, c_feat))
    return torch.from_numpy(np.concatenate([P, col], axis=1).astype(np.float32))
