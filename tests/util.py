"""Shared synthetic-input builders for the parity tests (seeded; SURVEY.md section 8d distributions)."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_cases():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "splat_*.npz")))


def load_golden(path):
    g = np.load(path)
    d = {k: g[k] for k in g.files}
    for k in ("W", "K", "ksize"):
        d[k] = int(d[k])
    for k in ("radius_px", "tau"):
        d[k] = float(d[k])
    d["accumulation"] = str(d["accumulation"])
    return d


def demo_cameras(B, kind="translate", seed=0, views=None):
    """Cameras of reference demo.py:36-96 (offset*origK, identity extrinsics) with a target pose:
    'translate' = circle of z_buffermodel.py:214, 'rotate' = direction L (z_buffermodel.py:229-240),
    'identity', 'behind' (pushes points through the camera plane)."""
    rng = np.random.default_rng(seed)
    offset = np.array([[2, 0, -1], [0, -2, 1], [0, 0, -1]], np.float32)
    origK = np.array([[1, 0, .5], [0, 1, .5], [0, 0, 1]], np.float32)
    P = np.eye(4, dtype=np.float32)
    P[:3, :3] = offset @ origK
    Pinv = np.linalg.inv(P).astype(np.float32)
    K = np.eye(4, dtype=np.float32)
    RT1 = np.repeat(P[None], B, 0)
    RT1inv = np.repeat(Pinv[None], B, 0)
    RT2 = RT1.copy()
    for b in range(B):
        if kind == "translate":
            n = int(rng.integers(0, 8)) if views is None else int(views[b])
            RT2[b, :3, 3] += (.35 * np.array([np.sin(2 * np.pi * n / 8), np.cos(2 * np.pi * n / 8),
                                              .4 * np.sin(2 * np.pi * (.25 + n / 8))])).astype(np.float32)
        elif kind == "rotate":
            th = -0.6 * (b + 1) / B
            M = np.eye(4, dtype=np.float32)
            M[0, 0] = np.cos(th); M[0, 2] = np.sin(th); M[2, 0] = -np.sin(th); M[2, 2] = np.cos(th)
            RT2[b] = M @ RT1[b]
        elif kind == "behind":
            RT2[b, 2, 3] += 2.0
    RT2inv = np.linalg.inv(RT2).astype(np.float32)
    Ks = np.repeat(K[None], B, 0)
    return Ks, Ks.copy(), RT1, RT1inv, RT2, RT2inv


def pack_mats(K, Kinv, RT1, RT1inv, RT2, RT2inv):
    return np.ascontiguousarray(
        np.stack([np.asarray(m, np.float32).reshape(-1, 16) for m in (K, Kinv, RT1, RT1inv, RT2, RT2inv)], 1))


def synthetic_view(B, W, C=3, kind="translate", seed=0, min_z=0.5, max_z=10.0, depth_mode="uniform"):
    rng = np.random.default_rng(seed)
    if depth_mode == "uniform":
        depth = rng.uniform(min_z, max_z, (B, 1, W, W)).astype(np.float32)
    elif depth_mode == "const":
        depth = np.full((B, 1, W, W), 2.0, np.float32)
    elif depth_mode == "smooth":  # low-frequency surface, like a depth network output
        yy, xx = np.mgrid[0:W, 0:W].astype(np.float32) / W
        ph = rng.uniform(0, 6.28, (B, 4)).astype(np.float32)
        depth = np.stack([2.5 + np.sin(3 * xx + p[0]) * np.cos(2 * yy + p[1]) + 0.5 * np.sin(7 * yy + p[2]) * xx
                          + 0.3 * np.cos(5 * xx + p[3]) for p in ph])[:, None].astype(np.float32)
    else:
        depth = (np.round(rng.uniform(min_z, 4.0, (B, 1, W, W)) * 4) / 4).astype(np.float32)
    feat = rng.uniform(-1, 1, (B, C, W, W)).astype(np.float32)
    mats = pack_mats(*demo_cameras(B, kind, seed))
    return depth, feat, mats
