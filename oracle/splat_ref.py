"""ctypes front-end to oracle/splat_oracle.c plus an independent numpy brute-force restatement.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Parity unpinned by the reference (no tests, PyTorch3D not vendored) -- see splat_oracle.c header.
Reference lines followed: models/projection/z_buffer_manipulator.py:38-83,184-266,
models/layers/z_buffer_layers.py:55-131, SURVEY.md Appendix A.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ACCUM = {"alphacomposite": 0, "wsum": 1, "wsumnorm": 2}


def build(force=False):
    so = os.path.join(_HERE, "_build", "libps_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def pack_mats(K, Kinv, RT1, RT1inv, RT2, RT2inv):
    """(B,4,4) x6 -> (B,6,16) f32 in forward_justpts argument order."""
    return _f32(np.stack([np.asarray(m, dtype=np.float32).reshape(-1, 16) for m in (K, Kinv, RT1, RT1inv, RT2, RT2inv)], 1))


def project(depth, mats, W, eps=1e-2, want_xyproj=False):
    """depth (B,P) or (B,1,W,W); mats (B,6,16) -> pts (B,P,3) [, xyproj (B,4,P)]."""
    depth = _f32(depth).reshape(-1, W * W)
    mats = _f32(mats)
    B = depth.shape[0]
    pts = np.empty((B, W * W, 3), np.float32)
    xyp = np.empty((B, 4, W * W), np.float32) if want_xyproj else None
    rc = lib().pso_project(_p(depth, ctypes.c_float), _p(mats, ctypes.c_float), B, W, ctypes.c_float(eps),
                           _p(pts, ctypes.c_float), _p(xyp, ctypes.c_float))
    assert rc == 0, rc
    return (pts, xyp) if want_xyproj else pts


def project_cloud(cloud, mats3, eps=1e-2):
    """cloud (B,4,P) homogeneous prior cloud; mats3 (B,3,16) = [K, RT2, RT3inv] -> pts (B,P,3), xyproj (B,4,P)."""
    cloud = _f32(cloud)
    mats3 = _f32(mats3)
    B, _, P = cloud.shape
    pts = np.empty((B, P, 3), np.float32)
    xyp = np.empty((B, 4, P), np.float32)
    rc = lib().pso_project_cloud(_p(cloud, ctypes.c_float), _p(mats3, ctypes.c_float), B, P, ctypes.c_float(eps),
                                 _p(pts, ctypes.c_float), _p(xyp, ctypes.c_float))
    assert rc == 0, rc
    return pts, xyp


def rasterize(pts, S, K, radius, naive=False, want_dist2=True):
    """pts (B,P,3) in project_pts' frame -> idx int32, zbuf f32, dist2 f32, each (B,S,S,K)."""
    pts = _f32(pts)
    B, P, _ = pts.shape
    idx = np.empty((B, S, S, K), np.int32)
    zbuf = np.empty((B, S, S, K), np.float32)
    d2 = np.empty((B, S, S, K), np.float32) if want_dist2 else None
    fn = lib().pso_rasterize_naive if naive else lib().pso_rasterize
    rc = fn(_p(pts, ctypes.c_float), B, P, S, K, ctypes.c_double(radius), _p(idx, ctypes.c_int32),
            _p(zbuf, ctypes.c_float), _p(d2, ctypes.c_float))
    assert rc == 0, rc
    return idx, zbuf, d2


def composite(idx, dist2, feat, radius, rad_pow=2, tau=1.0, accumulation="alphacomposite"):
    """feat (B,C,P) -> out (B,C,S,S)."""
    idx = np.ascontiguousarray(idx, np.int32)
    dist2 = _f32(dist2)
    feat = _f32(feat)
    B, S, _, K = idx.shape
    _, C, P = feat.shape
    out = np.empty((B, C, S, S), np.float32)
    rc = lib().pso_composite(_p(idx, ctypes.c_int32), _p(dist2, ctypes.c_float), _p(feat, ctypes.c_float), B, P, C, S, K,
                             ctypes.c_double(radius), rad_pow, ctypes.c_double(tau), ACCUM[accumulation],
                             _p(out, ctypes.c_float))
    assert rc == 0, rc
    return out


def bgmask(idx, ksize=13):
    idx = np.ascontiguousarray(idx, np.int32)
    B, S, _, K = idx.shape
    bg = np.empty((B, S, S), np.uint8)
    rc = lib().pso_bgmask(_p(idx, ctypes.c_int32), B, S, K, ksize, _p(bg, ctypes.c_uint8))
    assert rc == 0, rc
    return bg.astype(bool)


def splat(depth, feat, mats, W, S=None, K=128, radius_px=4.0, tau=1.0, rad_pow=2, bg_ksize=13,
          accumulation="alphacomposite"):
    """Whole S1-S5 path: forward_justpts.  Returns dict(out, bg, idx, zbuf, dist2, pts)."""
    S = S or W
    radius = float(radius_px) / float(S) * 2.0  # z_buffer_layers.py:77
    pts = project(depth, mats, W)
    idx, zbuf, d2 = rasterize(pts, S, K, radius)
    feat = _f32(feat).reshape(pts.shape[0], -1, W * W)
    out = composite(idx, d2, feat, radius, rad_pow, tau, accumulation)
    return dict(out=out, bg=bgmask(idx, bg_ksize), idx=idx, zbuf=zbuf, dist2=d2, pts=pts)


# ---------------------------------------------------------------------------------------------
# Independent numpy brute force of the published PyTorch3D point rasteriser (naive CPU semantics)
# and compositor; also used as the `pytorch3d` stand-in by tests/golden/make_splat_golden.py.
# ---------------------------------------------------------------------------------------------

def np_rasterize_points(points, S, radius, K):
    """points (N,P,3) *in PyTorch3D's frame* (i.e. after the layer negated x,y).  Returns
    idx (N,S,S,K) int64 packed, zbuf, dist2 -- rasterize_points' return triple."""
    points = np.asarray(points, np.float32)
    N, P, _ = points.shape
    rf = np.float32(radius)
    r2 = np.float32(rf * rf)
    idx = -np.ones((N, S, S, K), np.int64)
    zbuf = -np.ones((N, S, S, K), np.float32)
    d2o = -np.ones((N, S, S, K), np.float32)
    ii = np.arange(S, dtype=np.float32)
    ndc = np.float32(-1.0) + (np.float32(2.0) * ii + np.float32(1.0)) / np.float32(S)  # PixToNdc
    for n in range(N):
        px, py, pz = points[n, :, 0], points[n, :, 1], points[n, :, 2]
        front = pz >= 0
        for yi in range(S):
            yf = ndc[S - 1 - yi]
            dy = (py - yf).astype(np.float32)
            dy2 = (dy * dy).astype(np.float32)
            rowmask = front & (dy2 < r2)
            cand = np.nonzero(rowmask)[0]
            if cand.size == 0:
                continue
            for xi in range(S):
                xf = ndc[S - 1 - xi]
                dx = (px[cand] - xf).astype(np.float32)
                d2 = ((dx * dx).astype(np.float32) + dy2[cand]).astype(np.float32)
                hit = d2 < r2
                if not hit.any():
                    continue
                hp = cand[hit]
                order = np.lexsort((hp, pz[hp]))[:K]
                k = order.size
                idx[n, yi, xi, :k] = hp[order] + n * P
                zbuf[n, yi, xi, :k] = pz[hp][order]
                d2o[n, yi, xi, :k] = d2[hit][order]
    return idx, zbuf, d2o


def np_alpha_composite(pointsidx, alphas, features):
    """pointsidx (N,K,H,W) int64, alphas (N,K,H,W), features (C, sumP) -> (N,C,H,W)."""
    pointsidx = np.asarray(pointsidx)
    alphas = np.asarray(alphas, np.float32)
    features = np.asarray(features, np.float32)
    N, K, H, W = pointsidx.shape
    C = features.shape[0]
    out = np.zeros((N, C, H, W), np.float32)
    cum = np.ones((N, H, W), np.float32)
    for k in range(K):
        i = pointsidx[:, k]
        valid = i >= 0
        a = np.where(valid, alphas[:, k], np.float32(0)).astype(np.float32)
        f = features[:, np.where(valid, i, 0)]  # (C,N,H,W)
        for c in range(C):
            term = ((f[c] * cum).astype(np.float32) * a).astype(np.float32)
            out[:, c] = (out[:, c] + np.where(valid, term, np.float32(0))).astype(np.float32)
        cum = (cum * (np.float32(1) - a)).astype(np.float32)
    return out
