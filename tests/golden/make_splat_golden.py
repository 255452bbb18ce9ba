"""Generates tests/golden/splat_*.npz by running the REFERENCE's own splat glue on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_splat_golden.py

What runs unmodified from /root/reference:
  models/projection/z_buffer_manipulator.py  PtsManipulator.__init__ (xyzs grid), project_pts, forward_justpts
  models/layers/z_buffer_layers.py           RasterizePointsXYsBlending.forward (negation, radius, alpha
                                             formula, permutes, 13x13 background dilation, compositor call)
What is substituted (absent third-party dependency, PyTorch3D 0.4.0 / 0.2.0@e3819a49, docs/INSTALL.md:11,59,74):
  pytorch3d.structures.Pointclouds, pytorch3d.renderer.points.rasterize_points,
  pytorch3d.renderer.compositing.alpha_composite|weighted_sum|weighted_sum_norm
  -> oracle.splat_ref.np_rasterize_points / np_alpha_composite (numpy brute force of the published
     algorithm).  `Tensor.cuda()` is patched to a no-op (z_buffer_layers.py:107).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import splat_ref  # noqa: E402

import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
from _ref_import import use_reference_models  # noqa: E402  (called from __main__ only: importing this file must not rebind `models`)

REF = "/root/reference"
CAPTURE = {}


class _Pointclouds:
    def __init__(self, points, features=None):
        self._points = points
        self._features = features

    def points_padded(self):
        return self._points

    def features_packed(self):
        f = self._features
        return f.reshape(-1, f.shape[-1])


def _rasterize_points(pointclouds, image_size=256, radius=0.01, points_per_pixel=8, bin_size=None,
                      max_points_per_bin=None):
    pts = pointclouds.points_padded().detach().cpu().numpy()
    idx, zbuf, d2 = splat_ref.np_rasterize_points(pts, image_size, radius, points_per_pixel)
    CAPTURE["idx"], CAPTURE["zbuf"], CAPTURE["dist2"] = idx.astype(np.int32), zbuf, d2
    return torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(zbuf), torch.from_numpy(d2)


def _alpha_composite(pointsidx, alphas, pt_clds):
    return torch.from_numpy(splat_ref.np_alpha_composite(pointsidx.numpy(), alphas.numpy(), pt_clds.numpy()))


def _weighted_sum(pointsidx, alphas, pt_clds, norm=False):
    i = pointsidx.numpy()
    a = np.where(i >= 0, alphas.numpy(), 0).astype(np.float32)
    f = pt_clds.numpy()[:, np.where(i >= 0, i, 0)]  # (C,N,K,H,W)
    out = (f * a[None]).sum(2, dtype=np.float32).transpose(1, 0, 2, 3)
    if norm:
        out = out / np.maximum(a.sum(1, dtype=np.float32), np.float32(1e-4))[:, None]
    return torch.from_numpy(np.ascontiguousarray(out, np.float32))


def install_stubs():
    p3d = types.ModuleType("pytorch3d")
    st = types.ModuleType("pytorch3d.structures")
    st.Pointclouds = _Pointclouds
    rd = types.ModuleType("pytorch3d.renderer")
    comp = types.ModuleType("pytorch3d.renderer.compositing")
    comp.alpha_composite = _alpha_composite
    comp.weighted_sum = lambda i, a, f: _weighted_sum(i, a, f, False)
    comp.weighted_sum_norm = lambda i, a, f: _weighted_sum(i, a, f, True)
    rp = types.ModuleType("pytorch3d.renderer.points")
    rp.rasterize_points = _rasterize_points
    rd.compositing = comp
    rd.points = rp
    for name, mod in (("pytorch3d", p3d), ("pytorch3d.structures", st), ("pytorch3d.renderer", rd),
                      ("pytorch3d.renderer.compositing", comp), ("pytorch3d.renderer.points", rp)):
        sys.modules[name] = mod
    torch.Tensor.cuda = lambda self, *a, **k: self
    os.environ["DEBUG"] = "False"
    if REF not in sys.path:
        sys.path.insert(0, REF)


def cameras(B, rng, kind):
    """demo.py:36-96 camera (offset * origK, identity extrinsics) and a target pose."""
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
        if kind == "translate":  # z_buffermodel.py:214 circle translation
            n = rng.integers(0, 8)
            RT2[b, :3, 3] += (.35 * np.array([np.sin(2 * np.pi * n / 8), np.cos(2 * np.pi * n / 8),
                                              .4 * np.sin(2 * np.pi * (.25 + n / 8))])).astype(np.float32)
        elif kind == "rotate":  # z_buffermodel.py:229-240, direction L, rotation .6
            th = -0.6 * (b + 1) / B
            M = np.eye(4, dtype=np.float32)
            M[0, 0] = np.cos(th); M[0, 2] = np.sin(th); M[2, 0] = -np.sin(th); M[2, 2] = np.cos(th)
            RT2[b] = M @ RT1[b]
        elif kind == "behind":  # push some points behind the camera / onto the EPS plane
            RT2[b, 2, 3] += 2.0
    RT2inv = np.linalg.inv(RT2).astype(np.float32)
    Ks = np.repeat(K[None], B, 0)
    return Ks, Ks.copy(), RT1, RT1inv, RT2, RT2inv


def run_case(name, W, K, radius_px, kind, B=2, C=3, tau=1.0, accumulation="alphacomposite", ksize=13, seed=0,
             depth_mode="uniform"):
    from models.projection.z_buffer_manipulator import PtsManipulator
    rng = np.random.default_rng(seed)
    opt = types.SimpleNamespace(splatter="xyblending", learn_default_feature=True, radius=radius_px, pp_pixel=K,
                                rad_pow=2, tau=tau, accumulation=accumulation, background_smoothing_kernel_size=ksize)
    torch.manual_seed(0)
    pm = PtsManipulator(W, C=C, opt=opt)
    if depth_mode == "uniform":
        depth = rng.uniform(0.5, 10.0, (B, 1, W, W)).astype(np.float32)
    elif depth_mode == "const":  # all-equal z: ordering is pure tie-break
        depth = np.full((B, 1, W, W), 2.0, np.float32)
    else:  # quantised: many exact z ties
        depth = (np.round(rng.uniform(0.5, 4.0, (B, 1, W, W)) * 4) / 4).astype(np.float32)
    feat = rng.uniform(-1, 1, (B, C, W, W)).astype(np.float32)
    mats = cameras(B, rng, kind)
    t = [torch.from_numpy(m) for m in mats]
    with torch.no_grad():
        pts = pm.project_pts(torch.from_numpy(depth).view(B, 1, -1), *t)
        gen_fs, bg = pm.forward_justpts(torch.from_numpy(feat), torch.from_numpy(depth), *t)
    out = dict(depth=depth, feat=feat, mats=splat_ref.pack_mats(*mats), W=W, K=K, radius_px=radius_px, tau=tau,
               ksize=ksize, accumulation=accumulation, xyzs=pm.xyzs.numpy(),
               ref_pts=pts.permute(0, 2, 1).contiguous().numpy(), ref_gen_fs=gen_fs.numpy(), ref_bg=bg.numpy(),
               idx=CAPTURE["idx"], zbuf=CAPTURE["zbuf"], dist2=CAPTURE["dist2"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"splat_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "pts", out["ref_pts"].shape, "hits/pixel mean", (out["idx"] >= 0).sum(-1).mean(),
          "empty", (out["idx"][..., 0] < 0).mean(), os.path.getsize(path) // 1024, "KiB")


def run_cumulative(name, W=32, K=16, radius_px=3.0, n_views=3, seed=5):
    """gen_scene's growing cloud: n_views chained calls of the reference's UNMODIFIED forward_justpts_cumulative
    (z_buffer_manipulator.py:184-219) / project_pts_cumulative (:221-266), batch 1 like forward_scene
    (z_buffermodel.py:491-504,555-568): view v's source camera is view v-1's target, only the pixels under the
    previous background mask are appended, the prior cloud is the stored pre-division xy_proj."""
    from models.projection.z_buffer_manipulator import PtsManipulator
    rng = np.random.default_rng(seed)
    opt = types.SimpleNamespace(splatter="xyblending", learn_default_feature=True, radius=radius_px, pp_pixel=K,
                                rad_pow=2, tau=1.0, accumulation="alphacomposite", background_smoothing_kernel_size=5)
    torch.manual_seed(0)
    pm = PtsManipulator(W, C=3, opt=opt)
    Ks, Kinvs, RT1, RT1inv, _, _ = cameras(1, rng, "identity")
    out = dict(W=W, K=K, radius_px=radius_px, ksize=5, n_views=n_views)
    prior, fs_old, last_bg, last_out_inv = None, None, None, None
    src_rt, src_inv = RT1, RT1inv
    for v in range(n_views):
        depth = rng.uniform(1.0, 6.0, (1, 1, W, W)).astype(np.float32)
        feat = rng.uniform(-1, 1, (1, 3, W, W)).astype(np.float32)
        dst_rt = RT1.copy()
        # a target that keeps moving: translation + a small rotation about y, so each view uncovers new pixels
        th = 0.25 * (v + 1)
        M = np.eye(4, dtype=np.float32)
        M[0, 0] = np.cos(th); M[0, 2] = np.sin(th); M[2, 0] = -np.sin(th); M[2, 2] = np.cos(th)
        dst_rt[0] = M @ RT1[0]
        dst_rt[0, :3, 3] += np.array([0.3 * (v + 1), -0.1 * v, 0.05], np.float32)
        dst_inv = np.linalg.inv(dst_rt).astype(np.float32)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))
        with torch.no_grad():
            gen_fs, bg, cloud, src = pm.forward_justpts_cumulative(
                t(feat), t(depth), t(Ks), t(Kinvs), t(src_rt), t(src_inv), t(dst_rt), t(dst_inv),
                None if prior is None else prior.clone(), None if fs_old is None else fs_old.clone(),
                None if last_bg is None else last_bg.clone(), t(last_out_inv))
        for k, a in (("depth", depth), ("feat", feat), ("src_rt", src_rt), ("src_inv", src_inv), ("dst_rt", dst_rt),
                     ("dst_inv", dst_inv), ("gen_fs", gen_fs.numpy()), ("bg", bg.numpy()), ("cloud", cloud.numpy()),
                     ("src", src.numpy()), ("idx", CAPTURE["idx"]), ("zbuf", CAPTURE["zbuf"])):
            out["v%d_%s" % (v, k)] = a
        print(name, "view", v, "cloud", tuple(cloud.shape), "appended", 0 if last_bg is None else int(last_bg.sum()),
              "empty", float((CAPTURE["idx"][..., 0] < 0).mean()))
        prior, fs_old, last_bg, last_out_inv = cloud, src, bg, dst_inv
        src_rt, src_inv = dst_rt, dst_inv
    out["K_mat"], out["Kinv_mat"] = Ks, Kinvs
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    use_reference_models()
    install_stubs()
    if "--cumulative-only" in sys.argv:
        run_cumulative("cumul_w32_k16")
        sys.exit(0)
    run_case("w32_k8_translate", 32, 8, 4.0, "translate")
    run_case("w32_k64_rotate", 32, 64, 2.0, "rotate")
    run_case("w32_k16_behind", 32, 16, 4.0, "behind")
    run_case("w32_k32_ties", 32, 32, 4.0, "identity", depth_mode="const")
    run_case("w32_k32_quant", 32, 32, 4.0, "translate", depth_mode="quant")
    run_case("w32_k16_wsum", 32, 16, 3.0, "translate", accumulation="wsum", tau=2.0, ksize=5)
    run_case("w32_k16_wsumnorm", 32, 16, 3.0, "rotate", accumulation="wsumnorm")
    run_case("w48_k128_translate", 48, 128, 4.0, "translate", B=1)
    run_cumulative("cumul_w32_k16")
