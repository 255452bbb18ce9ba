"""torch.ops.pixelsynth_b200.* : thin registrations over the C ABI (include/pixelsynth_b200.h).

Each op validates shapes/dtypes/device (the reference's `assert`s / PyTorch3D's checks become
RuntimeError), allocates outputs and scratch with torch's caching allocator on the current device,
and enqueues the kernels on torch's current CUDA stream.  Only the CUDA dispatch key is registered:
CPU tensors raise, there is no fallback path.
"""
import torch

from . import _lib
from ._lib import check, ptr

NS = "pixelsynth_b200"
_libdef = torch.library.Library(NS, "DEF")
_impl = torch.library.Library(NS, "IMPL", "CUDA")

ACCUMULATION = {"alphacomposite": 0, "wsum": 1, "wsumnorm": 2}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t, name):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a float32 CUDA tensor, got {t.dtype} on {t.device}")
    return t.contiguous()


def _workspace(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# projection (reference: models/projection/z_buffer_manipulator.py:50-83, 221-266)
# ------------------------------------------------------------------------------------------------
_libdef.define("project_pts(Tensor depth, Tensor mats, int W, float eps, bool want_xyproj) -> (Tensor, Tensor)")


def _project_pts(depth, mats, W, eps, want_xyproj):
    depth = _f32c(depth, "depth").reshape(-1, W * W)
    mats = _f32c(mats, "mats")
    B = depth.shape[0]
    if mats.numel() != B * 96:
        raise RuntimeError(f"mats: expected (B,6,4,4) with B={B}, got {tuple(mats.shape)}")
    with torch.cuda.device(depth.device):
        pts = torch.empty((B, W * W, 3), dtype=torch.float32, device=depth.device)
        xyp = torch.empty((B, 4, W * W), dtype=torch.float32, device=depth.device) if want_xyproj else None
        check(_lib.lib().ps_project_pts(ptr(depth), ptr(mats), B, W, eps, ptr(pts), ptr(xyp), _stream()), "ps_project_pts")
    return pts, (xyp if want_xyproj else pts.new_empty(0))


_impl.impl("project_pts", _project_pts)

_libdef.define("project_cloud(Tensor cloud, Tensor mats3, float eps) -> (Tensor, Tensor)")


def _project_cloud(cloud, mats3, eps):
    cloud = _f32c(cloud, "cloud")
    mats3 = _f32c(mats3, "mats3")
    B, four, P = cloud.shape
    if four != 4 or mats3.numel() != B * 48:
        raise RuntimeError("project_cloud: expected cloud (B,4,P) and mats3 (B,3,4,4)")
    with torch.cuda.device(cloud.device):
        pts = torch.empty((B, P, 3), dtype=torch.float32, device=cloud.device)
        xyp = torch.empty((B, 4, P), dtype=torch.float32, device=cloud.device)
        check(_lib.lib().ps_project_cloud(ptr(cloud), ptr(mats3), B, P, eps, ptr(pts), ptr(xyp), _stream()),
              "ps_project_cloud")
    return pts, xyp


_impl.impl("project_cloud", _project_cloud)

# ------------------------------------------------------------------------------------------------
# rasterise + composite (reference: models/layers/z_buffer_layers.py:55-131)
# ------------------------------------------------------------------------------------------------
_libdef.define(
    "splat_points(Tensor pts, Tensor feat, int S, int K, float radius_px, float tau, int rad_pow, int accumulation, "
    "int bg_ksize, bool want_maps, bool want_dist2) -> (Tensor, Tensor, Tensor, Tensor, Tensor)")


def _alloc_outputs(B, C, S, K, device, want_maps, want_dist2):
    out = torch.empty((B, C, S, S), dtype=torch.float32, device=device)
    bg = torch.empty((B, S, S), dtype=torch.uint8, device=device)
    idx = torch.empty((B, S, S, K), dtype=torch.int32, device=device) if want_maps else None
    zbuf = torch.empty((B, S, S, K), dtype=torch.float32, device=device) if want_maps else None
    d2 = torch.empty((B, S, S, K), dtype=torch.float32, device=device) if want_dist2 else None
    return out, bg, idx, zbuf, d2


def _ret(out, bg, idx, zbuf, d2):
    e = out.new_empty(0)
    return out, bg.view(torch.bool), (idx if idx is not None else e.to(torch.int32)), (zbuf if zbuf is not None else e), \
        (d2 if d2 is not None else e)


def _splat_points(pts, feat, S, K, radius_px, tau, rad_pow, accumulation, bg_ksize, want_maps, want_dist2):
    pts = _f32c(pts, "pts")
    feat = _f32c(feat, "feat")
    if pts.dim() != 3 or pts.shape[2] != 3:
        raise RuntimeError(f"pts: expected (B,P,3), got {tuple(pts.shape)}")  # z_buffer_layers.py:68
    B, P, _ = pts.shape
    feat = feat.reshape(B, -1, P) if feat.dim() != 3 else feat
    if feat.shape[0] != B or feat.shape[2] != P:
        raise RuntimeError(f"feat: expected (B,C,P) matching pts, got {tuple(feat.shape)}")  # z_buffer_layers.py:69
    C = feat.shape[1]
    with torch.cuda.device(pts.device):
        L = _lib.lib()
        out, bg, idx, zbuf, d2 = _alloc_outputs(B, C, S, K, pts.device, want_maps, want_dist2)
        nb = L.ps_splat_workspace_bytes(B, P, S, radius_px)
        ws = _workspace(nb, pts.device)
        check(L.ps_splat_points(ptr(pts), ptr(feat), B, P, C, S, K, radius_px, tau, rad_pow, accumulation, bg_ksize,
                                ptr(out), ptr(bg), ptr(idx), ptr(zbuf), ptr(d2), ptr(ws), nb, _stream()),
              "ps_splat_points")
    return _ret(out, bg, idx, zbuf, d2)


_impl.impl("splat_points", _splat_points)

_libdef.define(
    "splat(Tensor depth, Tensor feat, Tensor mats, int W, int S, int K, float radius_px, float tau, int rad_pow, "
    "int accumulation, int bg_ksize, float eps, bool want_maps, bool want_dist2) -> (Tensor, Tensor, Tensor, Tensor, Tensor)")


def _splat(depth, feat, mats, W, S, K, radius_px, tau, rad_pow, accumulation, bg_ksize, eps, want_maps, want_dist2):
    depth = _f32c(depth, "depth").reshape(-1, W * W)
    B = depth.shape[0]
    feat = _f32c(feat, "feat").reshape(B, -1, W * W)
    mats = _f32c(mats, "mats")
    if mats.numel() != B * 96:
        raise RuntimeError(f"mats: expected (B,6,4,4) with B={B}, got {tuple(mats.shape)}")
    C = feat.shape[1]
    with torch.cuda.device(depth.device):
        L = _lib.lib()
        out, bg, idx, zbuf, d2 = _alloc_outputs(B, C, S, K, depth.device, want_maps, want_dist2)
        nb = L.ps_splat_fwd_workspace_bytes(B, W, S, radius_px)
        ws = _workspace(nb, depth.device)
        check(L.ps_splat_fwd(ptr(depth), ptr(feat), ptr(mats), B, W, C, S, K, radius_px, tau, rad_pow, accumulation,
                             bg_ksize, eps, ptr(out), ptr(bg), ptr(idx), ptr(zbuf), ptr(d2), ptr(ws), nb, _stream()),
              "ps_splat_fwd")
    return _ret(out, bg, idx, zbuf, d2)


_impl.impl("splat", _splat)


def pack_mats(K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2):
    """(B,4,4) x6 in forward_justpts argument order -> (B,6,4,4) contiguous f32."""
    ms = [K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2]
    ref = next(m for m in ms if m is not None)
    ms = [ref if m is None else m for m in ms]  # RTinv_cam2 may be None (forward_angle) and is never read
    return torch.stack([m.to(torch.float32) for m in ms], 1).contiguous()
