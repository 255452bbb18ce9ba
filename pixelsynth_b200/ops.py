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


# ------------------------------------------------------------------------------------------------
# PyTorch3D-shaped seam (reference: z_buffer_layers.py:81-84): rasterize_points -> (idx, zbuf, dist2)
# ------------------------------------------------------------------------------------------------
_libdef.define("rasterize_points_zbuf(Tensor pts, int S, float radius_px, int K, bool want_dist2) -> (Tensor, Tensor, Tensor)")


def _rasterize_points_zbuf(pts, S, radius_px, K, want_dist2):
    """pts (B,P,3) in project_pts' frame (x, y not yet negated, as ps_splat_points takes them) -> idx int32, zbuf f32,
    dist2 f32 (B,S,S,K), -1 padded, packed indices b*P+p: what pytorch3d.renderer.points.rasterize_points returns."""
    pts = _f32c(pts, "pts")
    if pts.dim() != 3 or pts.shape[2] != 3:
        raise RuntimeError(f"pts: expected (B,P,3), got {tuple(pts.shape)}")
    feat = torch.zeros((pts.shape[0], 1, pts.shape[1]), dtype=torch.float32, device=pts.device)
    _, _, idx, zbuf, d2 = _splat_points(pts, feat, S, K, radius_px, 1.0, 2, 0, 1, True, want_dist2)
    return idx, zbuf, d2


_impl.impl("rasterize_points_zbuf", _rasterize_points_zbuf)

# ------------------------------------------------------------------------------------------------
# cumulative cloud (reference: z_buffer_manipulator.py:184-266, forward_justpts_cumulative)
# ------------------------------------------------------------------------------------------------
_libdef.define(
    "splat_cumulative(Tensor depth, Tensor feat, Tensor mats, Tensor? prior_cloud, Tensor? prior_feat, "
    "Tensor? last_bg_mask, Tensor? RT3inv, int W, int S, int K, float radius_px, float tau, int rad_pow, "
    "int accumulation, int bg_ksize, float eps) -> (Tensor, Tensor, Tensor, Tensor)")


def _splat_cumulative(depth, feat, mats, prior_cloud, prior_feat, last_bg_mask, RT3inv, W, S, K, radius_px, tau, rad_pow,
                      accumulation, bg_ksize, eps):
    """-> (gen_fs (B,C,S,S), bg_mask (B,S,S) bool, new_cloud (B,4,P'), src (B,C,P')).  The current view contributes the
    pixels `last_bg_mask` marks (all of them for the first view); the prior cloud, kept pre-division in the previous
    target camera's frame, is carried over through RT2 . RT3inv.  Every image of the batch must select the same
    number of pixels (the reference's .view(bs, -1), z_buffer_manipulator.py:202)."""
    depth = _f32c(depth, "depth").reshape(-1, W * W)
    B = depth.shape[0]
    feat = _f32c(feat, "feat").reshape(B, -1, W * W)
    C = feat.shape[1]
    mats = _f32c(mats, "mats").reshape(B, 6, 4, 4)
    pts, xyp = _project_pts(depth, mats, W, eps, True)
    src = feat
    if prior_feat is not None:
        if last_bg_mask is None or prior_cloud is None or RT3inv is None:
            raise RuntimeError("splat_cumulative: prior_feat needs prior_cloud, last_bg_mask and RT3inv")
        sel = last_bg_mask.reshape(B, -1).to(torch.bool)
        pts = pts[sel].view(B, -1, 3)
        xyp = xyp.permute(0, 2, 1)[sel].view(B, -1, 4).permute(0, 2, 1)
        src = feat.permute(0, 2, 1)[sel].view(B, -1, C).permute(0, 2, 1)
        src = torch.cat([src, _f32c(prior_feat, "prior_feat").reshape(B, C, -1)], 2)
    if prior_cloud is not None:
        mats3 = torch.stack([mats[:, 0], mats[:, 4], _f32c(RT3inv, "RT3inv").reshape(B, 4, 4)], 1).contiguous()
        pts2, xyp2 = _project_cloud(_f32c(prior_cloud, "prior_cloud"), mats3, eps)
        pts = torch.cat([pts, pts2], 1)
        xyp = torch.cat([xyp, xyp2], 2)
    out, bg, _, _, _ = _splat_points(pts.contiguous(), src.contiguous(), S, K, radius_px, tau, rad_pow, accumulation,
                                     bg_ksize, False, False)
    return out, bg, xyp.contiguous(), src.contiguous()


_impl.impl("splat_cumulative", _splat_cumulative)

# ------------------------------------------------------------------------------------------------
# Networks.  A weight set is packed once (spectral norm / batch norm folded, bf16 or fp16 tiles in the kernels'
# layouts) and addressed by an integer handle, so the op signatures carry tensors and scalars only.
# ------------------------------------------------------------------------------------------------
_WEIGHTS = {}
_KINDS = ("depth_unet", "vqvae", "refine", "lmconv")


def register_weights(kind, state_dict, device="cuda", **kw):
    """Packs `state_dict` (the REFERENCE's key names for that sub-module: `pts_regressor.*`, `vqvae.*`, `projector.*`,
    `outpaint2.*` with the prefix stripped) for the kernels and returns the handle the network ops take."""
    from . import lmconv, nets

    if kind not in _KINDS:
        raise ValueError(f"register_weights: kind must be one of {_KINDS}, got {kind!r}")
    if not torch.cuda.is_available():
        raise RuntimeError("register_weights: packing uploads to the GPU; there is no CPU path")
    make = {"depth_unet": nets.UnetB200, "vqvae": nets.VQVAETopB200, "refine": nets.ResNetDecoderB200,
            "lmconv": lmconv.LmconvB200}[kind]
    runner = make(state_dict, device, **kw)
    handle = max(_WEIGHTS, default=0) + 1
    _WEIGHTS[handle] = (kind, runner)
    return handle


def register_runner(kind, runner):
    """Handle for an already-packed runner (ZbufferModelPts shares its networks with the ops this way)."""
    handle = max(_WEIGHTS, default=0) + 1
    _WEIGHTS[handle] = (kind, runner)
    return handle


def release_weights(handle):
    _WEIGHTS.pop(int(handle), None)


def _runner(handle, kind):
    ent = _WEIGHTS.get(int(handle))
    if ent is None or ent[0] != kind:
        raise RuntimeError(f"weights handle {handle} is not a registered {kind!r} weight set")
    return ent[1]


def _img(x, name, C):
    x = _f32c(x, name)
    if x.dim() != 4 or x.shape[1] != C or x.shape[2] != x.shape[3]:
        raise RuntimeError(f"{name}: expected (B,{C},S,S), got {tuple(x.shape)}")
    return x


_libdef.define("depth_unet(Tensor x, int weights, float min_z, float max_z) -> Tensor")


def _depth_unet(x, weights, min_z, max_z):
    """Unet.forward + sigmoid * (max_z - min_z) + min_z (architectures.py:230-279, z_buffermodel.py:304-308)."""
    with torch.cuda.device(x.device):
        return _runner(weights, "depth_unet").forward(_img(x, "x", 3), float(min_z), float(max_z))


_impl.impl("depth_unet", _depth_unet)

_libdef.define("vqvae_encode_top(Tensor x, int weights) -> Tensor")


def _vqvae_encode_top(x, weights):
    """VQVAETop.encode(x)[3] (vqvae.py:280-297): (B,3,256,256) f32 -> id_t (B,32,32) int64."""
    with torch.cuda.device(x.device):
        return _runner(weights, "vqvae").encode_top(_img(x, "x", 3))


_impl.impl("vqvae_encode_top", _vqvae_encode_top)

_libdef.define("vqvae_decode_code(Tensor ids, int weights) -> Tensor")


def _vqvae_decode_code(ids, weights):
    """VQVAETop.decode_code (vqvae.py:306-312): (B,32,32) int64 -> (B,3,256,256) f32."""
    if ids.dtype != torch.int64 or not ids.is_cuda or ids.dim() != 3:
        raise RuntimeError(f"ids: expected a (B,32,32) int64 CUDA tensor, got {tuple(ids.shape)} {ids.dtype} on {ids.device}")
    with torch.cuda.device(ids.device):
        return _runner(weights, "vqvae").decode_code(ids.contiguous())


_impl.impl("vqvae_decode_code", _vqvae_decode_code)

_libdef.define("refine_decode(Tensor x, Tensor background_mask, Tensor? noise, int weights) -> Tensor")


def _refine_decode(x, background_mask, noise, weights):
    """ResNetDecoder.forward(x, background_mask) with predict_residual (architectures.py:151-167).  noise (16,B,20):
    the z of the 16 LinearNoiseLayers in block order; None draws it with torch.randn like the reference."""
    x = _img(x, "x", 3)
    if background_mask.dtype != torch.bool or tuple(background_mask.shape) != (x.shape[0], x.shape[2], x.shape[3]):
        raise RuntimeError(f"background_mask: expected bool (B,S,S) matching x, got {tuple(background_mask.shape)} {background_mask.dtype}")
    if noise is not None and tuple(noise.shape) != (16, x.shape[0], 20):
        raise RuntimeError(f"noise: expected (16,{x.shape[0]},20), got {tuple(noise.shape)}")
    with torch.cuda.device(x.device):
        return _runner(weights, "refine").forward(x, background_mask, noise)


_impl.impl("refine_decode", _refine_decode)

_libdef.define("gen_order_masks(Tensor background_mask) -> (Tensor, Tensor, Tensor, Tensor)")


def _gen_order_masks(background_mask):
    """ZbufferModelPts.get_masks_for_batch (z_buffermodel.py:641-701) on the splat's (B,256,256) bool mask ->
    distances (B,32,32) i32, order (B,1024) i32 (cell r*32+c), tap-mask words (B,3,1024) i16 [A dil 1, B dil 1,
    B dil 2], sample_mask (B,32,32) bool.  Native HOST code (csrc/glue.cu): the mask crosses to the host once, as in the
    reference, and the four small results are returned as CPU tensors (the sampler's level scheduler reads them there)."""
    from . import lmconv

    if background_mask.dtype != torch.bool or background_mask.dim() != 3:
        raise RuntimeError(f"background_mask: expected bool (B,S,S), got {tuple(background_mask.shape)} {background_mask.dtype}")
    dist, order, words, smask = lmconv.glue_host(background_mask)
    return torch.from_numpy(dist), torch.from_numpy(order), torch.from_numpy(words.view("int16")), torch.from_numpy(smask)


_impl.impl("gen_order_masks", _gen_order_masks)

_libdef.define("lmconv_sample(Tensor codes, Tensor order, Tensor words, Tensor sample_mask, Tensor uniforms, "
               "float temperature, int weights) -> Tensor")


def _words_u16(words):
    w = words.detach().cpu().numpy()
    return w.view("uint16") if w.dtype.itemsize == 2 else w.astype("uint16")


def _lmconv_sample(codes, order, words, sample_mask, uniforms, temperature, weights):
    """sample() (lmconv/sample.py:8-73) as one launch: codes (B,32,32) int64 with the known cells -> codes with the
    sample_mask cells drawn in generation order; uniforms (B,>=n_sampled) f32 are the explicit categorical draws."""
    if codes.dtype != torch.int64 or not codes.is_cuda or codes.shape[-2:] != (32, 32):
        raise RuntimeError(f"codes: expected a (B,32,32) int64 CUDA tensor, got {tuple(codes.shape)} {codes.dtype} on {codes.device}")
    B = codes.shape[0]
    if order.numel() != B * 1024 or words.numel() != B * 3072 or sample_mask.numel() != B * 1024:
        raise RuntimeError("lmconv_sample: order (B,1024), words (B,3,1024), sample_mask (B,32,32) must match codes' batch")
    if uniforms.dim() != 2 or uniforms.shape[0] != B:
        raise RuntimeError(f"uniforms: expected (B,n) f32, got {tuple(uniforms.shape)}")
    with torch.cuda.device(codes.device):
        return _runner(weights, "lmconv").sample(codes, order, _words_u16(words), sample_mask, uniforms, float(temperature))


_impl.impl("lmconv_sample", _lmconv_sample)

_libdef.define("lmconv_logits(Tensor codes, Tensor order, Tensor words, int weights) -> Tensor")


def _lmconv_logits(codes, order, words, weights):
    """OurPixelCNN.forward under teacher forcing (lmconv/model.py:110-155): (B,32,32) int64 -> logits (B,512,32,32)."""
    if codes.dtype != torch.int64 or not codes.is_cuda or codes.shape[-2:] != (32, 32):
        raise RuntimeError(f"codes: expected a (B,32,32) int64 CUDA tensor, got {tuple(codes.shape)} {codes.dtype} on {codes.device}")
    with torch.cuda.device(codes.device):
        return _runner(weights, "lmconv").logits(codes, order, _words_u16(words))


_impl.impl("lmconv_logits", _lmconv_logits)

_libdef.define("combine(Tensor gen_fs, Tensor ar_sample, Tensor background_mask) -> Tensor")


def _combine(gen_fs, ar_sample, background_mask):
    """ZbufferModelPts.get_combined (z_buffermodel.py:703-708): gen_fs where foreground, ar_sample where background."""
    a, b = _f32c(gen_fs, "gen_fs"), _f32c(ar_sample, "ar_sample")
    if a.shape != b.shape or a.dim() != 4 or background_mask.dtype != torch.bool or \
            tuple(background_mask.shape) != (a.shape[0], a.shape[2], a.shape[3]):
        raise RuntimeError("combine: expected gen_fs, ar_sample (B,C,H,W) f32 and background_mask (B,H,W) bool")
    n, c, h, w = a.shape
    out = torch.empty_like(a)
    m = background_mask.contiguous().view(torch.uint8)
    with torch.cuda.device(a.device):
        check(_lib.lib().ps_combine(ptr(a), ptr(b), ptr(m), n, c, h * w, ptr(out), _stream()), "ps_combine")
    return out


_impl.impl("combine", _combine)
