"""Mirror of reference models/projection/z_buffer_manipulator.py:11-294 (PtsManipulator).

Same class / method names and argument meaning.  project_pts* run ps_project_pts / ps_project_cloud;
forward_justpts runs the fused ps_splat_fwd; forward_justpts_cumulative reproduces the growing-cloud
bookkeeping of the reference (only newly outpainted pixels are appended) around ps_splat_points.
"""
import torch
import torch.nn as nn

from .. import _ops_loaded  # noqa: F401
from ...ops import ACCUMULATION, pack_mats

EPS = 1e-2


def get_splatter(name, depth_values, opt=None, size=256, C=64, points_per_pixel=8):
    if name == "xyblending":
        from ..layers.z_buffer_layers import RasterizePointsXYsBlending

        return RasterizePointsXYsBlending(C, learn_feature=opt.learn_default_feature, radius=opt.radius, size=size,
                                          points_per_pixel=points_per_pixel, opts=opt)
    raise NotImplementedError()


class PtsManipulator(nn.Module):
    def __init__(self, W, C=64, opt=None):
        super().__init__()
        self.opt = opt
        self.W = W
        self.splatter = get_splatter(opt.splatter, None, opt, size=W, C=C, points_per_pixel=opt.pp_pixel)
        # the grid is generated inside the kernel; the buffer is kept so reference state_dicts load
        xs = torch.linspace(0, W - 1, W) / float(W - 1) * 2 - 1
        ys = torch.linspace(0, W - 1, W) / float(W - 1) * 2 - 1
        xs = xs.view(1, 1, 1, W).repeat(1, 1, W, 1)
        ys = ys.view(1, 1, W, 1).repeat(1, 1, 1, W)
        xyzs = torch.cat((xs, -ys, -torch.ones(xs.size()), torch.ones(xs.size())), 1).view(1, 4, -1)
        self.register_buffer("xyzs", xyzs)

    # -- z_buffer_manipulator.py:50-83 --------------------------------------------------------
    def project_pts(self, pts3D, K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2):
        mats = pack_mats(K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2)
        pts, _ = torch.ops.pixelsynth_b200.project_pts(pts3D, mats, self.W, EPS, False)
        return pts.permute(0, 2, 1)  # (B,3,P) like the reference's `sampler`

    # -- z_buffer_manipulator.py:85-107 -------------------------------------------------------
    def forward_justpts(self, src, pred_pts, K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2, return_maps=False):
        bs, c, w, h = src.size()
        o = self.opt
        mats = pack_mats(K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2)
        out, bg, idx, zbuf, d2 = torch.ops.pixelsynth_b200.splat(
            pred_pts, src, mats, self.W, self.W, int(o.pp_pixel), float(o.radius), float(o.tau), int(o.rad_pow),
            ACCUMULATION[o.accumulation], int(o.background_smoothing_kernel_size), EPS, return_maps, return_maps)
        if return_maps:
            return out, bg, idx, zbuf, d2
        return out, bg

    # -- z_buffer_manipulator.py:221-266 ------------------------------------------------------
    @staticmethod
    def _compact(sel, x):
        """Per-image stream compaction for a batch whose images keep DIFFERENT numbers of points: x (B,C,P), sel (B,P)
        bool -> (B,C,n_max) with image b's selected columns first, in order, and ZERO columns behind them.  A zero
        homogeneous point stays zero under every camera matrix, so |z| < EPS parks it at (-10, 10, 10) by the
        reference's own rule (z_buffer_manipulator.py:250-261): padding never reaches the rasteriser, and with equal
        counts (batch 1 always) the result is exactly the reference's boolean-mask `.view(bs, c, -1)`."""
        bs, c, _ = x.shape
        n = sel.sum(1)
        n_max = int(n.max())            # one small device->host read per view (the reference reads three)
        rank = sel.cumsum(1) - 1
        dst = torch.where(sel, rank, torch.full_like(rank, n_max))
        out = x.new_zeros((bs, c, n_max + 1))
        out.scatter_(2, dst.unsqueeze(1).expand(bs, c, -1), x)
        return out[:, :, :n_max].contiguous()

    def project_pts_cumulative(self, pts3D, K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2, prior_point_cloud=None,
                               last_background_mask=None, RTinv_cam3=None):
        """pts3D: full-grid depth (B,1,P); the reference passes the already-masked depth, here the mask is
        applied after projecting the full grid (per-point arithmetic is identical)."""
        mats = pack_mats(K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2)
        pts, xyp = torch.ops.pixelsynth_b200.project_pts(pts3D, mats, self.W, EPS, True)
        bs = pts.shape[0]
        if last_background_mask is not None:
            sel = last_background_mask.view(bs, -1)
            xyp = self._compact(sel, xyp)
            # the padded tail must come out parked like any |z| < EPS point: (x, y, z) = (-10, 10, 10)
            pts = self._compact(sel, pts.permute(0, 2, 1))
            tail = torch.arange(pts.shape[2], device=pts.device)[None, :] >= sel.sum(1)[:, None]
            park = pts.new_tensor([-10.0, 10.0, 10.0])[None, :, None]
            pts = torch.where(tail[:, None, :], park, pts).permute(0, 2, 1).contiguous()
        if prior_point_cloud is not None:
            mats3 = torch.stack([K, RT_cam2, RTinv_cam3], 1).to(torch.float32).contiguous()
            pts2, xyp2 = torch.ops.pixelsynth_b200.project_cloud(prior_point_cloud, mats3, EPS)
            pts = torch.cat([pts, pts2], 1)
            xyp = torch.cat([xyp, xyp2], 2)
        return pts.permute(0, 2, 1), xyp.contiguous()

    # -- z_buffer_manipulator.py:184-219 ------------------------------------------------------
    def forward_justpts_cumulative(self, src1, pred_pts, K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2,
                                   prior_point_cloud, src2, last_background_mask, RTinv_cam3):
        """Batches are allowed (BASELINE config 5): the images of a batch append different numbers of newly outpainted
        pixels, so clouds and features are zero-padded per image (see _compact); the reference's boolean-mask views only
        work when the counts agree, which is why it renders scenes one image at a time."""
        bs, c, w, h = src1.size()
        if last_background_mask is not None:
            last_background_mask = last_background_mask.view(bs, 1, -1)
        pred_pts = pred_pts.view(bs, 1, -1)
        src1 = src1.view(bs, c, -1)
        if src2 is not None:
            src1 = self._compact(last_background_mask.view(bs, -1), src1)
            src = torch.cat([src1, src2.view(bs, c, -1)], 2)
        else:
            src = src1
        pts3D, new_point_cloud = self.project_pts_cumulative(
            pred_pts, K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2, prior_point_cloud,
            last_background_mask if src2 is not None else None, RTinv_cam3)
        pointcloud = pts3D.permute(0, 2, 1).contiguous()
        result, background_mask = self.splatter(pointcloud, src)
        return result, background_mask, new_point_cloud, src
