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
    def project_pts_cumulative(self, pts3D, K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2, prior_point_cloud=None,
                               last_background_mask=None, RTinv_cam3=None):
        """pts3D: full-grid depth (B,1,P); the reference passes the already-masked depth, here the mask is
        applied after projecting the full grid (per-point arithmetic is identical)."""
        mats = pack_mats(K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2)
        pts, xyp = torch.ops.pixelsynth_b200.project_pts(pts3D, mats, self.W, EPS, True)
        bs = pts.shape[0]
        if last_background_mask is not None:
            sel = last_background_mask.view(bs, -1)
            pts = pts[sel].view(bs, -1, 3)
            xyp = xyp.permute(0, 2, 1)[sel].view(bs, -1, 4).permute(0, 2, 1)
        if prior_point_cloud is not None:
            mats3 = torch.stack([K, RT_cam2, RTinv_cam3], 1).to(torch.float32).contiguous()
            pts2, xyp2 = torch.ops.pixelsynth_b200.project_cloud(prior_point_cloud, mats3, EPS)
            pts = torch.cat([pts, pts2], 1)
            xyp = torch.cat([xyp, xyp2], 2)
        return pts.permute(0, 2, 1), xyp.contiguous()

    # -- z_buffer_manipulator.py:184-219 ------------------------------------------------------
    def forward_justpts_cumulative(self, src1, pred_pts, K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2,
                                   prior_point_cloud, src2, last_background_mask, RTinv_cam3):
        bs, c, w, h = src1.size()
        if last_background_mask is not None:
            last_background_mask = last_background_mask.view(bs, 1, -1)
        pred_pts = pred_pts.view(bs, 1, -1)
        src1 = src1.view(bs, c, -1)
        if src2 is not None:
            src1 = src1[last_background_mask.repeat(1, c, 1)].view(bs, c, -1)
            src = torch.cat([src1, src2.view(bs, c, -1)], 2)
        else:
            src = src1
        pts3D, new_point_cloud = self.project_pts_cumulative(
            pred_pts, K, K_inv, RT_cam1, RTinv_cam1, RT_cam2, RTinv_cam2, prior_point_cloud,
            last_background_mask if src2 is not None else None, RTinv_cam3)
        pointcloud = pts3D.permute(0, 2, 1).contiguous()
        result, background_mask = self.splatter(pointcloud, src)
        return result, background_mask, new_point_cloud, src
