"""Mirror of reference models/layers/z_buffer_layers.py:11-131 (RasterizePointsXYsBlending).

Same constructor and forward signature; PyTorch3D's rasterize_points + compositing and the ~10
elementwise kernels between them are replaced by one call to torch.ops.pixelsynth_b200.splat_points.
"""
import torch
from torch import nn

from .. import _ops_loaded  # noqa: F401  (registers torch.ops.pixelsynth_b200)
from ...ops import ACCUMULATION


class RasterizePointsXYsBlending(nn.Module):
    def __init__(self, C=64, learn_feature=True, radius=1.5, size=256, points_per_pixel=8, opts=None):
        super().__init__()
        # kept for state_dict compatibility; the reference never uses it in forward (z_buffer_layers.py:74-75)
        if learn_feature:
            self.register_parameter("default_feature", nn.Parameter(torch.randn(1, C, 1)))
        else:
            self.register_buffer("default_feature", torch.zeros(1, C, 1))
        self.radius = radius
        self.size = size
        self.points_per_pixel = points_per_pixel
        self.opts = opts

    def rasterize(self, pts3D, src, want_dist2=True):
        """Parity surface: returns (gen_fs, bg_mask, idx, zbuf, dist2) with the PyTorch3D-shaped maps."""
        bs = src.size(0)
        if src.dim() > 3:  # z_buffer_layers.py:57-62
            bs, c, w, _ = src.size()
            image_size = w
            pts3D = pts3D.permute(0, 2, 1)
            src = src.unsqueeze(2).repeat(1, 1, w, 1, 1).view(bs, c, -1)
        else:
            image_size = self.size
        assert pts3D.size(2) == 3
        assert pts3D.size(1) == src.size(2)
        o = self.opts
        return torch.ops.pixelsynth_b200.splat_points(
            pts3D, src, image_size, self.points_per_pixel, float(self.radius), float(o.tau), int(o.rad_pow),
            ACCUMULATION[o.accumulation], int(o.background_smoothing_kernel_size), True, want_dist2)

    def forward(self, pts3D, src):
        bs = src.size(0)
        if src.dim() > 3:
            bs, c, w, _ = src.size()
            image_size = w
            pts3D = pts3D.permute(0, 2, 1)
            src = src.unsqueeze(2).repeat(1, 1, w, 1, 1).view(bs, c, -1)
        else:
            image_size = self.size
        assert pts3D.size(2) == 3
        assert pts3D.size(1) == src.size(2)
        o = self.opts
        out, bg, _, _, _ = torch.ops.pixelsynth_b200.splat_points(
            pts3D, src, image_size, self.points_per_pixel, float(self.radius), float(o.tau), int(o.rad_pow),
            ACCUMULATION[o.accumulation], int(o.background_smoothing_kernel_size), False, False)
        return out, bg
