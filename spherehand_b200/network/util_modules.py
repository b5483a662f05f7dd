"""Drop-in mirror of the hot-path parts of /root/reference/network/util_modules.py: DepthNoise (:46-84),
HandSynthesizer (:86-122), RecoverXYZCoordinateFromHeatmap (:164-201).  Off-path classes of that file (DepthResample,
HeatmapVariance, PosePriorLoss, DepthSegmentation, TemporalSmoothnessLoss; flag-gated or unused) are not
mirrored.  `ResizeCropImage` (:383-424, the per-image scale augmentation; a "next" row of SURVEY.md §8f-2) runs as one
batched kernel instead of a Python loop with one interpolate per image.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..mesh.kinematicsTransformation import HandTransformationMat
from ..mesh.pointTransformation import RandScale
from ..mesh.render import DepthRender, Hand3DHeatmapRender


class DepthNoise(nn.Module):
    def __init__(self, width, height):
        super().__init__()
        self.sigma_x = 0.5
        self.sigma_y = 0.5
        self.sigma_z = 0.05
        u_grid, v_grid = np.meshgrid(np.arange(width), np.arange(height))
        self.register_buffer('u_grid', torch.from_numpy(u_grid).type(torch.long).unsqueeze(dim=0))
        self.register_buffer('v_grid', torch.from_numpy(v_grid).type(torch.long).unsqueeze(dim=0))

    def forward(self, dm, noise=None):
        """Pixel shuffle by rounded N(0, 0.5) offsets + N(0, 0.05) depth noise on the foreground (< 1.0) (:60-84).  The three
        N(0,1) draws are made with torch in the reference's order (shift_x, shift_y, z; :64,69,83) unless `noise`
        [3,B,H,W] is supplied (parity tests)."""
        dm = dm.contiguous().float()
        if noise is None:
            noise = torch.stack([torch.randn_like(dm), torch.randn_like(dm), torch.randn_like(dm)])
        return ops.depth_noise(dm, noise[0].contiguous(), noise[1].contiguous(), noise[2].contiguous(),
                               self.sigma_x, self.sigma_y, self.sigma_z)


class HandSynthesizer(nn.Module):
    """pose parameters [B,26] -> (depth map [B,S,S], uv heat-maps, depth heat-maps [B,41,h,h], key-points [B,41,4]), all
    detached: FK -> RandScale -> LBS -> project -> rasterise -> resize -> noise, and the gaussian target heat-maps."""

    def __init__(self, mesh, image_size, heatmap_size, uv_hm_scale, depth_scale, add_noise=True, out_heatmap=True):
        super().__init__()
        offset_mats = [bone['offset_matrix'].astype(np.float32) for bone in mesh['bones']]
        self.uv_hm_scale = uv_hm_scale
        self.depth_scale = depth_scale
        self.hand_skeleton_transform = HandTransformationMat(offset_mats)
        self.hm_render = Hand3DHeatmapRender(mesh['bones'], heatmap_size)
        self.dm_render = DepthRender(mesh, image_size)
        self.rand_scale = RandScale(0.1)
        self.depth_noiser = DepthNoise(image_size, image_size)
        self.add_noise = add_noise
        self.out_heatmap = out_heatmap

    @torch.no_grad()
    def forward(self, parameters):
        num_batch = parameters.shape[0]
        dev = parameters.device
        scales = self.rand_scale.draw(num_batch)                     # host draws, reference order (pointTransformation.py:140-142)
        rand_f_ratio = (torch.rand(num_batch) * 0.2 + 0.9).to(dev)   # util_modules.py:107
        transform_mats = self.hand_skeleton_transform(parameters, scales)      # FK and the scale product in one launch
        rendered = self.dm_render(transform_mats, rand_f_ratio)
        rendered = ops.scale(rendered, self.depth_scale, torch.empty_like(rendered))
        if self.add_noise:
            rendered = self.depth_noiser(rendered)
        if not self.out_heatmap:
            return rendered
        uv_hms, depth_hms, xyz_pts = self.hm_render(transform_mats, rand_f_ratio)
        if self.uv_hm_scale != 1.0:
            uv_hms = ops.scale(uv_hms, self.uv_hm_scale, torch.empty_like(uv_hms))
        depth_hms = ops.scale(depth_hms, self.depth_scale, torch.empty_like(depth_hms))
        return rendered.detach(), uv_hms.detach(), depth_hms.detach(), xyz_pts.detach()


class _SoftArgmaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv_hms, d_hms, depth_scale_inv):
        n, j, h, w = uv_hms.shape
        # the two inputs are normally the channel halves of one score tensor (create_network_and_criterion.py:111-123):
        # use it in place when they are adjacent views of the same contiguous buffer, else pack them
        es = uv_hms.element_size()
        adjacent = (uv_hms.dtype == torch.float32 and d_hms.dtype == torch.float32
                    and uv_hms.stride() == (2 * j * h * w, h * w, w, 1) and d_hms.stride() == uv_hms.stride()
                    and d_hms.data_ptr() == uv_hms.data_ptr() + j * h * w * es)
        if adjacent:
            score = torch.as_strided(uv_hms.detach(), (n, 2 * j, h, w), (2 * j * h * w, h * w, w, 1))
        else:
            score = torch.cat([uv_hms.detach(), d_hms.detach()], dim=1).contiguous().float()
        xyz, _ = ops.softargmax_fwd(score, j, depth_scale_inv=depth_scale_inv)
        ctx.save_for_backward(score)
        ctx.j, ctx.dsi = j, depth_scale_inv
        return xyz

    @staticmethod
    def backward(ctx, gxyz):
        (score,) = ctx.saved_tensors
        g = ops.softargmax_bwd(score, gxyz.contiguous().float(), ctx.j, depth_scale_inv=ctx.dsi)
        return g[:, :ctx.j], g[:, ctx.j:], None


class RecoverXYZCoordinateFromHeatmap(nn.Module):
    def __init__(self, width, height, depth_scale):
        super().__init__()
        self.depth_scale = 1.0 / depth_scale
        self.fx = width / 300.0
        self.fy = height / 300.0
        self.cx = width / 2
        self.cy = height / 2
        u_grid, v_grid = np.meshgrid(np.arange(width), np.arange(height))
        self.register_buffer('u_grid', torch.from_numpy(u_grid.reshape((1, 1, height, width))).type(torch.float))
        self.register_buffer('v_grid', torch.from_numpy(v_grid.reshape((1, 1, height, width))).type(torch.float))

    def forward(self, uv_hms, d_hms, is_shuffing=False):
        """u,v = sum softmax(20 hm) * grid; d = sum d_hm * relu(hm)/(sum relu(hm) + 1e-5); -> xyz [N,J,3] in mm (:182-201)."""
        assert uv_hms.shape[-1] == self.u_grid.shape[-1] and uv_hms.shape[-2] == self.u_grid.shape[-2], 'heat-map size mismatch'
        return _SoftArgmaxFunction.apply(uv_hms, d_hms, self.depth_scale)


class ResizeCropImage(nn.Module):
    def __init__(self):
        super().__init__()
        self.sigma_z = 0.05

    def forward(self, depth_maps, u_scales, v_scales):
        """depth_maps [N,H,W], u_scales / v_scales [N] -> [N,H,W].squeeze(): every image resized (nearest) by its own scale
        and pasted into the centre of an all-ones canvas (:388-424).  No gradient (it augments the network INPUT)."""
        n = depth_maps.shape[0]
        dm = depth_maps.detach().reshape(n, depth_maps.shape[-2], depth_maps.shape[-1]).contiguous().float()
        out = ops.resize_crop(dm, u_scales.detach().reshape(-1).contiguous().float(), v_scales.detach().reshape(-1).contiguous().float())
        return out.squeeze()
