"""Drop-in mirror of /root/reference/mesh/render.py for the hot path (same class names, constructor arguments, forward
signatures, return structures and registered-buffer names):

    BallRender (:10-53)              R2 core: one sphere per image
    HandBallPrimitiveRender (:56-90) R2 over the 41 key-point spheres (+ the per-sphere part maps the viewer shows)
    DataToModelLoss (:93-142)        CollisionLoss (:145-176)         BoneLengthLoss (:179-206)
    HeatmapRender (:210-248)         Hand3DHeatmapRender (:251-279)
    DepthRasterizationFunction (:282-287)   DepthRasterization (:289-312)   DepthRender (:315-331)   R1

Every forward runs CUDA kernels of libspherehand_b200.so through `spherehand_b200.ops`; the differentiable ones are
`torch.autograd.Function`s with the analytic backward of SURVEY.md §9.  There is no CPU path: CPU tensors raise.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import depth_rasterization, ops
from .kinematicsTransformation import keypoint_radii, keypoint_skin
from .pointTransformation import InverseOthographicalProjection, LinearBlendSkinning, OthographicalProjection, _f4, _no_grad_only


# ------------------------------------------------------------------------------------------------ R2
class SphereRenderFunction(torch.autograd.Function):
    """depth[n] = min over the J spheres of image n of BallRender's per-sphere depth (render.py:26-53 + :89).
    centres [N,J,3], radii [N,J] -> depth [N,H,W]; backward: SURVEY.md §9-A (gradient to the arg-min sphere only)."""

    @staticmethod
    def forward(ctx, centres, radii, height, width):
        sph = ops.pack_spheres(centres.detach(), radii.detach())
        depth, idx = ops.sphere_render_fwd(sph, height, width)
        ctx.save_for_backward(sph, idx)
        ctx.shapes = (centres.shape, radii.shape)
        ctx.mark_non_differentiable(idx)
        return depth, idx

    @staticmethod
    def backward(ctx, grad_depth, _grad_idx):
        sph, idx = ctx.saved_tensors
        g = ops.sphere_render_bwd(grad_depth.contiguous().float(), idx, sph)
        cshape, rshape = ctx.shapes
        gr = g[..., 3]
        if len(rshape) == 1 and rshape[0] == sph.shape[1] and sph.shape[0] != 1:
            gr = gr.sum(dim=0)                      # radii shared by every image
        return g[..., :3].reshape(cshape), gr.reshape(rshape), None, None


class BallRender(nn.Module):
    def __init__(self, width, height):
        super().__init__()
        self.width = width
        self.height = height
        x_grid, y_grid = np.meshgrid(np.arange(self.width), np.arange(self.height))
        self.register_buffer('x_grid', torch.from_numpy(x_grid).type(torch.float))
        self.register_buffer('y_grid', torch.from_numpy(y_grid).type(torch.float))
        # unused and uninitialised in the reference too (:21); kept so that criterion.state_dict() keys match
        self.dist_weight = torch.nn.Parameter(torch.zeros(1))

    def forward(self, xyz_centers, radiuses):
        """xyz_centers [N,>=3], radiuses [N] -> [N,H,W]: depth of sphere n alone, 100.0 off the sphere."""
        n = xyz_centers.shape[0]
        centres = xyz_centers[:, 0:3].reshape(n, 1, 3).float()
        depth, _ = SphereRenderFunction.apply(centres, radiuses.reshape(n, 1).float(), self.height, self.width)
        return depth


class HandBallPrimitiveRender(nn.Module):
    def __init__(self, bones, width, height):
        super().__init__()
        self.width = width
        self.height = height
        self.ball_renderer = BallRender(self.width, self.height)
        vertices, weights, indices = keypoint_skin(bones)
        self.num_vertices = len(vertices)
        self.lbs = LinearBlendSkinning(vertices, weights, indices)
        radiuses = torch.tensor(keypoint_radii({'bones': bones})).type(torch.float).unsqueeze(0)
        self.register_buffer('radiuses', radiuses)

    def forward(self, transformation_mats):
        skinned = self.lbs(transformation_mats)                               # [B,J,4]
        b = skinned.shape[0]
        radiuses = self.radiuses.repeat(b, 1)
        balls = self.ball_renderer(skinned.view(-1, 4), radiuses.view(-1))
        part_maps = balls.view(b, self.num_vertices, self.height, self.width)
        # the min over spheres is the fused kernel, not a reduction over the J part maps
        depth_maps, _ = SphereRenderFunction.apply(skinned[..., :3].contiguous(), radiuses, self.height, self.width)
        return part_maps, depth_maps


# ------------------------------------------------------------------------------------------------ loss heads
class _ScalarLossFunction(torch.autograd.Function):
    """A scalar loss whose kernel returns (value, d value / d joints) in one launch: backward just scales."""

    @staticmethod
    def forward(ctx, joints, value, grad):
        ctx.save_for_backward(grad)
        return value.reshape(()).clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


class DataToModelLoss(nn.Module):
    def __init__(self, width, height, mesh):
        super().__init__()
        u_grid, v_grid = np.meshgrid(np.arange(width), np.arange(height))
        x_grid = (u_grid.astype(np.float32) - width / 2) * 300.0 / width
        y_grid = (v_grid.astype(np.float32) - height / 2) * 300.0 / height
        self.register_buffer('x_grid', torch.from_numpy(x_grid).float().unsqueeze(dim=0))
        self.register_buffer('y_grid', torch.from_numpy(y_grid).float().unsqueeze(dim=0))
        self.width = width
        self.height = height
        radiuses = torch.tensor(keypoint_radii(mesh)).type(torch.float32)
        self.num_joints = len(radiuses)
        self.register_buffer('radiuses', radiuses.view(1, 1, 1, self.num_joints))
        self.relu = nn.ReLU()

    def forward(self, dms, joints):
        """dms [N,H,W] (mm, background > 99), joints [N,J,3] -> mean over all pixels of clamp(min_k ||P-c_k| - r_k|, 0, 50)."""
        n = dms.shape[0]
        j = joints.reshape(n, self.num_joints, 3)
        value, grad = ops.data_to_model_fwdbwd(dms.detach().reshape(n, self.height, self.width).contiguous().float(),
                                               j.detach().contiguous().float(), self.radiuses.reshape(-1).contiguous())
        return _ScalarLossFunction.apply(joints, value, grad.reshape(joints.shape))


def _pose_head(joints, flag, slot, min_dist=6.0):
    """CollisionLoss / BoneLengthLoss share the reference's view(B,-1,3) quirk: only rows 0..40 of every batch entry
    (view 0 of a [B,V,41,3] input) are seen (render.py:169-170, 197-198)."""
    b = joints.shape[0]
    flat = joints.reshape(b, -1, 3)
    if flat.shape[1] < 41:
        raise IndexError('index 40 is out of bounds: the pair tables address 41 key-points')
    v = flat.shape[1] // 41
    if flat.shape[1] != v * 41:
        v = 1
    jv = flat[:, :v * 41].detach().reshape(b, v, 41, 3).contiguous().float()
    cam = torch.eye(4, device=joints.device).repeat(b, v, 1, 1)
    losses, grads = ops.pose_losses_fwdbwd(cam, jv, flags=flag, min_dist=min_dist)
    g = torch.zeros_like(flat, dtype=torch.float32)
    g[:, :v * 41] = grads[slot].reshape(b, v * 41, 3)
    return _ScalarLossFunction.apply(joints, losses[slot], g.reshape(joints.shape))


class CollisionLoss(nn.Module):
    def __init__(self, min_dist=6):
        super().__init__()
        self.relu = nn.ReLU()
        self.min_dist = float(min_dist)
        self.min_sq_dist = min_dist ** 2
        joint_1, joint_2 = [], []
        for j_1 in range(11):                       # every finger sphere against every palm sphere (:153-157)
            for j_2 in range(11, 41):
                joint_1.append(j_1)
                joint_2.append(j_2)
        for j_1 in range(11, 41):                   # finger spheres of different fingers (:158-162)
            for j_2 in range(j_1 + 1, 41):
                if (j_1 - 11) // 6 != (j_2 - 11) // 6:
                    joint_1.append(j_1)
                    joint_2.append(j_2)
        self.register_buffer('joint_1', torch.tensor(joint_1).long())
        self.register_buffer('joint_2', torch.tensor(joint_2).long())

    def forward(self, joints):
        return _pose_head(joints, 2, 1, self.min_dist)


class BoneLengthLoss(nn.Module):
    def __init__(self):
        super().__init__()
        from . import bone_length
        median_length = torch.tensor(bone_length.uniform_length).float()
        self.register_buffer('joint_1', torch.tensor(bone_length.joint_1).long())
        self.register_buffer('joint_2', torch.tensor(bone_length.joint_2).long())
        self.register_buffer('max_length', ((median_length * 1.05) ** 2).unsqueeze(dim=0))
        self.register_buffer('min_length', ((median_length * 0.80) ** 2).unsqueeze(dim=0))
        self.relu = nn.ReLU()

    def forward(self, joints):
        return _pose_head(joints, 4, 2)


# ------------------------------------------------------------------------------------------------ heat-map targets
class HeatmapRender(nn.Module):
    def __init__(self, hm_size, sigma=1.0):
        super().__init__()
        self.sigma = sigma
        self.height = hm_size
        self.width = hm_size
        u_grid, v_grid = np.meshgrid(np.arange(self.width), np.arange(self.height))
        self.register_buffer('u_grid', torch.from_numpy(np.expand_dims(u_grid, axis=0)).type(torch.float32))
        self.register_buffer('v_grid', torch.from_numpy(np.expand_dims(v_grid, axis=0)).type(torch.float32))

    def forward(self, uvd_points):
        assert uvd_points.ndimension() == 3
        _no_grad_only(uvd_points, 'HeatmapRender')
        uv, d, _ = ops.heatmap_render(_f4(uvd_points), self.height, sigma=self.sigma)
        return uv, d


class Hand3DHeatmapRender(nn.Module):
    def __init__(self, bones, heatmap_size):
        super().__init__()
        self.width = heatmap_size
        self.height = heatmap_size
        self.hm_renderer = HeatmapRender(heatmap_size)
        self.camera = OthographicalProjection(self.width / 2, self.height / 2, self.width / 300, self.height / 300)
        self.inv_camera = InverseOthographicalProjection(self.width / 2, self.height / 2, self.width / 300, self.height / 300)
        vertices, weights, indices = keypoint_skin(bones)
        self.num_vertices = len(vertices)
        self.lbs = LinearBlendSkinning(vertices, weights, indices)

    def forward(self, transformation_mats, rand_f=None):
        """-> (uv heat-maps [B,J,h,w], depth heat-maps [B,J,h,w], xyz_points [B,J,4]); LBS + projection in one launch,
        heat-maps + inverse projection in another."""
        cam = (self.camera.cx, self.camera.cy, self.camera.fx, self.camera.fy)
        rf = None if rand_f is None else rand_f.reshape(-1).contiguous().float()
        uvd = self.lbs(transformation_mats, _mode=2 if rand_f is None else 1, _cam=cam, _rand_f=rf)
        return ops.heatmap_render(uvd, self.height, sigma=self.hm_renderer.sigma, cam=cam)


# ------------------------------------------------------------------------------------------------ R1
class DepthRasterizationFunction(torch.autograd.Function):
    """forward only, like the reference (:282-287): the z-buffer is used to synthesise training images, never differentiated."""

    @staticmethod
    def forward(ctx, width, height, face_vertices):
        depth_maps = depth_rasterization.forward(width, height, face_vertices)
        return ops.clamp_max(depth_maps, 100.0)


_LATTICE = {128: (5, 2, 1), 64: (10, 4, 2)}       # the source pixels of the 640 -> S bilinear resize


class DepthRasterization(nn.Module):
    def __init__(self, width, height, np_faces, right_hand=True):
        super().__init__()
        self.width = width
        self.height = height
        faces = np.array(np_faces, copy=True)          # the reference swaps columns IN PLACE on the caller's array (:298-300);
        if right_hand:                                 # a copy keeps a second construction from flipping the winding back
            faces[:, [0, 1]] = faces[:, [1, 0]]
        self.register_buffer('faces', torch.from_numpy(faces).type(torch.int64).view(-1))
        self.register_buffer('faces_i32', torch.from_numpy(np.ascontiguousarray(faces)).type(torch.int32), persistent=False)
        self.num_faces = len(faces)

    def forward(self, vertices):
        """vertices [B,Nv,4] (projected) -> [B,height,width]: rasterise at 640x640, clamp(max=100), bilinear resize (:306-312).
        For 64 / 128 outputs only the source pixels the resize reads are rasterised (exact, not an approximation)."""
        _no_grad_only(vertices, 'DepthRasterization')
        fv = ops.gather_faces(_f4(vertices), self.faces_i32)
        if self.width == self.height and self.width in _LATTICE:
            step, off, noff = _LATTICE[self.width]
            z = ops.tri_raster_lattice_fwd(fv, 640, step, off, noff)
            return ops.lattice_to_depth(z, self.width, noff, 1.0)
        rendered = DepthRasterizationFunction.apply(640, 640, fv).unsqueeze(dim=1)
        return torch.nn.functional.interpolate(rendered, size=(self.height, self.width), mode='bilinear', align_corners=False).squeeze(dim=1)


class DepthRender(nn.Module):
    def __init__(self, mesh, image_size):
        super().__init__()
        self.lbs = LinearBlendSkinning(mesh['vertices'], [b['weight_coeff'] for b in mesh['bones']],
                                       [b['weight_vertexid'] for b in mesh['bones']])
        self.camera = OthographicalProjection(320, 320, 640 / 300, 640 / 300)
        self.rasterizer = DepthRasterization(image_size, image_size, mesh['faces'])

    def forward(self, transformation_mats, rand_fx=None):
        cam = (self.camera.cx, self.camera.cy, self.camera.fx, self.camera.fy)
        rf = None if rand_fx is None else rand_fx.reshape(-1).contiguous().float()
        skinned_points = self.lbs(transformation_mats, _mode=2 if rand_fx is None else 1, _cam=cam, _rand_f=rf)
        return self.rasterizer(skinned_points)
