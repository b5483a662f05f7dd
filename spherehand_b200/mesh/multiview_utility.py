"""Drop-in mirror of /root/reference/mesh/multiview_utility.py for the hot path: MutualTransformation (:9-30),
MutualProjection (:32-77), MutualProjectionLoss (:80-130), MultiviewConsistencyLoss (:133-167).
`WeightedMultiviewConsistencyLoss` (:170) and `FuseMvPose` (:203) are constructed/imported by the reference but never
called in any epoch loop and are not mirrored.
"""
import torch
import torch.nn as nn

from .. import ops
from .kinematicsTransformation import keypoint_radii
from .render import BallRender, DataToModelLoss, SphereRenderFunction, _ScalarLossFunction


class MutualTransformation(nn.Module):
    """T[b,i,j] = inv_trans[b,j] @ trans[b,i] (:13-30).  O(B V^2) 4x4 products: host-side plumbing (batched matmul)."""

    def forward(self, trans_mats, inv_trans_mats):
        assert trans_mats.ndimension() == 4
        return torch.matmul(inv_trans_mats[:, None, :], trans_mats[:, :, None])


class MutualProjection(nn.Module):
    def __init__(self, img_size, mesh):
        super().__init__()
        self.height = img_size
        self.width = img_size
        self.mutual_trans_mat = MutualTransformation()
        self.ball_render = BallRender(self.width, self.height)
        radiuses = torch.tensor(keypoint_radii(mesh)).type(torch.float32)
        self.num_joints = len(radiuses)
        self.register_buffer('radiuses', radiuses.view(1, 1, 1, self.num_joints))

    def forward(self, camera_poses, inv_camera_poses, joints):
        """-> (depth_imgs [B,V,V,H,W], projected_points [B,V,V,J,3,1]): view i's spheres rendered in view j's frame (:55-77)."""
        b, v, j = joints.shape[:3]
        t = self.mutual_trans_mat(camera_poses, inv_camera_poses).detach()                       # [B,V,V,4,4]
        pts = torch.matmul(t[:, :, :, None, 0:3, 0:3], joints.reshape(b, v, 1, j, 3, 1)) + t[:, :, :, None, 0:3, 3].unsqueeze(-1)
        centres = pts.reshape(b * v * v, j, 3)
        radii = self.radiuses.reshape(1, j).expand(b * v * v, j)
        depth, _ = SphereRenderFunction.apply(centres, radii, self.height, self.width)
        return depth.view(b, v, v, self.height, self.width), pts


class _MutualProjectionLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, joints, cam, inv_cam, depth_maps, radii, is_mv):
        loss3, proj, grad = ops.mvproj_loss_fwdbwd(cam, inv_cam, joints.detach().contiguous().float(), depth_maps, radii, is_mv)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(proj)
        return loss3[0].clone(), proj

    @staticmethod
    def backward(ctx, g, _gproj):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None


class MutualProjectionLoss(nn.Module):
    """loss = 9*MSE(projected, real) + 500 * 9*DataToModel(real, projected joints)  (is_mv; the diagonal pairs x3 otherwise),
    value and d/d(joints) from ONE fused kernel pass (csrc/mvproj_loss.cu).  Returns (loss, projected_dms [B,V,V,H,W]);
    projected_dms is not differentiable on its own (the reference only ever visualises it, network/engine.py:229-260)."""

    def __init__(self, img_size, radiuses):
        super().__init__()
        self.mutual_projection = MutualProjection(img_size, radiuses)
        self.data_to_model_criterion = DataToModelLoss(img_size, img_size, radiuses)
        self.model_to_data_criterion = nn.MSELoss()
        self.model_beneth_surface_criterion = nn.MSELoss()
        self.num_joints = self.mutual_projection.num_joints
        self.relu = nn.ReLU()

    def forward(self, camera_poses, inv_camera_poses, joints, depth_maps, is_mv=True):
        if not is_mv and camera_poses.shape[1] != 3:
            raise IndexError('the single-view branch of the reference indexes views 0, 1, 2 (multiview_utility.py:107-127)')
        radii = self.mutual_projection.radiuses.reshape(-1).contiguous()
        return _MutualProjectionLossFunction.apply(joints, camera_poses.detach().contiguous().float(),
                                                   inv_camera_poses.detach().contiguous().float(),
                                                   depth_maps.detach().contiguous().float(), radii, bool(is_mv))


class MultiviewConsistencyLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.loss_func = torch.nn.MSELoss()

    def forward(self, camera_poses, joints, hm_weight=None):
        """camera_poses [B,V,4,4], joints [B,V,J,3] -> MSE between every view's canonical-frame joints and their per-coordinate
        median over views (gradient flows through the median's selected view)."""
        if hm_weight is not None:
            raise NotImplementedError('hm_weight is always None at the reference call site '
                                      '(network/create_network_and_criterion.py:225-229); the weighted branch is not on the hot path')
        losses, grads = ops.pose_losses_fwdbwd(camera_poses.detach().contiguous().float(), joints.detach().contiguous().float(), flags=1)
        return _ScalarLossFunction.apply(joints, losses[0], grads[0].reshape(joints.shape))
