"""Drop-in mirror of /root/reference/network/pose_denoiser.py: PoseDenoiser (:21-81, the eval-only joint denoiser, SURVEY.md
§8f-4), its index tables (:12-19) and `average_joint_error` (:84-95).  Same constructor arguments, buffers, `network.*`
state_dict keys (the reference's checkpoints load unchanged) and forward signature; the eval forward is ONE kernel
(`sh_pose_denoiser_fwd`: gather, 3 dense layers, 2 GroupNorm+ReLU, scatter) instead of 3 GEMMs + ~12 small kernels.

Training the denoiser (`train()` of the reference, an offline tool: forward in training mode adds noise and needs autograd) is
out of scope: a forward in training mode or on an input that requires grad raises.
"""
import torch
import torch.nn as nn

from .. import ops

key_points = list(range(11))
input_3d_points = list(range(11, 41))
input_2d_points = list(range(11))
input_indices = [i * 3 for i in input_3d_points] + [i * 3 + 1 for i in input_3d_points] + [i * 3 + 2 for i in input_3d_points]
input_indices += [i * 3 for i in input_2d_points] + [i * 3 + 1 for i in input_2d_points]
output_indices = []
for pt in key_points:
    output_indices += [pt * 3, pt * 3 + 1, pt * 3 + 2]


class PoseDenoiser(nn.Module):
    def __init__(self, input_indices=input_indices, output_indices=output_indices, model_path=None):
        super().__init__()
        self.input_fea = len(input_indices)
        self.output_fea = len(output_indices)
        self.scale_factor = 0.01
        self.register_buffer('input_indices', torch.tensor(input_indices).long())
        self.register_buffer('output_indices', torch.tensor(output_indices).long())
        self._make_network()
        self.criterion = nn.MSELoss()
        if model_path is not None:
            check_point = torch.load(model_path, map_location='cpu', weights_only=False)
            self.load_state_dict(check_point['network_state_dict'])
            for param in self.parameters():
                param.requires_grad = False

    def _make_network(self):
        # parameter holders with the reference's keys (network.0 / .1 / .3 / .4 / .6); the arithmetic runs in the kernel
        self.network = nn.Sequential(nn.Linear(self.input_fea, 256), nn.GroupNorm(16, 256), nn.ReLU(), nn.Linear(256, 256),
                                     nn.GroupNorm(16, 256), nn.ReLU(), nn.Linear(256, self.output_fea))

    def _blob(self):
        n = self.network
        parts = [n[0].weight.t(), n[0].bias, n[1].weight, n[1].bias, n[3].weight.t(), n[3].bias, n[4].weight, n[4].bias,
                 n[6].weight.t(), n[6].bias]
        return torch.cat([p.detach().float().contiguous().reshape(-1) for p in parts])

    def forward(self, fea):
        if self.training:
            raise NotImplementedError('PoseDenoiser training (noise injection + autograd, the reference\'s offline train()) is out of '
                                      'scope: call .eval()')
        if fea.requires_grad:
            raise NotImplementedError('PoseDenoiser is forward-only on the B200 path')
        is_skel = fea.ndimension() == 3
        shape = fea.shape
        flat = fea.reshape(shape[0], -1).contiguous().float() if is_skel else fea.contiguous().float()
        out = ops.pose_denoiser_fwd(flat, self.input_indices.to(torch.int32), self.output_indices.to(torch.int32), self._blob(),
                                    self.scale_factor)
        return out.reshape(shape) if is_skel else out

    def loss(self, gt_fea, est_fea):
        num_batch = gt_fea.shape[0]
        gt_fea = gt_fea.reshape(num_batch, -1)[:, self.output_indices]
        est_fea = est_fea.reshape(num_batch, -1)[:, self.output_indices]
        return self.criterion(gt_fea, est_fea)


def average_joint_error(gt_joints, est_joints, key_points):
    gt_joints = gt_joints.cpu()[:, key_points, :]
    est_joints = est_joints.cpu()[:, key_points, :]
    return float((gt_joints - est_joints).norm(dim=-1).mean())
