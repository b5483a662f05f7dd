"""Drop-in mirror of /root/reference/network/create_network_and_criterion.py: HeatmapEstimationNetwork (:27-144) and
MultiTaskLoss (:147-263) with the reference's constructor arguments, result-dict keys, loss-term names and public
`.weights` dict, built from the B200 modules of this package.  This is the module-by-module (autograd-driven) form of the
path — what `network/engine.py` of the reference calls unchanged; `spherehand_b200.engine.SelfSupTrainStep` is the same
arithmetic issued as one static CUDA-graph launch sequence.

The scale augmentation (real_aug=True, :41-51 / :94-102; a per-image Python loop of `F.interpolate` calls in the reference) runs
as one batched kernel (`ResizeCropImage`, util_modules.py) with the reference's random draws in the reference's order, so
the same seeds give the same scales.  Differences, all documented in DESIGN.md: `temporal_smooth_loss` (off by default,
cross-iteration state) raises if enabled; `domain_loss` has weight 0 in the reference (:179) and is reported as an exact 0
without reading the latents.
"""
import torch
import torch.nn as nn

from .hourglass import create_hourglass_network
from .pose_vae import PoseVae
from .util_modules import RecoverXYZCoordinateFromHeatmap, ResizeCropImage
from ..mesh.multiview_utility import MultiviewConsistencyLoss, MutualProjectionLoss
from ..mesh.render import BoneLengthLoss, CollisionLoss


class HeatmapEstimationNetwork(nn.Module):
    def __init__(self, heatmap_size, depth_scale, num_joints, num_stacks, real_aug=True):
        super().__init__()
        self.num_joints = num_joints
        self.hg = create_hourglass_network(num_joints * 2, num_stacks)
        self.xyz_recover = RecoverXYZCoordinateFromHeatmap(heatmap_size, heatmap_size, depth_scale)
        self.real_aug = real_aug
        self.resize_dm = ResizeCropImage() if real_aug else None

    def _augment(self, dms):
        """The reference's scale augmentation of the real views (:41-51, :94-102), same draws in the same order: one host
        uniform decides whether to augment at all, then a host draw of the common scale and two device draws of the u / v
        jitter.  dms [N,H,W] -> (dms, u_scales, v_scales); scales are None when nothing was resized."""
        if self.resize_dm is None or not self.training:
            return dms, None, None
        if torch.rand(1).item() < 0.5:
            return dms, None, None
        rnd_scale = torch.rand(dms.shape[0]).to(dms.device) * 0.2 + 0.75
        u = rnd_scale + torch.rand_like(rnd_scale) * 0.1 - 0.05
        v = rnd_scale + torch.rand_like(rnd_scale) * 0.1 - 0.05
        return self.resize_dm(dms, u, v).reshape(dms.shape), u, v

    @staticmethod
    def _unscale(xyz, u, v):
        """xyz[:, :, 0] /= u, xyz[:, :, 1] /= v (:60-62, :124-126), out of place (xyz comes from an autograd.Function)."""
        if u is None:
            return xyz
        inv = torch.stack([1.0 / u, 1.0 / v, torch.ones_like(u)], dim=-1).view(-1, 1, 3)
        return [p * inv for p in xyz]

    def _heads(self, output, lo, hi):
        uv = [o[lo:hi, :self.num_joints] for o in output]
        d = [o[lo:hi, self.num_joints:] for o in output]
        xyz = [self.xyz_recover(a, b) for a, b in zip(uv, d)]
        return uv, d, xyz

    def _foward_real(self, dms):
        num_real, num_view = dms.shape[0], dms.shape[1]
        dms, u, v = self._augment(dms.reshape(num_real * num_view, dms.shape[2], dms.shape[3]))
        output, _ = self.hg(dms)
        uv, d, xyz = self._heads(output, 0, num_real * num_view)
        xyz = self._unscale(xyz, u, v)
        j = self.num_joints
        result = {'real_uv_hms': [h.reshape(num_real, num_view, j, h.shape[-2], h.shape[-1]) for h in uv],
                  'real_d_hms': [h.reshape(num_real, num_view, j, h.shape[-2], h.shape[-1]) for h in d],
                  'real_xyz': [p.reshape(num_real, num_view, j, 3) for p in xyz]}
        if self.resize_dm is not None:
            result['real_resized_dms'] = dms
        return result

    def _foward_synthetic(self, dms):
        output, _ = self.hg(dms)
        uv, d, xyz = self._heads(output, 0, dms.shape[0])
        return {'synt_uv_hms': uv, 'synt_d_hms': d, 'synt_xyz': xyz}

    def forward(self, real_dms=None, synt_dms=None):
        if synt_dms is None:
            return self._foward_real(real_dms)
        if real_dms is None:
            return self._foward_synthetic(synt_dms)
        num_sync, num_real, num_view = synt_dms.shape[0], real_dms.shape[0], real_dms.shape[1]
        real_dms, u, v = self._augment(real_dms.reshape(num_real * num_view, real_dms.shape[2], real_dms.shape[3]))
        combined_output, combined_latent = self.hg(torch.cat([synt_dms, real_dms], dim=0))
        j = self.num_joints
        result = {}
        result['synt_uv_hms'], result['synt_d_hms'], result['synt_xyz'] = self._heads(combined_output, 0, num_sync)
        uv, d, xyz = self._heads(combined_output, num_sync, num_sync + num_real * num_view)
        xyz = self._unscale(xyz, u, v)
        result['real_uv_hms'] = [h.reshape(num_real, num_view, j, h.shape[-2], h.shape[-1]) for h in uv]
        result['real_d_hms'] = [h.reshape(num_real, num_view, j, h.shape[-2], h.shape[-1]) for h in d]
        result['real_xyz'] = [p.reshape(num_real, num_view, j, 3) for p in xyz]
        if self.resize_dm is not None:
            result['real_resized_dms'] = real_dms
        result['batch_synt_fea'] = [l[:num_sync] for l in combined_latent]
        result['batch_real_fea'] = [l[num_sync:] for l in combined_latent]
        return result


class MultiTaskLoss(nn.Module):
    def __init__(self, synthesized_loss, mv_projection_loss, mv_consistency_loss, temporal_smooth_loss, prior_loss,
                 collision_loss, bone_length_loss, constant, image_size=64, heatmap_size=16, pose_vae_path='mesh/model/pose_vae.pth'):
        super().__init__()
        if temporal_smooth_loss:
            raise NotImplementedError('TemporalSmoothnessLoss (--temporal, off by default, keeps cross-iteration state) is out of scope')
        self.synthesized_loss = nn.MSELoss() if synthesized_loss else None
        self.mv_projection_loss = MutualProjectionLoss(image_size, constant.mesh) if mv_projection_loss else None
        self.mv_consistency_loss = MultiviewConsistencyLoss() if mv_consistency_loss else None
        self.temporal_smooth_loss = None
        self.prior_loss = PoseVae(41 * 3, 32, pose_vae_path) if prior_loss else None
        self.collision_criterion = CollisionLoss() if collision_loss else None
        self.bone_length_criterion = BoneLengthLoss() if bone_length_loss else None
        self.domain_loss = nn.MSELoss()
        self.heatmap_size = heatmap_size
        self.weights = {'synt_hm': 1e3, 'synt_pt': 1e-1, 'mv_consistency': 1e-3, 'mv_projection': 1, 'temporal_smooth': 1.0,
                        'prior': 1e-2, 'hm_mean': 1e-2, 'domain': 0.0, 'collision': 1.0, 'bone_length': 1.0}

    @staticmethod
    def _mse(a, b=None):
        """nn.MSELoss on heat-map / point tensors: elementwise plumbing around the kernels (the fused step computes these
        sums inside sh_softargmax_fwd / sh_step_combine)."""
        return (a * a).mean() if b is None else ((a - b) ** 2).mean()

    def forward(self, result, synt_target=None, real_target=None):
        w = self.weights
        loss_terms = {}
        if self.synthesized_loss is not None and synt_target is not None:
            loss_terms['synt_uv'] = sum(w['synt_hm'] * self._mse(est, synt_target['uv_hms']) for est in result['synt_uv_hms'])
            target_z = synt_target['xyz_pts'][:, :, 2]
            loss_terms['synt_d'] = sum(w['synt_pt'] * self._mse(xyz[:, :, 2], target_z) for xyz in result['synt_xyz'])
        projected_dms = []
        is_mv = False if real_target is None else real_target.get('is_mv', True)
        if self.mv_projection_loss is not None and real_target is not None:
            loss_terms['mv_projection'] = 0
            for xyz in result['real_xyz']:
                cur, dm = self.mv_projection_loss(real_target['camera_poses'], real_target['inv_camera_poses'], xyz,
                                                  real_target['real_dms'], is_mv)
                loss_terms['mv_projection'] = loss_terms['mv_projection'] + cur * w['mv_projection']
                projected_dms.append(dm)
        if self.mv_consistency_loss is not None and real_target is not None:
            wc = w['mv_consistency'] if is_mv else 0
            loss_terms['mv_consistency'] = sum(wc * self.mv_consistency_loss(real_target['camera_poses'], xyz, None)
                                               for xyz in result['real_xyz'])
        if real_target is not None:
            if self.synthesized_loss is None:
                raise TypeError("'NoneType' object is not callable")     # the reference crashes here too (:160,235)
            loss_terms['uv_hm_mean'] = sum(w['hm_mean'] * self._mse(est) for est in result['real_uv_hms'])
        if self.prior_loss is not None:
            loss_terms['pose_prior'] = sum(w['prior'] * self.prior_loss.prior_loss(xyz / 100.0) for xyz in result['real_xyz'])
        if self.collision_criterion is not None:
            loss_terms['collision'] = sum(w['collision'] * self.collision_criterion(xyz) for xyz in result['real_xyz'])
        if self.bone_length_criterion is not None:
            loss_terms['bone_length'] = sum(w['bone_length'] * self.bone_length_criterion(xyz) for xyz in result['real_xyz'])
        if 'batch_synt_fea' in result and 'batch_real_fea' in result:
            if w['domain'] != 0.0:
                raise NotImplementedError('domain loss has weight 0 in the reference (:179); a non-zero weight is out of scope')
            loss_terms['domain_loss'] = torch.zeros((), device=result['batch_synt_fea'][0].device)
        return loss_terms, projected_dms


def combine_loss(loss_terms):
    loss = 0
    for _, l in loss_terms.items():
        loss = loss + l
    return loss
