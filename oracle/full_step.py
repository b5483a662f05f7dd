"""Oracle for the whole self-supervised train step.  TEST INFRASTRUCTURE (see oracle/__init__).

CPU restatement (torch fp32 + autograd, C for the triangle rasteriser) of the body of
`Engine._epoch_with_both`, /root/reference/network/engine.py:349-376:
    HandSynthesizer.forward                 network/util_modules.py:104-122
    HeatmapEstimationNetwork.forward        network/create_network_and_criterion.py:84-144   (real_aug off)
    MultiTaskLoss.forward                   network/create_network_and_criterion.py:183-263
    sum_loss_terms / backward / Adam step   network/engine.py:144-148, 375-376, 95-97
composed from the per-function oracles (oracle.synth / oracle.hourglass / oracle.losses), every random draw injected
(SURVEY §7.3-9).  Used as the checker of tests/test_gpu_step.py and as the timed CPU baseline of bench.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import hourglass as oh
from . import losses as ol
from . import synth as osy

WEIGHTS = {'synt_hm': 1e3, 'synt_pt': 1e-1, 'mv_consistency': 1e-3, 'mv_projection': 1.0, 'prior': 1e-2, 'hm_mean': 1e-2,
           'collision': 1.0, 'bone_length': 1.0}      # create_network_and_criterion.py:171-181


class HandTables:
    """Dense skinning tables (pointTransformation.py:27-32) from the arrays of tests/golden/hand_model.npz."""

    def __init__(self, a):
        v = np.asarray(a['vertices'], np.float64)
        nb = a['offset_mats'].shape[0]
        self.mesh = np.zeros((nb, v.shape[0], 4), np.float32)
        self.mesh[a['weight_bone'], a['weight_vertexid']] = (a['weight_coeff'][:, None] * v[a['weight_vertexid']]).astype(np.float32)
        kp = np.concatenate([np.asarray(a['keypoints'], np.float32)[:, :3], np.ones((len(a['keypoint_radius']), 1), np.float32)], 1)
        self.kp = np.zeros((nb, kp.shape[0], 4), np.float32)
        self.kp[a['keypoint_bone'], np.arange(kp.shape[0])] = kp
        self.radii = torch.from_numpy(np.asarray(a['keypoint_radius'], np.float32))
        self.offset_mats = torch.from_numpy(np.asarray(a['offset_mats'], np.float32))
        self.faces = osy.swap_faces_right_hand(np.asarray(a['faces']))


def synthesize(tables, poses, scales, rand_f, noise, S, depth_scale=0.01):
    """HandSynthesizer.forward with the draws injected -> (synt_dms [Ns,S,S], uv_hms, xyz_pts [Ns,41,4])."""
    mats = osy.rand_scale_apply(osy.forward_kinematics(poses, tables.offset_mats), scales)
    dm, _ = osy.depth_render(mats, tables.mesh, tables.faces, S, rand_f, fma=True)
    dm = osy.depth_noise(dm * depth_scale, noise[0], noise[1], noise[2])
    uv_hms, _d, xyz = osy.hand_heatmaps(mats, tables.kp, S // 4, rand_f)
    return dm, uv_hms, xyz


def loss_terms(sd, stacks, images, Ns, B, V, real, cams, inv_cams, uv_t, xyz_t, radii, vae_w, eps, is_mv=True,
               depth_scale=0.01, weights=None, round_bf16=False, use=('proj', 'cons', 'prior', 'collision', 'bone'),
               term_scale=None, aug=None):
    """Network + MultiTaskLoss -> (dict of weighted terms, list of projected_dms, list of real_xyz).
    term_scale(name) -> factor applied to each term as it is accumulated (names: the keys of the result plus
    'pose_prior_recon' / 'pose_prior_kld' for the two halves of the prior); None = 1 (the reference).
    aug = (u [B*V], v [B*V]): the scale augmentation was applied to the real rows of `images`; the recovered real joints are
    divided by it, xyz[:, :, 0] /= u, xyz[:, :, 1] /= v (create_network_and_criterion.py:124-126)."""
    w = dict(WEIGHTS)
    w.update(weights or {})
    if term_scale is not None:
        ts = term_scale
        w = dict(w)
        w['synt_hm'] *= ts('synt_uv'); w['synt_pt'] *= ts('synt_d'); w['mv_projection'] *= ts('mv_projection')
        w['mv_consistency'] *= ts('mv_consistency'); w['hm_mean'] *= ts('uv_hm_mean'); w['collision'] *= ts('collision')
        w['bone_length'] *= ts('bone_length')
    J = 41
    outs, _ = oh.hourglass_forward(images, sd, stacks, round_bf16=round_bf16)
    t = {k: 0.0 for k in ('synt_uv', 'synt_d', 'mv_projection', 'mv_consistency', 'uv_hm_mean', 'pose_prior', 'collision',
                          'bone_length')}
    projected, xyzs = [], []
    for si, o in enumerate(outs):
        xyz = ol.soft_argmax_xyz(o[:, :J], o[:, J:], depth_scale)
        if aug is not None:
            div = torch.stack([aug[0], aug[1], torch.ones_like(aug[0])], dim=-1)[:, None, :]
            xyz = torch.cat([xyz[:Ns], xyz[Ns:] / div], 0)
        if Ns:
            t['synt_uv'] = t['synt_uv'] + w['synt_hm'] * F.mse_loss(o[:Ns, :J], uv_t)
            t['synt_d'] = t['synt_d'] + w['synt_pt'] * F.mse_loss(xyz[:Ns, :, 2], xyz_t[:, :, 2])
        joints = xyz[Ns:].reshape(B, V, J, 3)
        xyzs.append(joints)
        if 'proj' in use:
            l, proj, _, _ = ol.mutual_projection_loss(cams, inv_cams, joints, real, radii, is_mv)
            t['mv_projection'] = t['mv_projection'] + w['mv_projection'] * l
            projected.append(proj)
        if 'cons' in use:
            t['mv_consistency'] = t['mv_consistency'] + (w['mv_consistency'] if is_mv else 0.0) * ol.multiview_consistency(cams, joints)
        t['uv_hm_mean'] = t['uv_hm_mean'] + w['hm_mean'] * (o[Ns:, :J] ** 2).mean()
        if 'prior' in use:
            rec, kld = ol.vae_prior_loss((joints / 100.0).reshape(B * V, J * 3), vae_w, eps[si], parts=True)
            if term_scale is not None:
                rec, kld = rec * term_scale('pose_prior_recon'), kld * term_scale('pose_prior_kld')
            t['pose_prior'] = t['pose_prior'] + w['prior'] * (rec + kld)
        if 'collision' in use:
            t['collision'] = t['collision'] + w['collision'] * ol.collision_loss(joints)
        if 'bone' in use:
            t['bone_length'] = t['bone_length'] + w['bone_length'] * ol.bone_length_loss(joints)
    return t, projected, xyzs


def train_step(sd, stacks, tables, vae_w, batch, S, opt_state=None, lr=1e-4, weight_decay=1e-5, is_mv=True,
               depth_scale=0.01, round_bf16=False, apply_update=True):
    """One whole step.  `sd`: dict name -> leaf tensors (requires_grad) updated IN PLACE by Adam.
    batch: dict(real [B,V,S,S], cams, inv_cams, poses [Ns,26], scales [Ns,3], rand_f [Ns], noise [3,Ns,S,S],
                eps [stacks,B*V,32][, aug_u, aug_v [B*V]: scale augmentation of the real views]).  Returns (terms dict of floats incl. 'total', grads dict, aux dict)."""
    real = batch['real']
    B, V = real.shape[:2]
    Ns = batch['poses'].shape[0]
    with torch.no_grad():
        synt, uv_t, xyz_t = synthesize(tables, batch['poses'], batch['scales'], batch['rand_f'], batch['noise'], S, depth_scale)
    real_in = real.reshape(B * V, S, S) * depth_scale
    aug = None
    if 'aug_u' in batch:          # HeatmapEstimationNetwork.forward with real_aug (create_network_and_criterion.py:94-102), draws injected
        aug = (batch['aug_u'], batch['aug_v'])
        real_in = osy.resize_crop(real_in, aug[0], aug[1]).reshape(B * V, S, S)
    images = torch.cat([synt, real_in], 0)
    for p in sd.values():
        p.grad = None
    terms, projected, xyzs = loss_terms(sd, stacks, images, Ns, B, V, real, batch['cams'], batch['inv_cams'], uv_t, xyz_t,
                                        tables.radii, vae_w, batch['eps'], is_mv, depth_scale, round_bf16=round_bf16, aug=aug)
    total = sum(v for v in terms.values() if torch.is_tensor(v))
    total.backward()
    grads = {k: p.grad.clone() for k, p in sd.items()}
    if apply_update:
        if opt_state is None:
            opt_state = {}
        if 'opt' not in opt_state:
            opt_state['opt'] = torch.optim.Adam(list(sd.values()), lr=lr, weight_decay=weight_decay)
        opt_state['opt'].step()
    out = {k: float(v) for k, v in terms.items()}
    out['total'] = float(total)
    return out, grads, dict(images=images, uv_t=uv_t, xyz_t=xyz_t, projected=projected, real_xyz=xyzs)
