"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU (build container only).

TEST INFRASTRUCTURE.  Usage:  python oracle/make_golden.py
The reference ships no golden vectors, so parity is pinned against outputs of the
reference code itself: every array written here is produced by a reference module imported from
/root/reference through oracle/_refshim.py (import-time shims only, no semantic edits), from inputs
drawn with the seeds recorded in each file.  The committed fixtures travel to the GPU box; the
reference does not.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')

from oracle import _refshim  # noqa: E402

_refshim.install()
import torch  # noqa: E402

from oracle.hourglass import det_state_dict, det_uniform  # noqa: E402

torch.set_num_threads(8)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s %8.1f KB' % (name, os.path.getsize(path) / 1024))


def rand_cams(B, V, seed):
    """camera_poses[b,0]=I, others a rotation <= 30 deg about a random axis, zero translation."""
    g = torch.Generator().manual_seed(seed)
    cams = torch.eye(4).repeat(B, V, 1, 1)
    for b in range(B):
        for v in range(1, V):
            axis = torch.randn(3, generator=g)
            axis = axis / axis.norm()
            ang = (torch.rand(1, generator=g) * 2 - 1) * np.pi / 6
            K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            cams[b, v, :3, :3] = torch.eye(3) + torch.sin(ang) * K + (1 - torch.cos(ang)) * (K @ K)
    return cams, torch.inverse(cams)


def main():
    os.makedirs(GOLD, exist_ok=True)
    from network.constants import Constant
    from mesh.render import (BallRender, HandBallPrimitiveRender, DataToModelLoss, CollisionLoss,
                             BoneLengthLoss, Hand3DHeatmapRender, DepthRender)
    from mesh.kinematicsTransformation import HandTransformationMat
    from mesh.pointTransformation import LinearBlendSkinning, RandScale
    from mesh.multiview_utility import MutualProjectionLoss, MultiviewConsistencyLoss
    from network.pose_vae import PoseVae
    from network.util_modules import RecoverXYZCoordinateFromHeatmap, DepthNoise, HandSynthesizer
    from network.hourglass import create_hourglass_network
    from network.create_network_and_criterion import HeatmapEstimationNetwork, MultiTaskLoss
    from dataset.joint_angle import JointAngleDataset

    mesh = Constant.mesh
    bones = mesh['bones']

    # ---------------------------------------------------------------- hand model (asset -> compact fixture)
    offs = np.stack([b['offset_matrix'].astype(np.float32) for b in bones])
    wid = np.concatenate([np.asarray(b['weight_vertexid'], np.int32) for b in bones])
    wco = np.concatenate([np.asarray(b['weight_coeff'], np.float64) for b in bones])
    wbone = np.concatenate([np.full(len(b['weight_vertexid']), i, np.int32) for i, b in enumerate(bones)])
    kp, kr, kb = [], [], []
    for i, b in enumerate(bones):
        for pt, r in b.get('keypoint', []):
            kp.append(np.asarray(pt, np.float64))
            kr.append(r)
            kb.append(i)
    save('hand_model', vertices=mesh['vertices'].astype(np.float64), faces=mesh['faces'].astype(np.int32),
         offset_mats=offs, weight_vertexid=wid, weight_coeff=wco, weight_bone=wbone,
         keypoints=np.stack(kp), keypoint_radius=np.asarray(kr, np.float64), keypoint_bone=np.asarray(kb, np.int32))

    # ---------------------------------------------------------------- poses
    torch.manual_seed(0)
    ds = JointAngleDataset()
    poses = torch.stack([ds[i] for i in range(6)])
    poses = torch.cat([torch.zeros(1, 26), poses])           # row 0 = rest pose
    fk = HandTransformationMat([o for o in offs])
    mats = fk(poses)
    save('fk', params=poses, mats=mats)

    # ---------------------------------------------------------------- R2: BallRender + min (config 1)
    for S in (64, 128):
        r = HandBallPrimitiveRender(bones, S, S)
        centres = r.lbs(mats[:3]).clone().requires_grad_(True)
        radii = r.radiuses.repeat(3, 1)
        balls = r.ball_renderer(centres.view(-1, 4), radii.view(-1)).view(3, 41, S, S)
        depth, idx = balls.min(dim=1)
        g = torch.Generator().manual_seed(5)
        gd = torch.randn(depth.shape, generator=g) * (depth < 100)
        (depth * gd).sum().backward()
        save('sphere_render_%d' % S, centres=centres, radii=r.radiuses[0], depth=depth,
             idx=torch.where(depth < 100, idx, torch.full_like(idx, 255)).to(torch.uint8),
             grad_depth=gd, grad_centres=centres.grad[..., :3])
    # random spheres, J=48 (config 2 shape, small N), radii as leaf for d/dr
    g = torch.Generator().manual_seed(1234)
    N, J, S = 4, 48, 128
    c = torch.cat([torch.rand(N, J, 2, generator=g) * 180 - 90, torch.rand(N, J, 1, generator=g) * 120 - 60], -1)
    rad41 = torch.tensor(kr, dtype=torch.float32)
    rad = torch.cat([rad41, torch.full((7,), 20.0)])
    c.requires_grad_(True)
    radl = rad.repeat(N, 1).clone().requires_grad_(True)
    balls = BallRender(S, S)(c.view(-1, 3), radl.view(-1)).view(N, J, S, S)
    depth, idx = balls.min(dim=1)
    gd = torch.randn(depth.shape, generator=g) * (depth < 100)
    (depth * gd).sum().backward()
    save('sphere_render_rand48', centres=c, radii=rad, depth=depth,
         idx=torch.where(depth < 100, idx, torch.full_like(idx, 255)).to(torch.uint8),
         grad_depth=gd, grad_centres=c.grad, grad_radii=radl.grad)

    # ---------------------------------------------------------------- multi-view losses
    radii_list = [float(x) for x in kr]
    for S, B in ((32, 2), (64, 2)):
        V = 3
        cams, inv = rand_cams(B, V, 11 + S)
        # joints: sphere centres of random poses, perturbed per view; real depth: spheres of the clean pose
        r = HandBallPrimitiveRender(bones, S, S)
        base = r.lbs(mats[1:1 + B])[..., :3]                                           # [B,41,3] canonical
        g = torch.Generator().manual_seed(3)
        real = []
        joints = []
        for v in range(V):
            pv = torch.einsum('bxy,bky->bkx', inv[:, v, :3, :3], base) + inv[:, v, None, :3, 3]
            balls = r.ball_renderer(pv.reshape(-1, 3), r.radiuses.repeat(B, 1).view(-1)).view(B, 41, S, S)
            real.append(balls.min(dim=1)[0])
            joints.append(pv + torch.randn(pv.shape, generator=g) * 4.0)
        real = torch.stack(real, 1)
        joints = torch.stack(joints, 1)
        crit = MutualProjectionLoss(S, radii_list)
        out = {}
        for is_mv in (True, False):
            j = joints.clone().requires_grad_(True)
            loss, proj = crit(cams, inv, j, real, is_mv)
            loss.backward()
            out['loss_mv%d' % is_mv] = loss
            out['grad_mv%d' % is_mv] = j.grad
            if is_mv:
                out['projected_dms'] = proj
        j = joints.reshape(B * V, 41, 3).clone().requires_grad_(True)
        d2m = DataToModelLoss(S, S, radii_list)(real.reshape(B * V, S, S), j)
        d2m.backward()
        jc = joints.clone().requires_grad_(True)
        cons = MultiviewConsistencyLoss()(cams, jc)
        cons.backward()
        jcol = joints.clone().requires_grad_(True)
        # squeeze the hand so collisions and bone-length violations actually fire
        col = CollisionLoss()(jcol * 0.5)
        col.backward()
        jb = joints.clone().requires_grad_(True)
        bl = BoneLengthLoss()(jb * 0.7) + BoneLengthLoss()(jb * 1.3)
        bl.backward()
        save('mv_losses_%d' % S, cams=cams, inv_cams=inv, joints=joints, real=real, radii=np.asarray(kr, np.float32),
             d2m=d2m, d2m_grad=j.grad, cons=cons, cons_grad=jc.grad, col=col, col_grad=jcol.grad,
             bone=bl, bone_grad=jb.grad, **out)

    # ---------------------------------------------------------------- VAE prior
    vae = PoseVae(41 * 3, 32, 'mesh/model/pose_vae.pth')
    save('pose_vae', **{k: v for k, v in vae.state_dict().items()})
    x = (joints.reshape(-1, 41, 3) / 100.0).clone().requires_grad_(True)
    torch.manual_seed(21)
    eps = torch.randn(x.shape[0], 32)
    torch.manual_seed(21)
    pl = vae.prior_loss(x)
    pl.backward()
    save('vae_prior', x=x, eps=eps, loss=pl, grad=x.grad)

    # ---------------------------------------------------------------- soft-argmax
    g = torch.Generator().manual_seed(9)
    uv = (torch.randn(3, 41, 16, 16, generator=g) * 0.3).requires_grad_(True)
    dh = (torch.randn(3, 41, 16, 16, generator=g) * 0.3).requires_grad_(True)
    xyz = RecoverXYZCoordinateFromHeatmap(16, 16, 0.01)(uv, dh)
    gx = torch.randn(xyz.shape, generator=g)
    (xyz * gx).sum().backward()
    save('softargmax', uv=uv, d=dh, xyz=xyz, gxyz=gx, guv=uv.grad, gd=dh.grad)

    # ---------------------------------------------------------------- synthetic branch pieces
    g = torch.Generator().manual_seed(17)
    scales = torch.rand(7, 3, generator=g) * 0.1 + 0.85
    rand_f = torch.rand(7, generator=g) * 0.2 + 0.9
    S4 = torch.eye(4).repeat(7, 1, 1)
    S4[:, 0, 0], S4[:, 1, 1], S4[:, 2, 2] = scales[:, 0], scales[:, 1], scales[:, 2]
    smats = S4[:, None] @ mats
    dr = DepthRender(copy.deepcopy(mesh), 128)
    verts = dr.camera(dr.lbs(smats), rand_f)
    fv = verts[:, dr.rasterizer.faces, 0:3].view(7, -1, 3, 3)
    for hm in (16, 32):
        h3 = Hand3DHeatmapRender(bones, hm)
        uvh, dhm, xyzp = h3(smats, rand_f)
        save('synth_heatmaps_%d' % hm, mats=smats, rand_f=rand_f, uv_hms=uvh[:3], d_hms=dhm[:3], xyz=xyzp)
    save('synth_geometry', params=poses, scales=scales, rand_f=rand_f, mats=smats,
         verts_checksum=verts.double().sum(dim=1), verts_head=verts[:, :64], faces=dr.rasterizer.faces.view(-1, 3).to(torch.int32),
         face_verts_b0=fv[1])
    dn = DepthNoise(64, 64)
    dm = torch.rand(2, 64, 64, generator=g) * 1.6
    torch.manual_seed(33)
    nx, ny, nz = torch.randn(2, 64, 64), torch.randn(2, 64, 64), torch.randn(2, 64, 64)
    torch.manual_seed(33)
    save('depth_noise', dm=dm, nx=nx, ny=ny, nz=nz, out=dn(dm.clone()))

    # ---------------------------------------------------------------- hourglass (deterministic weights)
    for stacks, S, N in ((1, 64, 2), (2, 64, 2)):
        net = create_hourglass_network(82, stacks)
        sd = det_state_dict(82, stacks, seed=7)
        net.load_state_dict(sd)
        x = torch.from_numpy(det_uniform(N * S * S, 99).reshape(N, S, S) * 0.5)
        outs, lats = net(x)
        gs = [torch.from_numpy(det_uniform(o.numel(), 100 + i).reshape(o.shape)) for i, o in enumerate(outs)]
        sum((o * gg).sum() for o, gg in zip(outs, gs)).backward()
        grads = dict(net.named_parameters())
        pick = ['conv1.weight', 'conv1.bias', 'bn1.weight', 'layer1.0.conv2.weight', 'layer1.0.downsample.0.weight',
                'layer2.0.bn2.weight', 'layer3.0.conv1.weight', 'hg.0.hg.0.3.0.conv2.weight', 'hg.0.hg.1.2.0.conv3.bias',
                'res.0.0.conv2.weight', 'fc.0.0.weight', 'fc.0.1.bias', 'score.0.weight', 'score.0.bias']
        if stacks == 2:
            pick += ['fc_.0.weight', 'score_.0.weight', 'score.1.weight', 'hg.1.hg.1.0.0.conv2.weight']
        extra = {('grad.' + k): grads[k].grad for k in pick}
        extra.update({('gradnorm.' + k): v.grad.double().norm() for k, v in grads.items()})
        save('hourglass_%dstack' % stacks, x=x, **{'score%d' % i: o for i, o in enumerate(outs)},
             **{'latent%d' % i: o for i, o in enumerate(lats)}, **extra)

    # ---------------------------------------------------------------- full loss (small): MultiTaskLoss terms
    S, hm, B, Ns, V = 64, 16, 2, 2, 3
    net = HeatmapEstimationNetwork(hm, 0.01, 41, 1, real_aug=False)
    net.hg.load_state_dict(det_state_dict(82, 1, seed=7))
    crit = MultiTaskLoss(True, True, True, False, True, True, True, Constant, image_size=S, heatmap_size=hm)
    data = np.load(os.path.join(GOLD, 'mv_losses_64.npz'))
    real = torch.from_numpy(data['real'])
    cams, inv = torch.from_numpy(data['cams']), torch.from_numpy(data['inv_cams'])
    h3 = Hand3DHeatmapRender(bones, hm)
    uvh, dhm, xyzp = h3(smats[1:1 + Ns], rand_f[1:1 + Ns])
    synt_dms = torch.from_numpy(det_uniform(Ns * S * S, 55).reshape(Ns, S, S) * 0.5 + 0.5)
    torch.manual_seed(77)
    eps = torch.randn(B * V, 32)
    torch.manual_seed(77)
    net.train()
    result = net(synt_dms=synt_dms, real_dms=real * 0.01)
    terms, proj = crit(result, real_target={'real_dms': real, 'camera_poses': cams, 'inv_camera_poses': inv, 'is_mv': True},
                       synt_target={'uv_hms': uvh, 'd_hms': dhm * 0.01, 'xyz_pts': xyzp})
    loss = sum(terms.values())
    loss.backward()
    gn = {('gradnorm.' + k): v.grad.double().norm() for k, v in net.hg.named_parameters()}
    save('full_step_small', synt_dms=synt_dms, real=real, cams=cams, inv_cams=inv, uv_hms=uvh, xyz_pts=xyzp, eps=eps,
         real_xyz=result['real_xyz'][0], synt_xyz=result['synt_xyz'][0],
         **{('term.' + k): v for k, v in terms.items()}, loss=loss,
         **{'grad.score.0.weight': net.hg.score[0].weight.grad, 'grad.conv1.weight': net.hg.conv1.weight.grad}, **gn)


if __name__ == '__main__':
    main()
