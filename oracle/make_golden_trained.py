"""Generate the trained-weight and BASELINE-shape fixtures by running the UNMODIFIED reference on CPU (build container only).

TEST INFRASTRUCTURE.  Usage:  python oracle/make_golden_trained.py
Everything written here is computed by reference modules imported from /root/reference through oracle/_refshim.py
(import-time shims only).  The one piece of the reference that cannot run on a CPU is its CUDA triangle rasteriser
(`depth_rasterization.forward`, mesh/cuda_kernel/depth_rasterization_cuda_kernel.cu); the stub module is bound to the C
restatement oracle/tri_raster.c (FMA-contracted variant), which tests/test_gpu_kernels.py proves bit-identical to the
reference's own compiled kernel on the GPU box.  Files (tests/golden/):

  trained_weights.npz        `hg.*` of /root/reference/pretrained/synthetic.pth (epoch 74, 1 stack; engine.py:446-460 loader keys)
  trained_hourglass_64.npz   HourglassNet(82, 1 stack) fwd + bwd on two reference-synthesised 64x64 depth maps, trained weights
  trained_step_64.npz        the whole step at the reference's native size: HandSynthesizer -> HeatmapEstimationNetwork(16, real_aug
                             off) -> MultiTaskLoss (all heads) -> backward, B=2 tuples x V=3 mesh-rendered views + Ns=2 poses
  trained_step_64_aug.npz    the same 64x64 step with the reference's scale augmentation on (HeatmapEstimationNetwork(real_aug=True),
                             create_network_and_criterion.py:94-102,124-126) and the drawn (u, v) scales recorded
  trained_hourglass_128.npz  2 stacks at 128x128 (BASELINE configs 3/4 shape), N=2; weights = oracle.hourglass.two_stack_from_trained
  trained_step_128.npz       the whole step at 128x128, 2 stacks, heatmap_size=32, B=2, V=3, Ns=2
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

from oracle import synth as osy  # noqa: E402
from oracle.hourglass import det_uniform, two_stack_from_trained  # noqa: E402

torch.set_num_threads(8)


def _tri_forward(width, height, vertices):
    return torch.from_numpy(osy.tri_raster(vertices.detach().cpu().numpy(), width, height, fma=True))


sys.modules['depth_rasterization'].forward = _tri_forward


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
            ang = float((torch.rand(1, generator=g) * 2 - 1) * np.pi / 6)
            K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            cams[b, v, :3, :3] = torch.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    return cams, torch.inverse(cams)


def main():
    from network.constants import Constant
    from mesh.render import DepthRender
    from mesh.kinematicsTransformation import HandTransformationMat
    from network.util_modules import HandSynthesizer
    from network.hourglass import create_hourglass_network
    from network.create_network_and_criterion import HeatmapEstimationNetwork, MultiTaskLoss
    from dataset.joint_angle import JointAngleDataset

    mesh = Constant.mesh
    ck = torch.load('pretrained/synthetic.pth')
    sd1 = {k[3:]: v.float() for k, v in ck['network_state_dict'].items() if k.startswith('hg.')}
    save('trained_weights', **sd1)

    torch.manual_seed(4)
    ds = JointAngleDataset()
    poses = torch.stack([ds[i] for i in range(8)])
    offs = [b['offset_matrix'].astype(np.float32) for b in mesh['bones']]
    fk = HandTransformationMat(offs)

    def synthesise(S, hm, p, seed):
        """reference HandSynthesizer with the draws recorded (same seed replayed in the reference's order)."""
        syn = HandSynthesizer(copy.deepcopy(mesh), image_size=S, heatmap_size=hm, uv_hm_scale=1.0, depth_scale=0.01)
        n = p.shape[0]
        torch.manual_seed(seed)
        sx, sy, sz = (torch.rand(n) * 0.1 + 0.90 - 0.1 / 2 for _ in range(3))          # pointTransformation.py:140-142
        rand_f = torch.rand(n) * 0.2 + 0.9                                               # util_modules.py:107
        noise = torch.stack([torch.randn(n, S, S) for _ in range(3)])                    # util_modules.py:64,69,83
        torch.manual_seed(seed)
        dms, uv, dh, xyz = syn(p)
        return dms, uv, dh, xyz, torch.stack([sx, sy, sz], 1), rand_f, noise

    def real_views(S, p, cams_inv):
        """"real" depth maps in mm (background 100.0): the mesh of pose b rendered from each of the V cameras with the reference's
        DepthRender (every bone matrix left-multiplied by inv_camera_poses[b,v]: skinning weights sum to one)."""
        B, V = cams_inv.shape[:2]
        dr = DepthRender(copy.deepcopy(mesh), S)
        mats = fk(p)                                                                     # [B,17,4,4]
        mv = cams_inv[:, :, None] @ mats[:, None]                                        # [B,V,17,4,4]
        dm = dr(mv.reshape(B * V, 17, 4, 4)).reshape(B, V, S, S)
        # the reference rasteriser's 1/z blending (.cu:103-109) leaves a handful of garbage pixels (down to -6e5 mm) where a
        # triangle crosses z = 0; a depth sensor has none, so for the REAL views they are set to background (input construction
        # only; the synthetic branch keeps whatever the reference produces)
        return torch.where(dm < -150.0, torch.full_like(dm, 100.0), dm)

    def hourglass_fixture(name, stacks, S, sd, x):
        net = create_hourglass_network(82, stacks)
        net.load_state_dict(sd)
        outs, lats = net(x)
        gs = [torch.from_numpy(det_uniform(o.numel(), 300 + i).reshape(o.shape)) for i, o in enumerate(outs)]
        sum((o * gg).sum() for o, gg in zip(outs, gs)).backward()
        named = dict(net.named_parameters())
        pick = ['conv1.weight', 'conv1.bias', 'bn1.weight', 'layer1.0.conv2.weight', 'layer2.0.bn2.weight',
                'layer3.0.conv1.weight', 'hg.0.hg.0.3.0.conv2.weight', 'hg.0.hg.1.2.0.conv3.bias', 'res.0.0.conv2.weight',
                'fc.0.0.weight', 'fc.0.1.bias', 'score.0.weight', 'score.0.bias']
        if stacks == 2:
            pick += ['fc_.0.weight', 'score_.0.weight', 'score.1.weight', 'hg.1.hg.1.0.0.conv2.weight']
        extra = {('grad.' + k): named[k].grad for k in pick}
        extra.update({('gradnorm.' + k): v.grad.double().norm() for k, v in named.items()})
        save(name, x=x, **{'score%d' % i: o for i, o in enumerate(outs)}, **{'latent%d' % i: o for i, o in enumerate(lats)}, **extra)

    def step_fixture(name, stacks, S, hm, sd, seed, aug=False):
        B, V, Ns = 2, 3, 2
        dms, uv, dh, xyz, scales, rand_f, noise = synthesise(S, hm, poses[4:4 + Ns], seed)
        cams, inv = rand_cams(B, V, seed + 1)
        real = real_views(S, poses[:B], inv)
        net = HeatmapEstimationNetwork(hm, 0.01, 41, stacks, real_aug=aug)
        net.hg.load_state_dict(sd)
        crit = MultiTaskLoss(True, True, True, False, True, True, True, Constant, image_size=S, heatmap_size=hm)
        net.train()
        extra = {}
        if aug:                                                                          # a seed whose first uniform says "augment"
            while True:
                torch.manual_seed(seed + 2)
                if torch.rand(1).item() >= 0.5:
                    break
                seed += 1
        torch.manual_seed(seed + 2)
        if aug:                                                                          # create_network_and_criterion.py:95-101
            torch.rand(1)
            rnd = torch.rand(B * V) * 0.2 + 0.75
            extra['aug_u'] = rnd + torch.rand_like(rnd) * 0.1 - 0.05
            extra['aug_v'] = rnd + torch.rand_like(rnd) * 0.1 - 0.05
        eps = torch.stack([torch.randn(B * V, 32) for _ in range(stacks)])               # pose_vae.py:51, one draw per stack output
        torch.manual_seed(seed + 2)
        result = net(synt_dms=dms, real_dms=real * 0.01)
        terms, proj = crit(result, real_target={'real_dms': real, 'camera_poses': cams, 'inv_camera_poses': inv, 'is_mv': True},
                           synt_target={'uv_hms': uv, 'd_hms': dh, 'xyz_pts': xyz})
        loss = sum(terms.values())
        loss.backward()
        named = dict(net.hg.named_parameters())
        gn = {('gradnorm.' + k): v.grad.double().norm() for k, v in named.items()}
        pick = ['conv1.weight', 'layer1.0.conv2.weight', 'hg.0.hg.1.2.0.conv3.weight', 'fc.0.0.weight', 'score.0.weight', 'score.0.bias']
        if stacks == 2:
            pick += ['score.1.weight', 'score_.0.weight']
        save(name, poses_real=poses[:B], poses_synt=poses[4:4 + Ns], scales=scales, rand_f=rand_f, noise=noise, eps=eps,
             synt_dms=dms, uv_hms=uv, d_hms=dh, xyz_pts=xyz, real=real, cams=cams, inv_cams=inv,
             **{('real_xyz%d' % i): r for i, r in enumerate(result['real_xyz'])},
             **{('synt_xyz%d' % i): r for i, r in enumerate(result['synt_xyz'])},
             **{('real_uv_hm_max%d' % i): r.amax(dim=(-1, -2)) for i, r in enumerate(result['real_uv_hms'])},
             **{('projected_dms%d' % i): p for i, p in enumerate(proj)},
             **{('term.' + k): v for k, v in terms.items()}, loss=loss,
             **{('grad.' + k): named[k].grad for k in pick}, **gn, **extra,
             **({'real_resized_dms': result['real_resized_dms']} if aug else {}))
        return dms

    dms64 = step_fixture('trained_step_64', 1, 64, 16, sd1, seed=50)
    hourglass_fixture('trained_hourglass_64', 1, 64, sd1, dms64)
    step_fixture('trained_step_64_aug', 1, 64, 16, sd1, seed=50, aug=True)
    sd2 = two_stack_from_trained(sd1)
    dms128 = step_fixture('trained_step_128', 2, 128, 32, sd2, seed=60)
    hourglass_fixture('trained_hourglass_128', 2, 128, sd2, dms128)


if __name__ == '__main__':
    main()
