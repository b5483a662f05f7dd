"""Synthetic, seeded stand-ins for the step's inputs (SURVEY §8d "synthetic NYU-shape"): there is no dataset on the
box.  Host-side plumbing, not part of the measured path except where bench.py says so.

  * random_poses      vectorised draw in the spirit of JointAngleDataset.__getitem__ (dataset/joint_angle.py:21-233):
                      palm Euler angles / translation with the reference's ranges, per-finger abduction + three flexions
  * random_cameras    camera_poses[b,0] = I, the others a rotation <= 30 deg about a random axis, zero translation,
                      inv_camera_poses = inverse (dataset/nyu_dataset.py:20-21)
  * synthetic_real_batch  "real" depth maps: the 41-sphere hand of a random pose rendered into every view with the
                      sphere renderer (foreground in mm, background 100.0 as dataset/utils.py:75,96 produces)
"""
import math

import torch

from . import ops


def random_poses(n, generator=None, device='cpu'):
    g = generator
    r = lambda *s: torch.rand(*s, generator=g)
    p = torch.zeros(n, 26)
    p[:, 0] = r(n) * 6.28 - 3.14
    p[:, 1] = -r(n) * 3.14
    p[:, 2] = r(n) * 6.28 - 3.14
    p[:, 3] = r(n) * 30 - 15
    p[:, 4] = r(n) * 30 - 15
    p[:, 5] = r(n) * 50 - 35
    spread = (r(n) - 0.35) / 1.55
    jitter = lambda: (r(n) * 10 - 5) * math.pi / 180
    # finger blocks: 6 index, 10 middle, 14 ring, 18 pinky, 22 thumb; (abduct, flex1, flex2, flex3) each
    for base, k in ((6, 1.55), (10, 0.75), (14, -0.75), (18, -2.2)):
        p[:, base] = k * (spread + jitter()) * 0.3
        closed = (r(n) < 0.5).float()
        for f in range(3):
            p[:, base + 1 + f] = closed * (r(n) * 0.9 + 0.5) + (1 - closed) * (r(n) * 0.4 - 0.1)
    p[:, 22] = r(n) * 0.6 - 0.3
    p[:, 23:26] = r(n, 3) * 0.7 - 0.1
    return p.to(device)


def random_cameras(B, V, generator=None, device='cpu', max_angle=math.pi / 6):
    g = generator
    cams = torch.eye(4).repeat(B, V, 1, 1)
    axis = torch.randn(B, V, 3, generator=g)
    axis = axis / axis.norm(dim=-1, keepdim=True)
    ang = (torch.rand(B, V, generator=g) * 2 - 1) * max_angle
    ang[:, 0] = 0
    K = torch.zeros(B, V, 3, 3)
    K[..., 0, 1], K[..., 0, 2] = -axis[..., 2], axis[..., 1]
    K[..., 1, 0], K[..., 1, 2] = axis[..., 2], -axis[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -axis[..., 1], axis[..., 0]
    s, c = torch.sin(ang)[..., None, None], torch.cos(ang)[..., None, None]
    cams[..., :3, :3] = torch.eye(3) + s * K + (1 - c) * (K @ K)
    inv = torch.inverse(cams)
    return cams.to(device).contiguous(), inv.to(device).contiguous()


def synthetic_real_batch(hand, B, V, S, generator=None):
    """-> (real_dms [B,V,S,S] mm, camera_poses, inv_camera_poses [B,V,4,4]) on hand.device, rendered by the CUDA kernels."""
    dev = hand.device
    poses = random_poses(B, generator, dev)
    poses[:, 3:6] *= 0.5                      # keep the hand inside every rotated view
    cams, inv = random_cameras(B, V, generator, dev)
    mats = ops.fk_fwd(poses, hand.offset_mats, hand.inv_offset_mats)
    centres = ops.lbs_fwd(mats, *hand.kp_csr, right_hand=True, mode=0)[..., :3]        # [B,J,3] canonical frame
    pv = torch.einsum('bvxy,bky->bvkx', inv[..., :3, :3], centres) + inv[:, :, None, :3, 3]
    sph = ops.pack_spheres(pv.reshape(B * V, -1, 3).contiguous(), hand.radii)
    depth, _ = ops.sphere_render_fwd(sph, S, S)
    return depth.view(B, V, S, S), cams, inv
