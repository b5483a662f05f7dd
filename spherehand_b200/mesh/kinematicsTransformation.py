"""Drop-in mirror of /root/reference/mesh/kinematicsTransformation.py for the hot path: HandTransformationMat
(:157-177, with AxisRotationMatrix :11-53, TranslationMatrix :56-68, FingerJoint :83-112, Finger :114-127, Palm :129-155
folded into ONE kernel launch, sh_fk_fwd) and SkeletonFK (:180-207).  Forward-only (the reference never
back-propagates through FK in any epoch loop)."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .pointTransformation import LinearBlendSkinning, RandScale, _no_grad_only


class HandTransformationMat(nn.Module):
    """offset_mats: 17 4x4 bone offset matrices.  forward(parameters [B,26]) -> [B,17,4,4]; parameter layout: [0:3] palm
    Euler x,y,z, [3:6] translation, then (abduct, flex1, flex2, flex3) per finger in bone order (:166,174)."""

    def __init__(self, offset_mats):
        super().__init__()
        off = np.stack([np.asarray(m, np.float32) for m in offset_mats])
        assert off.shape == (17, 4, 4), 'the hand model has 17 bones (palm, carpals, 5 x 3 finger bones)'
        self.register_buffer('offset_mats', torch.from_numpy(off))
        # torch.inverse in fp32, as FingerJoint.__init__ does (:87)
        self.register_buffer('inv_offset_mats', torch.inverse(torch.from_numpy(off)).contiguous())

    def forward(self, parameters, scales=None):
        _no_grad_only(parameters, 'HandTransformationMat')
        assert parameters.dim() == 2 and parameters.shape[1] == 26, 'parameters must be [B,26]'
        sc = None if scales is None else scales.to(parameters.device).contiguous().float()
        return ops.fk_fwd(parameters.contiguous().float(), self.offset_mats, self.inv_offset_mats, sc)


class SkeletonFK(nn.Module):
    """mesh dict -> forward(para [B,26]) -> the 41 key-points [B,41,4] after RandScale(0.2) (:180-207)."""

    def __init__(self, mesh):
        super().__init__()
        self.hand_skeleton_transform = HandTransformationMat([b['offset_matrix'].astype(np.float32) for b in mesh['bones']])
        self.rand_scale = RandScale(0.2)
        vertices, weights, indices = keypoint_skin(mesh['bones'])
        self.num_vertices = len(vertices)
        self.lbs = LinearBlendSkinning(vertices, weights, indices)

    def forward(self, para):
        scales = self.rand_scale.draw(para.shape[0])
        return self.lbs(self.hand_skeleton_transform(para, scales))


def keypoint_skin(bones):
    """The key-point 'mesh' the reference builds in HandBallPrimitiveRender / Hand3DHeatmapRender / SkeletonFK
    (mesh/render.py:62-75, 260-272): every key-point rigidly attached (weight 1) to its bone."""
    vertices, weights, indices = [], [], []
    for bone in bones:
        weights.append([])
        indices.append([])
        if 'keypoint' in bone:
            for pt, _ in bone['keypoint']:
                vertices.append(np.asarray([pt[0], pt[1], pt[2], 1.0], np.float32))
                weights[-1].append(1.0)
                indices[-1].append(len(vertices) - 1)
    return np.asarray(vertices).astype(np.float32), weights, indices


def keypoint_radii(mesh):
    """radii of the key-point spheres from a mesh dict, or a plain list (mesh/render.py:109-117)."""
    if type(mesh) == dict:
        return [r for bone in mesh['bones'] if 'keypoint' in bone for _, r in bone['keypoint']]
    if type(mesh) == list:
        return list(mesh)
    raise TypeError('mesh can only be list or dict')
