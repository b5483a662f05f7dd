"""Drop-in mirror of /root/reference/mesh/pointTransformation.py for the hot path: LinearBlendSkinning (:11-46),
OthographicalProjection (:69-99), InverseOthographicalProjection (:102-124), RandScale (:128-148).

Forward-only, like their use in the reference (they run under HandSynthesizer, whose outputs are detached,
network/util_modules.py:122): a tensor that requires grad raises instead of silently dropping the graph.
`RandOthographicalProjection` (:49-66) is dead code in the reference (it reads an undefined attribute) and is not mirrored.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops


def _no_grad_only(t, who):
    if torch.is_grad_enabled() and t.requires_grad:
        raise NotImplementedError('%s is forward-only on the B200 path (the reference only runs it under the detached '
                                  'synthetic branch, network/util_modules.py:104-122)' % who)


def _f4(points):
    """[B,N,3|4] -> contiguous fp32 [B,N,4] (w = 1 when absent)."""
    if points.shape[-1] == 4:
        return points.contiguous().float()
    pad = torch.ones(points.shape[:-1] + (1,), device=points.device, dtype=points.dtype)
    return torch.cat([points[..., :3], pad], dim=-1).contiguous().float()


class LinearBlendSkinning(nn.Module):
    """vertices [Nv,4] (homogeneous rest positions), skinning_weights / skinning_vertex_indices: one list per bone.
    forward(bone_transformations [B,nb,4,4]) -> [B,Nv,4] = sum_b T_b (w_b v), x mirrored for the right hand (:39-46).
    The dense [1,nb,Nv,4,1] `skin_vertices` buffer of the reference (98 % zeros) is replaced by a CSR of the non-zero
    (bone, vertex) weights; products w*v are formed in fp32 exactly as :30 does."""

    def __init__(self, vertices, skinning_weights, skinning_vertex_indices, right_hand=True):
        super().__init__()
        assert len(skinning_vertex_indices) == len(skinning_weights), 'vertex index and weight should be with the same size'
        for vert_idx, vert_weight in zip(skinning_vertex_indices, skinning_weights):
            assert len(vert_idx) == len(vert_weight), 'vertex index and weight should be the same size'
        vertices = np.asarray(vertices)
        nv = vertices.shape[0]
        vid = np.concatenate([np.asarray(i, np.int64).reshape(-1) for i in skinning_vertex_indices] or [np.zeros(0, np.int64)])
        wco = np.concatenate([np.asarray(w, np.float64).reshape(-1) for w in skinning_weights] or [np.zeros(0)])
        bid = np.concatenate([np.full(len(i), b, np.int64) for b, i in enumerate(skinning_vertex_indices)] or [np.zeros(0, np.int64)])
        # the reference ASSIGNS cur_skin_vert[i] = w * v (:29-30): a later duplicate (bone, vertex) entry overwrites
        key = bid * nv + vid
        _, last = np.unique(key[::-1], return_index=True)
        keep = np.sort(len(key) - 1 - last)
        vid, wco, bid = vid[keep], wco[keep], bid[keep]
        wv = (wco[:, None] * vertices[vid]).astype(np.float32)
        order = np.lexsort((bid, vid))
        vid, bid, wv = vid[order], bid[order], wv[order]
        row_ptr = np.zeros(nv + 1, np.int64)
        np.add.at(row_ptr, vid + 1, 1)
        self.register_buffer('row_ptr', torch.from_numpy(np.cumsum(row_ptr).astype(np.int32)))
        self.register_buffer('bone', torch.from_numpy(bid.astype(np.int32)))
        self.register_buffer('wv', torch.from_numpy(np.ascontiguousarray(wv)))
        self.num_vertices = nv
        self.right_hand = right_hand

    def forward(self, bone_transformations, _mode=0, _cam=(0.0, 0.0, 1.0, 1.0), _rand_f=None):
        _no_grad_only(bone_transformations, 'LinearBlendSkinning')
        mats = bone_transformations.contiguous().float()
        return ops.lbs_fwd(mats, self.row_ptr, self.bone, self.wv, right_hand=self.right_hand, mode=_mode, cam=_cam, rand_f=_rand_f)


class OthographicalProjection(nn.Module):
    def __init__(self, cx, cy, fx, fy):
        super().__init__()
        k_mat = torch.eye(4)
        k_mat[0, 0], k_mat[1, 1], k_mat[0, 3], k_mat[1, 3] = fx, fy, cx, cy
        self.cx, self.cy, self.fx, self.fy = cx, cy, fx, fy
        self.register_buffer('k_mat', k_mat.unsqueeze(0).float())

    def forward(self, xyz_points, rand_f=None):
        _no_grad_only(xyz_points, 'OthographicalProjection')
        pts = _f4(xyz_points)
        cam = (self.cx, self.cy, self.fx, self.fy)
        if rand_f is None:
            return ops.ortho_project(pts, 2, cam)
        return ops.ortho_project(pts, 1, cam, rand_f.reshape(-1).contiguous().float())


class InverseOthographicalProjection(nn.Module):
    def __init__(self, cx, cy, fx, fy):
        super().__init__()
        k_mat = torch.eye(4)
        k_mat[0, 0], k_mat[1, 1], k_mat[0, 3], k_mat[1, 3] = fx, fy, cx, cy
        self.cx, self.cy, self.fx, self.fy = cx, cy, fx, fy
        self.register_buffer('inv_k_mat', torch.inverse(k_mat).unsqueeze(0).float())

    def forward(self, uvd_points):
        _no_grad_only(uvd_points, 'InverseOthographicalProjection')
        return ops.ortho_project(_f4(uvd_points), 3, (self.cx, self.cy, self.fx, self.fy))


class RandScale(nn.Module):
    """Anisotropic random scale in [0.9 - s/2, 0.9 + s/2] left-multiplied on every bone matrix (:135-148).  The three
    draws are made with the HOST generator in the reference's order (x, y, z; :140-142), then shipped to the device."""

    def __init__(self, rand_scale):
        super().__init__()
        self.rand_scale = rand_scale
        self.register_buffer('identity_mat', torch.eye(4).unsqueeze(dim=0).float())

    def draw(self, batch_size):
        rs = self.rand_scale
        x = torch.rand(batch_size) * rs + 0.90 - rs / 2
        y = torch.rand(batch_size) * rs + 0.90 - rs / 2
        z = torch.rand(batch_size) * rs + 0.90 - rs / 2
        return torch.stack([x, y, z], dim=1)

    def forward(self, transform_mats, scales=None):
        _no_grad_only(transform_mats, 'RandScale')
        if scales is None:
            scales = self.draw(transform_mats.shape[0])
        return ops.rand_scale_apply(transform_mats.contiguous().float(), scales.to(transform_mats.device).contiguous().float())
