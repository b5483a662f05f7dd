"""Device-side packing of the hand model: the data the reference modules pull out of `mesh` (the dict loaded from
mesh/model/preprocessed_hand.pkl, network/constants.py:4-5) laid out for the kernels.

  * offset matrices + their inverses            (FingerJoint.__init__, mesh/kinematicsTransformation.py:86-89)
  * CSR of the non-zero skinning weights, pre-multiplied with the rest vertices in fp32 exactly as
    LinearBlendSkinning.__init__ does (mesh/pointTransformation.py:27-32)
  * the 41 key-point spheres (centres rigidly attached to their bone, radii) (mesh/render.py:62-79)
  * faces with the right-hand winding swap of DepthRasterization.__init__ (mesh/render.py:298-300) applied to a COPY
    (the reference mutates the caller's array)
"""
import numpy as np
import torch


def _csr(num_vertices, bone_ids, vertex_ids, wv):
    order = np.lexsort((bone_ids, vertex_ids))
    vertex_ids, bone_ids, wv = vertex_ids[order], bone_ids[order], wv[order]
    row_ptr = np.zeros(num_vertices + 1, np.int32)
    np.add.at(row_ptr, vertex_ids + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    return row_ptr, bone_ids.astype(np.int32), wv.astype(np.float32)


class HandModel:
    def __init__(self, vertices, faces, offset_mats, weight_vertexid, weight_coeff, weight_bone, keypoints,
                 keypoint_radius, keypoint_bone, device, right_hand=True):
        dev = torch.device(device)
        vertices = np.asarray(vertices)
        self.num_vertices = vertices.shape[0]
        self.num_keypoints = len(keypoint_radius)
        off = np.asarray(offset_mats, np.float32)
        self.offset_mats = torch.from_numpy(off).to(dev).contiguous()
        # torch.inverse in fp32, like the reference (kinematicsTransformation.py:87)
        self.inv_offset_mats = torch.inverse(torch.from_numpy(off)).to(dev).contiguous()
        wv = (np.asarray(weight_coeff)[:, None] * vertices[np.asarray(weight_vertexid)]).astype(np.float32)
        self.mesh_csr = tuple(torch.from_numpy(a).to(dev).contiguous()
                              for a in _csr(self.num_vertices, np.asarray(weight_bone), np.asarray(weight_vertexid), wv))
        kp = np.concatenate([np.asarray(keypoints, np.float32), np.ones((self.num_keypoints, 1), np.float32)], axis=1)
        self.kp_csr = tuple(torch.from_numpy(a).to(dev).contiguous()
                            for a in _csr(self.num_keypoints, np.asarray(keypoint_bone), np.arange(self.num_keypoints), kp))
        f = np.array(faces, copy=True).astype(np.int32)
        if right_hand:
            f[:, [0, 1]] = f[:, [1, 0]]
        self.faces = torch.from_numpy(np.ascontiguousarray(f)).to(dev)
        self.radii = torch.tensor(np.asarray(keypoint_radius, np.float32)).to(dev)
        self.device = dev

    @classmethod
    def from_arrays(cls, a, device):
        """`a`: mapping with the keys of tests/golden/hand_model.npz."""
        return cls(a['vertices'], a['faces'], a['offset_mats'], a['weight_vertexid'], a['weight_coeff'], a['weight_bone'],
                   a['keypoints'], a['keypoint_radius'], a['keypoint_bone'], device)

    @classmethod
    def from_mesh(cls, mesh, device):
        """`mesh`: the reference's dict {'vertices','faces','bones':[{'offset_matrix','weight_vertexid','weight_coeff',
        'keypoint'?}]} (mesh/preprocess.py output)."""
        bones = mesh['bones']
        wid = np.concatenate([np.asarray(b['weight_vertexid'], np.int64) for b in bones])
        wco = np.concatenate([np.asarray(b['weight_coeff'], np.float64) for b in bones])
        wb = np.concatenate([np.full(len(b['weight_vertexid']), i, np.int64) for i, b in enumerate(bones)])
        kp, kr, kb = [], [], []
        for i, b in enumerate(bones):
            for pt, r in b.get('keypoint', []):
                kp.append(np.asarray(pt, np.float64)[:3])
                kr.append(r)
                kb.append(i)
        return cls(mesh['vertices'], mesh['faces'], np.stack([np.asarray(b['offset_matrix'], np.float32) for b in bones]),
                   wid, wco, wb, np.stack(kp), np.asarray(kr), np.asarray(kb), device)
