"""Oracle for the synthetic-data branch.  TEST INFRASTRUCTURE (see oracle/__init__).

Restates (torch fp32 on CPU / numpy)
  * HandTransformationMat (FK)   /root/reference/mesh/kinematicsTransformation.py:11-177
  * RandScale                    /root/reference/mesh/pointTransformation.py:135-148
  * LinearBlendSkinning          /root/reference/mesh/pointTransformation.py:21-46
  * (Inverse)OthographicalProjection  /root/reference/mesh/pointTransformation.py:84-99,118-124
  * DepthRasterization / DepthRender  /root/reference/mesh/render.py:282-331
  * HeatmapRender / Hand3DHeatmapRender  /root/reference/mesh/render.py:226-248,274-279
  * DepthNoise                   /root/reference/network/util_modules.py:60-84
  * HandSynthesizer.forward      /root/reference/network/util_modules.py:104-122
The triangle kernel itself is restated in C (oracle/tri_raster.c) and called through ctypes.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

# abduction axes per finger slot, kinematicsTransformation.py:162-166 (bone order finger4,3,2,1,5)
_ABDUCT_AXES = [(0., 0., 1.), (0., 0., 1.), (0., -1., 0.), (0., -1., 0.), (0., 0., 1.)]


def _axis_rot(axis, ang):
    """Rodrigues rotation about a fixed unit axis, batched.  kinematicsTransformation.py:39-53."""
    x, y, z = [torch.tensor(a, dtype=torch.float32) for a in axis]
    c, s = torch.cos(ang), torch.sin(ang)
    i = 1 - c
    m = torch.eye(4).repeat(ang.shape[0], 1, 1)
    m[:, 0, 0] = x * x * i + c
    m[:, 0, 1] = x * y * i - z * s
    m[:, 0, 2] = x * z * i + y * s
    m[:, 1, 0] = x * y * i + z * s
    m[:, 1, 1] = y * y * i + c
    m[:, 1, 2] = y * z * i - x * s
    m[:, 2, 0] = x * z * i - y * s
    m[:, 2, 1] = y * z * i + x * s
    m[:, 2, 2] = z * z * i + c
    return m


def forward_kinematics(params, offset_mats):
    """params [B,26], offset_mats: 17 x (4,4) float32 -> [B,17,4,4].  kinematicsTransformation.py:169-177."""
    B = params.shape[0]
    off = [torch.from_numpy(np.asarray(o, np.float32)) for o in offset_mats]
    rot = _axis_rot((1., 0., 0.), params[:, 0])
    rot = _axis_rot((0., 1., 0.), params[:, 1]) @ rot
    rot = _axis_rot((0., 0., 1.), params[:, 2]) @ rot
    tr = torch.eye(4).repeat(B, 1, 1)
    tr[:, :3, 3] = params[:, 3:6]
    palm = tr @ rot
    mats = [palm, palm]                                            # carpals share the palm transform (:154)
    for f in range(5):
        ang = params[:, 6 + 4 * f: 10 + 4 * f]
        parent = palm
        for k in range(3):
            o = off[2 + 3 * f + k]
            if k == 0:
                local = _axis_rot(_ABDUCT_AXES[f], ang[:, 0]) @ _axis_rot((1., 0., 0.), ang[:, 1])
            else:
                local = _axis_rot((1., 0., 0.), ang[:, k + 1])
            g = (torch.inverse(o)[None] @ local) @ o[None]
            parent = parent @ g
            mats.append(parent)
    return torch.stack(mats, dim=1)


def rand_scale_apply(mats, scales):
    """scales [B,3] in [0.85,0.95] left-multiplied on all bones.  pointTransformation.py:138-148."""
    S = torch.eye(4).repeat(mats.shape[0], 1, 1)
    S[:, 0, 0], S[:, 1, 1], S[:, 2, 2] = scales[:, 0], scales[:, 1], scales[:, 2]
    return S[:, None] @ mats


def skin_table(vertices, bones, key='mesh'):
    """Dense [17,Nv,4] table of w_b * v (fp32), pointTransformation.py:27-32.

    key='mesh' uses bone['weight_vertexid'/'weight_coeff']; key='keypoint' builds the 41 sphere centres
    (render.py:62-75), each rigidly attached to its bone with weight 1."""
    if key == 'mesh':
        v = np.asarray(vertices)
        tab = np.zeros((len(bones), v.shape[0], 4), np.float32)
        for b, bone in enumerate(bones):
            for w, i in zip(bone['weight_coeff'], bone['weight_vertexid']):
                tab[b, i] = w * v[i]
        return tab
    pts, owner = [], []
    for b, bone in enumerate(bones):
        for pt, _ in bone.get('keypoint', []):
            pts.append(np.asarray([pt[0], pt[1], pt[2], 1.0], np.float32))
            owner.append(b)
    tab = np.zeros((len(bones), len(pts), 4), np.float32)
    for i, (p, b) in enumerate(zip(pts, owner)):
        tab[b, i] = p
    return tab


def keypoint_radii(bones):
    return np.asarray([r for bone in bones for _, r in bone.get('keypoint', [])], np.float32)


def lbs(mats, table, right_hand=True):
    """mats [B,17,4,4], table [17,Nv,4] -> [B,Nv,4].  pointTransformation.py:39-46."""
    tab = torch.from_numpy(table)
    out = torch.einsum('bkxy,kvy->bvx', mats, tab)
    if right_hand:
        out = out.clone()
        out[:, :, 0] *= -1
    return out


def ortho_project(xyz, cx, cy, fx, fy, rand_f=None):
    """pointTransformation.py:84-99."""
    out = torch.ones_like(xyz)
    if rand_f is None:
        out[..., 0] = xyz[..., 0] * fx + cx * xyz[..., 3]
        out[..., 1] = xyz[..., 1] * fy + cy * xyz[..., 3]
        out[..., 2] = xyz[..., 2]
        out[..., 3] = xyz[..., 3]
        return out
    f = rand_f.view(-1, 1)
    out[..., 0] = xyz[..., 0] * f * fx + cx
    out[..., 1] = xyz[..., 1] * f * fy + cy
    out[..., 2] = xyz[..., 2]
    return out


# ----------------------------------------------------------------------------------------------
# R1: the C restatement


_lib = None


def _tri_lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, 'libtri_raster_oracle.so')
        src = os.path.join(HERE, 'tri_raster.c')
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-fopenmp', src, '-o', so, '-lm'])
        _lib = ctypes.CDLL(so)
        _lib.oracle_tri_raster.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.oracle_tri_raster.restype = None
    return _lib


def tri_raster(face_vertices, width, height, fma=False, want_face=False):
    """face_vertices [B,F,3,3] fp32 -> z-buffer [B,H,W] fp32 (1000.0 where empty) (+ winning face id int32).

    depth_rasterization_cuda_kernel.cu:18-134.  `fma=True` evaluates the per-pixel expressions with the
    fused multiply-adds nvcc's default -fmad=true contraction produces (closest to the GPU binary)."""
    fv = np.ascontiguousarray(np.asarray(face_vertices, np.float32))
    B, Fn = fv.shape[:2]
    out = np.empty((B, height, width), np.float32)
    face = np.empty((B, height, width), np.int32)
    _tri_lib().oracle_tri_raster(fv.ctypes.data, B, Fn, width, height, out.ctypes.data, face.ctypes.data, int(fma))
    return (out, face) if want_face else out


def bilinear_resize_640(dm, size):
    """F.interpolate(bilinear, align_corners=False) from 640 to `size` (render.py:311)."""
    t = torch.as_tensor(dm)[:, None]
    return torch.nn.functional.interpolate(t, size=(size, size), mode='bilinear', align_corners=False)[:, 0]


def depth_render(mats, mesh_table, faces, size, rand_f=None, fma=False):
    """DepthRender.forward, render.py:328-331 (+ clamp(max=100) :286, resize :311).

    `faces` must already carry the right-hand column swap of render.py:298-300."""
    pts = ortho_project(lbs(mats, mesh_table), 320, 320, 640 / 300, 640 / 300, rand_f)
    fv = pts[:, torch.as_tensor(faces.reshape(-1).astype(np.int64)), 0:3].reshape(mats.shape[0], -1, 3, 3)
    z = np.minimum(tri_raster(fv.numpy(), 640, 640, fma=fma), np.float32(100.0))
    return bilinear_resize_640(z, size), fv


def swap_faces_right_hand(faces):
    f = np.array(faces, copy=True)
    f[:, 0], f[:, 1] = faces[:, 1], faces[:, 0]
    return f


def heatmap_render(uvd, hm_size, sigma=1.0):
    """uvd [B,J,>=3] -> uv_hm [B,J,h,w], d_hm [B,J,h,w].  render.py:226-248."""
    g = torch.arange(hm_size, dtype=torch.float32)
    du = (g.view(1, 1, 1, -1) - uvd[..., 0][..., None, None]) ** 2
    dv = (g.view(1, 1, -1, 1) - uvd[..., 1][..., None, None]) ** 2
    hm = torch.exp(-0.5 * sigma * (du + dv))
    d = torch.where(hm > 0.05, uvd[..., 2][..., None, None].expand_as(hm), torch.zeros_like(hm))
    return hm, d


def hand_heatmaps(mats, kp_table, hm_size, rand_f=None):
    """Hand3DHeatmapRender.forward, render.py:274-279."""
    c, f = hm_size / 2, hm_size / 300
    uvd = ortho_project(lbs(mats, kp_table), c, c, f, f, rand_f)
    hm, d = heatmap_render(uvd, hm_size)
    xyz = uvd.clone()
    xyz[..., 0] = (uvd[..., 0] - c * uvd[..., 3]) / f          # inverse K, pointTransformation.py:118-124
    xyz[..., 1] = (uvd[..., 1] - c * uvd[..., 3]) / f
    return hm, d, xyz


def depth_noise(dm, nx, ny, nz, sx=0.5, sy=0.5, sz=0.05):
    """DepthNoise.forward with the three randn draws injected (nx, ny, nz ~ N(0,1), shape of dm).

    util_modules.py:60-84: integer pixel shuffle by trunc(n*sigma+0.5), then z-noise on pixels < 1.0."""
    B, H, W = dm.shape
    u = torch.arange(W).view(1, 1, W)
    v = torch.arange(H).view(1, H, 1)
    su = torch.clamp((nx * sx + 0.5).long() + u, 0, W - 1)
    sv = torch.clamp((ny * sy + 0.5).long() + v, 0, H - 1)
    out = torch.gather(dm.reshape(B, -1), 1, (sv * W + su).reshape(B, -1)).reshape(B, H, W)
    return torch.where(out < 1.0, out + nz * sz, out)


def resize_crop(depth_maps, u_scales, v_scales):
    """ResizeCropImage.forward, network/util_modules.py:388-424, statement by statement (torch CPU): nearest-neighbour
    resize of every image by (v_scale, u_scale), centred paste into an all-ones canvas.  Reproduces the reference's
    indentation quirk: the paste sits inside the `else` of `if v_scale > 1`, so an image with v_scale > 1 stays all ones."""
    import torch.nn.functional as F
    height, width = depth_maps.shape[-2:]
    cropped = torch.ones_like(depth_maps)
    for idx, (dm, u_scale, v_scale) in enumerate(zip(depth_maps, u_scales, v_scales)):
        dm = dm.view(1, 1, height, width)
        new_size = (int(height * v_scale + 0.5), int(width * u_scale + 0.5))
        resized = F.interpolate(dm, new_size)
        if u_scale > 1:
            u_start, u_end = 0, width
            orig_u_start = (int(width * u_scale + 0.5) - width) // 2
            orig_u_end = orig_u_start + width
        else:
            orig_u_start, orig_u_end = 0, int(width * u_scale)
            u_start = (width - int(width * u_scale + 0.5)) // 2
            u_end = u_start + orig_u_end
        if v_scale > 1:
            pass
        else:
            orig_v_start, orig_v_end = 0, int(height * v_scale)
            v_start = (height - int(height * v_scale + 0.5)) // 2
            v_end = v_start + orig_v_end
            cropped[idx, v_start:v_end, u_start:u_end] = resized[0, 0, orig_v_start:orig_v_end, orig_u_start:orig_u_end]
    return cropped.squeeze()
