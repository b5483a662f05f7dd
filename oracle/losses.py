"""Oracle for the self-supervision loss heads.  TEST INFRASTRUCTURE (see oracle/__init__).

torch-fp32-on-CPU restatements (autograd supplies the reference gradients) of
  * MutualTransformation / MutualProjection / MutualProjectionLoss
        /root/reference/mesh/multiview_utility.py:13-30, 55-77, 90-130
  * DataToModelLoss            /root/reference/mesh/render.py:123-142
  * MultiviewConsistencyLoss   /root/reference/mesh/multiview_utility.py:138-167
  * CollisionLoss              /root/reference/mesh/render.py:145-176
  * BoneLengthLoss             /root/reference/mesh/render.py:179-206 (+ tables mesh/bone_length.py:36-56)
  * PoseVae.prior_loss         /root/reference/network/pose_vae.py:25-62, 81-89
  * RecoverXYZCoordinateFromHeatmap  /root/reference/network/util_modules.py:126-201
"""
import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# tables (restated from mesh/render.py:150-162 and mesh/bone_length.py:36-56)


def collision_pairs():
    """690 (i,j) pairs: 11 palm spheres x 30 finger spheres + cross-finger pairs. render.py:153-162."""
    a, b = [], []
    for i in range(11):
        for j in range(11, 41):
            a.append(i)
            b.append(j)
    for i in range(11, 41):
        for j in range(i + 1, 41):
            if (i - 11) // 6 != (j - 11) // 6:
                a.append(i)
                b.append(j)
    return np.asarray(a, np.int64), np.asarray(b, np.int64)


def bone_pairs():
    """35 (i,j) pairs + rest lengths in mm.  mesh/bone_length.py:36-56."""
    a = [3, 2, 3, 8, 2, 2, 9, 8, 4, 8, 7, 4, 6, 7, 0, 5, 7, 7, 6, 6]
    b = [2, 9, 8, 2, 4, 10, 10, 4, 10, 7, 4, 6, 10, 6, 5, 1, 0, 5, 5, 1]
    for f in range(5):
        for k in range(3):
            a.append(11 + 2 * k + 6 * f)
            b.append(12 + 2 * k + 6 * f)
    length = [25.212656021118164, 18.249488830566406, 27.5742244720459, 38.532264709472656,
              25.10819435119629, 31.173757553100586, 18.329626083374023, 19.15080451965332,
              16.209327697753906, 21.52261734008789, 32.740535736083984, 30.58920669555664,
              33.205970764160156, 11.672294616699219, 17.084707260131836, 17.084720611572266,
              16.697546005249023, 23.92103385925293, 20.87999725341797, 22.58038330078125,
              27.55999755859375, 15.471183776855469, 13.214692115783691, 21.748210906982422,
              13.021653175354004, 16.643720626831055, 18.83765983581543, 12.724685668945312,
              16.238431930541992, 18.04928970336914, 11.045844078063965, 11.320968627929688,
              30.078536987304688, 16.255985260009766, 19.434825897216797]
    return np.asarray(a, np.int64), np.asarray(b, np.int64), np.asarray(length, np.float32)


# ----------------------------------------------------------------------------------------------
# sphere renderer in torch (differentiable twin of oracle/sphere.py, used for loss gradients)


def _grids(size, dtype=torch.float32):
    u = torch.arange(size, dtype=dtype)
    g = (u - size / 2) * 300.0 / size
    return g


def ball_depth(centres, radii, size):
    """centres [...,J,3], radii [J] -> per-sphere depth [...,J,S,S].  render.py:26-53."""
    g = _grids(size, centres.dtype)
    cx = centres[..., 0][..., None, None]
    cy = centres[..., 1][..., None, None]
    cz = centres[..., 2][..., None, None]
    r = radii.to(centres.dtype).view(*([1] * (centres.dim() - 2)), -1, 1, 1)
    s = r * r - (g.view(1, -1) - cx) ** 2 - (g.view(-1, 1) - cy) ** 2
    fg = s > 0.01 if centres.dtype == torch.float64 else s > torch.tensor(1e-2, dtype=torch.float32)
    sq = torch.sqrt(torch.clamp(s, min=1e-2))
    return torch.where(fg, cz - sq, torch.full_like(sq, 100.0))


def mutual_transforms(cam, inv_cam):
    """T[b,i,j] = inv_cam[b,j] @ cam[b,i].  multiview_utility.py:19-29."""
    return torch.matmul(inv_cam[:, None, :, :, :], cam[:, :, None, :, :])


def mutual_projection(cam, inv_cam, joints, radii, size):
    """-> projected depth [B,V,V,S,S], projected joints [B,V,V,J,3].  multiview_utility.py:55-77."""
    T = mutual_transforms(cam, inv_cam).detach()
    R = T[..., :3, :3]
    t = T[..., :3, 3]
    p = torch.einsum('bijxy,biky->bijkx', R, joints) + t[:, :, :, None, :]
    depth = ball_depth(p, radii, size).min(dim=3).values
    return depth, p


def data_to_model(dms, joints, radii, size):
    """dms [N,S,S] (mm, background > 99), joints [N,J,3] -> scalar.  render.py:123-142."""
    g = _grids(size, joints.dtype)
    n = dms.shape[0]
    P = torch.stack([g.view(1, 1, -1).expand(n, size, size),
                     g.view(1, -1, 1).expand(n, size, size), dms], dim=-1)       # [N,S,S,3]
    d = (P[:, None] - joints[:, :, None, None, :]).norm(dim=-1)                    # [N,J,S,S]
    e = (d - radii.to(joints.dtype).view(1, -1, 1, 1)).abs()
    e = torch.where((dms > 99)[:, None].expand_as(e), torch.zeros_like(e), e)
    m = e.min(dim=1).values
    return torch.clamp(m, min=0, max=50).mean()


def mutual_projection_loss(cam, inv_cam, joints, depth_maps, radii, is_mv=True):
    """multiview_utility.py:90-130.  Returns (loss, projected_dms, model_to_data, data_to_model)."""
    B, V, J = joints.shape[:3]
    size = depth_maps.shape[-1]
    proj, p = mutual_projection(cam, inv_cam, joints, radii, size)
    real = depth_maps[:, None].expand(B, V, V, size, size)         # [b,i,j] compares with real view j
    if is_mv:
        m2d = F.mse_loss(proj, real) * 9
        d2m = data_to_model(real.reshape(-1, size, size), p.reshape(-1, J, 3), radii, size) * 9
    else:
        m2d = sum(F.mse_loss(proj[:, v, v], real[:, v, v]) for v in range(3)) * 3
        d2m = sum(data_to_model(real[:, v, v], p[:, v, v], radii, size) for v in range(3)) * 3
    return m2d + d2m * 500, proj, m2d, d2m


def multiview_consistency(cam, joints):
    """multiview_utility.py:138-167 with hm_weight=None (the only call-site form)."""
    q = torch.einsum('bvxy,bvky->bvkx', cam[..., :3, :3], joints) + cam[:, :, None, :3, 3]
    med = torch.median(q, dim=1).values
    return F.mse_loss(med[:, None].expand_as(q), q)


def collision_loss(xyz, min_dist=6.0):
    """xyz [B,V,41,3]; the reference flattens to (B,123,3) and indexes 0..40 => view 0 only.  render.py:168-176."""
    a, b = collision_pairs()
    j = xyz.reshape(xyz.shape[0], -1, 3)
    sq = ((j[:, a] - j[:, b]) ** 2).sum(-1)
    return F.relu(min_dist ** 2 - sq).sum()


def bone_length_loss(xyz):
    """render.py:196-206 (same view-0-only quirk)."""
    a, b, length = bone_pairs()
    length = torch.from_numpy(length)
    j = xyz.reshape(xyz.shape[0], -1, 3)
    sq = ((j[:, a] - j[:, b]) ** 2).sum(-1)
    lo = ((length * 0.80) ** 2)[None]
    hi = ((length * 1.05) ** 2)[None]
    return F.relu(lo - sq).mean() + F.relu(sq - hi).mean()


def vae_prior_loss(x, w, eps, parts=False):
    """pose_vae.py:81-89.  x [M,123] (already /100), w = dict of the state_dict tensors, eps [M,32].
    parts=True -> (reconstruction MSE [a batch MEAN], KLD [a batch SUM]) separately (data-parallel reduction rules)."""
    def lin(h, name):
        return F.linear(h, w[name + '.weight'], w[name + '.bias'])

    def gn(h, name):
        return F.group_norm(h, 16, w[name + '.weight'], w[name + '.bias'])

    h = F.relu(gn(lin(x, 'base.0'), 'base.1'))
    h = F.relu(gn(lin(h, 'base.3'), 'base.4'))
    mu = lin(h, 'mu')
    logvar = lin(h, 'logvar')
    z = eps * (torch.exp(0.5 * logvar) * 0.1) + mu
    h = F.relu(gn(lin(z, 'decoder.0'), 'decoder.1'))
    h = F.relu(gn(lin(h, 'decoder.3'), 'decoder.4'))
    recon = lin(h, 'decoder.6')
    kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    if parts:
        return F.mse_loss(x, recon), kld
    return F.mse_loss(x, recon) + kld


def soft_argmax_xyz(uv_hms, d_hms, depth_scale=0.01):
    """util_modules.py:182-201.  uv_hms, d_hms [N,J,h,w] -> xyz [N,J,3] in mm."""
    n, j, h, w = uv_hms.shape
    p = F.softmax((uv_hms * 20.0).reshape(n * j, h * w), dim=1).reshape(n, j, h, w)
    ug = torch.arange(w, dtype=uv_hms.dtype).view(1, 1, 1, w)
    vg = torch.arange(h, dtype=uv_hms.dtype).view(1, 1, h, 1)
    rl = F.relu(uv_hms).reshape(n, j, h * w)
    nrm = (rl / (rl.sum(-1, keepdim=True) + 1e-5)).reshape(n, j, h, w)
    u = (p * ug).reshape(n, j, -1).sum(2)
    v = (p * vg).reshape(n, j, -1).sum(2)
    d = (d_hms * nrm).reshape(n, j, -1).sum(2)
    x = (u - w / 2) / (w / 300.0)
    y = (v - h / 2) / (h / 300.0)
    z = d * (1.0 / depth_scale)
    return torch.stack([x, y, z], dim=-1)


def pose_denoiser_forward(sd, fea, input_indices, output_indices, scale_factor=0.01):
    """PoseDenoiser.forward in eval mode, network/pose_denoiser.py:56-73, statement by statement (torch fp32 CPU).
    sd: state_dict with the reference's keys (network.0/1/3/4/6.*)."""
    import torch.nn.functional as F
    is_skel = fea.ndimension() == 3
    shape = fea.shape
    if is_skel:
        fea = fea.reshape(shape[0], -1)
    x = fea[:, input_indices] * scale_factor
    x = F.relu(F.group_norm(F.linear(x, sd['network.0.weight'], sd['network.0.bias']), 16, sd['network.1.weight'], sd['network.1.bias']))
    x = F.relu(F.group_norm(F.linear(x, sd['network.3.weight'], sd['network.3.bias']), 16, sd['network.4.weight'], sd['network.4.bias']))
    y = F.linear(x, sd['network.6.weight'], sd['network.6.bias']) / scale_factor
    out = fea.clone()
    out[:, output_indices] = y
    return out.reshape(shape) if is_skel else out
