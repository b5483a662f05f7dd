"""Tensor-level wrappers over the C ABI: one Python function per kernel entry, torch used only for device
memory and streams.  Every function raises if its inputs are not CUDA/contiguous/of the expected dtype, mirroring
the CHECK_CUDA / CHECK_CONTIGUOUS of the reference shim (mesh/cuda_kernel/depth_rasterization_cuda.cpp:11-13)."""
import torch

from . import _lib

_call = _lib.call


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype=torch.float32, name='tensor'):
    if not isinstance(t, torch.Tensor):
        raise TypeError('%s must be a torch.Tensor' % name)
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor' % name)
    if not t.is_contiguous():
        raise RuntimeError('%s must be contiguous' % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError('%s must be %s (got %s)' % (name, dtype, t.dtype))
    return t.data_ptr()


def _opt(t, dtype=torch.float32, name='tensor'):
    return None if t is None else _chk(t, dtype, name)


BF16 = torch.bfloat16

# ------------------------------------------------------------------------------------------------ renderers


def pack_spheres(centres, radii):
    """centres [N,J,>=3], radii [J] | [N,J] | [N*J] -> float4 spheres [N,J,4] (cx,cy,cz,r)."""
    n, j = centres.shape[:2]
    r = radii.reshape(-1, j).expand(n, j) if radii.numel() != n * j else radii.reshape(n, j)
    return torch.cat([centres[..., :3], r.unsqueeze(-1).to(centres.dtype)], dim=-1).contiguous().float()


def sphere_render_fwd(spheres, H, W):
    n, j = spheres.shape[:2]
    depth = torch.empty((n, H, W), device=spheres.device, dtype=torch.float32)
    idx = torch.empty((n, H, W), device=spheres.device, dtype=torch.uint8)
    _call('sh_sphere_render_fwd', _chk(spheres, name='spheres'), n, j, H, W, depth.data_ptr(), idx.data_ptr(), _stream())
    return depth, idx


def sphere_render_bwd(grad_depth, idx, spheres):
    n, j = spheres.shape[:2]
    H, W = idx.shape[-2:]
    g = torch.empty_like(spheres)
    _call('sh_sphere_render_bwd', _chk(grad_depth, name='grad_depth'), _chk(idx, torch.uint8, 'idx'),
          _chk(spheres, name='spheres'), n, j, H, W, g.data_ptr(), _stream())
    return g


def tri_raster_fwd(face_vertices, width, height):
    _chk(face_vertices, name='vertices')
    b, f = face_vertices.shape[0], face_vertices.shape[1]
    out = torch.empty((b, height, width), device=face_vertices.device, dtype=torch.float32)
    _call('sh_tri_raster_fwd', face_vertices.data_ptr(), b, f, width, height, out.data_ptr(), _stream())
    return out


def tri_raster_lattice_fwd(face_vertices, size, step, off0, noff):
    b, f = face_vertices.shape[0], face_vertices.shape[1]
    o = (size // step) * noff
    out = torch.empty((b, o, o), device=face_vertices.device, dtype=torch.float32)
    _call('sh_tri_raster_lattice_fwd', _chk(face_vertices, name='vertices'), b, f, size, size, step, off0, noff,
          out.data_ptr(), o, o, _stream())
    return out


# ------------------------------------------------------------------------------------------------ loss heads


def mvproj_loss_fwdbwd(cam, inv_cam, joints, real, radii, is_mv=True):
    B, V, J = joints.shape[:3]
    H, W = real.shape[-2:]
    dev = joints.device
    proj = torch.empty((B, V, V, H, W), device=dev, dtype=torch.float32)
    loss3 = torch.empty(3, device=dev, dtype=torch.float32)
    grad = torch.empty((B, V, J, 3), device=dev, dtype=torch.float32)
    nbytes = _lib.lib().sh_mvproj_scratch_bytes(B, V, J)
    scratch = torch.empty(nbytes // 4 + 4, device=dev, dtype=torch.float32)
    _call('sh_mvproj_loss_fwdbwd', _chk(cam, name='camera_poses'), _chk(inv_cam, name='inv_camera_poses'),
          _chk(joints, name='joints'), _chk(real, name='depth_maps'), _chk(radii, name='radii'), B, V, J, H, W,
          int(bool(is_mv)), proj.data_ptr(), loss3.data_ptr(), grad.data_ptr(), scratch.data_ptr(), _stream())
    return loss3, proj, grad


def pose_losses_fwdbwd(cam, joints, flags=7, min_dist=6.0):
    B, V, J = joints.shape[:3]
    dev = joints.device
    losses = torch.empty(3, device=dev, dtype=torch.float32)
    grads = torch.empty((3, B, V, J, 3), device=dev, dtype=torch.float32)
    scratch = torch.empty(8, device=dev, dtype=torch.float64)
    _call('sh_pose_losses_fwdbwd', _chk(cam, name='camera_poses'), _chk(joints, name='joints'), B, V, J, flags,
          float(min_dist), losses.data_ptr(), grads.data_ptr(), scratch.data_ptr(), _stream())
    return losses, grads


def vae_blob_from_state_dict(sd, device):
    """Pack PoseVae's state_dict (network/pose_vae.py:26-46) into the blob layout of csrc/pose_vae.cu."""
    parts = []
    for lin, gn in (('base.0', 'base.1'), ('base.3', 'base.4'), ('mu', None), ('logvar', None),
                    ('decoder.0', 'decoder.1'), ('decoder.3', 'decoder.4'), ('decoder.6', None)):
        w = sd[lin + '.weight'].detach().float().cpu()
        parts += [w.reshape(-1), w.t().contiguous().reshape(-1), sd[lin + '.bias'].detach().float().cpu().reshape(-1)]
        if gn:
            parts += [sd[gn + '.weight'].detach().float().cpu().reshape(-1), sd[gn + '.bias'].detach().float().cpu().reshape(-1)]
    blob = torch.cat(parts).contiguous()
    assert blob.numel() == _lib.lib().sh_vae_blob_floats(), (blob.numel(), _lib.lib().sh_vae_blob_floats())
    return blob.to(device)


def vae_prior_fwdbwd(x, eps, blob, M_mean=None):
    M = x.shape[0]
    dev = x.device
    loss3 = torch.empty(3, device=dev, dtype=torch.float32)
    grad = torch.empty((M, 123), device=dev, dtype=torch.float32)
    scratch = torch.empty(4, device=dev, dtype=torch.float64)
    _call('sh_vae_prior_fwdbwd', _chk(x, name='x'), _chk(eps, name='eps'), _chk(blob, name='weights'), M,
          M if M_mean is None else M_mean, loss3.data_ptr(), grad.data_ptr(), scratch.data_ptr(), _stream())
    return loss3, grad


def softargmax_fwd(score, J, Ns=0, target_uv=None, depth_scale_inv=100.0, want_sse=False, want_aux=False):
    """-> (xyz [N,J,3], sse [2] f64 | None[, aux [N,J,8]: the per-(sample, joint) scalars softargmax_bwd_nhwc needs])."""
    N, C, h, w = score.shape
    xyz = torch.empty((N, J, 3), device=score.device, dtype=torch.float32)
    sse = torch.empty(2, device=score.device, dtype=torch.float64) if want_sse else None
    if want_aux:
        aux = torch.empty((N, J, 8), device=score.device, dtype=torch.float32)
        _call('sh_softargmax_fwd_aux', _chk(score, name='score'), N, Ns, J, C, h, w, float(depth_scale_inv),
              _opt(target_uv, name='target_uv'), xyz.data_ptr(), None if sse is None else sse.data_ptr(), aux.data_ptr(), _stream())
        return xyz, sse, aux
    _call('sh_softargmax_fwd', _chk(score, name='score'), N, Ns, J, C, h, w, float(depth_scale_inv),
          _opt(target_uv, name='target_uv'), xyz.data_ptr(), None if sse is None else sse.data_ptr(), _stream())
    return xyz, sse


def softargmax_bwd_nhwc(score, gxyz, aux, J, Ns=0, target_uv=None, depth_scale_inv=100.0, c_synt=0.0, c_real=0.0, Cp=128):
    """d loss / d score as the bf16 NHWC tensor [N,h,w,Cp] of the network's backward pass (channels >= 2J zero)."""
    N, C, h, w = score.shape
    out = torch.empty((N, h, w, Cp), device=score.device, dtype=BF16)
    _call('sh_softargmax_bwd_nhwc', _chk(score, name='score'), _chk(gxyz, name='gxyz'), _chk(aux, name='aux'), N, Ns, J, C, h, w,
          float(depth_scale_inv), _opt(target_uv, name='target_uv'), float(c_synt), float(c_real), out.data_ptr(), Cp, _stream())
    return out


def softargmax_bwd(score, gxyz, J, Ns=0, target_uv=None, depth_scale_inv=100.0, c_synt=0.0, c_real=0.0, out=None):
    N, C, h, w = score.shape
    g = torch.zeros_like(score) if out is None else out
    _call('sh_softargmax_bwd', _chk(score, name='score'), _chk(gxyz, name='gxyz'), N, Ns, J, C, h, w,
          float(depth_scale_inv), _opt(target_uv, name='target_uv'), float(c_synt), float(c_real), _chk(g, name='gscore'),
          _stream())
    return g


# ------------------------------------------------------------------------------------------------ synthetic branch


def fk_fwd(params, offset_mats, inv_offset_mats, scales=None):
    B = params.shape[0]
    mats = torch.empty((B, 17, 4, 4), device=params.device, dtype=torch.float32)
    _call('sh_fk_fwd', _chk(params, name='params'), _opt(scales, name='scales'), _chk(offset_mats, name='offset_mats'),
          _chk(inv_offset_mats, name='inv_offset_mats'), B, mats.data_ptr(), _stream())
    return mats


def lbs_fwd(mats, row_ptr, bone, wv, right_hand=True, mode=0, cam=(0.0, 0.0, 1.0, 1.0), rand_f=None):
    B = mats.shape[0]
    Nv = row_ptr.numel() - 1
    out = torch.empty((B, Nv, 4), device=mats.device, dtype=torch.float32)
    _call('sh_lbs_fwd', _chk(mats, name='mats'), _chk(row_ptr, torch.int32, 'row_ptr'), _chk(bone, torch.int32, 'bone'),
          _chk(wv, name='wv'), B, Nv, int(right_hand), mode, float(cam[0]), float(cam[1]), float(cam[2]), float(cam[3]),
          _opt(rand_f, name='rand_f'), out.data_ptr(), _stream())
    return out


def gather_faces(points, faces):
    B, Nv = points.shape[:2]
    F = faces.shape[0]
    fv = torch.empty((B, F, 3, 3), device=points.device, dtype=torch.float32)
    _call('sh_gather_faces', _chk(points, name='points'), _chk(faces, torch.int32, 'faces'), B, Nv, F, fv.data_ptr(), _stream())
    return fv


def lattice_to_depth(z, S, noff, depth_scale):
    B = z.shape[0]
    dm = torch.empty((B, S, S), device=z.device, dtype=torch.float32)
    _call('sh_lattice_to_depth', _chk(z, name='z'), B, S, noff, float(depth_scale), dm.data_ptr(), _stream())
    return dm


def depth_noise(dm, nx, ny, nz, sx=0.5, sy=0.5, sz=0.05, out=None):
    B, H, W = dm.shape
    out = torch.empty_like(dm) if out is None else out
    _call('sh_depth_noise', _chk(dm, name='dm'), _chk(nx, name='nx'), _chk(ny, name='ny'), _chk(nz, name='nz'), B, H, W,
          sx, sy, sz, _chk(out, name='out'), _stream())
    return out


def heatmap_render(uvd, hm, sigma=1.0, uv_scale=1.0, depth_scale=1.0, cam=None):
    B, J = uvd.shape[:2]
    cam = cam or (hm / 2, hm / 2, hm / 300, hm / 300)
    uv = torch.empty((B, J, hm, hm), device=uvd.device, dtype=torch.float32)
    d = torch.empty_like(uv)
    xyz = torch.empty((B, J, 4), device=uvd.device, dtype=torch.float32)
    _call('sh_heatmap_render', _chk(uvd, name='uvd'), B, J, hm, float(sigma), float(uv_scale), float(depth_scale),
          float(cam[0]), float(cam[1]), float(cam[2]), float(cam[3]), uv.data_ptr(), d.data_ptr(), xyz.data_ptr(), _stream())
    return uv, d, xyz


# ------------------------------------------------------------------------------------------------ hourglass layers


def conv_fwd(x, w, bias, N, H, W, Cin, Cout, cout_pad, taps, y=None, y_ld=0, y_nchw=None, residual=None, stats=None,
             groups=0, gn=None):
    """gn = (stats [N,G,2], gamma, beta, G, eps): x is the RAW GroupNorm input and relu(groupnorm(x)) is formed in shared memory
    (1x1 layers, images of >= 64 pixels)."""
    if gn is not None:
        gst, gamma, beta, G, eps = gn
        if taps != 1:
            raise RuntimeError('conv_fwd: only 1x1 layers take a fused input GroupNorm')
        _call('sh_conv_fwd_gn', _chk(x, BF16, 'x'), _chk(gst, name='gn_stats'), _chk(gamma, name='gamma'), _chk(beta, name='beta'), G, eps,
              _chk(w, BF16, 'w'), _opt(bias, name='bias'), _opt(residual, BF16, 'residual'), N, H, W, Cin, Cout, cout_pad,
              _opt(y, BF16, 'y'), y_ld, _opt(y_nchw, name='y_nchw'), _opt(stats, name='stats'), groups, _stream())
        return
    _call('sh_conv_fwd', _chk(x, BF16, 'x'), _chk(w, BF16, 'w'), _opt(bias, name='bias'), _opt(residual, BF16, 'residual'),
          N, H, W, Cin, Cout, cout_pad, taps, _opt(y, BF16, 'y'), y_ld, _opt(y_nchw, name='y_nchw'), _opt(stats, name='stats'),
          groups, _stream())


def conv_wgrad(dy, x, N, H, W, x_C, Cin, dy_C, Cout, taps, dw, gn=None):
    """gn: as in conv_fwd -- the gradient is taken against relu(groupnorm(x))."""
    if gn is not None:
        gst, gamma, beta, G, eps = gn
        if taps != 1:
            raise RuntimeError('conv_wgrad: only 1x1 layers take a fused input GroupNorm')
        _call('sh_conv_wgrad_gn', _chk(dy, BF16, 'dy'), _chk(x, BF16, 'x'), _chk(gst, name='gn_stats'), _chk(gamma, name='gamma'),
              _chk(beta, name='beta'), G, eps, N, H, W, x_C, Cin, dy_C, Cout, _chk(dw, name='dw'), _stream())
        return
    _call('sh_conv_wgrad', _chk(dy, BF16, 'dy'), _chk(x, BF16, 'x'), N, H, W, x_C, Cin, dy_C, Cout, taps, _chk(dw, name='dw'),
          _stream())


def conv_wgrad3x3(dy, x, N, H, W, x_C, Cin, dy_C, Cout, scratch):
    _call('sh_conv_wgrad3x3', _chk(dy, BF16, 'dy'), _chk(x, BF16, 'x'), N, H, W, x_C, Cin, dy_C, Cout, _chk(scratch, name='scratch'),
          _stream())


def unpack_wgrad_batch(table, scratch, grad):
    _call('sh_unpack_wgrad_batch', _chk(table, torch.int32, 'table'), table.shape[0], _chk(scratch, name='scratch'),
          _chk(grad, name='grad'), _stream())


def gn_relu_fwd(x, stats_in, gamma, beta, N, HW, C, G, y, stats_out=None, G_out=16, eps=1e-5):
    _call('sh_gn_relu_fwd', _chk(x, BF16, 'x'), _chk(stats_in, name='stats'), _chk(gamma, name='gamma'), _chk(beta, name='beta'),
          N, HW, C, G, eps, _chk(y, BF16, 'y'), _opt(stats_out, name='stats_out'), G_out, _stream())


def gn_relu_bwd_scratch_words(N, G):
    return _lib.lib().sh_gn_relu_bwd_scratch_words(N, G)


def gn_relu_bwd_scratch(N, G, device):
    """Scratch tensor sh_gn_relu_bwd needs (per-(n,g) partial sums + per-sample arrive counters)."""
    return torch.empty(_lib.lib().sh_gn_relu_bwd_scratch_words(N, G), device=device, dtype=torch.float32)


def gn_relu_bwd(da, x, stats_in, gamma, beta, N, HW, C, G, red, dgamma, dbeta, dx, addend=None, colsum=None, eps=1e-5,
                prezeroed=False):
    """prezeroed: `red` is already all zero (a slice of an arena the caller cleared once): the entry skips its own memset."""
    if red.numel() < _lib.lib().sh_gn_relu_bwd_scratch_words(N, G):
        raise RuntimeError('gn_relu_bwd: scratch too small (use ops.gn_relu_bwd_scratch)')
    _call('sh_gn_relu_bwd_prezeroed' if prezeroed else 'sh_gn_relu_bwd', _chk(da, BF16, 'da'), _chk(x, BF16, 'x'), _chk(stats_in, name='stats'), _chk(gamma, name='gamma'),
          _chk(beta, name='beta'), _opt(addend, BF16, 'addend'), N, HW, C, G, eps, _chk(red, name='red'),
          _chk(dgamma, name='dgamma'), _chk(dbeta, name='dbeta'), _chk(dx, BF16, 'dx'), _opt(colsum, name='colsum'), _stream())


def maxpool_fwd(x, N, H, W, C, y, stats_out=None, G_out=16):
    _call('sh_maxpool_fwd', _chk(x, BF16, 'x'), N, H, W, C, _chk(y, BF16, 'y'), _opt(stats_out, name='stats_out'), G_out, _stream())


def maxpool_bwd(dy, x, N, H, W, C, dx, addend=None, colsum=None):
    _call('sh_maxpool_bwd', _chk(dy, BF16, 'dy'), _chk(x, BF16, 'x'), _opt(addend, BF16, 'addend'), N, H, W, C,
          _chk(dx, BF16, 'dx'), _opt(colsum, name='colsum'), _stream())


def upsample_add_fwd(up1, low, N, h, w, C, y, stats_out=None, G_out=16):
    _call('sh_upsample_add_fwd', _chk(up1, BF16, 'up1'), _chk(low, BF16, 'low'), N, h, w, C, _chk(y, BF16, 'y'),
          _opt(stats_out, name='stats_out'), G_out, _stream())


def upsample_bwd(dy, N, h, w, C, dlow, colsum=None):
    _call('sh_upsample_bwd', _chk(dy, BF16, 'dy'), N, h, w, C, _chk(dlow, BF16, 'dlow'), _opt(colsum, name='colsum'), _stream())


def add(a, b, N, HW, C, y, c=None, stats_out=None, G_out=16, colsum=None):
    _call('sh_add', _chk(a, BF16, 'a'), _chk(b, BF16, 'b'), _opt(c, BF16, 'c'), N, HW, C, _chk(y, BF16, 'y'),
          _opt(stats_out, name='stats_out'), G_out, _opt(colsum, name='colsum'), _stream())


def colsum(x, N, HW, C, out):
    _call('sh_colsum', _chk(x, BF16, 'x'), N, HW, C, _chk(out, name='colsum'), _stream())


def stem_conv_fwd(img, w, b, N, S, y, stats_out=None, G_out=4):
    _call('sh_stem_conv_fwd', _chk(img, name='img'), _chk(w, name='w'), _chk(b, name='b'), N, S, _chk(y, BF16, 'y'),
          _opt(stats_out, name='stats_out'), G_out, _stream())


def stem_conv_wgrad(img, dy, N, S, dw, db):
    _call('sh_stem_conv_wgrad', _chk(img, name='img'), _chk(dy, BF16, 'dy'), N, S, _chk(dw, name='dw'), _chk(db, name='db'), _stream())


def nchw_to_nhwc(x, N, C, HW, Cp, y):
    _call('sh_nchw_to_nhwc', _chk(x, name='x'), N, C, HW, Cp, _chk(y, BF16, 'y'), _stream())


def nhwc_to_nchw(x, N, C, HW, y):
    _call('sh_nhwc_to_nchw', _chk(x, BF16, 'x'), N, C, HW, _chk(y, name='y'), _stream())


def pack_weights(w, Cout, Cin, taps, cout_pad, cin_pad, wf, wb=None, b_rows=0, b_cols=0):
    _call('sh_pack_weights', _chk(w, name='w'), Cout, Cin, taps, cout_pad, cin_pad, b_rows, b_cols, _chk(wf, BF16, 'wf'),
          _opt(wb, BF16, 'wb'), _stream())


def pack_weights_batch(flat, table, arena):
    _call('sh_pack_weights_batch', _chk(flat, name='flat'), _chk(table, torch.int32, 'table'), table.shape[0],
          _chk(arena, BF16, 'arena'), _stream())


def unpack_wgrad(dw, Cout, Cin, taps, cout_ld, cin_ld, grad):
    _call('sh_unpack_wgrad', _chk(dw, name='dw'), Cout, Cin, taps, cout_ld, cin_ld, _chk(grad, name='grad'), _stream())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    _call('sh_adam_step', _chk(p, name='p'), _chk(g, name='g'), _chk(m, name='m'), _chk(v, name='v'), p.numel(), lr, beta1,
          beta2, eps, weight_decay, step, grad_scale, _stream())


def adam_step_dev(p, g, m, v, lr_dev, step_dev, beta1, beta2, eps, weight_decay, grad_scale=1.0):
    _call('sh_adam_step_dev', _chk(p, name='p'), _chk(g, name='g'), _chk(m, name='m'), _chk(v, name='v'), p.numel(),
          _chk(lr_dev, name='lr'), _chk(step_dev, torch.int32, 'step'), beta1, beta2, eps, weight_decay, grad_scale, _stream())


# ------------------------------------------------------------------------------------------------ train-step glue
import ctypes as _ct


def step_combine(xyz, Ns, M, J, hw, weights8, gxyz, terms9, g_mvproj=None, g_pose3=None, g_prior=None, target_xyz4=None,
                 loss_mv3=None, loss_pose3=None, loss_prior3=None, sse2=None, mean_scale=1.0):
    w = (_ct.c_float * 8)(*[float(x) for x in weights8])
    _call('sh_step_combine', _opt(g_mvproj, name='g_mvproj'), _opt(g_pose3, name='g_pose3'), _opt(g_prior, name='g_prior'),
          _chk(xyz, name='xyz'), _opt(target_xyz4, name='target_xyz4'), _opt(loss_mv3, name='loss_mv3'),
          _opt(loss_pose3, name='loss_pose3'), _opt(loss_prior3, name='loss_prior3'), _opt(sse2, torch.float64, 'sse2'),
          Ns, M, J, hw, _ct.cast(w, _ct.c_void_p), float(mean_scale), _chk(gxyz, name='gxyz'), _chk(terms9, name='terms9'),
          _stream())


def data_to_model_fwdbwd(dms, joints, radii):
    N, H, W = dms.shape
    J = joints.shape[1]
    dev = dms.device
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    grad = torch.empty((N, J, 3), device=dev, dtype=torch.float32)
    scratch = torch.empty(_lib.lib().sh_data_to_model_scratch_bytes(N, J) // 4 + 4, device=dev, dtype=torch.float32)
    _call('sh_data_to_model_fwdbwd', _chk(dms, name='dms'), _chk(joints, name='joints'), _chk(radii, name='radii'), N, J, H, W,
          loss.data_ptr(), grad.data_ptr(), scratch.data_ptr(), _stream())
    return loss, grad


def ortho_project(points, mode, cam, rand_f=None):
    B, Nv = points.shape[:2]
    out = torch.empty_like(points)
    _call('sh_ortho_project', _chk(points, name='points'), B, Nv, mode, float(cam[0]), float(cam[1]), float(cam[2]), float(cam[3]),
          _opt(rand_f, name='rand_f'), out.data_ptr(), _stream())
    return out


def rand_scale_apply(mats, scales):
    B, nmat = mats.shape[:2]
    out = torch.empty_like(mats)
    _call('sh_rand_scale_apply', _chk(mats, name='mats'), _chk(scales, name='scales'), B, nmat, out.data_ptr(), _stream())
    return out


def clamp_max(x, max_value):
    out = torch.empty_like(x)
    _call('sh_clamp_max', _chk(x, name='x'), x.numel(), float(max_value), out.data_ptr(), _stream())
    return out


def unscale_xy(xyz, u_scales, v_scales):
    """In place on fp32 [M,J,3]: xyz[..., 0] /= u[m], xyz[..., 1] /= v[m]."""
    M, J = xyz.shape[0], xyz.shape[1]
    _call('sh_unscale_xy', _chk(xyz, name='xyz'), _chk(u_scales, name='u_scales'), _chk(v_scales, name='v_scales'), M, J, _stream())
    return xyz


def resize_crop(depth_maps, u_scales, v_scales, out=None):
    N, H, W = depth_maps.shape
    if out is None:
        out = torch.empty_like(depth_maps)
    _chk(out, name='out')
    _call('sh_resize_crop', _chk(depth_maps, name='depth_maps'), _chk(u_scales, name='u_scales'), _chk(v_scales, name='v_scales'),
          N, H, W, out.data_ptr(), _stream())
    return out


def pose_denoiser_fwd(fea, in_idx, out_idx, blob, scale):
    """fea [M,n_fea] fp32, in_idx / out_idx int32 -> out [M,n_fea] (sh_pose_denoiser_fwd)."""
    M, n_fea = fea.shape
    n_in, n_out = in_idx.numel(), out_idx.numel()
    if blob.numel() != _lib.lib().sh_pose_denoiser_blob_floats(n_in, n_out):
        raise RuntimeError('pose_denoiser_fwd: weight blob has the wrong size')
    out = torch.empty_like(fea)
    _call('sh_pose_denoiser_fwd', _chk(fea, name='fea'), _chk(in_idx, torch.int32, 'in_idx'), _chk(out_idx, torch.int32, 'out_idx'),
          _chk(blob, name='blob'), M, n_fea, n_in, n_out, float(scale), out.data_ptr(), _stream())
    return out


def sample_poses(u, offsets):
    """u: fp32 uniforms (any shape, contiguous), offsets int32 [n] -> poses [n,26] (sh_sample_poses)."""
    n = offsets.numel()
    out = torch.empty((n, 26), device=u.device, dtype=torch.float32)
    _call('sh_sample_poses', _chk(u, name='u'), _chk(offsets, torch.int32, 'offsets'), n, out.data_ptr(), _stream())
    return out


def scale(x, s, out):
    _call('sh_scale', _chk(x, name='x'), float(s), x.numel(), _chk(out, name='out'), _stream())
    return out
