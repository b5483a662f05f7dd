"""ncu target: ONE launch (after one warm-up launch) of every kernel north_star names, at the big shapes of the BASELINE step, in a
fixed order (the warm-up launch is the even, the profiled one the odd instance of each kernel):

    ncu --set full --clock-control none --import-source on \
        -k regex:"conv_fwd_kernel|wgrad1x1_kernel|wgrad3x3_kernel|gn_relu_(fwd|bwd)_kernel|sphere_render_(fwd|bwd)_kernel|tri_raster_kernel|mvproj_main_kernel|stem_conv_fwd_kernel" \
        -o gpurun_out/r2_full_shapes python tools/ncu_shapes.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spherehand_b200 import data, ops                              # noqa: E402
from spherehand_b200.model import HandModel                        # noqa: E402

DEV, BF16, N = 'cuda', torch.bfloat16, 256


def conv(H, Cin, Cout, taps, res):
    cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
    x = torch.randn(N, H, H, Cin, device=DEV).to(BF16)
    r = torch.randn(N, H, H, Cout, device=DEV).to(BF16) if res else None
    y = torch.empty(N, H, H, Cout, device=DEV, dtype=BF16)
    k = 3 if taps == 9 else 1
    w = torch.randn(Cout, Cin, k, k, device=DEV) * 0.05
    wf = torch.empty((taps, cout_pad, Cin), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, taps, cout_pad, Cin, wf)
    st = torch.zeros(N, 16, 2, device=DEV)
    for _ in range(2):
        ops.conv_fwd(x, wf, torch.zeros(Cout, device=DEV), N, H, H, Cin, Cout, cout_pad, taps, y=y, y_ld=Cout, residual=r, stats=st, groups=16)


def conv_gn(H, Cin, Cout, res):
    """the 1x1 layer with the GroupNorm + ReLU in front of it folded in (sh_conv_fwd_gn) and its weight gradient (sh_conv_wgrad_gn)"""
    cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
    x = torch.randn(N, H, H, Cin, device=DEV).to(BF16)
    v = x.float().reshape(N, H * H, 16, Cin // 16)
    gn_in = (torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous(), torch.rand(Cin, device=DEV) + 0.5,
             torch.randn(Cin, device=DEV) * 0.1, 16, 1e-5)
    r = torch.randn(N, H, H, Cout, device=DEV).to(BF16) if res else None
    y = torch.empty(N, H, H, Cout, device=DEV, dtype=BF16)
    dy = torch.randn(N, H, H, Cout, device=DEV).to(BF16)
    w = torch.randn(Cout, Cin, 1, 1, device=DEV) * 0.05
    wf = torch.empty((1, cout_pad, Cin), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, 1, cout_pad, Cin, wf)
    st = torch.zeros(N, 16, 2, device=DEV)
    dw = torch.zeros_like(w)
    for _ in range(2):
        ops.conv_fwd(x, wf, torch.zeros(Cout, device=DEV), N, H, H, Cin, Cout, cout_pad, 1, y=y, y_ld=Cout, residual=r, stats=st, groups=16, gn=gn_in)
    for _ in range(2):
        ops.conv_wgrad(dy, x, N, H, H, Cin, Cin, Cout, Cout, 1, dw, gn=gn_in)


def stem():
    img = torch.rand(N, 128, 128, device=DEV)
    w = torch.randn(64, 1, 5, 5, device=DEV) * 0.1
    y = torch.empty(N, 64, 64, 64, device=DEV, dtype=BF16)
    st = torch.zeros(N, 4, 2, device=DEV)
    for _ in range(2):
        ops.stem_conv_fwd(img, w, torch.zeros(64, device=DEV), N, 128, y, st, 4)


def wgrad(H, Cin, Cout, taps):
    x = torch.randn(N, H, H, Cin, device=DEV).to(BF16)
    dy = torch.randn(N, H, H, Cout, device=DEV).to(BF16)
    k = 3 if taps == 9 else 1
    dw = torch.zeros(Cout, Cin, k, k, device=DEV)
    scratch = torch.zeros(9 * Cout * Cin, device=DEV)
    for _ in range(2):
        if taps == 9:
            ops.conv_wgrad3x3(dy, x, N, H, H, Cin, Cin, Cout, Cout, scratch)
        else:
            ops.conv_wgrad(dy, x, N, H, H, Cin, Cin, Cout, Cout, taps, dw)


def gn(H, C, addend):
    x = torch.randn(N, H, H, C, device=DEV).to(BF16)
    da = torch.randn(N, H, H, C, device=DEV).to(BF16)
    ad = torch.randn(N, H, H, C, device=DEV).to(BF16) if addend else None
    y = torch.empty_like(x)
    v = x.float().reshape(N, H * H, 16, C // 16)
    st = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1
    red = ops.gn_relu_bwd_scratch(N, 16, DEV)
    dg, db, cs = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    for _ in range(2):
        ops.gn_relu_fwd(x, st, gamma, beta, N, H * H, C, 16, y)
    for _ in range(2):
        ops.gn_relu_bwd(da, x, st, gamma, beta, N, H * H, C, 16, red, dg, db, y, ad, cs)


def renderers():
    hm = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'hand_model.npz')))
    hand = HandModel.from_arrays(hm, DEV)
    g = torch.Generator().manual_seed(1234)
    Nn, J, S = 256, 48, 128
    c = torch.cat([torch.rand(Nn, J, 2, generator=g) * 180 - 90, torch.rand(Nn, J, 1, generator=g) * 120 - 60], -1).to(DEV)
    r = torch.cat([hand.radii.cpu(), torch.full((J - hand.radii.numel(),), 20.0)]).to(DEV)
    sph = ops.pack_spheres(c, r)
    for _ in range(2):
        depth, idx = ops.sphere_render_fwd(sph, S, S)
    gd = torch.randn_like(depth) * (idx != 255)
    for _ in range(2):
        ops.sphere_render_bwd(gd, idx, sph)
    B = 64
    poses = data.random_poses(B, g, DEV)
    mats = ops.fk_fwd(poses, hand.offset_mats, hand.inv_offset_mats)
    pts = ops.lbs_fwd(mats, *hand.mesh_csr, right_hand=True, mode=2, cam=(320.0, 320.0, 640 / 300, 640 / 300))
    fv = ops.gather_faces(pts, hand.faces)
    for _ in range(2):
        ops.tri_raster_fwd(fv, 640, 640)
    real, cams, inv = data.synthetic_real_batch(hand, B, 3, S, g)
    joints = torch.randn(B, 3, 41, 3, device=DEV) * 40
    for _ in range(2):
        ops.mvproj_loss_fwdbwd(cams, inv, joints, real, hand.radii, True)


if __name__ == '__main__':
    conv(32, 256, 128, 1, False)
    conv(32, 128, 256, 1, True)
    conv(32, 128, 128, 9, False)
    wgrad(32, 256, 128, 1)
    wgrad(32, 128, 128, 9)
    gn(32, 256, True)
    gn(32, 128, False)
    renderers()
    conv_gn(32, 256, 128, False)
    conv_gn(32, 128, 256, True)
    stem()
    torch.cuda.synchronize()
    print('done')
