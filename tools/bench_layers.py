"""Per-shape micro-benchmark of the hourglass layer kernels (GPU box).  Reports device time per launch (CUDA events,
back-to-back launches over a ring of buffers larger than L2), algorithmic TFLOP/s and GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spherehand_b200 import ops
DEV = 'cuda'
BF16 = torch.bfloat16
N = int(os.environ.get('N', 256))
RING = 3


def timeit(fn, reps=12):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def conv_case(H, Cin, Cout, taps, res, stats=os.environ.get('STATS', '1') == '1'):
    cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
    xs = [torch.randn(N, H, H, Cin, device=DEV).to(BF16) for _ in range(RING)]
    rs = [torch.randn(N, H, H, Cout, device=DEV).to(BF16) for _ in range(RING)] if res else None
    ys = [torch.empty(N, H, H, Cout, device=DEV, dtype=BF16) for _ in range(RING)]
    w = torch.randn(Cout, Cin, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device=DEV) * 0.05
    wf = torch.empty((taps, cout_pad, Cin), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, taps, cout_pad, Cin, wf)
    b = torch.randn(Cout, device=DEV)
    st = torch.zeros(N, 16, 2, device=DEV)

    def fn(i):
        k = i % RING
        ops.conv_fwd(xs[k], wf, b, N, H, H, Cin, Cout, cout_pad, taps, y=ys[k], y_ld=Cout, residual=rs[k] if res else None,
                     stats=st if stats else None, groups=16)
    us = timeit(fn)
    flops = 2.0 * N * H * H * Cin * Cout * taps
    byts = N * H * H * 2 * (Cin + Cout * (2 if res else 1))
    print('conv %dx%d %3d->%3d @%3d res=%d : %8.1f us  %7.1f TF/s  %7.0f GB/s' % (3 if taps == 9 else 1, 3 if taps == 9 else 1, Cin, Cout, H, res, us, flops / us * 1e-6, byts / us * 1e-3), flush=True)


def wgrad_case(H, Cin, Cout, taps):
    xs = [torch.randn(N, H, H, Cin, device=DEV).to(BF16) for _ in range(RING)]
    dys = [torch.randn(N, H, H, Cout, device=DEV).to(BF16) for _ in range(RING)]
    k = 3 if taps == 9 else 1
    dw = torch.zeros(Cout, Cin, k, k, device=DEV)

    scratch = torch.zeros(9 * Cout * Cin, device=DEV)

    def fn(i):
        j = i % RING
        if taps == 9 and H >= 16:
            ops.conv_wgrad3x3(dys[j], xs[j], N, H, H, Cin, Cin, Cout, Cout, scratch)
        else:
            ops.conv_wgrad(dys[j], xs[j], N, H, H, Cin, Cin, Cout, Cout, taps, dw)
    us = timeit(fn)
    flops = 2.0 * N * H * H * Cin * Cout * taps
    byts = N * H * H * 2 * (Cin + Cout)
    print('wgrad %dx%d %3d->%3d @%3d      : %8.1f us  %7.1f TF/s  %7.0f GB/s' % (k, k, Cin, Cout, H, us, flops / us * 1e-6, byts / us * 1e-3), flush=True)


def gn_case(H, C, G=16):
    xs = [torch.randn(N, H, H, C, device=DEV).to(BF16) for _ in range(RING)]
    das = [torch.randn(N, H, H, C, device=DEV).to(BF16) for _ in range(RING)]
    ads = [torch.randn(N, H, H, C, device=DEV).to(BF16) for _ in range(RING)]
    ys = [torch.empty(N, H, H, C, device=DEV, dtype=BF16) for _ in range(RING)]
    v = xs[0].float().reshape(N, H * H, G, C // G)
    st = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1
    us = timeit(lambda i: ops.gn_relu_fwd(xs[i % RING], st, gamma, beta, N, H * H, C, G, ys[i % RING]))
    byts = N * H * H * C * 2 * 2
    print('gn_relu_fwd C=%3d @%3d         : %8.1f us  %7.0f GB/s' % (C, H, us, byts / us * 1e-3), flush=True)
    red = ops.gn_relu_bwd_scratch(N, G, DEV)
    dg, db, cs = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    us = timeit(lambda i: ops.gn_relu_bwd(das[i % RING], xs[i % RING], st, gamma, beta, N, H * H, C, G, red, dg, db, ys[i % RING], ads[i % RING], cs))
    byts = N * H * H * C * 2 * 4       # minimum traffic of a fused single pass: da, x, addend in, dx out
    print('gn_relu_bwd C=%3d @%3d (+add)  : %8.1f us  %7.0f GB/s (vs 4-tensor minimum)' % (C, H, us, byts / us * 1e-3), flush=True)


def raster_cases():
    """The renderers at BASELINE.json's configs: R2 fwd+bwd (config 2: N=256, J=48, 128^2; in-situ N=576, J=41), R1 at the
    pybind boundary (B=64 meshes of 3382 faces -> 640^2) and on the resize lattice, against the HBM roofline."""
    import json
    import numpy as np
    from spherehand_b200.model import HandModel
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = json.load(open(os.path.join(root, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(root, 'MEASURED_PEAKS.json')) else 6650.0
    g = torch.Generator(device='cpu').manual_seed(1234)
    for Nn, J in ((256, 48), (576, 41), (4096, 41)):
        S = 128
        c = torch.cat([torch.rand(Nn, J, 2, generator=g) * 180 - 90, torch.rand(Nn, J, 1, generator=g) * 120 - 60], -1).to(DEV)
        r = (torch.rand(J, generator=g) * 16 + 8).to(DEV)
        sph = ops.pack_spheres(c, r)
        depth, idx = ops.sphere_render_fwd(sph, S, S)
        gd = torch.randn_like(depth) * (idx != 255)
        fwd_b = Nn * (16 * J + 5 * S * S)
        bwd_b = Nn * (5 * S * S + 28 * J)
        us_f = timeit(lambda i: ops.sphere_render_fwd(sph, S, S), reps=30)
        us_b = timeit(lambda i: ops.sphere_render_bwd(gd, idx, sph), reps=30)
        print('R2 sphere render N=%4d J=%d 128^2 : fwd %6.1f us %6.0f GB/s (%4.1f%% of measured HBM)   bwd %6.1f us %6.0f GB/s (%4.1f%%)   fwd+bwd %5.1f%%'
              % (Nn, J, us_f, fwd_b / us_f * 1e-3, 100 * fwd_b / us_f * 1e-3 / peak, us_b, bwd_b / us_b * 1e-3,
                 100 * bwd_b / us_b * 1e-3 / peak, 100 * (fwd_b + bwd_b) / (us_f + us_b) * 1e-3 / peak), flush=True)
    hm = dict(np.load(os.path.join(root, 'tests', 'golden', 'hand_model.npz')))
    hand = HandModel.from_arrays(hm, DEV)
    B = 64
    from spherehand_b200 import data
    poses = data.random_poses(B, g, DEV)
    mats = ops.fk_fwd(poses, hand.offset_mats, hand.inv_offset_mats)
    pts = ops.lbs_fwd(mats, *hand.mesh_csr, right_hand=True, mode=2, cam=(320.0, 320.0, 640 / 300, 640 / 300))
    fv = ops.gather_faces(pts, hand.faces)
    F = fv.shape[1]
    us = timeit(lambda i: ops.tri_raster_fwd(fv, 640, 640), reps=10)
    byts = B * (36 * F + 4 * 640 * 640)
    print('R1 triangle raster B=%d F=%d -> 640^2 (pybind boundary): %7.1f us %6.0f GB/s (%4.1f%% of measured HBM)' % (B, F, us, byts / us * 1e-3, 100 * byts / us * 1e-3 / peak), flush=True)
    us = timeit(lambda i: ops.tri_raster_lattice_fwd(fv, 640, 5, 2, 1), reps=10)
    print('R1 on the 640->128 resize lattice (what the train step runs)  : %7.1f us  (%.1f us per mesh)' % (us, us / B), flush=True)


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'conv'):
        conv_case(32, 128, 128, 9, False)
        conv_case(32, 256, 128, 1, False)
        conv_case(32, 128, 256, 1, True)
        conv_case(32, 256, 256, 1, False)
        conv_case(16, 128, 128, 9, False)
        conv_case(16, 128, 256, 1, True)
        conv_case(8, 128, 128, 9, False)
        conv_case(64, 64, 64, 9, False)
        conv_case(64, 64, 128, 1, True)
        conv_case(64, 64, 64, 1, False)
    if which in ('all', 'wgrad'):
        wgrad_case(32, 128, 128, 9)
        wgrad_case(32, 256, 128, 1)
        wgrad_case(32, 128, 256, 1)
        wgrad_case(16, 128, 128, 9)
        wgrad_case(64, 64, 64, 9)
        wgrad_case(64, 64, 128, 1)
    if which in ('all', 'raster'):
        raster_cases()
    if which in ('all', 'gn'):
        gn_case(32, 256)
        gn_case(32, 128)
        gn_case(64, 64)
        gn_case(64, 128)
        gn_case(16, 256)
