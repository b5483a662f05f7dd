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
    if which in ('all', 'gn'):
        gn_case(32, 256)
        gn_case(32, 128)
        gn_case(64, 64)
        gn_case(64, 128)
        gn_case(16, 256)
