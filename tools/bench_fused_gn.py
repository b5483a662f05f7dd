"""Device time of the fused-GroupNorm 1x1 convolution / weight gradient against the two-pass path (gn_relu_fwd + conv / wgrad) at the
bottleneck shapes of the BASELINE step; CUDA events around 20 back-to-back calls rotating over 3 buffer sets (> L2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spherehand_b200 import ops                              # noqa: E402

DEV, BF16, N = 'cuda', torch.bfloat16, 256


def timed(fn, calls=21):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(calls):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / calls


def case(H, Cin, Cout, res):
    cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
    sets = []
    for _ in range(3):
        x = torch.randn(N, H, H, Cin, device=DEV).to(BF16)
        v = x.float().reshape(N, H * H, 16, Cin // 16)
        st_in = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()
        sets.append(dict(x=x, st=st_in, a=torch.empty_like(x), r=torch.randn(N, H, H, Cout, device=DEV).to(BF16) if res else None,
                         y=torch.empty(N, H, H, Cout, device=DEV, dtype=BF16), dy=torch.randn(N, H, H, Cout, device=DEV).to(BF16)))
    gamma, beta = torch.rand(Cin, device=DEV) + 0.5, torch.randn(Cin, device=DEV) * 0.1
    w = torch.randn(Cout, Cin, 1, 1, device=DEV) * 0.05
    wf = torch.empty((1, cout_pad, Cin), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, 1, cout_pad, Cin, wf)
    bias = torch.zeros(Cout, device=DEV)
    st = torch.zeros(N, 16, 2, device=DEV)
    dw = torch.zeros_like(w)

    def gn(i):
        s = sets[i % 3]
        ops.gn_relu_fwd(s['x'], s['st'], gamma, beta, N, H * H, Cin, 16, s['a'])

    def conv(i, fused):
        s = sets[i % 3]
        ops.conv_fwd(s['x'] if fused else s['a'], wf, bias, N, H, H, Cin, Cout, cout_pad, 1, y=s['y'], y_ld=Cout, residual=s['r'], stats=st,
                     groups=16, gn=(s['st'], gamma, beta, 16, 1e-5) if fused else None)

    def wg(i, fused):
        s = sets[i % 3]
        ops.conv_wgrad(s['dy'], s['x'] if fused else s['a'], N, H, H, Cin, Cin, Cout, Cout, 1, dw, gn=(s['st'], gamma, beta, 16, 1e-5) if fused else None)

    t_gn = timed(gn)
    print('%3dx%-3d %3d->%-3d res=%d | gn_relu_fwd %6.1f | conv %6.1f  conv_gn %6.1f (two-pass %6.1f) | wgrad %6.1f  wgrad_gn %6.1f' % (
        H, H, Cin, Cout, res, t_gn, timed(lambda i: conv(i, False)), timed(lambda i: conv(i, True)),
        t_gn + timed(lambda i: conv(i, False)), timed(lambda i: wg(i, False)), timed(lambda i: wg(i, True))), flush=True)


for shape in [(32, 256, 128, False), (32, 128, 256, True), (16, 256, 128, False), (16, 128, 256, True), (64, 64, 64, False), (64, 64, 128, True),
              (8, 256, 128, False), (8, 128, 256, True)]:
    case(*shape)
