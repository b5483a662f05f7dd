"""ncu target: the fused-GroupNorm 1x1 convolution / weight gradient at the big 32x32 shapes (warm-up launch + profiled launch each).

    ncu --set full --clock-control none --import-source on -k regex:"conv_fwd_kernel|wgrad1x1_kernel" -o gpurun_out/fused_gn python tools/ncu_fused_gn.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spherehand_b200 import ops                              # noqa: E402

DEV, BF16, N = 'cuda', torch.bfloat16, 256


def run(H, Cin, Cout, res, fused):
    cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
    x = torch.randn(N, H, H, Cin, device=DEV).to(BF16)
    v = x.float().reshape(N, H * H, 16, Cin // 16)
    st_in = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()
    gamma, beta = torch.rand(Cin, device=DEV) + 0.5, torch.randn(Cin, device=DEV) * 0.1
    gn = (st_in, gamma, beta, 16, 1e-5) if fused else None
    r = torch.randn(N, H, H, Cout, device=DEV).to(BF16) if res else None
    y = torch.empty(N, H, H, Cout, device=DEV, dtype=BF16)
    dy = torch.randn(N, H, H, Cout, device=DEV).to(BF16)
    w = torch.randn(Cout, Cin, 1, 1, device=DEV) * 0.05
    wf = torch.empty((1, cout_pad, Cin), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, 1, cout_pad, Cin, wf)
    st = torch.zeros(N, 16, 2, device=DEV)
    dw = torch.zeros_like(w)
    for _ in range(2):
        ops.conv_fwd(x, wf, torch.zeros(Cout, device=DEV), N, H, H, Cin, Cout, cout_pad, 1, y=y, y_ld=Cout, residual=r, stats=st, groups=16, gn=gn)
    for _ in range(2):
        ops.conv_wgrad(dy, x, N, H, H, Cin, Cin, Cout, Cout, 1, dw, gn=gn)
    torch.cuda.synchronize()


for fused in (True, False):
    run(32, 256, 128, False, fused)
    run(32, 128, 256, True, fused)
