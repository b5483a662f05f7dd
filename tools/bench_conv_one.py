"""One convolution shape, a handful of launches (ncu target).  Usage: python tools/bench_conv_one.py H Cin Cout taps [res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spherehand_b200 import ops
H, Cin, Cout, taps = (int(v) for v in sys.argv[1:5])
res = len(sys.argv) > 5 and sys.argv[5] == '1'
N, BF16 = 256, torch.bfloat16
cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
x = torch.randn(N, H, H, Cin, device='cuda').to(BF16)
r = torch.randn(N, H, H, Cout, device='cuda').to(BF16) if res else None
y = torch.empty(N, H, H, Cout, device='cuda', dtype=BF16)
k = 3 if taps == 9 else 1
w = torch.randn(Cout, Cin, k, k, device='cuda') * 0.05
wf = torch.empty((taps, cout_pad, Cin), device='cuda', dtype=BF16)
ops.pack_weights(w, Cout, Cin, taps, cout_pad, Cin, wf)
b = torch.randn(Cout, device='cuda')
st = torch.zeros(N, 16, 2, device='cuda')
for _ in range(4):
    ops.conv_fwd(x, wf, b, N, H, H, Cin, Cout, cout_pad, taps, y=y, y_ld=Cout, residual=r, stats=st, groups=16)
torch.cuda.synchronize()
