"""One GroupNorm-backward shape, a handful of launches (ncu target).  Usage: python tools/bench_gn_one.py H C"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spherehand_b200 import ops
H, C, G, N = int(sys.argv[1]), int(sys.argv[2]), 16, 256
BF16 = torch.bfloat16
x = torch.randn(N, H, H, C, device='cuda').to(BF16)
da = torch.randn(N, H, H, C, device='cuda').to(BF16)
ad = torch.randn(N, H, H, C, device='cuda').to(BF16)
y = torch.empty_like(x)
v = x.float().reshape(N, H * H, G, C // G)
st = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()
gamma, beta = torch.rand(C, device='cuda') + 0.5, torch.randn(C, device='cuda') * 0.1
red = ops.gn_relu_bwd_scratch(N, G, 'cuda')
dg, db, cs = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
for _ in range(4):
    ops.gn_relu_bwd(da, x, st, gamma, beta, N, H * H, C, G, red, dg, db, y, ad, cs)
torch.cuda.synchronize()
