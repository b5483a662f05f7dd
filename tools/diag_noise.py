"""Run-to-run reproducibility of the hourglass kernels (GPU box): the same forward + backward twice on the same inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spherehand_b200.network.hourglass import create_hourglass_network
from oracle.hourglass import det_state_dict, det_uniform, two_stack_from_trained
DEV = 'cuda'
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

def run(net, x, w):
    net.zero_grad()
    o, lat = net(x)
    (o[0] * w).sum().backward()
    return [t.detach().clone() for t in o], {k: p.grad.clone() for k, p in net.named_parameters()}, net

for name in ('det', 'trained'):
    sd = det_state_dict(82, 1, seed=7) if name == 'det' else {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, 'trained_weights.npz')).items()}
    net = create_hourglass_network(82, 1).to(DEV)
    net.load_state_dict(sd)
    if name == 'det':
        x = torch.from_numpy(det_uniform(2 * 64 * 64, 1).reshape(2, 64, 64)).to(DEV)
    else:
        x = torch.from_numpy(np.load(os.path.join(G, 'trained_hourglass_64.npz'))['x']).to(DEV)
    w = torch.from_numpy(det_uniform(2 * 82 * 16 * 16, 3).reshape(2, 82, 16, 16)).to(DEV)
    o1, g1, _ = run(net, x, w)
    o2, g2, _ = run(net, x, w)
    print(name, 'forward scores bitwise equal:', torch.equal(o1[0], o2[0]), 'max diff', float((o1[0] - o2[0]).abs().max()), 'of', float(o1[0].abs().max()))
    rows = sorted(((float((g1[k] - g2[k]).norm() / g1[k].norm().clamp_min(1e-30)), k) for k in g1), reverse=True)
    print(name, 'run-to-run gradient l2 diff: worst', rows[:5], 'median', rows[len(rows) // 2])
    # per-layer: where does the forward first differ?  tap the tape of two runs
    net.zero_grad()
    with torch.no_grad():
        pass
# GroupNorm statistics from the conv epilogue: reproducible?
from spherehand_b200 import ops
N, H, C = 8, 32, 256
xs = torch.randn(N, H, H, C, device=DEV).to(torch.bfloat16)
wgt = torch.randn(128, C, 1, 1, device=DEV) * 0.05
wf = torch.empty((1, 128, C), device=DEV, dtype=torch.bfloat16)
ops.pack_weights(wgt, 128, C, 1, 128, C, wf)
b = torch.randn(128, device=DEV)
outs = []
for r in range(3):
    st = torch.zeros(N, 16, 2, device=DEV)
    y = torch.empty(N, H, H, 128, device=DEV, dtype=torch.bfloat16)
    ops.conv_fwd(xs, wf, b, N, H, H, C, 128, 128, 1, y=y, y_ld=128, stats=st, groups=16)
    outs.append((y.clone(), st.clone()))
print('conv y bitwise equal:', torch.equal(outs[0][0], outs[1][0]), 'stats bitwise equal:', torch.equal(outs[0][1], outs[1][1]),
      'stats rel diff', float(((outs[0][1] - outs[1][1]).abs() / outs[0][1].abs().clamp_min(1e-20)).max()))
