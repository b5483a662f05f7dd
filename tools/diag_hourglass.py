"""Diagnostic (GPU box): is the hourglass gradient error bf16 rounding noise or a bug?
Compares  ours  vs  torch fp32  vs  torch fp32 with bf16 rounding emulated at the same materialisation points."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import hourglass as oh
from spherehand_b200.network.hourglass import create_hourglass_network
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = 'cuda'


class Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x): return x.bfloat16().float()
    @staticmethod
    def backward(ctx, g): return g.bfloat16().float()


def emul_forward(x, sd, stacks):
    """oracle.hourglass with rounding points; weights rounded to bf16 (convs), GN params fp32."""
    q = Q.apply
    def gn(t, p, n, g=16): return q(F.relu(F.group_norm(t, g, sd[p + n + '.weight'], sd[p + n + '.bias'])))
    def conv(t, name, pad=0, res=None, quant=True):
        y = F.conv2d(t, q(sd[name + '.weight']), sd[name + '.bias'], padding=pad)
        if res is not None: y = y + res
        return q(y) if quant else y
    def bott(x, p):
        o = conv(gn(x, p, 'bn1'), p + 'conv1')
        o = conv(gn(o, p, 'bn2'), p + 'conv2', 1)
        a3 = gn(o, p, 'bn3')
        res = conv(x, p + 'downsample.0') if (p + 'downsample.0.weight') in sd else x
        return conv(a3, p + 'conv3', res=res)
    def hg(n, x, p):
        up1 = bott(x, '%shg.%d.0.0.' % (p, n - 1))
        low1 = bott(F.max_pool2d(x, 2, 2), '%shg.%d.1.0.' % (p, n - 1))
        if n > 1: low2, lat = hg(n - 1, low1, p)
        else:
            low2 = bott(low1, '%shg.%d.3.0.' % (p, n - 1)); lat = low2
        low3 = bott(low2, '%shg.%d.2.0.' % (p, n - 1))
        return q(up1 + F.interpolate(low3, scale_factor=2, mode='bilinear', align_corners=False)), lat
    x = x[:, None]
    x = q(F.conv2d(x, sd['conv1.weight'], sd['conv1.bias'], stride=2, padding=2))
    x = q(F.relu(F.group_norm(x, 4, sd['bn1.weight'], sd['bn1.bias'])))
    x = bott(x, 'layer1.0.'); x = F.max_pool2d(x, 2, 2); x = bott(x, 'layer2.0.'); x = bott(x, 'layer3.0.')
    outs = []
    for i in range(stacks):
        y, lat = hg(2, x, 'hg.%d.' % i)
        y = bott(y, 'res.%d.0.' % i)
        y = conv(y, 'fc.%d.0' % i)
        y = q(F.relu(F.group_norm(y, 16, sd['fc.%d.1.weight' % i], sd['fc.%d.1.bias' % i])))
        score = conv(y, 'score.%d' % i, quant=False)
        outs.append(score)
        if i < stacks - 1:
            t = conv(y, 'fc_.%d' % i, res=x)
            x = conv(q(score), 'score_.%d' % i, res=t)
    return outs


def nrm(a, b): return float((a - b).double().norm() / b.double().norm().clamp_min(1e-30))
def mx(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(stacks, N, S, upstream):
    sd0 = {k: v.to(DEV) for k, v in oh.det_state_dict(82, stacks, seed=7).items()}
    torch.manual_seed(0)
    x = (torch.randn(N, S, S, device=DEV) * 0.3).clamp(-1, 1)
    tgt = torch.rand(N, 82, S // 4, S // 4, device=DEV) * 0.1
    def loss_of(outs):
        if upstream == 'noise':
            return sum((o * torch.from_numpy(oh.det_uniform(o.numel(), 100 + i).reshape(o.shape)).to(DEV)).sum() for i, o in enumerate(outs))
        return sum(((o - tgt) ** 2).mean() * 1e3 for o in outs)
    res = {}
    for mode in ('fp32', 'emul'):
        sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
        outs = oh.hourglass_forward(x, sd, stacks)[0] if mode == 'fp32' else emul_forward(x, sd, stacks)
        loss_of(outs).backward()
        res[mode] = ({k: v.grad for k, v in sd.items()}, [o.detach() for o in outs])
    net = create_hourglass_network(82, stacks).to(DEV)
    net.load_state_dict(sd0)
    outs, _ = net(x)
    loss_of(outs).backward()
    res['ours'] = ({k: p.grad for k, p in net.named_parameters()}, [o.detach() for o in outs])
    print('== stacks=%d N=%d S=%d upstream=%s' % (stacks, N, S, upstream))
    for i in range(stacks):
        print(' score%d: ours-fp32 %.4f  emul-fp32 %.4f  ours-emul %.4f (norm-rel)' % (
            i, nrm(res['ours'][1][i], res['fp32'][1][i]), nrm(res['emul'][1][i], res['fp32'][1][i]), nrm(res['ours'][1][i], res['emul'][1][i])))
    rows = []
    for k in sd0:
        go, gf, ge = res['ours'][0][k], res['fp32'][0][k], res['emul'][0][k]
        rows.append((k, nrm(go, gf), nrm(ge, gf), nrm(go, ge), mx(go, gf), mx(ge, gf)))
    rows.sort(key=lambda r: -r[1])
    print(' %-34s %9s %9s %9s %9s %9s' % ('param', 'o-f nrm', 'e-f nrm', 'o-e nrm', 'o-f max', 'e-f max'))
    for r in rows[:12] + rows[-4:]:
        print(' %-34s %9.4f %9.4f %9.4f %9.4f %9.4f' % r)
    a = np.array([r[1] for r in rows]); b = np.array([r[2] for r in rows])
    print(' median o-f %.4f  e-f %.4f ; max o-f %.4f e-f %.4f' % (np.median(a), np.median(b), a.max(), b.max()))


if __name__ == '__main__':
    run(1, 4, 64, 'noise')
    run(2, 4, 128, 'noise')
    run(2, 4, 128, 'mse')
