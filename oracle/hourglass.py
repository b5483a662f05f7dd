"""Oracle for the stacked-hourglass network.  TEST INFRASTRUCTURE (see oracle/__init__).

A functional torch-fp32 restatement (F.conv2d / F.group_norm on CPU; autograd gives the reference
gradients) of
  * Bottleneck.forward           /root/reference/network/hourglass.py:23-41
  * Hourglass._hour_glass_forward /root/reference/network/hourglass.py:68-82
  * HourglassNet.forward         /root/reference/network/hourglass.py:147-173
driven directly by a state_dict with the reference's key names (`conv1.weight`, `layer1.0.bn1.weight`,
`hg.0.hg.1.2.0.conv2.weight`, `res.0.0...`, `fc.0.0/.1`, `score.0`, `fc_.0`, `score_.0`).
"""
import numpy as np
import torch
import torch.nn.functional as F


def param_shapes(num_outputs=82, num_stacks=1):
    """Ordered {name: shape} of the reference module's parameters (hourglass.py:89-120)."""
    shapes = {}

    def conv(name, cout, cin, k):
        shapes[name + '.weight'] = (cout, cin, k, k)
        shapes[name + '.bias'] = (cout,)

    def gn(name, c):
        shapes[name + '.weight'] = (c,)
        shapes[name + '.bias'] = (c,)

    def bottleneck(name, cin, planes, down):
        gn(name + '.bn1', cin)
        conv(name + '.conv1', planes, cin, 1)
        gn(name + '.bn2', planes)
        conv(name + '.conv2', planes, planes, 3)
        gn(name + '.bn3', planes)
        conv(name + '.conv3', planes * 2, planes, 1)
        if down:
            conv(name + '.downsample.0', planes * 2, cin, 1)

    conv('conv1', 64, 1, 5)
    gn('bn1', 64)
    bottleneck('layer1.0', 64, 64, True)
    bottleneck('layer2.0', 128, 128, True)
    bottleneck('layer3.0', 256, 128, False)
    for i in range(num_stacks):
        for d in range(2):
            for r in range(4 if d == 0 else 3):
                bottleneck('hg.%d.hg.%d.%d.0' % (i, d, r), 256, 128, False)
    for i in range(num_stacks):
        bottleneck('res.%d.0' % i, 256, 128, False)
    for i in range(num_stacks):
        conv('fc.%d.0' % i, 256, 256, 1)
        gn('fc.%d.1' % i, 256)
    for i in range(num_stacks):
        conv('score.%d' % i, num_outputs, 256, 1)
    for i in range(num_stacks - 1):
        conv('fc_.%d' % i, 256, 256, 1)
    for i in range(num_stacks - 1):
        conv('score_.%d' % i, 256, num_outputs, 1)
    return shapes


def _splitmix(idx, seed):
    z = (idx + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def det_uniform(n, seed):
    """Version-independent deterministic U(-1,1) stream (splitmix64), for fixtures that must be
    regenerated bit-identically on any box."""
    with np.errstate(over='ignore'):
        z = _splitmix(np.arange(1, n + 1, dtype=np.uint64), seed)
    return ((z >> np.uint64(40)).astype(np.float64) / float(1 << 24) * 2.0 - 1.0).astype(np.float32)


def det_state_dict(num_outputs=82, num_stacks=1, seed=7):
    """Deterministic weights: conv ~ U(-1,1)*sqrt(3/fan_in) (unit-gain), GN weight 1+0.1u, biases 0.1u."""
    sd = {}
    for i, (name, shp) in enumerate(param_shapes(num_outputs, num_stacks).items()):
        n = int(np.prod(shp))
        u = det_uniform(n, seed * 1000 + i).reshape(shp)
        if len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            u = u * np.float32(np.sqrt(3.0 / fan_in))
        elif name.endswith('.weight'):
            u = np.float32(1.0) + np.float32(0.1) * u
        else:
            u = np.float32(0.1) * u
        sd[name] = torch.from_numpy(np.ascontiguousarray(u.astype(np.float32)))
    return sd


class _RoundBf16(torch.autograd.Function):
    """Round-to-bf16 in the forward AND of the gradient in the backward: models a tensor materialised in bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _ident(x):
    return x


def _bottleneck(x, sd, p, q=_ident):
    def gn(t, n, g=16):
        return q(F.relu(F.group_norm(t, g, sd[p + n + '.weight'], sd[p + n + '.bias'])))

    def conv(t, n, pad=0, res=None):
        y = F.conv2d(t, q(sd[p + n + '.weight']), sd[p + n + '.bias'], padding=pad)
        return q(y if res is None else y + res)

    out = conv(gn(x, 'bn1'), 'conv1')
    out = conv(gn(out, 'bn2'), 'conv2', 1)
    a3 = gn(out, 'bn3')
    res = conv(x, 'downsample.0') if (p + 'downsample.0.weight') in sd else x
    return conv(a3, 'conv3', res=res)


def _hourglass(n, x, sd, p, q=_ident):
    up1 = _bottleneck(x, sd, '%shg.%d.0.0.' % (p, n - 1), q)
    low1 = _bottleneck(F.max_pool2d(x, 2, stride=2), sd, '%shg.%d.1.0.' % (p, n - 1), q)
    if n > 1:
        low2, latent = _hourglass(n - 1, low1, sd, p, q)
    else:
        low2 = _bottleneck(low1, sd, '%shg.%d.3.0.' % (p, n - 1), q)
        latent = low2
    low3 = _bottleneck(low2, sd, '%shg.%d.2.0.' % (p, n - 1), q)
    up2 = F.interpolate(low3, scale_factor=2, mode='bilinear', align_corners=False)
    return q(up1 + up2), latent


def hourglass_forward(x, sd, num_stacks=1, prefix='', round_bf16=False):
    """x [N,S,S] or [N,1,S,S] -> (list of score [N,num_outputs,S/4,S/4], list of latent [N,256,S/16,S/16]).

    round_bf16=False is the reference (fp32 everywhere).  round_bf16=True evaluates the SAME graph in fp32 arithmetic
    but rounds to bf16 every tensor (and its gradient) that the B200 implementation materialises in bf16 — conv
    weights, conv outputs (after the fused residual add), GroupNorm+ReLU outputs, the up-sample sum — which is the
    yardstick for "no worse than bf16 storage allows" in tests/test_gpu_nn.py."""
    p = prefix
    q = _RoundBf16.apply if round_bf16 else _ident
    if x.dim() == 3:
        x = x[:, None]
    x = q(F.conv2d(x, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'], stride=2, padding=2))
    x = q(F.relu(F.group_norm(x, 4, sd[p + 'bn1.weight'], sd[p + 'bn1.bias'])))
    x = _bottleneck(x, sd, p + 'layer1.0.', q)
    x = F.max_pool2d(x, 2, stride=2)
    x = _bottleneck(x, sd, p + 'layer2.0.', q)
    x = _bottleneck(x, sd, p + 'layer3.0.', q)
    outs, latents = [], []
    for i in range(num_stacks):
        y, latent = _hourglass(2, x, sd, '%shg.%d.' % (p, i), q)
        y = _bottleneck(y, sd, '%sres.%d.0.' % (p, i), q)
        y = q(F.conv2d(y, q(sd['%sfc.%d.0.weight' % (p, i)]), sd['%sfc.%d.0.bias' % (p, i)]))
        y = q(F.relu(F.group_norm(y, 16, sd['%sfc.%d.1.weight' % (p, i)], sd['%sfc.%d.1.bias' % (p, i)])))
        score = F.conv2d(y, q(sd['%sscore.%d.weight' % (p, i)]), sd['%sscore.%d.bias' % (p, i)])
        outs.append(score)
        latents.append(latent)
        if i < num_stacks - 1:
            t = q(x + F.conv2d(y, q(sd['%sfc_.%d.weight' % (p, i)]), sd['%sfc_.%d.bias' % (p, i)]))
            x = q(t + F.conv2d(q(score), q(sd['%sscore_.%d.weight' % (p, i)]), sd['%sscore_.%d.bias' % (p, i)]))
    return outs, latents


def two_stack_from_trained(sd1, seed=11, gain=0.05):
    """A 2-stack state_dict built from the reference's trained 1-stack weights (pretrained/synthetic.pth): the trunk and stack 0
    are the trained tensors, stack 1 (`hg.1`, `res.1`, `fc.1`, `score.1`) is a copy of stack 0 and the two inter-stack
    re-injection convolutions (`fc_.0`, `score_.0`, hourglass.py:169-172 — absent from a 1-stack checkpoint) are small
    deterministic weights (unit-gain U(-1,1) * `gain`, zero bias), so both stacks produce the peaked heat-maps of a trained
    network.  A recipe, not an asset: fixtures at BASELINE.json's 2-stack shape are generated from and tested with it."""
    sd = {}
    for i, (name, shp) in enumerate(param_shapes(82, 2).items()):
        src = name
        for mod in ('hg.1.', 'res.1.', 'fc.1.', 'score.1.'):
            if name.startswith(mod):
                src = mod[:-2] + '0.' + name[len(mod):]
        if src in sd1:
            sd[name] = torch.as_tensor(sd1[src]).clone().float()
        elif len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            u = det_uniform(int(np.prod(shp)), seed * 1000 + i).reshape(shp) * np.float32(gain * np.sqrt(3.0 / fan_in))
            sd[name] = torch.from_numpy(np.ascontiguousarray(u.astype(np.float32)))
        else:
            sd[name] = torch.zeros(shp)
    return sd
