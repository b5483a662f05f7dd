"""GPU parity tests for the hourglass layer kernels and the whole network (run with -m gpu).

Floating-point kernels: the comparison is against plain PyTorch fp32 (TF32 disabled) on the same operands —
for single layers on operands already rounded to bf16 (so only accumulation order and the bf16 output rounding
differ: tolerance 1e-2 of the tensor's max), for the whole network against the fp32 oracle / the golden fixtures
of the reference (tolerance 2e-2 on heat-maps, 5e-2 on gradients: bf16 operand contract)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BF16 = torch.bfloat16

if torch.cuda.is_available():
    from spherehand_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def nhwc(x):      # fp32 NCHW -> bf16 NHWC
    return x.permute(0, 2, 3, 1).contiguous().to(BF16)


def nchw(x):      # bf16 NHWC -> fp32 NCHW
    return x.float().permute(0, 3, 1, 2).contiguous()


def stats_of(x_nhwc, G):
    N, H, W, C = x_nhwc.shape
    v = x_nhwc.float().reshape(N, H * W, G, C // G)
    return torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).contiguous()


@pytest.mark.parametrize('N,H,Cin,Cout,taps', [(2, 32, 128, 128, 9), (3, 16, 256, 128, 1), (2, 64, 64, 64, 9), (5, 8, 128, 256, 1),
                                               (4, 4, 128, 128, 9), (37, 8, 128, 128, 9), (2, 32, 256, 82, 1), (2, 32, 64, 128, 1), (3, 16, 128, 128, 9), (1, 64, 64, 64, 9)])
def test_conv_fwd_dgrad_wgrad(N, H, Cin, Cout, taps):
    torch.manual_seed(N * 100 + H)
    k = 3 if taps == 9 else 1
    x = torch.randn(N, Cin, H, H, device=DEV).to(BF16).float()
    w = (torch.randn(Cout, Cin, k, k, device=DEV) / (Cin * taps) ** 0.5)
    b = torch.randn(Cout, device=DEV)
    res = torch.randn(N, Cout, H, H, device=DEV).to(BF16).float()
    wq = w.to(BF16).float()
    ref = F.conv2d(x, wq, b, padding=k // 2) + res
    cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
    y_ld = cout_pad if Cout % 8 else Cout
    wf = torch.empty((taps, cout_pad, Cin), device=DEV, dtype=BF16)
    b_rows = (Cin + 127) // 128 * 128 if Cin > 64 else 64
    b_cols = (Cout + 63) // 64 * 64
    wb = torch.empty((taps, b_rows, b_cols), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, taps, cout_pad, Cin, wf, wb, b_rows, b_cols)
    y = torch.zeros((N, H, H, y_ld), device=DEV, dtype=BF16)
    y32 = torch.empty((N, Cout, H, H), device=DEV)
    G = 16 if Cout % 16 == 0 and 32 % (Cout // 16) == 0 else 0
    st = torch.zeros((N, 16, 2), device=DEV)
    use_res = Cout % 8 == 0
    ops.conv_fwd(nhwc(x), wf, b, N, H, H, Cin, Cout, cout_pad, taps, y=y, y_ld=y_ld, y_nchw=y32,
                 residual=nhwc(res) if use_res else None, stats=st if G else None, groups=16)
    if not use_res:
        ref = ref - res
    torch.cuda.synchronize()
    assert rel_err(y32.cpu(), ref.cpu()) < 2e-3                          # fp32 output: accumulation order only
    assert rel_err(nchw(y)[:, :Cout].cpu(), ref.cpu()) < 1e-2            # bf16 output rounding
    if y_ld > Cout:
        assert (y[..., Cout:] == 0).all()
    if G:
        assert rel_err(st.cpu(), stats_of(y, 16).cpu()) < 1e-3           # statistics describe the stored tensor
    # ---- data gradient: same kernel on flipped / transposed weights
    dy = torch.randn(N, Cout, H, H, device=DEV).to(BF16).float()
    dyp = torch.zeros((N, H, H, b_cols), device=DEV, dtype=BF16)
    dyp[..., :Cout] = nhwc(dy)
    dx = torch.empty((N, H, H, Cin), device=DEV, dtype=BF16)
    ops.conv_fwd(dyp, wb, None, N, H, H, b_cols, Cin, b_rows, taps, y=dx, y_ld=Cin)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, wq, dy, padding=k // 2)
    assert rel_err(nchw(dx).cpu(), ref_dx.cpu()) < 1e-2
    # ---- weight gradient (reference layout [Cout,Cin,k,k]); fp32 atomics over the pixel split
    dw = torch.zeros_like(w)
    ops.conv_wgrad(dyp, nhwc(x), N, H, H, Cin, Cin, b_cols, Cout, taps, dw)
    ref_dw = torch.nn.grad.conv2d_weight(x, w.shape, dy, padding=k // 2)
    assert rel_err(dw.cpu(), ref_dw.cpu()) < 2e-3
    if taps == 9:
        # [9][Cout][Cin] scratch (accumulating; W >= 16: kernel-row kernel, narrower: per-tap kernel with vector reductions) + batched
        # unpack into the reference layout (+=)
        scratch = torch.zeros(9 * Cout * Cin, device=DEV)
        ops.conv_wgrad3x3(dyp, nhwc(x), N, H, H, Cin, Cin, b_cols, Cout, scratch)
        dw2 = torch.ones_like(w)
        ops.unpack_wgrad_batch(torch.tensor([[0, 0, Cout, Cin]], dtype=torch.int32, device=DEV), scratch, dw2.view(-1))
        assert rel_err((dw2 - 1).cpu(), ref_dw.cpu()) < 2e-3


@pytest.mark.parametrize('N,H,Cin,Cout,res', [(3, 32, 128, 64, False), (5, 8, 256, 128, False), (2, 64, 64, 128, True), (7, 16, 128, 256, True),
                                              (37, 8, 64, 128, True), (1, 16, 256, 128, False)])
def test_conv1x1_with_fused_input_groupnorm(N, H, Cin, Cout, res):
    """sh_conv_fwd_gn / sh_conv_wgrad_gn (GroupNorm + ReLU applied to the operand tiles in shared memory) against the two-pass
    path sh_gn_relu_fwd -> sh_conv_fwd / sh_conv_wgrad on the same inputs: the forward is BIT-identical (same scale / shift
    expression, same bf16 rounding of the normalised operand, same MMA schedule); the weight gradient differs by atomics order only."""
    torch.manual_seed(N * 131 + H + Cin)
    x = (torch.randn(N, H, H, Cin, device=DEV) * 1.7 + 0.4).to(BF16)
    gamma = torch.rand(Cin, device=DEV) + 0.5
    beta = torch.randn(Cin, device=DEV) * 0.3
    stx = stats_of(x, 16)
    w = torch.randn(Cout, Cin, 1, 1, device=DEV) / Cin ** 0.5
    b = torch.randn(Cout, device=DEV)
    r = torch.randn(N, H, H, Cout, device=DEV).to(BF16) if res else None
    cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
    b_rows = (Cin + 127) // 128 * 128 if Cin > 64 else 64
    b_cols = (Cout + 63) // 64 * 64
    wf = torch.empty((1, cout_pad, Cin), device=DEV, dtype=BF16)
    wb = torch.empty((1, b_rows, b_cols), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, 1, cout_pad, Cin, wf, wb, b_rows, b_cols)
    a = torch.empty_like(x)
    ops.gn_relu_fwd(x, stx, gamma, beta, N, H * H, Cin, 16, a)
    out = {}
    for fused in (False, True):
        y = torch.zeros((N, H, H, Cout), device=DEV, dtype=BF16)
        st = torch.zeros((N, 16, 2), device=DEV)
        ops.conv_fwd(x if fused else a, wf, b, N, H, H, Cin, Cout, cout_pad, 1, y=y, y_ld=Cout, residual=r, stats=st, groups=16,
                     gn=(stx, gamma, beta, 16, 1e-5) if fused else None)
        dy = torch.randn(N, H, H, Cout, generator=torch.Generator(device=DEV).manual_seed(5), device=DEV).to(BF16)
        dw = torch.zeros_like(w)
        ops.conv_wgrad(dy, x if fused else a, N, H, H, Cin, Cin, Cout, Cout, 1, dw, gn=(stx, gamma, beta, 16, 1e-5) if fused else None)
        torch.cuda.synchronize()
        out[fused] = (y, st, dw)
    assert torch.equal(out[True][0], out[False][0]), (out[True][0].float() - out[False][0].float()).abs().max()
    assert rel_err(out[True][1].cpu(), out[False][1].cpu()) < 1e-4
    assert rel_err(out[True][2].cpu(), out[False][2].cpu()) < 1e-4
    ref = F.conv2d(nchw(a), w.to(BF16).float(), b) + (nchw(r) if res else 0)
    assert rel_err(nchw(out[True][0]).cpu(), ref.cpu()) < 1e-2
    # the raw operand was left untouched in HBM
    assert torch.equal(stats_of(x, 16), stx)


@pytest.mark.parametrize('N,H,W,Cin,Cout', [(40, 32, 32, 128, 128), (19, 16, 16, 128, 128), (5, 64, 64, 64, 64), (75, 16, 8, 256, 128),
                                            (2, 32, 32, 128, 82)])
def test_conv3x3_halo_mode(N, H, W, Cin, Cout):
    """3x3 without a residual on images >= 16x8 takes the halo path (one TMA halo box per k-block, nine shifted UMMA
    descriptors, two pixel tiles per weight tile): CTAs with 1, 2 and 3 tiles, odd tile counts, image borders, 2 and 4 k-blocks."""
    torch.manual_seed(N + H + Cin)
    x = torch.randn(N, Cin, H, W, device=DEV).to(BF16).float()
    w = torch.randn(Cout, Cin, 3, 3, device=DEV) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, device=DEV)
    ref = F.conv2d(x, w.to(BF16).float(), b, padding=1)
    cout_pad = (Cout + 127) // 128 * 128 if Cout > 64 else 64
    y_ld = cout_pad if Cout % 8 else Cout
    wf = torch.empty((9, cout_pad, Cin), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, 9, cout_pad, Cin, wf)
    y = torch.zeros((N, H, W, y_ld), device=DEV, dtype=BF16)
    y32 = torch.empty((N, Cout, H, W), device=DEV)
    G = 16 if Cout % 16 == 0 else 0
    st = torch.zeros((N, 16, 2), device=DEV)
    ops.conv_fwd(nhwc(x), wf, b, N, H, W, Cin, Cout, cout_pad, 9, y=y, y_ld=y_ld, y_nchw=y32, stats=st if G else None, groups=16)
    torch.cuda.synchronize()
    assert rel_err(y32.cpu(), ref.cpu()) < 2e-3
    assert rel_err(nchw(y)[:, :Cout].cpu(), ref.cpu()) < 1e-2
    if G:
        assert rel_err(st.cpu(), stats_of(y, 16).cpu()) < 1e-3
    # the nine-box path (SH_CONV_HALO=0, read per call) on the same input: fp32 results equal up to accumulation order
    os.environ['SH_CONV_HALO'] = '0'
    try:
        y32b = torch.empty_like(y32)
        ops.conv_fwd(nhwc(x), wf, b, N, H, W, Cin, Cout, cout_pad, 9, y=torch.empty_like(y), y_ld=y_ld, y_nchw=y32b)
        torch.cuda.synchronize()
    finally:
        del os.environ['SH_CONV_HALO']
    assert rel_err(y32.cpu(), y32b.cpu()) < 1e-5


@pytest.mark.parametrize('N,H,C,G', [(3, 16, 128, 16), (2, 32, 256, 16), (2, 32, 64, 16), (2, 32, 64, 4),
                                     (5, 4, 256, 16), (3, 8, 256, 16), (4, 8, 128, 16)])     # the last three: one block per sample
def test_gn_relu_fwd_bwd(N, H, C, G):
    torch.manual_seed(C + G)
    x = (torch.randn(N, C, H, H, device=DEV) * 2 + 0.5).to(BF16).float().requires_grad_(True)
    gamma = (torch.rand(C, device=DEV) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device=DEV) * 0.3).requires_grad_(True)
    ref = F.relu(F.group_norm(x, G, gamma, beta))
    da = torch.randn_like(ref).to(BF16).float()
    add = torch.randn_like(ref).to(BF16).float()
    ref.backward(da)
    xh = nhwc(x.detach())
    st = stats_of(xh, G)
    y = torch.empty_like(xh)
    st_out = torch.zeros((N, 16, 2), device=DEV)
    ops.gn_relu_fwd(xh, st, gamma.detach(), beta.detach(), N, H * H, C, G, y, st_out, 16)
    assert rel_err(nchw(y).cpu(), ref.detach().cpu()) < 1e-2
    assert rel_err(st_out.cpu(), stats_of(y, 16).cpu()) < 1e-3
    red = ops.gn_relu_bwd_scratch(N, G, DEV)
    dg, db, cs = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dx = torch.empty_like(xh)
    ops.gn_relu_bwd(nhwc(da), xh, st, gamma.detach(), beta.detach(), N, H * H, C, G, red, dg, db, dx, nhwc(add), cs)
    assert rel_err(nchw(dx).cpu(), (x.grad + add).cpu()) < 1.5e-2
    assert rel_err(dg.cpu(), gamma.grad.cpu()) < 1e-2 and rel_err(db.cpu(), beta.grad.cpu()) < 1e-2
    assert rel_err(cs.cpu(), nchw(dx).sum(dim=(0, 2, 3)).cpu()) < 1e-3


def test_pool_upsample_add_stem_adam():
    torch.manual_seed(3)
    N, H, C = 3, 16, 128
    x = torch.randn(N, C, 2 * H, 2 * H, device=DEV).to(BF16).float().requires_grad_(True)
    ref = F.max_pool2d(x, 2, 2)
    y = torch.empty((N, H, H, C), device=DEV, dtype=BF16)
    st = torch.zeros((N, 16, 2), device=DEV)
    ops.maxpool_fwd(nhwc(x.detach()), N, H, H, C, y, st, 16)
    assert torch.equal(nchw(y), ref.detach())
    assert rel_err(st.cpu(), stats_of(y, 16).cpu()) < 1e-3
    dy = torch.randn_like(ref).to(BF16).float()
    ref.backward(dy)
    dx = torch.empty((N, 2 * H, 2 * H, C), device=DEV, dtype=BF16)
    ops.maxpool_bwd(nhwc(dy), nhwc(x.detach()), N, H, H, C, dx)
    assert torch.equal(nchw(dx), x.grad)
    # up-sample + add
    low = torch.randn(N, C, H, H, device=DEV).to(BF16).float().requires_grad_(True)
    up1 = torch.randn(N, C, 2 * H, 2 * H, device=DEV).to(BF16).float()
    ref = up1 + F.interpolate(low, scale_factor=2, mode='bilinear', align_corners=False)
    y2 = torch.empty((N, 2 * H, 2 * H, C), device=DEV, dtype=BF16)
    ops.upsample_add_fwd(nhwc(up1), nhwc(low.detach()), N, H, H, C, y2)
    assert rel_err(nchw(y2).cpu(), ref.detach().cpu()) < 1e-2
    dy2 = torch.randn_like(ref).to(BF16).float()
    ref.backward(dy2)
    dlow = torch.empty((N, H, H, C), device=DEV, dtype=BF16)
    cs = torch.zeros(C, device=DEV)
    ops.upsample_bwd(nhwc(dy2), N, H, H, C, dlow, cs)
    assert rel_err(nchw(dlow).cpu(), low.grad.cpu()) < 1e-2
    assert rel_err(cs.cpu(), nchw(dlow).sum(dim=(0, 2, 3)).cpu()) < 1e-3
    # add (3 operands) + colsum
    a, b, c = [torch.randn(N, H, H, C, device=DEV).to(BF16) for _ in range(3)]
    y3 = torch.empty_like(a)
    ops.add(a, b, N, H * H, C, y3, c=c)
    assert rel_err(y3.float().cpu(), (a.float() + b.float() + c.float()).cpu()) < 1e-2
    cs2 = torch.zeros(C, device=DEV)
    ops.colsum(a, N, H * H, C, cs2)
    assert rel_err(cs2.cpu(), a.float().sum(dim=(0, 1, 2)).cpu()) < 1e-4
    # stem conv 5x5 s2 + its weight gradient
    S = 64
    img = torch.randn(N, 1, S, S, device=DEV)
    w = (torch.randn(64, 1, 5, 5, device=DEV) * 0.2).requires_grad_(True)
    bias = torch.randn(64, device=DEV).requires_grad_(True)
    ref = F.conv2d(img, w, bias, stride=2, padding=2)
    ys = torch.empty((N, S // 2, S // 2, 64), device=DEV, dtype=BF16)
    st4 = torch.zeros((N, 4, 2), device=DEV)
    ops.stem_conv_fwd(img[:, 0].contiguous(), w.detach(), bias.detach(), N, S, ys, st4, 4)
    assert rel_err(nchw(ys).cpu(), ref.detach().cpu()) < 1e-2
    assert rel_err(st4.cpu(), stats_of(ys, 4).cpu()) < 1e-3
    dys = torch.randn_like(ref).to(BF16).float()
    ref.backward(dys)
    dw, dbias = torch.zeros_like(w), torch.zeros_like(bias)
    ops.stem_conv_wgrad(img[:, 0].contiguous(), nhwc(dys), N, S, dw, dbias)
    assert rel_err(dw.cpu(), w.grad.cpu()) < 1e-3 and rel_err(dbias.cpu(), bias.grad.cpu()) < 1e-3
    # layout conversions
    t = torch.randn(N, 82, 8, 8, device=DEV)
    tp = torch.empty((N, 8, 8, 128), device=DEV, dtype=BF16)
    ops.nchw_to_nhwc(t, N, 82, 64, 128, tp)
    assert torch.equal(tp[..., :82], nhwc(t)) and (tp[..., 82:] == 0).all()
    back = torch.empty((N, 128, 8, 8), device=DEV)
    ops.nhwc_to_nchw(tp, N, 128, 64, back)
    assert torch.equal(back, nchw(tp))
    # Adam == torch.optim.Adam(weight_decay) step for step
    p = torch.randn(1000, device=DEV)
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3, weight_decay=1e-5)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(1000, device=DEV)
        p_ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-5, step)
        assert rel_err(p.cpu(), p_ref.detach().cpu()) < 1e-5


@pytest.mark.parametrize('stacks', [1, 2])
def test_hourglass_golden(stacks):
    """Whole network, forward + backward, against the reference's own outputs (deterministic weights)."""
    from spherehand_b200.network.hourglass import create_hourglass_network
    from oracle.hourglass import det_state_dict, det_uniform
    g = golden('hourglass_%dstack' % stacks)
    net = create_hourglass_network(82, stacks).to(DEV)
    net.load_state_dict(det_state_dict(82, stacks, seed=7))
    x = torch.from_numpy(g['x']).to(DEV)
    outs, lats = net(x)
    errs = []
    for i, o in enumerate(outs):
        errs.append(rel_err(o.detach().cpu(), g['score%d' % i]))
        errs.append(rel_err(lats[i].cpu(), g['latent%d' % i]))
    print('hourglass %d-stack forward rel errs (score, latent per stack):' % stacks, ['%.4f' % e for e in errs])
    assert max(errs) < 2e-2
    # ---- gradients.  A deep bf16 network has inherent rounding noise in its gradients (tools/diag_hourglass.py), so the
    # yardstick is torch evaluating the SAME graph in fp32 arithmetic with bf16 rounding at the same materialisation
    # points (oracle.hourglass round_bf16=True): our error against the fp32 reference must not exceed that
    # emulation's error (x2.5 + 2e-2 slack: both are noise realisations, sums over as few as 16 pixels at the 4x4 level), parameter by parameter, for a white-noise upstream gradient (the golden
    # fixture's) and for the MSE heat-map loss the reference trains with.
    from oracle.hourglass import hourglass_forward
    sd0 = {k: v.to(DEV) for k, v in det_state_dict(82, stacks, seed=7).items()}
    tgt = torch.from_numpy(det_uniform(outs[0].numel(), 55).reshape(outs[0].shape)).to(DEV).abs() * 0.1

    def loss_of(os_, kind):
        if kind == 'noise':
            return sum((o * torch.from_numpy(det_uniform(o.numel(), 100 + i).reshape(o.shape)).to(DEV)).sum() for i, o in enumerate(os_))
        return sum(((o - tgt) ** 2).mean() * 1e3 for o in os_)

    def nrm(a, b):
        return float((a - b).double().norm() / b.double().norm().clamp_min(1e-30))

    for kind in ('noise', 'mse'):
        grads = {}
        for mode in ('fp32', 'emul'):
            sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
            loss_of(hourglass_forward(x, sd, stacks, round_bf16=(mode == 'emul'))[0], kind).backward()
            grads[mode] = {k: v.grad for k, v in sd.items()}
        net.zero_grad()
        outs, _ = net(x)
        loss_of(outs, kind).backward()
        ours = {k: p.grad for k, p in net.named_parameters()}
        if kind == 'noise':     # the fp32 torch-on-GPU reference agrees with the committed fixture of the reference itself
            for k in g:
                if k.startswith('grad.'):
                    assert rel_err(grads['fp32'][k[5:]].cpu(), g[k]) < 5e-3
        worst = []
        for k in sd0:
            e_ours, e_emul = nrm(ours[k], grads['fp32'][k]), nrm(grads['emul'][k], grads['fp32'][k])
            worst.append((e_ours - 2.5 * e_emul, k, e_ours, e_emul))
        worst.sort(reverse=True)
        print('hourglass %d-stack %s: worst (ours, emulated-bf16) norm-rel grad errors:' % (stacks, kind),
              [(k, '%.4f' % a, '%.4f' % b) for _, k, a, b in worst[:4]])
        assert worst[0][0] < 2e-2, worst[0]
        med = float(np.median([w[2] for w in worst]))
        assert med < (0.25 if kind == 'noise' else 0.03)


def test_fused_groupnorm_conv_is_stable_under_reordered_loads():
    """The fused kernel's operand-transform warps and its MMA issuer meet the phases of the pipeline barriers in a fixed order; TMA
    loads of the operand (x, here L2-resident) and of the residual (evicted before every launch) complete out of order, which is
    what once let a transform thread pass a barrier on a stale parity.  300 launches, every output bit-identical to the first."""
    torch.manual_seed(11)
    N, H, Cin, Cout = 64, 32, 128, 256
    x = (torch.randn(N, H, H, Cin, device=DEV) * 1.3).to(BF16)
    gamma, beta = torch.rand(Cin, device=DEV) + 0.5, torch.randn(Cin, device=DEV) * 0.3
    stx = stats_of(x, 16)
    w = torch.randn(Cout, Cin, 1, 1, device=DEV) / Cin ** 0.5
    b = torch.randn(Cout, device=DEV)
    r = torch.randn(N, H, H, Cout, device=DEV).to(BF16)
    wf = torch.empty((1, 256, Cin), device=DEV, dtype=BF16)
    ops.pack_weights(w, Cout, Cin, 1, 256, Cin, wf)
    flush = torch.empty(160 << 20, device=DEV, dtype=torch.uint8)
    hog_a, hog_b = torch.empty(256 << 20, device=DEV, dtype=torch.uint8), torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
    side = torch.cuda.Stream()
    first = None
    for it in range(300):
        if it % 3 != 2:
            flush.zero_()                                   # evict everything ...
            x.add_(0)                                       # ... then bring x (and not the residual) back into L2
        if it % 2:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                   # HBM contention from a copy running beside the convolution
                hog_b.copy_(hog_a)
        y = torch.zeros((N, H, H, Cout), device=DEV, dtype=BF16)
        ops.conv_fwd(x, wf, b, N, H, H, Cin, Cout, 256, 1, y=y, y_ld=Cout, residual=r, groups=16, gn=(stx, gamma, beta, 16, 1e-5))
        if first is None:
            first = y
        elif it % 10 == 0 or it == 299:
            assert torch.equal(y, first), it
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()


def test_hourglass_with_fused_groupnorm_equals_two_pass_network():
    """The whole 2-stack network with GroupNorm + ReLU folded into the bottleneck 1x1 layers (default) against the two-pass network
    (SH_FUSE_GN=0 path).  Each fused convolution is bit-identical to its two-pass form on the same operands
    (test_conv1x1_with_fused_input_groupnorm); through the network the GroupNorm statistics are fp32 atomic sums whose order differs
    from run to run, so the yardstick for heat-maps and parameter gradients is the run-to-run difference of the two-pass network."""
    from spherehand_b200.network import hourglass as hg
    from oracle.hourglass import det_state_dict, det_uniform
    net = hg.create_hourglass_network(82, 2).to(DEV)
    net.load_state_dict(det_state_dict(82, 2, seed=3))
    x = torch.from_numpy(det_uniform(5 * 128 * 128, 4).reshape(5, 128, 128)).to(DEV)
    w = [torch.from_numpy(det_uniform(5 * 82 * 32 * 32, 5 + i).reshape(5, 82, 32, 32)).to(DEV) for i in range(2)]
    def run(fused):
        old, hg.FUSE_GN = hg.FUSE_GN, fused
        try:
            net.zero_grad()
            o, _ = net(x)
            (o[0] * w[0] + o[1] * w[1]).sum().backward()
        finally:
            hg.FUSE_GN = old
        return [t.detach().clone() for t in o], torch.cat([p.grad.flatten() for p in net.parameters()])
    o_f, g_f = run(True)
    o_u, g_u = run(False)
    o_u2, g_u2 = run(False)
    nrm = lambda a, b: float((a - b).double().norm() / b.double().norm())
    for i in range(2):
        floor = nrm(o_u2[i], o_u[i])
        print('stack %d heat-maps, fused vs two-pass: %.3e (two-pass run-to-run: %.3e)' % (i, nrm(o_f[i], o_u[i]), floor))
        assert nrm(o_f[i], o_u[i]) < 2 * floor + 1e-3
    floor = nrm(g_u2, g_u)
    print('fused vs two-pass gradients: %.3e (two-pass run-to-run: %.3e)' % (nrm(g_f, g_u), floor))
    assert nrm(g_f, g_u) < 2 * floor + 1e-3


def test_hourglass_backward_uses_its_own_forward_tape():
    """Two forwards (different batch sizes) before a backward, and a no-grad forward in between: each autograd node must
    back-propagate through the activations of ITS forward (the tape used to be the most recent forward's)."""
    from spherehand_b200.network.hourglass import create_hourglass_network
    from oracle.hourglass import det_state_dict, det_uniform
    net = create_hourglass_network(82, 1).to(DEV)
    net.load_state_dict(det_state_dict(82, 1, seed=7))
    x1 = torch.from_numpy(det_uniform(2 * 64 * 64, 1).reshape(2, 64, 64)).to(DEV)
    x2 = torch.from_numpy(det_uniform(3 * 64 * 64, 2).reshape(3, 64, 64)).to(DEV)
    w = torch.from_numpy(det_uniform(2 * 82 * 16 * 16, 3).reshape(2, 82, 16, 16)).to(DEV)
    def l2(a, b):
        return float((a - b).double().norm() / b.double().norm().clamp_min(1e-30))
    runs = []
    for _ in range(4):
        net.zero_grad()
        o, _ = net(x1)
        (o[0] * w).sum().backward()
        runs.append({k: p.grad.clone() for k, p in net.named_parameters()})
    want = runs[0]
    # run-to-run noise of the same pass (fp32 atomics order -> bf16 flips), worst of three repeats: single small tensors move by several %
    floor = {k: max(l2(r[k], want[k]) for r in runs[1:]) for k in want}
    net.zero_grad()
    o1, _ = net(x1)
    o2, _ = net(x2)                                   # a later forward with another batch size
    with torch.no_grad():
        net(x2)                                       # and an inference pass
    (o1[0] * w).sum().backward()
    worst = max((l2(p.grad, want[k]) - 3 * floor[k], k) for k, p in net.named_parameters())
    print('own-tape backward vs a plain forward/backward: worst excess over 3x the run-to-run floor', worst, 'largest floor', max(floor.values()))
    assert worst[0] < 5e-2, worst
    allg = lambda d: torch.cat([d[k].flatten() for k in want])
    agg, agg_floor = l2(allg({k: p.grad for k, p in net.named_parameters()}), allg(want)), max(l2(allg(r), allg(want)) for r in runs[1:])
    assert agg < 2 * agg_floor + 1e-2, (agg, agg_floor)
    # a wrong tape (the batch-3 forward's) would not even have the right shapes; a stale one of the same shape would be ~100 % off
    (o2[0] * 0.5).sum().backward()                    # the second node still has its own tape
    with pytest.raises(RuntimeError):
        (o2[0] * 0.5).sum().backward()
