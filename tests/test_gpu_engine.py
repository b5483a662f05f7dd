"""GPU parity tests of the fused train step (spherehand_b200.engine.SelfSupTrainStep — the product path bench.py times):
loss terms and parameter gradients against the CPU oracle of the whole step (oracle/full_step.py, itself pinned against
the reference's own step in tests/test_oracle_golden.py), CUDA-graph replay against the eager launch sequence, the fused
Adam against torch.optim.Adam, and size-independent properties at BASELINE.json's full size (256 images of 128x128,
2 stacks).  Tolerances: the hourglass runs bf16 operands / fp32 accumulation, so network-dependent quantities
are compared at 5e-2; everything that does not pass through the network (Adam arithmetic, term bookkeeping) at 1e-5."""
import os

import numpy as np
import pytest
import torch

from conftest import golden, check_grads

pytestmark = pytest.mark.gpu
DEV = 'cuda'

if torch.cuda.is_available():
    from spherehand_b200 import data, ops
    from spherehand_b200.engine import SelfSupTrainStep, TERM_NAMES
    from spherehand_b200.model import HandModel
    from spherehand_b200.network.hourglass import create_hourglass_network
from oracle import full_step as ofs, hourglass as oh


def same_terms(a, b):
    """Two runs of the same step: fp32 atomics order differs between launches, single-ulp bf16 flips follow, and the flat
    heat-maps of a random-weight network amplify them: ~0.3 % on the smooth terms, a few % on the two hinge terms (sums of
    relu over a few active sphere pairs / bones)."""
    a, b = a.detach().cpu().numpy().astype(np.float64), b.detach().cpu().numpy().astype(np.float64)
    total = max(abs(a[-1]), abs(b[-1]))
    for i, k in enumerate(TERM_NAMES):
        hinge = k in ('collision', 'bone_length')
        # a hinge term can switch single pairs on or off: its absolute slack is tied to the size of the whole loss
        if abs(a[i] - b[i]) > (8e-2 if hinge else 1.5e-2) * max(abs(a[i]), abs(b[i])) + (2e-3 * total if hinge else 1e-3):
            print('term', k, a[i], b[i])
            return False
    return True


def make_step(hand_model, B, V, Ns, S, stacks, use_graph, seed=7, lr=1e-4):
    hand = HandModel.from_arrays(hand_model, DEV)
    vae_sd = {k: torch.from_numpy(v) for k, v in golden('pose_vae').items()}
    blob = ops.vae_blob_from_state_dict(vae_sd, DEV)
    net = create_hourglass_network(82, stacks).to(DEV)
    sd0 = oh.det_state_dict(82, stacks, seed=seed)
    net.load_state_dict(sd0)
    step = SelfSupTrainStep(net, hand, blob, B, V, Ns, S, lr=lr, use_graph=use_graph)
    gen = torch.Generator().manual_seed(1)
    real, cams, inv = data.synthetic_real_batch(hand, B, V, S, gen)
    poses = data.random_poses(Ns, gen)
    step.load_batch(real, cams, inv, poses)
    torch.manual_seed(3)
    step.draw_randoms()
    return step, sd0, vae_sd, dict(real=real.cpu(), cams=cams.cpu(), inv_cams=inv.cpu(), poses=poses.cpu())


def test_train_step_terms_and_gradients_vs_oracle(hand_model):
    B, V, Ns, S, stacks = 2, 3, 2, 64, 1
    step, sd0, vae_sd, batch = make_step(hand_model, B, V, Ns, S, stacks, use_graph=False)
    flat0 = step.net._flat.clone()
    terms = step.step(is_mv=True).cpu().numpy()
    assert np.isfinite(terms).all()
    batch.update(scales=step.scales.cpu(), rand_f=step.rand_f.cpu(), noise=step.noise.cpu(), eps=step.vae_eps.cpu())
    out = {}
    for mode in ('fp32', 'emul'):
        sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
        out[mode] = ofs.train_step(sd, stacks, ofs.HandTables(hand_model), vae_sd, batch, S, apply_update=False, round_bf16=(mode == 'emul'))
    ref, grads = out['emul'][0], out['fp32'][1]
    for k, v in zip(TERM_NAMES, terms):
        assert abs(v - ref[k]) <= 5e-2 * abs(ref[k]) + 1e-3, (k, float(v), ref[k])
    assert abs(terms[-1] - terms[:-1].sum()) <= 1e-5 * abs(terms[-1])                      # 'total' is the sum of the 8 terms
    # parameter gradients (the flat buffer the all-reduce and Adam see) against the fp32 oracle's autograd, held to what bf16
    # activations allow: the error of torch evaluating the same graph with bf16 rounding at the kernels' materialisation points
    # (random weights: flat heat-maps, the soft-argmax amplifies the rounding; the tight joint / term bounds are on trained weights,
    # tests/test_gpu_trained.py)
    ours = {name: step.net.grad_view(p).detach().float().cpu() for name, p in step.net.named_parameters()}
    # (random weights: two noise realisations of a 20 % effect, aggregates within 1.5 x; trained weights hold 1.25 x)
    check_grads(ours, grads, out['emul'][1], 'fused step, deterministic random weights', cos_min=0.95, agg=1.5)
    # the fused Adam on OUR gradient == torch.optim.Adam(lr, weight_decay=1e-5) on the same gradient (engine.py:95-97 of the reference)
    p_ref = flat0.clone().requires_grad_(True)
    p_ref.grad = step.net._flat_grad.clone()
    torch.optim.Adam([p_ref], lr=1e-4, weight_decay=1e-5).step()
    assert float((step.net._flat - p_ref.detach()).abs().max()) <= 2e-7                    # one fp32 ulp of a weight; the update is 1e-4


def test_graph_replay_equals_eager_over_steps(hand_model):
    B, V, Ns, S, stacks = 2, 3, 2, 64, 2
    eager, _, _, _ = make_step(hand_model, B, V, Ns, S, stacks, use_graph=False)
    graph, _, _, _ = make_step(hand_model, B, V, Ns, S, stacks, use_graph=True)
    for it in range(3):
        # same weights and optimiser state at the start of every step (two trajectories drift apart through the noise
        # described in same_terms; what is compared is one step at a time, multi-view and single-view)
        eager.net._flat.copy_(graph.net._flat); eager.adam_m.copy_(graph.adam_m); eager.adam_v.copy_(graph.adam_v)
        eager.step_dev.copy_(graph.step_dev)
        for st in (eager, graph):
            torch.manual_seed(100 + it)
            st.draw_randoms()
        te = eager.step(is_mv=(it != 1)).clone()
        tg = graph.step(is_mv=(it != 1)).clone()
        # same kernels, same inputs; fp32 atomics order differs between runs -> bf16 flips -> small relative differences
        assert same_terms(te, tg), (it, te.tolist(), tg.tolist())
    assert graph.launches_per_step and graph.launches_per_step > 100
    d = (eager.net._flat - graph.net._flat).abs().max()
    assert float(d) <= 2 * 1e-4 * (1 + 1e-3)             # one Adam step of at most lr each, in either direction


def test_full_size_step_properties(hand_model):
    """BASELINE.json config 3 (B = 64 tuples x 3 views + 64 synthetic poses = 256 images of 128x128, 2 stacks): properties that
    hold at any size.  (1) every term finite, total = sum; (2) the first Adam step moves no parameter by more than lr;
    (3) a second replay of the graph on the same inputs with the same weights reproduces the terms; (4) the single-view
    step (is_mv=False) only changes the two multi-view terms' weighting, not the synthetic / prior terms."""
    B, V, Ns, S, stacks = 64, 3, 64, 128, 2
    step, _, _, _ = make_step(hand_model, B, V, Ns, S, stacks, use_graph=True, lr=1e-4)
    flat0 = step.net._flat.clone()
    m0, v0, s0 = step.adam_m.clone(), step.adam_v.clone(), step.step_dev.clone()
    t1 = step.step(is_mv=True).clone()
    assert torch.isfinite(t1).all()
    assert abs(float(t1[-1] - t1[:-1].sum())) <= 1e-5 * abs(float(t1[-1]))
    dp = (step.net._flat - flat0).abs()
    assert float(dp.max()) <= 1e-4 * (1 + 1e-3) and float(dp.max()) > 0.5e-4
    assert torch.isfinite(step.net._flat).all()
    # rewind the weights and optimiser state, replay: same inputs -> same terms (up to atomics order)
    step.net._flat.copy_(flat0); step.adam_m.copy_(m0); step.adam_v.copy_(v0); step.step_dev.copy_(s0)
    t2 = step.step(is_mv=True).clone()
    assert same_terms(t1, t2), (t1.tolist(), t2.tolist())
    step.net._flat.copy_(flat0); step.adam_m.copy_(m0); step.adam_v.copy_(v0); step.step_dev.copy_(s0)
    t3 = step.step(is_mv=False).clone()
    names = list(TERM_NAMES)
    for k in ('synt_uv', 'synt_d', 'pose_prior', 'uv_hm_mean', 'collision', 'bone_length'):
        i = names.index(k)
        hinge = k in ('collision', 'bone_length')
        assert abs(float(t3[i] - t1[i])) <= (8e-2 if hinge else 1.5e-2) * abs(float(t1[i])) + (2e-3 * abs(float(t1[-1])) if hinge else 1e-3), k


def test_checkpoint_roundtrip_in_the_reference_format(hand_model, tmp_path):
    """SelfSupTrainStep.checkpoint / load_checkpoint speak `Engine.save_model` / `load_model`'s format (network/engine.py:437-459):
    after two steps the dict survives torch.save / torch.load, a stock torch.optim.Adam over a fresh copy of the network loads
    its optimiser part unchanged (what the reference would do), and a second fused step object restored from it holds the same
    weights, moments, step count and learning rate bit for bit."""
    B, V, Ns, S, stacks = 2, 3, 2, 64, 1
    a, _, _, _ = make_step(hand_model, B, V, Ns, S, stacks, use_graph=False, lr=3e-4)
    for _ in range(2):
        a.step(is_mv=True)
    ck = a.checkpoint(epoch=7)
    path = str(tmp_path / 'model_7.pth')
    torch.save(ck, path)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {'epoch', 'network_state_dict', 'optimizer_state_dict'} and ck['epoch'] == 7
    keys = set(ck['network_state_dict'])
    assert {'xyz_recover.u_grid', 'xyz_recover.v_grid'} <= keys and all(k.startswith('hg.') or k.startswith('xyz_recover.') for k in keys)
    assert ck['network_state_dict']['xyz_recover.u_grid'][0, 0, 3, 5] == 5 and ck['network_state_dict']['xyz_recover.v_grid'][0, 0, 3, 5] == 3
    # reference side: stock Adam over a stock copy of the parameters
    ref_net = create_hourglass_network(82, stacks).to(DEV)
    ref_net.load_state_dict({k[3:]: v for k, v in ck['network_state_dict'].items() if k.startswith('hg.')})
    ref_opt = torch.optim.Adam(ref_net.parameters(), lr=1.0, weight_decay=1e-5)
    ref_opt.load_state_dict(ck['optimizer_state_dict'])
    assert ref_opt.param_groups[0]['lr'] == 3e-4 and ref_opt.param_groups[0]['weight_decay'] == 1e-5
    p0 = next(iter(ref_net.parameters()))
    assert int(float(ref_opt.state[p0]['step'])) == 2 and ref_opt.state[p0]['exp_avg'].shape == p0.shape
    # fused side: a new step object restored from the checkpoint
    b, _, _, _ = make_step(hand_model, B, V, Ns, S, stacks, use_graph=False, seed=99, lr=1e-4)
    assert not torch.equal(a.net._flat, b.net._flat)
    b.load_checkpoint(ck)
    assert torch.equal(a.net._flat, b.net._flat) and torch.equal(a.adam_m, b.adam_m) and torch.equal(a.adam_v, b.adam_v)
    assert int(b.step_dev.item()) == 2 and b.lr == 3e-4 and abs(float(b.lr_dev.item()) - 3e-4) < 1e-9
    tb = b.step(is_mv=True).clone()
    ta = a.step(is_mv=True).clone()
    assert same_terms(ta, tb)


def test_prefetch_commit_equals_load_batch(hand_model):
    """The double-buffered input path (prefetch_batch on a copy stream + commit_batch) delivers the same buffers as load_batch,
    batch after batch, also when the next prefetch is issued before the current step has run."""
    B, V, Ns, S, stacks = 2, 3, 2, 64, 1
    step, _, _, batch = make_step(hand_model, B, V, Ns, S, stacks, use_graph=False)
    with pytest.raises(RuntimeError):
        step.commit_batch()
    gen = torch.Generator().manual_seed(9)
    batches = []
    for _ in range(3):
        batches.append([(t + torch.rand(t.shape, generator=gen) * 0.01).pin_memory() for t in (batch['real'], batch['cams'], batch['inv_cams'], batch['poses'])])
    n = step.prefetch_batch(*batches[0])
    assert n == sum(t.numel() * 4 for t in batches[0])
    for k in range(3):
        step.commit_batch()
        if k + 1 < 3:
            step.prefetch_batch(*batches[k + 1])           # overlaps the step below
        step.draw_randoms()
        step.step(is_mv=True)
        for buf, want in zip((step.real, step.cams, step.inv_cams, step.poses), batches[k]):
            assert torch.equal(buf.cpu(), want)


def test_two_shards_sum_to_the_global_step(hand_model):
    """Data-parallel semantics ON THE KERNELS: the global batch split into two rank shards, each run through the fused
    step with world_size=2 (mean_scale = 1/2 on the batch-MEAN terms, M_mean = global rows in the VAE prior, one shard through the bucketed
    reduction path and one through the single all-reduce, collective replaced by the identity), gradients and terms SUMMED by hand == the unsharded world_size=1 step on the whole
    batch.  Trained weights, so the comparison is tight: same kernels on the same images, only the fp32 summation order differs."""
    from oracle.hourglass import two_stack_from_trained  # noqa: F401  (weights recipe only)
    B, V, Ns, S, stacks = 4, 3, 4, 64, 1
    w1 = {k: torch.from_numpy(v) for k, v in golden('trained_weights').items()}
    hand = HandModel.from_arrays(hand_model, DEV)
    vae_sd = {k: torch.from_numpy(v) for k, v in golden('pose_vae').items()}
    blob = ops.vae_blob_from_state_dict(vae_sd, DEV)
    gen = torch.Generator().manual_seed(11)
    real, cams, inv = data.synthetic_real_batch(hand, B, V, S, gen)
    poses = data.random_poses(Ns, gen)
    torch.manual_seed(5)
    draws = dict(scales=torch.rand(Ns, 3, device=DEV) * 0.1 + 0.85, rand_f=torch.rand(Ns, device=DEV) * 0.2 + 0.9,
                 noise=torch.randn(3, Ns, S, S, device=DEV), eps=torch.randn(stacks, B * V, 32, device=DEV))

    def run(lo_b, hi_b, lo_s, hi_s, world, bucketed=False):
        net = create_hourglass_network(82, stacks).to(DEV)
        net.load_state_dict(w1)
        calls = []
        st = SelfSupTrainStep(net, hand, blob, hi_b - lo_b, V, hi_s - lo_s, S, lr=0.0, use_graph=False, world_size=world,
                              allreduce=lambda t: calls.append(t.numel()), bucketed=bucketed)
        st.load_batch(real[lo_b:hi_b], cams[lo_b:hi_b], inv[lo_b:hi_b], poses[lo_s:hi_s])
        st.scales.copy_(draws['scales'][lo_s:hi_s]); st.rand_f.copy_(draws['rand_f'][lo_s:hi_s])
        st.noise.copy_(draws['noise'][:, lo_s:hi_s]); st.vae_eps.copy_(draws['eps'][:, lo_b * V:hi_b * V])
        terms = st.step(is_mv=True).clone()
        return net._flat_grad.clone(), terms, calls, net
    g_all, t_all, calls, net = run(0, B, 0, Ns, 1)
    assert calls == []
    g_again, t_again, _, _ = run(0, B, 0, Ns, 1)         # the run-to-run noise floor of the same step (fp32 atomics order)
    g0, t0, c0, _ = run(0, B // 2, 0, Ns // 2, 2, bucketed=True)
    g1, t1, c1, _ = run(B // 2, B, Ns // 2, Ns, 2, bucketed=False)
    assert sum(c0) == g0.numel() and len(c0) >= 3          # bucketed: hg.0 under the trunk's backward, then the two end slices
    assert c1 == [g1.numel()]                              # default: one all-reduce of the whole flat gradient
    gs, ts = g0 + g1, (t0 + t1).cpu().numpy()
    ta = t_all.cpu().numpy()
    tn = t_again.cpu().numpy()
    for k, a, b, c in zip(TERM_NAMES, ts, ta, tn):
        # the same global step run twice already differs by up to ~1 % per term (measured: tools/diag_noise.py -- the order of the
        # fp32 statistics atomics changes single bf16 roundings and the deep network amplifies them); the sharded sum may differ
        # from the global step by no more than a few times that
        assert abs(a - b) <= 3 * abs(c - b) + (5e-2 if k in ('collision', 'bone_length') else 1e-2) * abs(b) + 1e-6, (k, a, b, c)
    # same kernels on the same images; what differs is the order of the fp32 atomics (GroupNorm statistics, weight-gradient sums),
    # which flips single bf16 roundings downstream (the run-to-run noise of same_terms)
    err = float((gs - g_all).double().norm() / g_all.double().norm())
    per = max(float((gs[o:o + n] - g_all[o:o + n]).double().norm() / g_all[o:o + n].double().norm().clamp_min(1e-20))
              for p in net.parameters() for o, n in (net._offsets[id(p)],) if float(g_all[o:o + n].norm()) > 1e-6)
    floor = float((g_again - g_all).double().norm() / g_all.double().norm())
    print('two shards vs global batch: whole-gradient l2 %.2e, worst tensor %.2e; the same global step twice: %.2e' % (err, per, floor))
    assert err < 2.5 * floor + 1e-3


def test_step_lr_matches_torch_scheduler(hand_model):
    """StepLR of the fused optimiser == torch.optim.lr_scheduler.StepLR (engine.py:98-99: step_size = epochs // 3, gamma 0.1); the rate
    reaches the device-side buffer the captured graph reads."""
    step, _, _, _ = make_step(hand_model, 1, 3, 1, 64, 1, use_graph=False, lr=1e-3)
    sched = step.step_lr(75 // 3, 0.1)
    p = torch.nn.Parameter(torch.zeros(1))
    ref_opt = torch.optim.Adam([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.StepLR(ref_opt, step_size=75 // 3, gamma=0.1)
    for epoch in range(75):
        ref_opt.step()
        ref.step()
        sched.step()
        assert abs(step.lr - ref_opt.param_groups[0]['lr']) <= 1e-12 * 1e-3, epoch
        assert abs(float(step.lr_dev.item()) - step.lr) <= 1e-7 * step.lr
    assert abs(step.lr - 1e-6) < 1e-12
