"""GPU parity tests of the drop-in nn.Module / autograd.Function surfaces (spherehand_b200.mesh, spherehand_b200.network,
spherehand_b200.depth_rasterization): same constructor arguments and forward signatures as the reference, results against
the golden fixtures generated from the reference and against the CPU oracle, gradients through torch autograd."""
import numpy as np
import pytest
import torch

from conftest import golden, mesh_dict, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'

if torch.cuda.is_available():
    import spherehand_b200
    from spherehand_b200.mesh import kinematicsTransformation as kt
    from spherehand_b200.mesh import multiview_utility as mv
    from spherehand_b200.mesh import pointTransformation as pt
    from spherehand_b200.mesh import render
    from spherehand_b200.network import create_network_and_criterion as cnc
    from spherehand_b200.network import pose_vae, util_modules
from oracle import full_step as ofs, hourglass as oh, losses, sphere, synth


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to(DEV).to(dtype).contiguous()


def test_ball_render_and_hand_primitive_render(hand_model):
    g = golden('sphere_render_64')
    c, r = g['centres'][..., :3], g['radii']
    N, J = c.shape[:2]
    br = render.BallRender(64, 64).to(DEV)
    assert set(br.state_dict()) == {'dist_weight', 'x_grid', 'y_grid'}          # the reference's keys (render.py:16-21)
    centres = cu(c.reshape(N * J, 3)).requires_grad_(True)
    radii = cu(np.tile(r, N)).requires_grad_(True)
    balls = br(centres, radii)                                                   # [N*J,64,64], one sphere per image
    ob = sphere.ball_render(c.reshape(N * J, 3), np.tile(r, N), 64, 64)
    assert np.array_equal(balls.detach().cpu().numpy(), ob)
    # min over spheres == the fused renderer == the golden reference output
    assert np.abs(balls.detach().view(N, J, 64, 64).min(dim=1).values.cpu().numpy() - g['depth']).max() <= 8e-6
    # autograd through the per-sphere maps (what HandBallPrimitiveRender / MutualProjection do in the reference)
    gd = cu(g['grad_depth'])
    depth, idx = balls.view(N, J, 64, 64).min(dim=1)
    (depth * gd).sum().backward()
    assert rel_err(centres.grad.view(N, J, 3).cpu(), g['grad_centres']) < 1e-4
    _, orad = sphere.sphere_render_backward(g['grad_depth'], sphere.sphere_render(c, r, 64, 64)[1], c, r, 64, 64)
    assert rel_err(radii.grad.view(N, J).cpu(), orad) < 1e-4
    # HandBallPrimitiveRender: FK matrices -> (part maps, depth)
    mesh = mesh_dict(hand_model)
    hb = render.HandBallPrimitiveRender(mesh['bones'], 64, 64).to(DEV)
    fk = golden('fk')
    part, dep = hb(cu(fk['mats'][:2]))
    assert part.shape == (2, 41, 64, 64) and dep.shape == (2, 64, 64)
    assert torch.equal(part.min(dim=1).values, dep)
    t = ofs.HandTables(hand_model)
    oc = synth.lbs(torch.from_numpy(fk['mats'][:2]), t.kp)[..., :3].numpy()
    od, _ = sphere.sphere_render(oc, hand_model['keypoint_radius'].astype(np.float32), 64, 64)
    assert np.abs(dep.cpu().numpy() - od).max() < 1e-3 and ((dep.cpu().numpy() < 100) != (od < 100)).mean() < 1e-3


def test_loss_modules_against_golden_and_oracle():
    g = golden('mv_losses_32')
    t = {k: cu(v) for k, v in g.items() if v.dtype == np.float32 and v.ndim > 0}
    radii = [float(x) for x in g['radii']]
    mpl = mv.MutualProjectionLoss(32, radii).to(DEV)
    for is_mv in (True, False):
        joints = t['joints'].clone().requires_grad_(True)
        loss, proj = mpl(t['cams'], t['inv_cams'], joints, t['real'], is_mv)
        (loss * 2.0).backward()
        assert rel_err(loss.item(), g['loss_mv%d' % is_mv]) < 1e-4
        assert rel_err(joints.grad.cpu() / 2.0, g['grad_mv%d' % is_mv]) < 1e-4
        assert rel_err(proj.cpu(), g['projected_dms']) < 1e-4 and not proj.requires_grad
    # MutualProjection alone (depth images + projected points)
    dimg, pts = mpl.mutual_projection(t['cams'], t['inv_cams'], t['joints'])
    assert rel_err(dimg.cpu(), g['projected_dms']) < 1e-4 and pts.shape == (2, 3, 3, 41, 3, 1)
    # DataToModelLoss on the 6 (b, j) images with view-0... the fixture pairs real[b,j] with joints[b,j]
    d2m = render.DataToModelLoss(32, 32, radii).to(DEV)
    j2 = t['joints'].reshape(6, 41, 3).clone().requires_grad_(True)
    l = d2m(t['real'].reshape(6, 32, 32), j2)
    l.backward()
    assert rel_err(l.item(), g['d2m']) < 1e-4 and rel_err(j2.grad.cpu(), g['d2m_grad']) < 1e-4
    # consistency / collision / bone length
    for mod, key, scale in ((mv.MultiviewConsistencyLoss(), 'cons', 1.0), (render.CollisionLoss(), 'col', 0.5)):
        j = (t['joints'] * scale).clone().requires_grad_(True)
        args = (t['cams'], j) if key == 'cons' else (j,)
        l = mod.to(DEV)(*args)
        l.backward()
        assert rel_err(l.item(), g[key]) < 1e-4 and rel_err(j.grad.cpu() * scale, g[key + '_grad']) < 1e-4
    bl = render.BoneLengthLoss().to(DEV)
    assert set(bl.state_dict()) == {'joint_1', 'joint_2', 'max_length', 'min_length'}
    ja, jb = (t['joints'] * 0.7).clone().requires_grad_(True), (t['joints'] * 1.3).clone().requires_grad_(True)
    la, lb = bl(ja), bl(jb)
    (la + lb).backward()
    assert rel_err(la.item() + lb.item(), g['bone']) < 1e-4
    assert rel_err((ja.grad * 0.7 + jb.grad * 1.3).cpu(), g['bone_grad']) < 1e-4
    # [B,41,3] input (one view) and the reference's "view 0 only" quirk on [B,V,41,3]
    j1 = t['joints'][:, 0].clone().contiguous()
    assert abs(render.CollisionLoss().to(DEV)(j1 * 0.5).item() - g['col']) < 1e-4 * abs(g['col'])
    with pytest.raises(NotImplementedError):
        mv.MultiviewConsistencyLoss()(t['cams'], t['joints'], torch.ones(2, 3, 41, 1, device=DEV))
    with pytest.raises(TypeError):
        render.DataToModelLoss(32, 32, 'not a mesh')


def test_synth_modules_against_golden_and_oracle(hand_model):
    mesh = mesh_dict(hand_model)
    fk = golden('fk')
    htm = kt.HandTransformationMat([b['offset_matrix'].astype(np.float32) for b in mesh['bones']]).to(DEV)
    assert rel_err(htm(cu(fk['params'])).cpu(), fk['mats']) < 1e-5
    geo = golden('synth_geometry')
    rs = pt.RandScale(0.1).to(DEV)
    scaled = rs(htm(cu(geo['params'])), torch.from_numpy(geo['scales']))
    assert rel_err(scaled.cpu(), geo['mats']) < 1e-5
    torch.manual_seed(3)
    drawn = rs.draw(5)
    torch.manual_seed(3)
    ref = torch.stack([torch.rand(5) * 0.1 + 0.90 - 0.05 for _ in range(3)], dim=1)       # pointTransformation.py:140-142
    assert torch.equal(drawn, ref)
    # DepthRender (LBS + projection + rasteriser + resize) against the oracle pipeline, 64 and 128
    faces_before = mesh['faces'].copy()
    t = ofs.HandTables(hand_model)
    for S in (64, 128):
        dr = render.DepthRender(mesh, S).to(DEV)
        dm = dr(cu(geo['mats'][:3]), cu(geo['rand_f'][:3]))
        od, _ = synth.depth_render(torch.from_numpy(geo['mats'][:3]), t.mesh, t.faces, S, torch.from_numpy(geo['rand_f'][:3]), fma=True)
        assert dm.shape == (3, S, S)
        # vertices differ in the last ulp between the CUDA and the CPU skinning: a few silhouette samples flip
        assert (np.abs(dm.cpu().numpy() - od.numpy()) > 1e-3).mean() < 1e-2
    assert np.array_equal(mesh['faces'], faces_before)        # unlike the reference, construction does not mutate the caller's faces
    # generic size: full 640^2 rasterisation + bilinear resize, and the drop-in pybind entry
    dr = render.DepthRender(mesh, 80).to(DEV)
    assert dr(cu(geo['mats'][:1]), cu(geo['rand_f'][:1])).shape == (1, 80, 80)
    import spherehand_b200.depth_rasterization as drz
    z = drz.forward(32, 16, torch.zeros((2, 0, 3, 3), device=DEV))
    assert z.shape == (2, 16, 32) and (z == 1000).all()
    with pytest.raises(RuntimeError):
        drz.forward(16, 16, torch.zeros((1, 4, 3, 3)))
    f = render.DepthRasterizationFunction.apply(16, 16, torch.zeros((1, 0, 3, 3), device=DEV))
    assert (f == 100).all()                                                       # clamp(max=100) (render.py:286)
    # heat-map targets
    for hm in (16, 32):
        g = golden('synth_heatmaps_%d' % hm)
        h3 = render.Hand3DHeatmapRender(mesh['bones'], hm).to(DEV)
        uv, d, xyz = h3(cu(g['mats']), cu(g['rand_f']))
        assert rel_err(uv[:3].cpu(), g['uv_hms']) < 1e-4 and rel_err(xyz.cpu(), g['xyz']) < 1e-4
    # projection modules on their own
    p = cu(geo['verts_head'][:2])
    cam = pt.OthographicalProjection(8.0, 8.0, 16 / 300, 16 / 300).to(DEV)
    inv = pt.InverseOthographicalProjection(8.0, 8.0, 16 / 300, 16 / 300).to(DEV)
    q = cam(p)
    ref = torch.matmul(cam.k_mat, p.view(-1, 4, 1)).view(2, -1, 4)
    assert rel_err(q.cpu(), ref.cpu()) < 1e-6 and rel_err(inv(q).cpu(), p.cpu()) < 1e-5
    rf = cu([0.93, 1.07])
    q1 = cam(p, rf)
    assert rel_err(q1[..., 0].cpu(), (p[..., 0] * rf[:, None] * cam.fx + cam.cx).cpu()) < 1e-6 and (q1[..., 3] == 1).all()
    # DepthNoise with injected draws, and HandSynthesizer end to end (noise off: its draws are made on the device)
    gn = golden('depth_noise')
    dn = util_modules.DepthNoise(64, 64).to(DEV)
    out = dn(cu(gn['dm']), torch.stack([cu(gn['nx']), cu(gn['ny']), cu(gn['nz'])]))
    assert np.abs(out.cpu().numpy() - gn['out']).max() < 1e-6
    hs = util_modules.HandSynthesizer(mesh, 64, 16, 1.0, 0.01, add_noise=False).to(DEV)
    torch.manual_seed(11)
    dm, uv, dh, xyz = hs(cu(geo['params'][:4]))
    torch.manual_seed(11)
    scales = torch.stack([torch.rand(4) * 0.1 + 0.85 for _ in range(3)], dim=1)
    rand_f = torch.rand(4) * 0.2 + 0.9
    mats = synth.rand_scale_apply(synth.forward_kinematics(torch.from_numpy(geo['params'][:4]), t.offset_mats), scales)
    od, _ = synth.depth_render(mats, t.mesh, t.faces, 64, rand_f, fma=True)
    ouv, _, oxyz = synth.hand_heatmaps(mats, t.kp, 16, rand_f)
    assert (np.abs(dm.cpu().numpy() - od.numpy() * 0.01) > 1e-5).mean() < 1e-2
    assert rel_err(uv.cpu(), ouv) < 1e-4 and rel_err(xyz.cpu(), oxyz) < 1e-4
    assert not dm.requires_grad
    with pytest.raises(NotImplementedError):
        htm(cu(fk['params']).requires_grad_(True))                                # forward-only, loudly


def test_resize_crop_image_against_golden_and_oracle():
    """ResizeCropImage (util_modules.py:383-424) as one batched kernel: bit-exact against the reference fixture and, on
    random scales in the range MultiTaskLoss draws (create_network_and_criterion.py:99-101) and beyond, against the oracle."""
    g = golden('resize_crop')
    rc = util_modules.ResizeCropImage().to(DEV)
    for S in (64, 128):
        out = rc(cu(g['dm%d' % S]), cu(g['u%d' % S]), cu(g['v%d' % S]))
        assert np.array_equal(out.cpu().numpy(), g['out%d' % S])
    gen = torch.Generator().manual_seed(5)
    for (n, H, W) in ((192, 128, 128), (7, 48, 80), (1, 64, 64)):
        dm = torch.rand(n, H, W, generator=gen)
        u = torch.rand(n, generator=gen) * 0.7 + 0.5
        v = torch.rand(n, generator=gen) * 0.7 + 0.5
        out = rc(dm.to(DEV), u.to(DEV), v.to(DEV))
        assert torch.equal(out.cpu(), synth.resize_crop(dm, u, v))
    assert rc(torch.zeros(0, 8, 8, device=DEV), torch.zeros(0, device=DEV), torch.zeros(0, device=DEV)).numel() == 0
    with pytest.raises(RuntimeError):
        rc(torch.rand(2, 8, 8), torch.ones(2), torch.ones(2))                      # host tensors: no CPU fallback


def test_joint_angle_dataset_against_reference_fixture():
    """The batched pose sampler: same generator state -> the reference's poses bit for bit (256 consecutive `__getitem__`
    results of the unmodified reference), generator left where the reference leaves it; independent-stream mode against the
    oracle restatement; the poses drive the FK kernel unchanged."""
    from spherehand_b200.dataset import joint_angle as ja
    from oracle import poses as oposes
    g = golden('joint_angle')
    n = g['poses'].shape[0]
    ds = ja.JointAngleDataset()
    assert len(ds) == 400000 and ds.num_parameter == 26
    torch.manual_seed(int(g['seed']))
    p = ds.sample_batch(n)
    assert p.is_cuda and np.array_equal(p.cpu().numpy(), g['poses'])
    assert np.array_equal(torch.rand(4).numpy(), g['after'])
    torch.manual_seed(int(g['seed']))
    assert np.array_equal(ds[0].cpu().numpy(), g['poses'][0]) and np.array_equal(ds[1].cpu().numpy(), g['poses'][1])
    gen = torch.Generator().manual_seed(5)
    q = ds.sample_batch(1000, generator=gen, sequential=False).cpu().numpy()
    u = torch.rand(1000 * ja.MAX_UNIFORMS, generator=torch.Generator().manual_seed(5)).numpy()
    for i in (0, 1, 499, 999):
        assert np.array_equal(q[i], oposes.joint_angle_getitem(oposes._Stream(u, i * ja.MAX_UNIFORMS)))
    assert np.isfinite(q).all() and q[:, 1].max() <= 0 and q[:, 1].min() >= -3.1400001
    assert ds.sample_batch(0).shape == (0, 26)


def test_shard_batch_loader_prefetch():
    """ShardBatchLoader over the committed shards: device batches equal direct indexing (sequential and seeded-shuffle order),
    staging buffers are pinned, batches stay valid while the next one is prefetched, several epochs reuse the buffers."""
    import os
    from spherehand_b200.dataset import nyu_dataset as nd
    ds = nd.create_nyu_dataset(os.path.join(os.path.dirname(__file__), 'golden', 'shard'))
    ld = nd.ShardBatchLoader(ds, 3, device=DEV, shuffle=False)
    assert len(ld) == 2 and all(h.is_pinned() for slot in ld._host for h in slot)
    for epoch in range(2):
        seen = []
        for b, (dm, jp, cp, icp) in enumerate(ld):
            assert dm.is_cuda and dm.shape == (3, 3, 16, 16) and cp.shape == (3, 3, 4, 4)
            seen.append((dm, jp, cp, icp))                       # keep the previous batch alive across the next request
            for k in range(3):
                want = ds[b * 3 + k]
                for got, w in zip((dm, jp, cp, icp), want):
                    assert np.array_equal(got[k].cpu().numpy(), np.asarray(w, np.float32))
        assert len(seen) == 2
        assert np.array_equal(seen[0][0][0].cpu().numpy(), np.asarray(ds[0][0]))       # batch 0 still intact after batch 1 arrived
    gen = torch.Generator().manual_seed(3)
    order = torch.randperm(len(ds), generator=torch.Generator().manual_seed(3)).tolist()
    for b, (dm, jp, cp, icp) in enumerate(nd.ShardBatchLoader(ds, 4, device=DEV, shuffle=True, generator=gen)):
        for k in range(4):
            assert np.array_equal(icp[k].cpu().numpy(), ds[order[b * 4 + k]][3].astype(np.float32))
    with pytest.raises(ValueError):
        nd.ShardBatchLoader(ds, 9, device=DEV)
    # an ODD number of batches per epoch (here one) with a consumer whose reads are still queued when the next epoch starts
    # (the host runs ahead of the stream under graph replay): the new epoch's copies must wait for them
    gen = torch.Generator().manual_seed(8)
    ref_gen = torch.Generator().manual_seed(8)
    ld1 = nd.ShardBatchLoader(ds, 5, device=DEV, shuffle=True, generator=gen)
    got, want = [], []
    for epoch in range(4):
        order = torch.randperm(len(ds), generator=ref_gen).tolist()[:5]
        for dm, jp, cp, icp in ld1:
            torch.cuda._sleep(30_000_000)                      # ~15 ms of queued GPU work in front of the consumer's read
            got.append(dm.clone())
            want.append(np.stack([np.asarray(ds[i][0], np.float32) for i in order]))
    torch.cuda.synchronize()
    for g_, w_ in zip(got, want):
        assert np.array_equal(g_.cpu().numpy(), w_)


def test_pose_denoiser_against_golden_and_oracle():
    """PoseDenoiser eval forward as one kernel: the reference's output on the fixture (3-D and 2-D inputs), the oracle on a
    larger random batch with a ragged row count, untouched coordinates copied bit for bit."""
    from spherehand_b200.network import pose_denoiser as pd
    g = golden('pose_denoiser')
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd.')}
    net = pd.PoseDenoiser().to(DEV).eval()
    net.load_state_dict(sd)
    out3 = net(cu(g['joints']))
    assert out3.shape == (7, 41, 3) and rel_err(out3.cpu(), g['out3']) < 1e-5
    assert rel_err(net(cu(g['joints']).reshape(7, -1)[:3].contiguous()).cpu(), g['out2']) < 1e-5
    keep = np.setdiff1d(np.arange(123), g['output_indices'])
    assert np.array_equal(out3.reshape(7, -1).cpu().numpy()[:, keep], g['joints'].reshape(7, -1)[:, keep])
    x = torch.randn(1001, 123, generator=torch.Generator().manual_seed(9)) * 50
    want = losses.pose_denoiser_forward(sd, x, torch.from_numpy(g['input_indices']), torch.from_numpy(g['output_indices']))
    assert rel_err(net(x.to(DEV)).cpu(), want) < 1e-5
    assert net(torch.zeros(0, 123, device=DEV)).shape == (0, 123)


def test_network_heads_and_full_criterion(hand_model):
    # soft-argmax head through autograd
    g = golden('softargmax')
    score = torch.cat([cu(g['uv']), cu(g['d'])], dim=1).contiguous().requires_grad_(True)
    rec = util_modules.RecoverXYZCoordinateFromHeatmap(16, 16, 0.01).to(DEV)
    xyz = rec(score[:, :41], score[:, 41:])
    (xyz * cu(g['gxyz'])).sum().backward()
    assert rel_err(xyz.detach().cpu(), g['xyz']) < 1e-4
    assert rel_err(score.grad[:, :41].cpu(), g['guv']) < 1e-4 and rel_err(score.grad[:, 41:].cpu(), g['gd']) < 1e-4
    uv2, d2 = cu(g['uv']).requires_grad_(True), cu(g['d']).requires_grad_(True)      # separate (non-adjacent) tensors
    (rec(uv2, d2) * cu(g['gxyz'])).sum().backward()
    assert rel_err(uv2.grad.cpu(), g['guv']) < 1e-4 and rel_err(d2.grad.cpu(), g['gd']) < 1e-4
    # VAE prior with the reference's layer names
    vae = pose_vae.PoseVae(123, 32)
    w = {k: torch.from_numpy(v) for k, v in golden('pose_vae').items()}
    assert set(vae.state_dict()) == set(w)
    vae.load_state_dict(w)
    vae.to(DEV)
    gp = golden('vae_prior')
    x = cu(gp['x']).requires_grad_(True)
    l = vae.prior_loss(x, cu(gp['eps']))
    l.backward()
    assert rel_err(l.item(), gp['loss']) < 1e-4 and rel_err(x.grad.cpu(), gp['grad']) < 1e-4
    assert torch.isfinite(vae.prior_loss(x.detach())).item()                          # eps drawn internally
    # whole criterion, module by module, against the reference's own numbers (fixture generated from the reference)
    f = golden('full_step_small')

    class Constant:
        mesh = mesh_dict(hand_model)
    net = cnc.HeatmapEstimationNetwork(16, 0.01, 41, 1, real_aug=False).to(DEV)
    net.hg.load_state_dict(oh.det_state_dict(82, 1, seed=7))
    crit = cnc.MultiTaskLoss(True, True, True, False, True, True, True, Constant(), image_size=64, heatmap_size=16,
                             pose_vae_path=None).to(DEV)
    crit.prior_loss.load_state_dict(w)
    crit.prior_loss.to(DEV)
    assert crit.weights['synt_hm'] == 1e3 and crit.weights['domain'] == 0.0
    eps = cu(f['eps'])
    crit.prior_loss.prior_loss = (lambda fn: (lambda x: fn(x, eps)))(crit.prior_loss.prior_loss)      # inject the draw
    result = net(real_dms=cu(f['real']) * 0.01, synt_dms=cu(f['synt_dms']))
    assert set(result) >= {'synt_uv_hms', 'synt_d_hms', 'synt_xyz', 'real_uv_hms', 'real_d_hms', 'real_xyz', 'batch_synt_fea', 'batch_real_fea'}
    assert result['real_xyz'][0].shape == (2, 3, 41, 3) and result['real_uv_hms'][0].shape == (2, 3, 41, 16, 16)
    terms, proj = crit(result, {'uv_hms': cu(f['uv_hms']), 'xyz_pts': cu(f['xyz_pts'])},
                       {'camera_poses': cu(f['cams']), 'inv_camera_poses': cu(f['inv_cams']), 'real_dms': cu(f['real']), 'is_mv': True})
    assert proj[0].shape == (2, 3, 3, 64, 64)
    for k in ('synt_uv', 'synt_d', 'mv_projection', 'mv_consistency', 'uv_hm_mean', 'pose_prior', 'domain_loss'):
        ref = float(f['term.' + k])
        assert abs(float(terms[k].detach()) - ref) <= 5e-2 * abs(ref) + 1e-3, (k, float(terms[k]), ref)      # bf16 hourglass contract
    # the hinge terms (sums of relu over a few active pairs) amplify the bf16 noise of the network's joints: check the
    # criterion itself on OUR joints against the oracle (1e-4), and the reference's value only loosely
    jx = result['real_xyz'][0].detach().cpu()
    assert rel_err(float(terms['collision'].detach()), float(losses.collision_loss(jx))) < 1e-4
    assert rel_err(float(terms['bone_length'].detach()), float(losses.bone_length_loss(jx))) < 1e-4
    # random-weight heat-maps are flat, so softmax(20 hm) turns a one-ulp bf16 difference in a few activations into a visible
    # move of single joints: bound the typical deviation tightly and the worst joint loosely
    # (so only the TYPICAL deviation is asserted here; the per-joint bound -- every joint within 1e-2 of the coordinate range -- is
    # held on the reference's trained weights in tests/test_gpu_trained.py, where the heat-maps are peaked)
    dj = np.abs(jx.numpy() - f['real_xyz'])
    assert dj.mean() < 0.03 * np.abs(f['real_xyz']).max()
    total = cnc.combine_loss(terms)
    total.backward()
    gw = net.hg.score[0].weight.grad
    ref_gw = torch.from_numpy(f['grad.score.0.weight'])
    cos = torch.nn.functional.cosine_similarity(gw.cpu().flatten(), ref_gw.flatten(), dim=0).item()
    assert cos > 0.97, cos      # direction sanity only: bf16 network + hinge / arg-min terms (network gradients: test_gpu_nn.py)
    # the criterion's gradient wiring in isolation: same joints on both sides, d(total)/d(real_xyz) against CPU autograd
    leaf = result['real_xyz'][0].detach().clone().requires_grad_(True)
    r2 = {'real_xyz': [leaf], 'real_uv_hms': [result['real_uv_hms'][0].detach()]}
    rt = {'camera_poses': cu(f['cams']), 'inv_camera_poses': cu(f['inv_cams']), 'real_dms': cu(f['real']), 'is_mv': True}
    t2, _ = crit(r2, None, rt)
    cnc.combine_loss(t2).backward()
    jo = jx.clone().requires_grad_(True)
    cams, inv, real = (torch.from_numpy(f[k]) for k in ('cams', 'inv_cams', 'real'))
    lo = (losses.mutual_projection_loss(cams, inv, jo, real, torch.from_numpy(hand_model['keypoint_radius'].astype(np.float32)), True)[0]
          + 1e-3 * losses.multiview_consistency(cams, jo) + 1e-2 * losses.vae_prior_loss((jo / 100.0).reshape(6, 123), w, torch.from_numpy(f['eps']))
          + losses.collision_loss(jo) + losses.bone_length_loss(jo))
    lo.backward()
    # (untrained-network joints lie far from the observed hand: a handful of silhouette / arg-min knife-edge pixels differ
    #  between torch-CPU and the kernel; the per-head tests above hold 1e-4 on the reference's fixtures)
    assert rel_err(leaf.grad.cpu(), jo.grad) < 1e-3
    assert rel_err(float(sum(v for k, v in t2.items() if k != 'uv_hm_mean').detach()), float(lo.detach())) < 1e-4
    # ---- scale augmentation (real_aug=True, reference :41-62): same draws in the same order as the reference, resized views
    # from the batched ResizeCropImage kernel, joints divided by the scales
    aug = cnc.HeatmapEstimationNetwork(16, 0.01, 41, 1, real_aug=True).to(DEV)
    aug.hg.load_state_dict(oh.det_state_dict(82, 1, seed=7))
    assert set(aug.state_dict()) == set(net.state_dict())                               # ResizeCropImage holds no state
    real = cu(f['real']) * 0.01
    n_img = real.shape[0] * real.shape[1]
    seed = next(sd for sd in range(100) if torch.manual_seed(sd) and torch.rand(1).item() >= 0.5)
    torch.manual_seed(seed)
    out_aug = aug.train()(real_dms=real)
    torch.manual_seed(seed)
    assert torch.rand(1).item() >= 0.5
    rnd = torch.rand(n_img).to(DEV) * 0.2 + 0.75
    u = rnd + torch.rand_like(rnd) * 0.1 - 0.05
    v = rnd + torch.rand_like(rnd) * 0.1 - 0.05
    want = synth.resize_crop(real.reshape(n_img, 64, 64).cpu(), u.cpu(), v.cpu())
    assert torch.equal(out_aug['real_resized_dms'].cpu(), want)
    out_ref = aug.eval()(real_dms=want.to(DEV).reshape(real.shape))                      # eval: no augmentation, no division
    inv = torch.stack([1 / u, 1 / v, torch.ones_like(u)], dim=-1).view(real.shape[0], real.shape[1], 1, 3)
    # two runs of the random-weight network differ by single-ulp bf16 flips that the flat heat-maps amplify (see above), so the
    # end-to-end comparison is loose (without the division the mean deviation would be ~20 %); the division itself is exact
    dj = (out_aug['real_xyz'][0].detach() - out_ref['real_xyz'][0].detach() * inv).abs().cpu().numpy()
    ref_max = float(out_ref['real_xyz'][0].detach().abs().max())
    assert dj.mean() < 0.03 * ref_max
    p3 = torch.randn(n_img, 41, 3, device=DEV)
    assert torch.allclose(aug._unscale([p3], u, v)[0], torch.stack([p3[..., 0] / u[:, None], p3[..., 1] / v[:, None], p3[..., 2]], -1), rtol=1e-6)
    assert aug._unscale([p3], None, None)[0] is p3
    torch.manual_seed(next(sd for sd in range(100) if torch.manual_seed(sd) and torch.rand(1).item() < 0.5))
    assert torch.equal(aug.train()(real_dms=real)['real_resized_dms'], real.reshape(n_img, 64, 64))   # the 'no augmentation' draw


def test_install_registers_reference_import_names():
    import sys
    saved = {k: sys.modules.get(k) for k in ('depth_rasterization', 'mesh', 'mesh.render', 'mesh.cuda_kernel', 'network',
                                              'network.hourglass', 'mesh.multiview_utility')}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        spherehand_b200.install()
        import depth_rasterization
        from mesh.cuda_kernel import depth_rasterization as dr2
        from mesh.render import BallRender, DepthRender  # noqa: F401
        from network.hourglass import create_hourglass_network  # noqa: F401
        assert depth_rasterization is dr2 and depth_rasterization.forward is spherehand_b200.depth_rasterization.forward
    finally:
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
