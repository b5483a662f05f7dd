"""The N>1 path on CPU: two gloo ranks, each computing the gradient of its shard with the data-parallel reduction rules of
spherehand_b200.parallel (the oracle stands in for the kernels as the gradient engine), all-reduced with
parallel.allreduce_gradients, against the single-process gradient on the global batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _global_batch():
    f = dict(np.load(os.path.join(GOLD, 'full_step_small.npz')))
    t = lambda k: torch.from_numpy(f[k])
    B, V, S = 2, 3, 32
    g = torch.Generator().manual_seed(5)
    # a smaller problem than the fixture (32x32, 1 stack) so that the CPU suite stays fast
    real = torch.nn.functional.interpolate(t('real'), size=(S, S), mode='nearest')
    images_synt = torch.nn.functional.interpolate(t('synt_dms')[:, None], size=(S, S), mode='nearest')[:, 0]
    uv_t = torch.rand(2, 41, S // 4, S // 4, generator=g)
    return dict(real=real, cams=t('cams'), inv_cams=t('inv_cams'), synt=images_synt, uv_t=uv_t, xyz_t=t('xyz_pts'),
                eps=torch.randn(1, B * V, 32, generator=g))


def _grad(batch, world, rank=None):
    """Flat gradient of the (rank-scaled) loss on `batch` with the oracle network + heads."""
    sys.path.insert(0, ROOT)
    from oracle import full_step as ofs, hourglass as oh
    from spherehand_b200 import parallel
    hm = dict(np.load(os.path.join(GOLD, 'hand_model.npz')))
    vae = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, 'pose_vae.npz')).items()}
    radii = torch.from_numpy(hm['keypoint_radius'].astype(np.float32))
    sd = {k: v.clone().requires_grad_(True) for k, v in oh.det_state_dict(82, 1, seed=7).items()}
    B, V = batch['real'].shape[:2]
    Ns = batch['synt'].shape[0]
    S = batch['real'].shape[-1]
    images = torch.cat([batch['synt'], batch['real'].reshape(B * V, S, S) * 0.01], 0)
    scale = (lambda name: parallel.term_gradient_scale(name, world)) if world > 1 else None
    terms, _, _ = ofs.loss_terms(sd, 1, images, Ns, B, V, batch['real'], batch['cams'], batch['inv_cams'], batch['uv_t'],
                                 batch['xyz_t'], radii, vae, batch['eps'], term_scale=scale)
    total = sum(v for v in terms.values() if torch.is_tensor(v))
    total.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in sd.values()])
    return flat, torch.stack([torch.as_tensor(float(v)) for v in terms.values()])


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    torch.set_num_threads(2)
    from spherehand_b200 import parallel
    r, w, _ = parallel.init_from_env('gloo')
    assert (r, w) == (rank, world)
    g = _global_batch()
    shard = parallel.shard_batch(dict(real=g['real'], cams=g['cams'], inv_cams=g['inv_cams'], poses=torch.zeros(2, 26), eps=g['eps']), rank, world)
    lo, hi = parallel.shard_bounds(2, rank, world)
    shard.update(synt=g['synt'][lo:hi], uv_t=g['uv_t'][lo:hi], xyz_t=g['xyz_t'][lo:hi])
    flat, terms = _grad(shard, world, rank)
    parallel.allreduce_gradients(flat, world)
    parallel.allreduce_terms(terms, world)
    if rank == 0:
        torch.save(dict(flat=flat, terms=terms), out)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gradient_equals_global_batch_gradient(tmp_path):
    out = str(tmp_path / 'dp.pt')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.set_num_threads(4)
    ref_flat, ref_terms = _grad(_global_batch(), 1)
    err = (got['flat'] - ref_flat).abs().max() / ref_flat.abs().max()
    assert err < 1e-4, float(err)
    assert torch.allclose(got['terms'], ref_terms, rtol=1e-4, atol=1e-5), (got['terms'], ref_terms)


def test_reduction_rules_and_sharding():
    sys.path.insert(0, ROOT)
    from spherehand_b200 import parallel
    assert parallel.term_gradient_scale('collision', 8) == 1.0 and parallel.term_gradient_scale('pose_prior_kld', 8) == 1.0
    assert parallel.term_gradient_scale('mv_projection', 8) == 0.125 and parallel.mean_scale(4) == 0.25
    with pytest.raises(KeyError):
        parallel.term_gradient_scale('nonsense', 2)
    assert parallel.shard_bounds(512, 3, 8) == (192, 256)
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 0, 4)
    with pytest.raises(ValueError):
        parallel.shard_bounds(8, 4, 4)
    b = dict(real=torch.zeros(4, 3, 8, 8), cams=torch.zeros(4, 3, 4, 4), inv_cams=torch.zeros(4, 3, 4, 4), poses=torch.arange(8.)[:, None].repeat(1, 26),
             scales=torch.zeros(8, 3), rand_f=torch.zeros(8), noise=torch.zeros(3, 8, 8, 8), eps=torch.arange(24.)[None, :, None].repeat(2, 1, 32))
    s = parallel.shard_batch(b, 1, 2)
    assert s['real'].shape[0] == 2 and s['poses'][:, 0].tolist() == [4., 5., 6., 7.] and s['noise'].shape == (3, 4, 8, 8)
    assert s['eps'].shape == (2, 6, 32) and s['eps'][0, :, 0].tolist() == [6., 7., 8., 9., 10., 11.]
    t = torch.ones(3)
    assert parallel.allreduce_gradients(t, 1) is t                        # world 1: no collective, no process group needed
