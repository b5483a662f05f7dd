"""Data-parallel check on real GPUs (torchrun, one rank per GPU, NCCL): the gradient all-reduce between the backward pass and Adam
(default) and in buckets underneath the backward pass, eager and with the step replayed as CUDA-graph segments.  Every rank: identical replicas, its own shard; after a step (eager, then graph replay) the all-reduced flat gradient
must be bit-identical on every rank and equal to the sum of the ranks' local gradients (computed by a second replica whose
collective is the identity and summed with one plain all_reduce); weights stay identical across ranks over several steps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spherehand_b200 import data, ops, parallel                      # noqa: E402
from spherehand_b200.engine import SelfSupTrainStep, TERM_NAMES      # noqa: E402
from spherehand_b200.model import HandModel                          # noqa: E402
from spherehand_b200.network.hourglass import create_hourglass_network   # noqa: E402
from oracle.hourglass import two_stack_from_trained                   # noqa: E402  (weights recipe of the tests)

GOLD = os.path.join(ROOT, 'tests', 'golden')


def main():
    rank, world, local = parallel.init_from_env('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    hand = HandModel.from_arrays(dict(np.load(os.path.join(GOLD, 'hand_model.npz'))), dev)
    vae_sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, 'pose_vae.npz')).items()}
    blob = ops.vae_blob_from_state_dict(vae_sd, dev)
    w2 = two_stack_from_trained({k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, 'trained_weights.npz')).items()})
    B, V, Ns, S, stacks = 4, 3, 4, 64, 2
    gen = torch.Generator().manual_seed(100 + rank)
    ok = True

    def make(allreduce=None, use_graph=False, bucketed=False):
        net = create_hourglass_network(82, stacks).to(dev)
        net.load_state_dict(w2)
        st = SelfSupTrainStep(net, hand, blob, B, V, Ns, S, lr=1e-4, world_size=world, use_graph=use_graph, allreduce=allreduce,
                              bucketed=bucketed)
        return st

    real, cams, inv = data.synthetic_real_batch(hand, B, V, S, gen)
    poses = data.random_poses(Ns, gen)
    for mode in ('eager', 'graph', 'graph+buckets'):
        a = make(use_graph=(mode != 'eager'), bucketed=mode.endswith('buckets'))      # the real thing: NCCL all-reduce(s)
        b = make(allreduce=lambda t: t, use_graph=False)      # same step, collective = identity -> the local gradient
        for st in (a, b):
            st.load_batch(real, cams, inv, poses)
        for it in range(3):
            torch.manual_seed(7 + it + 10 * rank)
            a.draw_randoms()
            for name in ('scales', 'rand_f', 'noise', 'vae_eps'):
                getattr(b, name).copy_(getattr(a, name))
            b.net._flat.copy_(a.net._flat); b.adam_m.copy_(a.adam_m); b.adam_v.copy_(a.adam_v); b.step_dev.copy_(a.step_dev)
            a.step(is_mv=True)
            b.step(is_mv=True)
            torch.cuda.synchronize()
            ga = a.net._flat_grad.clone()
            want = b.net._flat_grad.clone()
            dist.all_reduce(want, op=dist.ReduceOp.SUM)
            # identical on every rank?
            ref = ga.clone()
            dist.broadcast(ref, src=0)
            same = bool(torch.equal(ref, ga))
            # a and b are two runs of the same local step (run-to-run noise of the fp32 atomics, amplified by the network): compare at
            # that level, and exactly for the reduction itself
            err = float((ga - want).double().norm() / want.double().norm())
            wa = a.net._flat.clone()
            dist.broadcast(wa, src=0)
            same_w = bool(torch.equal(wa, a.net._flat))
            terms = a.loss_dict(reduce=True)
            if rank == 0:
                print('%s step %d: reduced gradient identical on all ranks: %s; vs sum of local gradients l2 %.3e; weights identical: %s; '
                      'global loss %.4f' % (mode, it, same, err, same_w, terms['total']), flush=True)
            ok = ok and same and same_w and err < 0.25 and np.isfinite(terms['total'])
        if mode != 'eager' and rank == 0:
            print('launches per step', a.launches_per_step, flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    good = flag.item() == 1.0
    if rank == 0:
        print('DP CHECK', 'OK' if good else 'FAILED', flush=True)
    del a, b
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print('barrier passed', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if good else 1)


if __name__ == '__main__':
    main()
