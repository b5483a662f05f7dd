"""Data parallelism of the train step: one process per GPU, batch sharded over ranks, ONE all-reduce(SUM) of the flat
fp32 gradient per step.  The reference has no distributed code at all; this is the only collective.

Host-side rules, independent of the device (tested with the gloo backend on CPU, tests/test_parallel_gloo.py):

  * every multi-view tuple and every synthetic pose is an independent unit (GroupNorm is per sample, no loss couples
    samples), so ranks get equal contiguous shards and no activation ever crosses ranks;
  * the step's loss mixes reductions: batch-MEAN terms (synt_uv, synt_d, mv_projection, mv_consistency, uv_hm_mean,
    bone_length, the VAE reconstruction MSE) and batch-SUM terms (collision, render.py:176; the VAE KLD, pose_vae.py:61).
    With equal shards, mean(global) = (1/R) sum_r mean(shard_r) and sum(global) = sum_r sum(shard_r): each rank scales the
    gradient of its MEAN terms by 1/R and leaves SUM terms alone; the all-reduce(SUM) then yields exactly the gradient a
    single GPU would compute on the global batch.
"""
import os

import torch
import torch.distributed as dist

MEAN_TERMS = ('synt_uv', 'synt_d', 'mv_projection', 'mv_consistency', 'uv_hm_mean', 'bone_length', 'pose_prior_recon')
SUM_TERMS = ('collision', 'pose_prior_kld')


def term_gradient_scale(name, world_size):
    """Factor a rank applies to the gradient (and value) of loss term `name` before the SUM all-reduce."""
    if name in MEAN_TERMS:
        return 1.0 / world_size
    if name in SUM_TERMS:
        return 1.0
    raise KeyError('unknown loss term %r (MEAN: %s; SUM: %s)' % (name, MEAN_TERMS, SUM_TERMS))


def mean_scale(world_size):
    return term_gradient_scale('mv_projection', world_size)


def shard_bounds(n, rank, world_size):
    """[lo, hi) of rank's shard of n units.  Shards must be equal: unequal shards would turn the mean of shard means into
    a weighted mean and break parity with the single-GPU step."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('bad rank %d / world size %d' % (rank, world_size))
    if n % world_size:
        raise ValueError('%d units do not split evenly over %d ranks' % (n, world_size))
    per = n // world_size
    return rank * per, (rank + 1) * per


_TUPLE_KEYS = ('real', 'cams', 'inv_cams')       # [B, V, ...]: shard over tuples
_POSE_KEYS = ('poses', 'scales', 'rand_f')       # [Ns, ...]: shard over synthetic poses


def shard_batch(batch, rank, world_size):
    """Global batch dict -> this rank's shard.  Keys: real [B,V,S,S], cams/inv_cams [B,V,4,4], poses [Ns,26], and the
    pre-drawn randoms scales [Ns,3], rand_f [Ns], noise [3,Ns,S,S], eps [stacks,B*V,32]."""
    out = {}
    B, V = batch['real'].shape[:2]
    lo_b, hi_b = shard_bounds(B, rank, world_size)
    lo_s, hi_s = shard_bounds(batch['poses'].shape[0], rank, world_size)
    for k, v in batch.items():
        if k in _TUPLE_KEYS:
            out[k] = v[lo_b:hi_b]
        elif k in _POSE_KEYS:
            out[k] = v[lo_s:hi_s]
        elif k == 'noise':
            out[k] = v[:, lo_s:hi_s]
        elif k == 'eps':
            out[k] = v[:, lo_b * V:hi_b * V]
        else:
            out[k] = v
    return out


def allreduce_gradients(flat_grad, world_size, group=None):
    """The step's single collective: SUM over ranks of the flat fp32 gradient, in place (NCCL over NVLink 5 / NVSwitch on the
    GPU box; 16.8 MB for the 2-stack network)."""
    if world_size > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def allreduce_terms(terms, world_size, group=None):
    """Per-rank weighted loss terms (already scaled per term_gradient_scale) -> global-batch values on every rank."""
    if world_size > 1:
        dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=group)
    return terms


def init_from_env(backend=None):
    """torchrun / torch.distributed.run environment -> (rank, world_size, local_rank); initialises the default group when
    WORLD_SIZE > 1 (nccl on CUDA, gloo otherwise; rendezvous on 127.0.0.1 unless MASTER_ADDR is set)."""
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kw['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank
