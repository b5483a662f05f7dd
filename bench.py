#!/usr/bin/env python
"""Benchmark of the hot path: BASELINE.json's metric — train-step images/sec (128x128 depth, self-sup loss).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the body of Engine._epoch_with_both (network/engine.py:349-376) over one batch of synthetic
input: per GPU B=64 multi-view tuples x V=3 views + Ns=64 synthetic poses = 256 images through the 2-stack hourglass at
128x128 (BASELINE.json configs[3]; under N GPUs the global batch is N x that, configs[4], weak scaling).
value = images/s over all ranks with inputs resident in HBM; e2e = the same through the public step API with the batch
copied from pinned host memory and the loss terms read back every step.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')

B, V, NS, S, STACKS, J = 64, 3, 64, 128, 2, 41
IMAGES_PER_STEP = NS + B * V
METRIC = 'train-step images/sec (128x128 depth, self-sup loss)'
WORKLOAD = 'full self-supervised train step: B=64 tuples x V=3 views + Ns=64 synthetic poses per GPU (256 images), 128x128 depth, 2-stack hourglass, J=41, all loss heads, Adam'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if val.strip() == 'Active':
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_arm(steps, warmup, tuples=1, synt=1, threads=None):
    """The oracle port of the whole step (oracle/full_step.py: torch-fp32 + C rasteriser on the host) on a bounded sample."""
    import numpy as np
    import torch
    from oracle import full_step as ofs, hourglass as oh
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    hm = dict(np.load(os.path.join(GOLD, 'hand_model.npz')))
    tables = ofs.HandTables(hm)
    vae = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, 'pose_vae.npz')).items()}
    g = torch.Generator().manual_seed(0)
    sd = {k: v.clone().requires_grad_(True) for k, v in oh.det_state_dict(82, STACKS, seed=7).items()}
    # inputs: sphere-rendered "real" views of random poses (oracle renderer), random cameras
    from oracle import losses as ol, synth as osy
    poses = torch.zeros(tuples, 26)
    poses[:, :3] = torch.rand(tuples, 3, generator=g) - 0.5
    mats = osy.forward_kinematics(poses, tables.offset_mats)
    centres = osy.lbs(mats, tables.kp)[..., :3]
    cams = torch.eye(4).repeat(tuples, V, 1, 1)
    inv = cams.clone()
    real = ol.ball_depth(centres[:, None].expand(tuples, V, J, 3), tables.radii, S).min(dim=2).values
    batch = dict(real=real, cams=cams, inv_cams=inv, poses=torch.zeros(synt, 26), scales=torch.full((synt, 3), 0.9),
                 rand_f=torch.ones(synt), noise=torch.randn(3, synt, S, S, generator=g),
                 eps=torch.randn(STACKS, tuples * V, 32, generator=g))
    state = {}
    for _ in range(warmup):
        ofs.train_step(sd, STACKS, tables, vae, batch, S, opt_state=state)
    t0 = time.perf_counter()
    for _ in range(steps):
        ofs.train_step(sd, STACKS, tables, vae, batch, S, opt_state=state)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    imgs = synt + tuples * V
    return dict(value=imgs / dt, unit='images/s', cores=cores, kind='port',
                sample='%d step(s) of %d tuple(s) x %d views + %d synthetic pose(s) (%d images) at 128x128, 2 stacks: '
                       'oracle/full_step.py (torch fp32 CPU + C rasteriser), %.2f s/step' % (steps, tuples, V, synt, imgs, dt)), dt


def run_reference(args, rank):
    if rank != 0:
        return
    base, dt = cpu_arm(max(args.steps, 1), min(args.warmup, 1))
    line = dict(impl='reference', metric=METRIC, value=base['value'], unit='images/s', n_gpus=args.gpus, steps=args.steps,
                warmup=min(args.warmup, 1), ms_per_step=dt * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', config=dict(workload=WORKLOAD, sample=base['sample']), cpu_baseline=base,
                e2e=dict(value=base['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
def conv_work(name, a):
    """(flops, algorithmic HBM bytes, class) of one C-ABI convolution call."""
    if name == 'sh_conv_fwd':      # (x,w,bias,res,N,H,W,Cin,Cout,cout_pad,taps,y,y_ld,y_nchw,stats,groups,stream)
        px = float(a[4] * a[5] * a[6])
        flops = 2.0 * px * a[7] * a[8] * a[10]
        byts = px * (2 * a[7] + (2 * a[12] if a[11] else 0) + (2 * a[8] if a[3] else 0) + (4 * a[8] if a[13] else 0))
        return flops, byts, '3x3' if a[10] == 9 else '1x1'
    if name == 'sh_conv_wgrad':    # (dy,x,N,H,W,x_C,Cin,dy_C,Cout,taps,dw,stream)
        px = float(a[2] * a[3] * a[4])
        return 2.0 * px * a[6] * a[8] * a[9], px * 2 * (a[5] + a[7]), '3x3' if a[9] == 9 else '1x1'
    if name == 'sh_conv_wgrad3x3':  # (dy,x,N,H,W,x_C,Cin,dy_C,Cout,scratch,stream)
        px = float(a[2] * a[3] * a[4])
        return 2.0 * px * a[6] * a[8] * 9, px * 2 * (a[5] + a[7]), '3x3'
    return 0.0, 0.0, ''


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from spherehand_b200 import _lib, data, ops
    from spherehand_b200.engine import SelfSupTrainStep, TERM_NAMES
    from spherehand_b200.model import HandModel
    from spherehand_b200.network.hourglass import create_hourglass_network

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    hand = HandModel.from_arrays(dict(np.load(os.path.join(GOLD, 'hand_model.npz'))), dev)
    vae_sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, 'pose_vae.npz')).items()}
    blob = ops.vae_blob_from_state_dict(vae_sd, dev)
    torch.manual_seed(0)                                   # identical replicas
    net = create_hourglass_network(2 * J, STACKS).to(dev)
    step = SelfSupTrainStep(net, hand, blob, B, V, NS, S, lr=1e-4, world_size=world)
    gen = torch.Generator().manual_seed(1234 + rank)       # each rank its own shard of the global batch
    real, cams, inv = data.synthetic_real_batch(hand, B, V, S, gen)
    poses = data.random_poses(NS, gen)
    host = [t.cpu().pin_memory() for t in (real, cams, inv, poses)]
    h2d = step.load_batch(*host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def resident_step():
        step.draw_randoms()
        step.step(is_mv=True)

    def e2e_step():
        # double-buffered input pipeline: this step's batch crossed the host link underneath the previous step (prefetch_batch,
        # copy stream); commit_batch moves it into the buffers the graph reads, then the NEXT batch's copies start.  One batch
        # (h2d bytes) crosses per step, inside the timed region.
        step.commit_batch()
        step.prefetch_batch(*host)
        step.draw_randoms()
        return step.step(is_mv=True).cpu()                 # D2H of the 9 loss terms: synchronises every step

    for _ in range(max(args.warmup, 3)):
        resident_step()
    sampler = ClockSampler(local_rank)
    ms = timed(resident_step, args.steps)
    clocks = sampler.stop()
    step.prefetch_batch(*host)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    terms = step.terms.cpu().tolist()
    launches = step.launches_per_step * args.steps

    line = dict(metric=METRIC, value=world * IMAGES_PER_STEP * args.steps / (ms * 1e-3), unit='images/s', n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='bf16', data='synthetic',
                config=dict(workload=WORKLOAD, global_images_per_step=world * IMAGES_PER_STEP, tuples_per_s=world * B * args.steps / (ms * 1e-3),
                            parallelism='dp%d' % world, precision='bf16 operands / fp32 accumulate in the hourglass, fp32 everywhere else',
                            l2='not flushed: per-step activation working set (>= 10 GB) >> 126 MB L2', real_aug='off',
                            loss_total=terms[8]),
                clocks=clocks, gpu_launches=launches,
                e2e=dict(value=world * IMAGES_PER_STEP * args.steps / (ms_e2e * 1e-3), unit='images/s', h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=36, ms_per_step=ms_e2e / args.steps))

    # ---- kernel shares + roofline of the dominant kernel family: eager (un-graphed) steps with CUDA events around every
    #      C-ABI call on the launching stream (the graph replays above cannot be instrumented per kernel).  EVERY rank runs
    #      these steps (they contain the gradient all-reduce, a collective); only rank 0 records and reports.
    step.use_graph = False
    reps = 3
    for _ in range(2):
        resident_step()
    torch.cuda.synchronize()
    if rank == 0:
        _lib.PROFILE = []
    for _ in range(reps):
        resident_step()
    torch.cuda.synchronize()
    step.use_graph = True
    if rank == 0:
        pk = peaks()
        prof, _lib.PROFILE = _lib.PROFILE, None
        fam, cls = {}, {}
        for name, a, e0, e1 in prof:
            ms_call = e0.elapsed_time(e1)
            d = fam.setdefault(name, dict(ms=0.0, calls=0))
            d['ms'] += ms_call
            d['calls'] += 1
            flops, byts, kind = conv_work(name, a)
            if kind:
                c = cls.setdefault(('sh_conv_wgrad' if name == 'sh_conv_wgrad3x3' else name, kind), dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
                c['ms'] += ms_call; c['calls'] += 1; c['flops'] += flops; c['bytes'] += byts
        total_ms = sum(d['ms'] for d in fam.values())
        shares = {k: round(d['ms'] / total_ms, 4) for k, d in sorted(fam.items(), key=lambda kv: -kv[1]['ms'])[:8]}

        def roof(name, kind):
            c = cls.get((name, kind))
            if not c:
                return None
            kern = {'sh_conv_fwd': 'conv_fwd_kernel (tcgen05 implicit-GEMM convolution, forward + data-gradient launches)',
                    'sh_conv_wgrad': 'wgrad kernels (tcgen05 weight gradient)'}[name]
            if kind == '1x1':      # 85 flop/B against a ridge of ~214 flop/B: HBM-bound
                ach, peak, unit, bound, src = c['bytes'] / (c['ms'] * 1e-3) / 1e9, pk['hbm'], 'GB/s', 'hbm', pk['src'] + ' hbm_gbs (copy)'
            else:
                ach, peak, unit, bound, src = c['flops'] / (c['ms'] * 1e-3) / 1e12, pk['tf_sust'], 'TFLOP/s', 'tensor', pk['src'] + ' bf16_tflops_sustained'
            return dict(kernel='%s, %s layers' % (kern, kind), bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak, peak_source=src,
                        traffic=None, avg_launch_us=c['ms'] * 1e3 / c['calls'], launches_per_step=c['calls'] // reps,
                        step_share=round(c['ms'] / total_ms, 4), algorithmic_bytes_per_launch=c['bytes'] / c['calls'],
                        flops_per_launch=c['flops'] / c['calls'])

        # dominant kernel = conv_fwd_kernel; its launches fall in two roofline classes -- the headline is the class with the
        # larger share of the step, the others are listed beside it
        cands = [r for r in (roof('sh_conv_fwd', '1x1'), roof('sh_conv_fwd', '3x3')) if r]
        cands.sort(key=lambda r: -r['step_share'])
        line['roofline'] = cands[0]
        line['roofline']['step_ms_eager_sum'] = total_ms / reps
        line['roofline']['shares'] = shares
        tpath = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
        if os.path.exists(tpath):          # dram__bytes_read+write per launch of the same class from the committed ncu capture
            tr = json.load(open(tpath))
            key = '1x1' if line['roofline']['bound'] == 'hbm' else '3x3'
            if key in tr:
                line['roofline']['traffic'] = tr[key]['dram_bytes_per_launch']
                line['roofline']['traffic_source'] = tr[key]['source']
        line['roofline_other'] = cands[1:] + [r for r in (roof('sh_conv_wgrad', '1x1'), roof('sh_conv_wgrad', '3x3')) if r]
        if world == 1:
            base, _ = cpu_arm(1, 1)
            line['cpu_baseline'] = base
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus == 1:
            world, rank, local_rank = 1, 0, 0
        else:
            raise SystemExit('--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)' % (args.gpus, args.gpus, world))
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
