#!/usr/bin/env python
"""Benchmark of the hot path: BASELINE.json's metric — train-step images/sec (128x128 depth, self-sup loss).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the body of Engine._epoch_with_both (network/engine.py:349-376) over one batch of synthetic
input: per GPU B=64 multi-view tuples x V=3 views + Ns=64 synthetic poses = 256 images through the 2-stack hourglass at
128x128 (BASELINE.json configs[3]; under N GPUs the global batch is N x that, configs[4], weak scaling).
value = images/s over all ranks with inputs resident in HBM; e2e = the same through the public step API with the batch
copied from pinned host memory and the loss terms read back every step.  The K-step timed block is repeated REPEATS times
(`repeats`: every block's ms/step, min, max); value / e2e are the MEDIAN block.  Prints ONE JSON line (rank 0).

Beside the headline (rank 0, N=1 only; `--legs 0` skips them):
  roofline / roofline_other   every kernel family north_star names, from an eager step instrumented per C-ABI call (convolution
                              classes, weight gradients, GroupNorm fwd / bwd, pool / up-sample / add, MutualProjectionLoss) and
                              from CUDA-graph replays over buffer rings larger than L2 (sphere renderer R2 fwd / bwd at BASELINE
                              config 2 and in situ, triangle rasteriser R1 at the pybind boundary)
  gpu_reference               the UNMODIFIED reference's own GPU path (oracle/ref_gpu.py --mode stock: eager PyTorch + its CUDA
                              rasteriser, staged under oracle/_ref/ by oracle/build_ref.py) on the same box, same step, same
                              batch shape: strict fp32 (TF32 off), torch's default flags (cuDNN TF32 convolutions), and its
                              native 64x64 / 1-stack size; vs_reference_gpu = value / gpu_reference.value
  e2e_dropin                  the reference's loop body over `spherehand_b200.install()`-ed modules (module-by-module autograd path)
  cpu_baseline                the CPU port of the step (oracle/full_step.py) on two bounded samples (8 and 32 images)
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
REPEATS = 5
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
    """The reference's own CPU path of the step = the oracle port (the reference cannot be pip-installed: DESIGN.md §5), all host
    threads, each step a bounded sample (2 tuples x 3 views + 2 synthetic poses = 8 images) of the workload."""
    if rank != 0:
        return
    base, dt = cpu_arm(max(args.steps, 1), args.warmup, tuples=2, synt=2)
    line = dict(impl='reference', metric=METRIC, value=base['value'], unit='images/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', config=dict(workload=WORKLOAD, sample=base['sample']), cpu_baseline=base,
                e2e=dict(value=base['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
def call_work(name, a):
    """(flops, algorithmic HBM bytes, family) of one C-ABI call (per-unit figures: SURVEY.md §8d); family '' =
    not a roofline family.  `a` are the call's arguments in the order of include/spherehand_b200.h."""
    if name == 'sh_conv_fwd':      # (x,w,bias,res,N,H,W,Cin,Cout,cout_pad,taps,y,y_ld,y_nchw,stats,groups,stream)
        px = float(a[4] * a[5] * a[6])
        flops = 2.0 * px * a[7] * a[8] * a[10]
        byts = px * (2 * a[7] + (2 * a[12] if a[11] else 0) + (2 * a[8] if a[3] else 0) + (4 * a[8] if a[13] else 0))
        return flops, byts, 'conv3x3' if a[10] == 9 else 'conv1x1'
    if name == 'sh_conv_fwd_gn':   # (x,gn_stats,gamma,beta,G,eps,w,bias,res,N,H,W,Cin,Cout,cout_pad,y,y_ld,y_nchw,stats,groups,stream)
        px = float(a[9] * a[10] * a[11])
        byts = px * (2 * a[12] + (2 * a[16] if a[15] else 0) + (2 * a[13] if a[8] else 0) + (4 * a[13] if a[17] else 0))
        return 2.0 * px * a[12] * a[13], byts, 'conv1x1'
    if name == 'sh_conv_wgrad_gn':  # (dy,x,gn_stats,gamma,beta,G,eps,N,H,W,x_C,Cin,dy_C,Cout,dw,stream)
        px = float(a[7] * a[8] * a[9])
        return 2.0 * px * a[11] * a[13], px * 2 * (a[10] + a[12]), 'wgrad1x1'
    if name == 'sh_conv_wgrad':    # (dy,x,N,H,W,x_C,Cin,dy_C,Cout,taps,dw,stream)
        px = float(a[2] * a[3] * a[4])
        return 2.0 * px * a[6] * a[8] * a[9], px * 2 * (a[5] + a[7]), 'wgrad3x3' if a[9] == 9 else 'wgrad1x1'
    if name == 'sh_conv_wgrad3x3':  # (dy,x,N,H,W,x_C,Cin,dy_C,Cout,scratch,stream)
        px = float(a[2] * a[3] * a[4])
        return 2.0 * px * a[6] * a[8] * 9, px * 2 * (a[5] + a[7]), 'wgrad3x3'
    if name == 'sh_gn_relu_fwd':   # (x,stats_in,gamma,beta,N,HW,C,G,eps,y,stats_out,G_out,stream): x in, y out
        return 0.0, 4.0 * a[4] * a[5] * a[6], 'gn_relu_fwd'
    if name in ('sh_gn_relu_bwd', 'sh_gn_relu_bwd_prezeroed'):   # (da,x,stats,gamma,beta,addend,N,HW,C,...): da, x (, addend) in, dx out
        return 0.0, 2.0 * a[6] * a[7] * a[8] * (3 + (1 if a[5] else 0)), 'gn_relu_bwd'
    if name == 'sh_mvproj_loss_fwdbwd':   # (cam,inv,joints,real,radii,B,V,J,H,W,is_mv,...): 14 HW + 36 J + 128 B per view pair
        pairs = a[5] * a[6] * (a[6] if a[10] else 1)
        return 0.0, float(pairs) * (14.0 * a[8] * a[9] + 36.0 * a[7] + 128.0), 'mvproj'
    if name == 'sh_maxpool_fwd':   # (x,N,H,W,C,y,...): H, W of the pooled map
        return 0.0, 2.0 * a[1] * a[2] * a[3] * a[4] * 5, 'pool_upsample_add'
    if name == 'sh_maxpool_bwd':   # (dy,x,addend,N,H,W,C,dx,...): dy + x (4x) + addend (4x) in, dx (4x) out
        return 0.0, 2.0 * a[3] * a[4] * a[5] * a[6] * (1 + 4 + (4 if a[2] else 0) + 4), 'pool_upsample_add'
    if name == 'sh_upsample_add_fwd':   # (up1,low,N,h,w,C,y,...): up1 (4x) + low in, y (4x) out
        return 0.0, 2.0 * a[2] * a[3] * a[4] * a[5] * 9, 'pool_upsample_add'
    if name == 'sh_upsample_bwd':  # (dy,N,h,w,C,dlow,...)
        return 0.0, 2.0 * a[1] * a[2] * a[3] * a[4] * 5, 'pool_upsample_add'
    if name == 'sh_add':           # (a,b,c,N,HW,C,y,...)
        return 0.0, 2.0 * a[3] * a[4] * a[5] * (3 + (1 if a[2] else 0)), 'pool_upsample_add'
    return 0.0, 0.0, ''


FAMILY_KERNELS = {
    'conv1x1': ('conv_fwd_kernel (tcgen05 implicit-GEMM convolution, forward + data-gradient launches; incl. the launches that apply the '
                'preceding GroupNorm + ReLU to their operand tiles in shared memory), 1x1 layers', 'hbm'),
    'conv3x3': ('conv_fwd_kernel (tcgen05 implicit-GEMM convolution, forward + data-gradient launches), 3x3 layers', 'tensor'),
    'wgrad1x1': ('wgrad1x1_kernel / wgrad_kernel (tcgen05 weight gradient; incl. the fused-GroupNorm-operand launches), 1x1 layers', 'hbm'),
    'wgrad3x3': ('wgrad3x3_kernel / wgrad_kernel (tcgen05 weight gradient), 3x3 layers', 'tensor'),
    'gn_relu_fwd': ('gn_relu_fwd_kernel (GroupNorm + ReLU apply where it is not folded into the consuming 1x1 layer; statistics come from '
                    'the producer)', 'hbm'),
    'gn_relu_bwd': ('gn_relu_bwd_kernel (GroupNorm + ReLU backward, one pass, fused residual add + bias column sum)', 'hbm'),
    'pool_upsample_add': ('maxpool / upsample+add / add kernels, forward and backward', 'hbm'),
    'mvproj': ('mvproj_main_kernel (MutualProjectionLoss fwd+bwd: transform, sphere render, both loss terms, analytic gradient)', 'hbm'),
}


def graph_us(fn, calls, reps=REPEATS):
    """Median device time (us) of one call of fn(i): `calls` calls captured in a CUDA graph (no host launch cost between them),
    replayed `reps` times, CUDA events on the launching stream."""
    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(calls):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(calls):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / calls)
    return statistics.median(out), min(out), max(out)


def renderer_rooflines(hand, pk, dev):
    """R2 (sphere renderer fwd / bwd) at BASELINE config 2 (N=256, J=48, 128^2, seed 1234) and at the in-situ shape of the step
    (N = B*V^2 = 576, J=41), R1 (triangle rasteriser) at the pybind boundary (B=64 meshes -> 640^2): device time per launch from
    CUDA-graph replays, each call on the next set of a buffer ring larger than the 126 MB L2."""
    import torch
    from spherehand_b200 import data, ops
    rows = []
    g = torch.Generator().manual_seed(1234)
    for Nn, Jj, what in ((256, 48, 'BASELINE config 2'), (B * V * V, J, 'in-situ shape of the train step')):
        ring = max(2, int(200e6 // (Nn * S * S * 5)) + 1)
        c = torch.cat([torch.rand(Nn, Jj, 2, generator=g) * 180 - 90, torch.rand(Nn, Jj, 1, generator=g) * 120 - 60], -1).to(dev)
        r = torch.cat([hand.radii.cpu(), torch.full((Jj - hand.radii.numel(),), 20.0)]).to(dev) if Jj > hand.radii.numel() else hand.radii
        sph = ops.pack_spheres(c, r)
        depth = [torch.empty((Nn, S, S), device=dev) for _ in range(ring)]
        idx = [torch.empty((Nn, S, S), device=dev, dtype=torch.uint8) for _ in range(ring)]
        gs = torch.empty_like(sph)
        st = torch.cuda.current_stream

        def fwd(i):
            _lib_call('sh_sphere_render_fwd', sph.data_ptr(), Nn, Jj, S, S, depth[i % ring].data_ptr(), idx[i % ring].data_ptr(), st().cuda_stream)
        for i in range(ring):
            fwd(i)
        gd = [torch.randn((Nn, S, S), device=dev) * (idx[i] != 255) for i in range(ring)]

        def bwd(i):
            _lib_call('sh_sphere_render_bwd', gd[i % ring].data_ptr(), idx[i % ring].data_ptr(), sph.data_ptr(), Nn, Jj, S, S, gs.data_ptr(), st().cuda_stream)
        fb, bb = Nn * (16 * Jj + 5 * S * S), Nn * (5 * S * S + 28 * Jj)
        for name, fn, byts in (('sphere_render_fwd_kernel', fwd, fb), ('sphere_render_bwd_kernel', bwd, bb)):
            us, lo, hi = graph_us(fn, 4 * ring)
            rows.append(dict(kernel='%s (R2), N=%d J=%d 128x128: %s' % (name, Nn, Jj, what), bound='hbm', achieved=byts / us * 1e-3, peak=pk['hbm'],
                             unit='GB/s', frac=byts / us * 1e-3 / pk['hbm'], avg_launch_us=us, min_us=lo, max_us=hi, algorithmic_bytes_per_launch=byts,
                             timing='CUDA-graph replay of %d launches over a ring of %d buffer sets (> L2)' % (4 * ring, ring)))
        del depth, idx, gd
    poses = data.random_poses(B, g, dev)
    mats = ops.fk_fwd(poses, hand.offset_mats, hand.inv_offset_mats)
    pts = ops.lbs_fwd(mats, *hand.mesh_csr, right_hand=True, mode=2, cam=(320.0, 320.0, 640 / 300, 640 / 300))
    fv = ops.gather_faces(pts, hand.faces)
    F = fv.shape[1]
    outs = [torch.empty((B, 640, 640), device=dev) for _ in range(2)]
    us, lo, hi = graph_us(lambda i: _lib_call('sh_tri_raster_fwd', fv.data_ptr(), B, F, 640, 640, outs[i % 2].data_ptr(), torch.cuda.current_stream().cuda_stream), 4)
    byts = B * (36 * F + 4 * 640 * 640)
    rows.append(dict(kernel='tri_raster_kernel (R1) at the pybind boundary: B=%d meshes x %d faces -> 640x640' % (B, F), bound='hbm',
                     achieved=byts / us * 1e-3, peak=pk['hbm'], unit='GB/s', frac=byts / us * 1e-3 / pk['hbm'], avg_launch_us=us, min_us=lo, max_us=hi,
                     algorithmic_bytes_per_launch=byts, timing='CUDA-graph replay of 4 launches over 2 output buffers (210 MB)'))
    us, lo, hi = graph_us(lambda i: ops.tri_raster_lattice_fwd(fv, 640, 5, 2, 1), 8)
    rows.append(dict(kernel='tri_raster_kernel on the 640->128 resize lattice (what the train step runs), B=%d' % B, bound='latency/atomics',
                     avg_launch_us=us, min_us=lo, max_us=hi, us_per_mesh=us / B))
    return rows


def _lib_call(name, *args):
    from spherehand_b200 import _lib
    return _lib.call(name, *args)


def harness(*argv, timeout=900):
    """oracle/ref_gpu.py in a subprocess (its own import world) -> its JSON line, or {'unavailable': why}."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_gpu.py'), *argv], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {'unavailable': 'timed out after %d s' % timeout}
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    if not lines:
        return {'unavailable': (p.stderr.strip().splitlines() or ['no output'])[-1][:300]}
    return json.loads(lines[-1])


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
    step = SelfSupTrainStep(net, hand, blob, B, V, NS, S, lr=1e-4, world_size=world, real_aug=bool(args.real_aug), bucketed=bool(args.bucketed))
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
        if args.sample_poses:
            step.sample_poses()                            # JointAngleDataset.__getitem__ as one batched kernel (engine.py:326-329)
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

    warm = max(args.warmup, 3)
    for _ in range(warm):
        resident_step()
    sampler = ClockSampler(local_rank)
    blocks = [timed(resident_step, args.steps) for _ in range(REPEATS)]      # REPEATS blocks of exactly K steps each
    clocks = sampler.stop()
    ms = statistics.median(blocks)
    step.prefetch_batch(*host)
    for _ in range(2):
        e2e_step()
    blocks_e2e = [timed(e2e_step, args.steps) for _ in range(REPEATS)]
    ms_e2e = statistics.median(blocks_e2e)
    terms = step.terms.cpu().tolist()
    launches = step.launches_per_step * args.steps

    ips = lambda t: world * IMAGES_PER_STEP * args.steps / (t * 1e-3)
    line = dict(metric=METRIC, value=ips(ms), unit='images/s', n_gpus=world,
                steps=args.steps, warmup=warm, ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='bf16', data='synthetic',
                config=dict(workload=WORKLOAD, global_images_per_step=world * IMAGES_PER_STEP, tuples_per_s=world * B * args.steps / (ms * 1e-3),
                            parallelism='dp%d' % world, gradient_allreduce=('bucketed under the backward pass' if step.bucketed else 'one call after the backward pass') if world > 1 else 'none', precision='bf16 operands / fp32 accumulate in the hourglass, fp32 everywhere else',
                            l2='not flushed: per-step activation working set (>= 10 GB) >> 126 MB L2',
                            real_aug='on: the reference\'s scale augmentation of the real views (create_network_and_criterion.py:94-102) inside the step, '
                                     'drawn per step (50 % of the steps resize)' if args.real_aug else 'off',
                            sample_poses='on-device JointAngleDataset sampler every step' if args.sample_poses else 'fixed poses resident in HBM',
                            loss_total=terms[8]),
                repeats=dict(n=REPEATS, block_steps=args.steps, value='median block', ms_per_step=[b / args.steps for b in blocks],
                             images_per_s_min=ips(max(blocks)), images_per_s_max=ips(min(blocks)),
                             e2e_ms_per_step=[b / args.steps for b in blocks_e2e]),
                clocks=clocks, gpu_launches=launches,
                e2e=dict(value=ips(ms_e2e), unit='images/s', h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=36, ms_per_step=ms_e2e / args.steps, min=ips(max(blocks_e2e)), max=ips(min(blocks_e2e))))

    # ---- kernel shares + rooflines: eager (un-graphed) steps with CUDA events around every C-ABI call on the launching stream
    #      (the graph replays above cannot be instrumented per kernel).  EVERY rank runs these steps (they contain the gradient
    #      all-reduce, a collective); only rank 0 records and reports.
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
            flops, byts, kind = call_work(name, a)
            if kind:
                c = cls.setdefault(kind, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
                c['ms'] += ms_call; c['calls'] += 1; c['flops'] += flops; c['bytes'] += byts
        total_ms = sum(d['ms'] for d in fam.values())
        shares = {k: round(d['ms'] / total_ms, 4) for k, d in sorted(fam.items(), key=lambda kv: -kv[1]['ms'])[:10]}

        def roof(kind):
            c = cls.get(kind)
            if not c:
                return None
            kern, bound = FAMILY_KERNELS[kind]
            if bound == 'hbm':
                ach, peak, unit, src = c['bytes'] / (c['ms'] * 1e-3) / 1e9, pk['hbm'], 'GB/s', pk['src'] + ' hbm_gbs (copy)'
            else:
                ach, peak, unit, src = c['flops'] / (c['ms'] * 1e-3) / 1e12, pk['tf_sust'], 'TFLOP/s', pk['src'] + ' bf16_tflops_sustained'
            r = dict(kernel=kern, bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak, peak_source=src,
                     traffic=None, avg_launch_us=c['ms'] * 1e3 / c['calls'], launches_per_step=c['calls'] // reps,
                     step_share=round(c['ms'] / total_ms, 4), algorithmic_bytes_per_launch=c['bytes'] / c['calls'],
                     flops_per_launch=c['flops'] / c['calls'])
            if bound == 'tensor':
                r['frac_of_burst'] = ach / pk['tf_burst']
            return r

        # headline = the dominant KERNEL, conv_fwd_kernel (the tcgen05 implicit-GEMM convolution: ~1/3 of the step over its forward and
        # data-gradient launches), reported on the roofline class that holds the larger share of the step; every other named
        # kernel family beside it, largest share first
        cands = [r for r in (roof(k) for k in FAMILY_KERNELS) if r]
        cands.sort(key=lambda r: (not r['kernel'].startswith('conv_fwd_kernel'), -r['step_share']))
        cands = cands[:1] + sorted(cands[1:], key=lambda r: -r['step_share'])
        line['roofline'] = cands[0]
        line['roofline']['step_ms_eager_sum'] = total_ms / reps
        line['roofline']['shares'] = shares
        tpath = os.path.join(ROOT, 'profiles', 'kernel_traffic.json')
        traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
        for r in cands:                    # dram__bytes_read+write per launch of the same family from the committed ncu capture
            key = next((k for k, (kern, _) in FAMILY_KERNELS.items() if kern == r['kernel']), None)
            if key in traffic:
                r['traffic'] = traffic[key]['dram_bytes_per_launch']
                r['traffic_source'] = traffic[key]['source']
        line['roofline_other'] = cands[1:]
        if world == 1 and args.legs:
            try:
                line['roofline_other'] += renderer_rooflines(hand, pk, dev)
            except Exception as e:             # a measurement leg must not take the headline down with it
                line['roofline_other'].append(dict(kernel='renderer micro-benchmarks', error=repr(e)[:300]))
    if world == 1 and args.legs and rank == 0:
        del step
        torch.cuda.empty_cache()
        ref = harness('--mode', 'stock', '--S', str(S), '--stacks', str(STACKS), '--B', str(B), '--Ns', str(NS), '--steps', '5', '--warmup', '2', '--tf32', 'off')
        ref['tf32_default_flags'] = harness('--mode', 'stock', '--S', str(S), '--stacks', str(STACKS), '--B', str(B), '--Ns', str(NS), '--steps', '5', '--warmup', '2',
                                            '--tf32', 'default')
        ref['with_per_step_sync'] = harness('--mode', 'stock', '--S', str(S), '--stacks', str(STACKS), '--B', str(B), '--Ns', str(NS), '--steps', '5', '--warmup', '2',
                                            '--tf32', 'off', '--sync', '1')
        ref['native_64x64_1stack'] = harness('--mode', 'stock', '--S', '64', '--stacks', '1', '--B', str(B), '--Ns', str(NS), '--steps', '10', '--warmup', '3', '--tf32', 'off')
        line['gpu_reference'] = ref
        if 'value' in ref:
            best = max(ref['value'], ref['tf32_default_flags'].get('value', 0.0))
            line['vs_reference_gpu'] = dict(value=line['value'] / ref['value'], e2e=line['e2e']['value'] / ref['value'],
                                            vs_its_fastest_setting=line['e2e']['value'] / best,
                                            note='images/s of this arm / images/s of the unmodified reference on the same GPU, same step, same batch shape '
                                                 '(north_star target: >= 10)')
        line['e2e_dropin'] = harness('--mode', 'dropin', '--S', str(S), '--stacks', str(STACKS), '--B', str(B), '--Ns', str(NS), '--steps', '10', '--warmup', '3')
        base, _ = cpu_arm(1, 1, tuples=2, synt=2)
        big, _ = cpu_arm(1, 0, tuples=8, synt=8)
        base['larger_sample'] = dict(value=big['value'], sample=big['sample'])
        line['cpu_baseline'] = base
    if rank == 0:
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
    ap.add_argument('--legs', type=int, default=1, help='0: skip the reference-GPU / drop-in / renderer / CPU legs (N=1 only)')
    ap.add_argument('--real_aug', type=int, default=1, help="the reference's scale augmentation of the real views inside the timed step (Engine's default: "
                    'HeatmapEstimationNetwork(real_aug=True)); 0 = off')
    ap.add_argument('--bucketed', type=int, default=0, help='1: all-reduce the gradient in buckets underneath the backward pass instead of once after it (N > 1; measured slower)')
    ap.add_argument('--sample_poses', type=int, default=0, help='1: draw the synthetic poses on the device every step')
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
