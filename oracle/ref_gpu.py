#!/usr/bin/env python
"""Timing / parity harness for the reference's OWN GPU path.  TEST + MEASUREMENT INFRASTRUCTURE (never imported by the product).

Runs the body of `Engine._epoch_with_both` (network/engine.py:349-376: synthesise -> zero_grad -> network -> criterion -> sum ->
backward -> Adam.step) as a stand-alone loop over the reference's modules, constructed the way `Engine.__init__` constructs
them (engine.py:54-70, 95-97), in one of two import worlds:

    --mode stock    the UNMODIFIED reference: oracle/_ref/reference/ on sys.path, its CUDA rasteriser compiled by
                    oracle/build_ref.py registered as `depth_rasterization`.  Eager PyTorch fp32, TF32 off (torch defaults).
    --mode dropin   the same loop, same import statements, after `spherehand_b200.install(reference_root=...)`: every mirrored
                    module resolves to the B200 kernels, everything else (constants.py, ...) to the reference's files.

Import-time compatibility shims only (np.float, a matplotlib stub for mesh/bone_length.py, torch.load defaults for the
pickled assets, CWD = reference root for its relative asset paths).  Prints ONE JSON line.  bench.py runs it in a subprocess
(rank 0, N=1) for its `gpu_reference` / `e2e_dropin` keys; tests/test_gpu_dropin.py compares the two worlds' loss terms.
"""
import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_TREE = os.path.join(HERE, '_ref', 'reference')


def setup(mode):
    import numpy as np
    import torch
    if not os.path.isdir(os.path.join(REF_TREE, 'network')):
        raise SystemExit(json.dumps({'impl': 'gpu_reference', 'unavailable': 'oracle/_ref/reference not staged (oracle/build_ref.py)'}))
    if not hasattr(np, 'float'):
        np.float = float
    for name in ('matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    _load = torch.load

    def load(f, *a, **k):
        k.setdefault('weights_only', False)
        return _load(f, *a, **k)
    torch.load = load
    try:                                  # cv2 >= 4.5 rejects the float pixel coordinates network/util_vis.py:21 passes to cv2.circle
        import cv2
        _circle = cv2.circle
        cv2.circle = lambda img, c, *a, **k: _circle(img, (int(c[0]), int(c[1])), *a, **k)
    except ImportError:
        pass
    os.chdir(REF_TREE)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if mode == 'stock':
        sys.path.insert(0, REF_TREE)
        sys.path.insert(0, HERE)
        import build_ref
        ext = build_ref.load()
        if ext is None:
            raise SystemExit(json.dumps({'impl': 'gpu_reference', 'unavailable': 'oracle/_ref/depth_rasterization_ref.so missing'}))
        sys.modules['depth_rasterization'] = ext
    else:
        import spherehand_b200
        spherehand_b200.install(reference_root=REF_TREE)


def make_inputs(B, V, Ns, S, seed, dev):
    """Seeded inputs built with the (installed or stock) reference-named modules: sphere-rendered views of random poses as the
    'real' depth maps (mm, background 100.0), random camera rotations <= 30 degrees, JointAngleDataset poses for the synthetic branch."""
    import numpy as np
    import torch
    from network.constants import Constant
    from mesh.kinematicsTransformation import HandTransformationMat
    from mesh.render import HandBallPrimitiveRender
    from dataset.joint_angle import JointAngleDataset
    torch.manual_seed(seed)
    ds = JointAngleDataset()
    poses = torch.stack([ds[i] for i in range(B + Ns)]).float()
    g = torch.Generator().manual_seed(seed + 1)
    cams = torch.eye(4).repeat(B, V, 1, 1)
    axis = torch.randn(B, V, 3, generator=g)
    axis = axis / axis.norm(dim=-1, keepdim=True)
    ang = (torch.rand(B, V, generator=g) * 2 - 1) * np.pi / 6
    ang[:, 0] = 0
    K = torch.zeros(B, V, 3, 3)
    K[..., 0, 1], K[..., 0, 2] = -axis[..., 2], axis[..., 1]
    K[..., 1, 0], K[..., 1, 2] = axis[..., 2], -axis[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -axis[..., 1], axis[..., 0]
    cams[..., :3, :3] = torch.eye(3) + torch.sin(ang)[..., None, None] * K + (1 - torch.cos(ang))[..., None, None] * (K @ K)
    inv = torch.inverse(cams)
    offs = [b['offset_matrix'].astype(np.float32) for b in Constant.mesh['bones']]
    fk = HandTransformationMat(offs).to(dev)
    rnd = HandBallPrimitiveRender(Constant.mesh['bones'], S, S).to(dev)
    with torch.no_grad():
        mats = fk(poses[:B].to(dev))
        mv = (inv.to(dev)[:, :, None] @ mats[:, None]).reshape(B * V, 17, 4, 4).contiguous()
        chunks = [rnd(mv[i:i + 48])[1] for i in range(0, B * V, 48)]
        real = torch.cat(chunks).reshape(B, V, S, S).contiguous()
    return real, cams.to(dev), inv.to(dev), poses[B:].to(dev)


def run_engine(args):
    """The reference's UNMODIFIED `Engine` (network/engine.py:53-139) constructed from run_engine.py's options and one epoch of
    `_epoch_with_both` (engine.py:318-435: its own DataLoaders, RunningAverage, visualisation dump, log file) on a few generated
    64x64 shards in the writer's format (dataset/nyu_generator.py:100-118).  Prints the averaged loss terms the engine logged."""
    import pickle
    import re
    import tempfile
    import numpy as np
    import torch
    dev = torch.device('cuda', 0)
    tmp = tempfile.mkdtemp(prefix='sh_engine_')
    n, V, S = args.B, 3, 64
    real, cams, inv, _ = make_inputs(n, V, 1, S, args.seed + 10, dev)
    for split in ('train', 'test'):
        d = os.path.join(tmp, 'data', split)
        os.makedirs(d)
        dms = real.cpu().numpy().astype(np.float32)
        with open(os.path.join(d, 'mv_data_0_shape.pkl'), 'wb') as f:
            pickle.dump({'dms': dms.shape}, f)
        fp = np.memmap(os.path.join(d, 'mv_data_0_dms.bat'), dtype='float32', mode='w+', shape=dms.shape)
        fp[:] = dms
        fp.flush()
        del fp
        np.save(os.path.join(d, 'mv_data_0_joint_poses.npy'), np.zeros((n, V, 36, 3), np.float32))
        np.save(os.path.join(d, 'mv_data_0_camera_poses.npy'), cams.cpu().numpy().astype(np.float32))
    opts = argparse.Namespace(synthesize=True, mv_projection=True, mv_consistency=True, temporal=False, collision=True, bone_length=True,
                              prior=True, mode='Train', model_dir=os.path.join(tmp, 'models'), initial_model='pretrained/synthetic.pth',
                              restore_from_model=None, restore_from_epoch=-1, num_stacks=1, epoch=3, dataset_dir=os.path.join(tmp, 'data'),
                              depth_resample=0, lr=args.lr, tag='dropin')
    from network.engine import Engine, Mode
    torch.manual_seed(args.seed + 20)
    engine = Engine(opts)
    t0 = time.perf_counter()
    engine._epoch_with_both(Mode.Train, 0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    line = open(engine.log_file).read().strip().splitlines()[0]
    terms = {k: float(v) for k, v in re.findall(r'(\w+): (-?\d+\.\d+)', line.split('loss:')[1].split(', lr')[0])}
    metric = float(re.search(r'avg_joint_error: (-?\d+\.\d+)', line).group(1))
    images = [f for f in os.listdir(engine.image_dir)]
    engine.save_model(0)
    ck = torch.load(os.path.join(engine.model_path, 'model_0.pth'), map_location='cpu')
    print(json.dumps(dict(impl='engine', mode=args.mode, engine_file=sys.modules['network.engine'].__file__,
                          network_module=type(engine.network.hg).__module__, criterion_module=type(engine.criterion.mv_projection_loss).__module__,
                          synthesizer_module=type(engine.hand_synthesizer).__module__, terms=terms, avg_joint_error=metric,
                          images_written=images, checkpoint_keys=len(ck['network_state_dict']), epoch_s=dt, log_line=line)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--engine', type=int, default=0, help='1: run the reference\'s own Engine for one epoch instead of the timing loop')
    ap.add_argument('--mode', default='stock', choices=['stock', 'dropin'])
    ap.add_argument('--S', type=int, default=128)
    ap.add_argument('--stacks', type=int, default=2)
    ap.add_argument('--B', type=int, default=64)
    ap.add_argument('--Ns', type=int, default=64)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--sync', type=int, default=0, help='1: float() every loss term every step like RunningAverage (engine.py:40,371)')
    ap.add_argument('--lr', type=float, default=1e-4)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--real_aug', type=int, default=1)
    ap.add_argument('--tf32', default='off', choices=['off', 'default'], help="'default': torch's own flags, as the reference runs (cuDNN TF32 convolutions)")
    ap.add_argument('--weights', default='', help="'trained': pretrained/synthetic.pth (1 stack only)")
    ap.add_argument('--sections', type=int, default=0, help='1: after the timed loop, time the sections of one step (with synchronisations)')
    ap.add_argument('--dump', default='', help='write the first step\'s loss terms + joints to this .npz')
    args = ap.parse_args()
    setup(args.mode)
    if args.engine:
        return run_engine(args)
    import numpy as np
    import torch
    if args.tf32 == 'off':                         # strict fp32; 'default' leaves torch's flags alone (cuDNN convolutions: TF32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    from network.constants import Constant
    from network.create_network_and_criterion import HeatmapEstimationNetwork, MultiTaskLoss
    from network.util_modules import HandSynthesizer
    constant = Constant()
    V, S, hm = 3, args.S, args.S // 4
    B, Ns = args.B, args.Ns
    note = ''
    while True:
        try:
            torch.manual_seed(args.seed)
            network = HeatmapEstimationNetwork(hm, constant.depth_scale, constant.num_joint, args.stacks, real_aug=bool(args.real_aug)).cuda()
            if args.weights == 'trained':
                ck = torch.load('pretrained/synthetic.pth', map_location='cuda')
                network.load_state_dict(ck['network_state_dict'] if hm == 16 else
                                        {k: v for k, v in ck['network_state_dict'].items() if not k.startswith('xyz_recover.')}, strict=(hm == 16))
            criterion = MultiTaskLoss(True, True, True, False, True, True, True, constant, image_size=S, heatmap_size=hm).cuda()
            synth = HandSynthesizer(constant.mesh, image_size=S, heatmap_size=hm, uv_hm_scale=constant.uv_hm_scale,
                                    depth_scale=constant.depth_scale).cuda()
            optimizer = torch.optim.Adam(network.parameters(), lr=args.lr, weight_decay=1e-5)
            orig_real, cams, inv, poses = make_inputs(B, V, Ns, S, args.seed + 10, dev)
            network.train()
            first = {}

            def step(record=False):
                real_dms = orig_real * constant.depth_scale                                        # engine.py:337
                synt_dms, uv_hms, d_hms, synt_xyz = synth(poses)                                   # :350-351
                optimizer.zero_grad()                                                              # :355
                result = network(synt_dms=synt_dms, real_dms=real_dms)                             # :356
                real_target = {'real_dms': orig_real, 'camera_poses': cams, 'inv_camera_poses': inv, 'is_mv': True}
                synt_target = {'uv_hms': uv_hms, 'd_hms': d_hms, 'xyz_pts': synt_xyz}
                loss_terms, ball_dms = criterion(result, real_target=real_target, synt_target=synt_target)   # :366-367
                loss = 0
                for _, l in loss_terms.items():                                                    # sum_loss_terms, :144-148
                    loss = loss + l
                if args.sync:
                    _ = {k: float(v) for k, v in loss_terms.items()}                               # RunningAverage.append, :40
                loss.backward()                                                                    # :375
                optimizer.step()                                                                   # :376
                if record:
                    first.update({'term.' + k: float(v) for k, v in loss_terms.items()})
                    first['loss'] = float(loss)
                    for i, x in enumerate(result['real_xyz']):
                        first['real_xyz%d' % i] = x.detach().float().cpu().numpy()
                    first['ball_dms0'] = ball_dms[0].detach().float().cpu().numpy() if len(ball_dms) else np.zeros(0)
                return loss

            torch.manual_seed(args.seed + 20)
            step(record=True)
            for _ in range(max(args.warmup - 1, 0)):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
            for i in range(args.steps):
                step()
                marks[i].record()
            e1.record()
            torch.cuda.synchronize()
            per_step = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
            wall = time.perf_counter() - t0
            ms = e0.elapsed_time(e1) / max(args.steps, 1)
            sections = {}
            if args.sections:
                def tick(name, t=[None]):
                    torch.cuda.synchronize()
                    now = time.perf_counter()
                    if t[0] is not None and name:
                        sections[name] = sections.get(name, 0.0) + (now - t[0]) * 1e3 / 3
                    t[0] = now
                for _ in range(3):
                    tick('')
                    real_dms = orig_real * constant.depth_scale
                    synt_dms, uv_hms, d_hms, synt_xyz = synth(poses)
                    tick('synthesizer')
                    optimizer.zero_grad()
                    result = network(synt_dms=synt_dms, real_dms=real_dms)
                    tick('network forward')
                    loss_terms, ball_dms = criterion(result, real_target={'real_dms': orig_real, 'camera_poses': cams, 'inv_camera_poses': inv, 'is_mv': True},
                                                     synt_target={'uv_hms': uv_hms, 'd_hms': d_hms, 'xyz_pts': synt_xyz})
                    loss = sum(loss_terms.values())
                    tick('criterion')
                    loss.backward()
                    tick('backward')
                    optimizer.step()
                    tick('optimizer.step')
            break
        except torch.cuda.OutOfMemoryError:
            if B <= 1:
                raise
            del network, criterion, synth, optimizer
            torch.cuda.empty_cache()
            note = 'out of memory at B=%d; ' % B + note
            B, Ns = B // 2, max(Ns // 2, 1)
    images = Ns + B * V
    out = dict(impl='gpu_reference' if args.mode == 'stock' else 'dropin', mode=args.mode, value=images / (ms * 1e-3), unit='images/s',
               ms_per_step=ms, wall_ms_per_step=wall * 1e3 / max(args.steps, 1), images_per_step=images, B=B, V=V, Ns=Ns, S=S, stacks=args.stacks,
               steps=args.steps, warmup=args.warmup, per_step_sync=bool(args.sync), real_aug=bool(args.real_aug), dtype=('f32, TF32 off' if args.tf32 == 'off' else 'f32, torch default flags (cuDNN convolutions in TF32)') if args.mode == 'stock' else 'bf16 hourglass / f32 heads',
               peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, note=note + ('body of network/engine.py:349-376 over the %s modules' % (
                   'unmodified reference (oracle/_ref/reference, eager PyTorch + its own CUDA rasteriser)' if args.mode == 'stock'
                   else 'reference-named modules of spherehand_b200.install()')),
               torch=torch.__version__, sections_ms=sections, per_step_ms=[round(x, 2) for x in per_step], terms={k: v for k, v in first.items() if k.startswith('term.') or k == 'loss'})
    if args.dump:
        np.savez(args.dump, **{k: np.asarray(v) for k, v in first.items()})
    print(json.dumps(out))


if __name__ == '__main__':
    main()
