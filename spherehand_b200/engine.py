"""The self-supervised train step on the B200 kernels: body of `Engine._epoch_with_both`
(/root/reference/network/engine.py:349-376) — synthetic branch -> zero_grad -> network -> losses -> sum -> backward ->
Adam step — issued as one static sequence of C-ABI kernel launches and replayed as a CUDA graph.

    synthetic branch (no grad)   HandSynthesizer.forward            network/util_modules.py:104-122
    network                      HeatmapEstimationNetwork.forward   network/create_network_and_criterion.py:84-144
    losses                       MultiTaskLoss.forward              network/create_network_and_criterion.py:183-263
    optimiser                    Adam(lr, weight_decay=1e-5)        network/engine.py:95-97

torch is used for device memory, streams, RNG draws (same draws, same order as the reference: SURVEY §7.3-9), CUDA-graph
capture and the NCCL gradient all-reduce; every arithmetic kernel is in libspherehand_b200.so.  There is no CPU path.

Scale augmentation (`real_aug=True`, what the reference's Engine trains with: HeatmapEstimationNetwork(real_aug=True),
create_network_and_criterion.py:94-102,124-126): with probability 1/2 per step every real view is resized by its own
(u, v) scale and pasted on an all-ones canvas before the network, and the recovered x / y of its joints are divided by
(u, v).  The draws live in static device buffers (ones when the step does not augment), the resize and the two divisions
(values and gradient) are kernels of the static launch sequence, so the captured graph replays either case.

Data parallelism: tuples and synthetic poses shard across ranks; the flat fp32 gradient is all-reduced (SUM)
once, between the backward pass and Adam (two CUDA-graph segments with the NCCL call between their replays).
`bucketed=True` instead all-reduces in buckets that follow the backward pass (stack k's parameters are final when its
backward ends; their all-reduce runs on a communication stream underneath the rest of the pass, graph segments cut at the
bucket boundaries).  Measured on two B200s the bucketed form is SLOWER (13.80 vs 13.67 ms per step, 13.52 on one GPU): the
NCCL kernels take SMs away from persistent kernels that were launched with one CTA per SM, which costs more than the 0.15 ms
the single exposed all-reduce does.  Batch-MEAN loss terms are
pre-scaled by 1/world_size and batch-SUM terms (collision, VAE KLD) are not, so the summed gradient equals the single-GPU
gradient at the global batch and the summed terms (`loss_dict(reduce=True)`) are its loss values.
"""
import torch

from . import _lib, ops, parallel
from .network.hourglass import HourglassNet

LOSS_WEIGHTS = {          # MultiTaskLoss.weights, create_network_and_criterion.py:171-181
    'synt_hm': 1e3, 'synt_pt': 1e-1, 'mv_consistency': 1e-3, 'mv_projection': 1.0, 'temporal_smooth': 1.0,
    'prior': 1e-2, 'hm_mean': 1e-2, 'domain': 0.0, 'collision': 1.0, 'bone_length': 1.0,
}
TERM_NAMES = ('synt_uv', 'synt_d', 'mv_projection', 'mv_consistency', 'uv_hm_mean', 'pose_prior', 'collision',
              'bone_length', 'total')

_LATTICE = {128: (5, 2, 1), 64: (10, 4, 2)}      # 640 -> S bilinear resize == these source samples


class SelfSupTrainStep:
    """One replica of the train step for B multi-view tuples x V views + Ns synthetic poses at S x S depth maps."""

    def __init__(self, net, hand, vae_blob, B, V, Ns, S, depth_scale=0.01, lr=1e-4, weight_decay=1e-5,
                 weights=None, use_prior=True, use_collision=True, use_bone_length=True, use_mv_projection=True,
                 use_mv_consistency=True, world_size=1, process_group=None, use_graph=True, real_aug=False,
                 allreduce=None, bucketed=False, overlap_heads=True):
        if not isinstance(net, HourglassNet):
            raise TypeError('net must be a spherehand_b200 HourglassNet')
        if S not in _LATTICE:
            raise ValueError('depth size must be 64 or 128 (the 640 -> S resize lattice is tabulated for those)')
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('SelfSupTrainStep needs a CUDA (sm_100a) device; there is no CPU fallback')
        self.net, self.hand, self.vae_blob, self.dev = net, hand, vae_blob, dev
        self.B, self.V, self.Ns, self.S, self.J = B, V, Ns, S, hand.num_keypoints
        self.hm = S // 4
        self.N = Ns + B * V
        self.depth_scale = depth_scale
        self.weights = dict(LOSS_WEIGHTS)
        self.weights.update(weights or {})
        self.flags = dict(prior=use_prior and vae_blob is not None, collision=use_collision, bone=use_bone_length,
                          proj=use_mv_projection, cons=use_mv_consistency)
        self.world_size, self.pg = world_size, process_group
        # allreduce(tensor): the collective (default: torch.distributed SUM over `process_group`); a test may inject its own
        self._allreduce_fn = allreduce if allreduce is not None else (
            lambda t: parallel.allreduce_gradients(t, self.world_size, self.pg))
        self.bucketed = bucketed and world_size > 1
        self.real_aug = bool(real_aug)
        self.use_graph = use_graph
        self.betas, self.eps_adam, self.weight_decay = (0.9, 0.999), 1e-8, weight_decay
        f32 = dict(device=dev, dtype=torch.float32)
        # ---- static input buffers (H2D targets)
        self.real = torch.full((B, V, S, S), 100.0, **f32)             # mm, background 100.0
        self.cams = torch.eye(4, **f32).repeat(B, V, 1, 1).contiguous()
        self.inv_cams = self.cams.clone()
        self.poses = torch.zeros((Ns, 26), **f32)
        # ---- static RNG buffers (drawn on the host side of the graph, every step)
        self.scales = torch.full((Ns, 3), 0.9, **f32)
        self.rand_f = torch.ones((Ns,), **f32)
        self.noise = torch.zeros((3, Ns, S, S), **f32)
        self.vae_eps = torch.zeros((net.num_stacks, B * V, 32), **f32)
        self.aug_u = torch.ones((B * V,), **f32)                        # scale augmentation of the real views (ones = none)
        self.aug_v = torch.ones((B * V,), **f32)
        self.augmented = False
        self._real_scaled = torch.empty((B * V, S, S), **f32) if self.real_aug else None
        # ---- state
        self.images = torch.zeros((self.N, S, S), **f32)
        self.terms = torch.zeros(9, **f32)
        net.flatten_parameters()
        self.adam_m = torch.zeros_like(net._flat)
        self.adam_v = torch.zeros_like(net._flat)
        self.lr = float(lr)
        self.lr_dev = torch.tensor([lr], **f32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.projected_dms = None
        self._graphs = {}
        self._staging = None                                   # prefetch_batch / commit_batch
        self._side = [torch.cuda.Stream(device=dev) for _ in range(2)]     # the VAE prior runs here, under the projection loss
        self._heads_stream = torch.cuda.Stream(device=dev)     # heads of stack k < last run here, under stack k+1's forward
        self.overlap_heads = overlap_heads
        self._comm = torch.cuda.Stream(device=dev)             # gradient buckets are all-reduced here, under the backward pass
        self.launches_per_step = None

    # ------------------------------------------------------------------ host-side plumbing
    def set_lr(self, lr):
        self.lr = float(lr)
        self.lr_dev.fill_(lr)

    def load_batch(self, real_dms, camera_poses, inv_camera_poses, pose_parameters, non_blocking=True):
        """Host (pinned) or device tensors -> the static device buffers."""
        self.real.copy_(real_dms, non_blocking=non_blocking)
        self.cams.copy_(camera_poses, non_blocking=non_blocking)
        self.inv_cams.copy_(inv_camera_poses, non_blocking=non_blocking)
        self.poses.copy_(pose_parameters, non_blocking=non_blocking)
        return real_dms.numel() * 4 + 2 * camera_poses.numel() * 4 + pose_parameters.numel() * 4

    def prefetch_batch(self, real_dms, camera_poses, inv_camera_poses, pose_parameters):
        """Start the host->device copies of the NEXT batch (pinned host tensors) on a copy stream into staging buffers; they run
        underneath the current step.  `commit_batch()` then moves them into the static buffers the captured graph reads
        (device-to-device, microseconds).  Returns the bytes that cross the host link."""
        if self._staging is None:
            self._staging = [torch.empty_like(t) for t in (self.real, self.cams, self.inv_cams, self.poses)]
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._staged, self._committed = torch.cuda.Event(), torch.cuda.Event()
            self._committed.record(torch.cuda.current_stream(self.dev))
        self._copy_stream.wait_event(self._committed)          # the previous commit has finished reading the staging buffers
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._staging, (real_dms, camera_poses, inv_camera_poses, pose_parameters)):
                dst.copy_(src, non_blocking=True)
            self._staged.record(self._copy_stream)
        self._has_staged = True
        return sum(t.numel() * 4 for t in self._staging)

    def commit_batch(self):
        """Make the prefetched batch the current one (see prefetch_batch)."""
        if not getattr(self, '_has_staged', False):
            raise RuntimeError('commit_batch: no batch has been prefetched')
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(self._staged)
        for dst, src in zip((self.real, self.cams, self.inv_cams, self.poses), self._staging):
            dst.copy_(src, non_blocking=True)
        self._committed.record(main)
        self._has_staged = False

    def sample_poses(self, generator=None, sequential=True):
        """Fill the synthetic-branch poses on the device with the batched JointAngleDataset sampler (dataset/joint_angle.py of the
        reference: what its DataLoader workers draw one `torch.rand(1)` at a time) instead of copying them from the host."""
        from .dataset.joint_angle import JointAngleDataset
        self.poses.copy_(JointAngleDataset(self.dev).sample_batch(self.Ns, generator, sequential))

    def draw_randoms(self, generator=None):
        """The step's random draws, in the reference's order: RandScale x,y,z (pointTransformation.py:140-142),
        rand_f_ratio (util_modules.py:107), DepthNoise shift_x, shift_y, z (util_modules.py:64,69,83), VAE eps per stack
        (pose_vae.py:51)."""
        g = generator
        self.scales.uniform_(0.85, 0.95, generator=g)
        self.rand_f.uniform_(0.9, 1.1, generator=g)
        self.noise.normal_(generator=g)
        if self.real_aug:
            # HeatmapEstimationNetwork.forward (:94-102): one host uniform decides whether this step augments, then the common
            # scale and the u / v jitter per real view
            self.augmented = bool(torch.rand(1).item() >= 0.5)
            if self.augmented:
                self.aug_u.uniform_(0.0, 1.0, generator=g).mul_(0.2).add_(0.75)                  # rnd_scale
                self.aug_v.copy_(self.aug_u)
                self.aug_u.add_(torch.rand(self.aug_u.shape, device=self.dev, generator=g) * 0.1 - 0.05)
                self.aug_v.add_(torch.rand(self.aug_v.shape, device=self.dev, generator=g) * 0.1 - 0.05)
            else:
                self.aug_u.fill_(1.0)
                self.aug_v.fill_(1.0)
        self.vae_eps.normal_(generator=g)

    # ------------------------------------------------------------------ the step
    def _forward_backward(self, is_mv, on_bucket=None):
        """on_bucket(lo, hi): called by the backward pass as soon as flat_grad[lo:hi] is final (data-parallel runs)."""
        net, hand = self.net, self.hand
        B, V, Ns, S, J, hm, N = self.B, self.V, self.Ns, self.S, self.J, self.hm, self.N
        M = B * V
        w = self.weights
        ms = parallel.mean_scale(self.world_size)
        # ---- synthetic branch (HandSynthesizer.forward): FK -> scale -> LBS -> project -> rasterise -> resize -> noise
        step, off, noff = _LATTICE[S]
        mats = ops.fk_fwd(self.poses, hand.offset_mats, hand.inv_offset_mats, self.scales)
        pts = ops.lbs_fwd(mats, *hand.mesh_csr, right_hand=True, mode=1, cam=(320.0, 320.0, 640 / 300, 640 / 300),
                          rand_f=self.rand_f)
        fv = ops.gather_faces(pts, hand.faces)
        z = ops.tri_raster_lattice_fwd(fv, 640, step, off, noff)
        dm = ops.lattice_to_depth(z, S, noff, self.depth_scale)
        ops.depth_noise(dm, self.noise[0], self.noise[1], self.noise[2], out=self.images[:Ns])
        uvd = ops.lbs_fwd(mats, *hand.kp_csr, right_hand=True, mode=1, cam=(hm / 2, hm / 2, hm / 300, hm / 300),
                          rand_f=self.rand_f)
        uv_t, _d_t, xyz_t = ops.heatmap_render(uvd, hm)
        # ---- real branch input: real_dms * depth_scale (engine.py:337), then the scale augmentation (a scale of 1 is the identity)
        if self.real_aug:
            ops.scale(self.real, self.depth_scale, self._real_scaled)
            ops.resize_crop(self._real_scaled, self.aug_u, self.aug_v, out=self.images[Ns:])
        else:
            ops.scale(self.real, self.depth_scale, self.images[Ns:])
        # ---- network + heads.  The heads of a stack need only that stack's scores, and their result (d loss / d score) is not
        # needed before the backward pass: the heads of every stack but the last therefore run on their own stream underneath the
        # NEXT stack's forward pass (fork / join events = graph edges); the last stack's heads run in line.
        self.terms.zero_()
        w8 = (w['synt_hm'], w['synt_pt'], w['mv_projection'], w['mv_consistency'] if is_mv else 0.0, w['hm_mean'],
              w['prior'], w['collision'], w['bone_length'])
        ctx = dict(is_mv=is_mv, w=w, w8=w8, ms=ms, uv_t=uv_t, xyz_t=xyz_t)
        n_stacks = net.num_stacks
        gscores, projected, joins = [None] * n_stacks, [None] * n_stacks, []
        main = torch.cuda.current_stream(self.dev)

        def on_score(si, score):
            if self.overlap_heads and si < n_stacks - 1:
                fork = torch.cuda.Event()
                fork.record(main)
                self._heads_stream.wait_event(fork)
                with torch.cuda.stream(self._heads_stream):
                    gscores[si], projected[si] = self._heads(si, score, ctx)
                    done = torch.cuda.Event()
                    done.record(self._heads_stream)
                joins.append(done)
            else:
                gscores[si], projected[si] = self._heads(si, score, ctx)
        net.run_forward(self.images, on_score=on_score, want_latents=False)
        for done in joins:
            main.wait_event(done)
        projected = [p for p in projected if p is not None]
        self.projected_dms = projected
        net.run_backward(gscores, on_bucket=on_bucket)

    def _heads(self, si, score, ctx):
        """MultiTaskLoss.forward for the output of stack `si` (create_network_and_criterion.py:183-263) on the current stream:
        soft-argmax, every loss head, the weighted terms, and d loss / d score.  -> (gscore, projected_dms | None)."""
        hand = self.hand
        B, V, Ns, J, hm = self.B, self.V, self.Ns, self.J, self.hm
        M, hw = B * V, hm * hm
        w, ms, uv_t, is_mv = ctx['w'], ctx['ms'], ctx['uv_t'], ctx['is_mv']
        fused_layout = J == 41 and score.shape[1] == 2 * J       # the NHWC backward kernel's build (the hand model's 41 spheres)
        if fused_layout:
            xyz, sse, aux = ops.softargmax_fwd(score, J, Ns, uv_t, 1.0 / self.depth_scale, want_sse=True, want_aux=True)
        else:
            xyz, sse = ops.softargmax_fwd(score, J, Ns, uv_t, 1.0 / self.depth_scale, want_sse=True)
        if self.real_aug:
            ops.unscale_xy(xyz[Ns:], self.aug_u, self.aug_v)         # xyz[:, :, 0] /= u, xyz[:, :, 1] /= v  (:124-126)
        joints = xyz[Ns:].view(B, V, J, 3)
        loss_mv = g_mv = loss_p = g_p = loss_v = g_v = proj = None
        joined = None
        here = torch.cuda.current_stream(self.dev)
        if self.flags['prior']:
            # The VAE prior is a latency chain of 14 small dense layers on M/4 CTAs: it goes out FIRST, on a side stream, and
            # runs underneath the projection loss (which fills the GPU) instead of in front of it.  Inside the CUDA graph the
            # fork / join events become graph edges.
            side = self._side[si % len(self._side)]
            fork = torch.cuda.Event()
            fork.record(here)
            side.wait_event(fork)
            with torch.cuda.stream(side):
                x = torch.empty((M, J * 3), device=self.dev, dtype=torch.float32)
                ops.scale(xyz[Ns:], 0.01, x)
                loss_v, g_v = ops.vae_prior_fwdbwd(x, self.vae_eps[si], self.vae_blob, M_mean=M * self.world_size)
                joined = torch.cuda.Event()
                joined.record(side)
        if self.flags['proj']:
            loss_mv, proj, g_mv = ops.mvproj_loss_fwdbwd(self.cams, self.inv_cams, joints, self.real, hand.radii, is_mv)
        pf = (1 if self.flags['cons'] else 0) | (2 if self.flags['collision'] else 0) | (4 if self.flags['bone'] else 0)
        if pf:
            loss_p, g_p = ops.pose_losses_fwdbwd(self.cams, joints, flags=pf)
        if joined is not None:
            here.wait_event(joined)
        gxyz = torch.empty_like(xyz)
        ops.step_combine(xyz, Ns, M, J, hw, ctx['w8'], gxyz, self.terms, g_mvproj=g_mv, g_pose3=g_p, g_prior=g_v,
                         target_xyz4=ctx['xyz_t'], loss_mv3=loss_mv, loss_pose3=loss_p, loss_prior3=loss_v, sse2=sse,
                         mean_scale=ms)
        if self.real_aug:
            ops.unscale_xy(gxyz[Ns:], self.aug_u, self.aug_v)        # backward of the two divisions
        c_synt = 2.0 * w['synt_hm'] * ms / (Ns * J * hw) if Ns else 0.0
        c_real = 2.0 * w['hm_mean'] * ms / (M * J * hw) if M else 0.0
        if fused_layout:
            # d loss / d score straight into the bf16 NHWC layout the network's backward consumes (no fp32 NCHW round trip)
            gscore = ops.softargmax_bwd_nhwc(score, gxyz, aux, J, Ns, uv_t, 1.0 / self.depth_scale, c_synt=c_synt, c_real=c_real)
        else:
            gscore = torch.empty_like(score)
            if score.shape[1] != 2 * J:
                gscore.zero_()
            ops.softargmax_bwd(score, gxyz, J, Ns, uv_t, 1.0 / self.depth_scale, c_synt=c_synt, c_real=c_real, out=gscore)
        return gscore, proj

    def _reduce_bucket(self, lo, hi):
        """Called by the backward pass when flat_grad[lo:hi] is final: all-reduce it on the communication stream."""
        main = torch.cuda.current_stream(self.dev)
        self._comm.wait_stream(main)
        with torch.cuda.stream(self._comm):
            self._allreduce_fn(self.net._flat_grad[lo:hi])

    def _optimizer(self):
        net = self.net
        ops.adam_step_dev(net._flat, net._flat_grad, self.adam_m, self.adam_v, self.lr_dev, self.step_dev,
                          self.betas[0], self.betas[1], self.eps_adam, self.weight_decay)

    def _allreduce(self):
        if self.world_size > 1 and not self.bucketed:
            self._allreduce_fn(self.net._flat_grad)

    def _eager_step(self, is_mv):
        self._forward_backward(is_mv, on_bucket=self._reduce_bucket if self.bucketed else None)
        if self.bucketed:
            torch.cuda.current_stream(self.dev).wait_stream(self._comm)      # every bucket has been reduced
        self._allreduce()
        self._optimizer()

    def _capture(self, is_mv):
        """Warm up eagerly on a side stream (lazy CUDA init, cudaFuncSetAttribute, allocator), then capture the step as a list of
        CUDA-graph SEGMENTS with the gradient collectives between them: [('graph', g) | ('reduce', lo, hi), ...].  One GPU: a single
        segment.  Data parallel: the backward pass is cut where a bucket of the flat gradient becomes final; at replay the bucket's
        all-reduce is issued (an ordinary NCCL call on the communication stream) and the next segment is replayed on top of it, so
        the collective runs underneath the rest of the backward pass.  NCCL stays OUT of the captured graphs on purpose: with
        the collectives captured (one graph for everything) a later `dist.barrier()` never returned on this stack."""
        import gc
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        step0 = self.step_dev.clone()
        flat0, m0, v0 = self.net._flat.clone(), self.adam_m.clone(), self.adam_v.clone()
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager_step(is_mv)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # the warm-up steps must not count as training
        self.net._flat.copy_(flat0); self.adam_m.copy_(m0); self.adam_v.copy_(v0); self.step_dev.copy_(step0)
        items, state = [], dict(pool=None)
        count = _lib.lib().sh_launch_count

        def begin():
            g = torch.cuda.CUDAGraph()
            ctx = torch.cuda.graph(g) if state['pool'] is None else torch.cuda.graph(g, pool=state['pool'])
            ctx.__enter__()
            state.update(g=g, ctx=ctx, n0=count())

        def end():
            state['ctx'].__exit__(None, None, None)
            if state['pool'] is None:
                state['pool'] = state['g'].pool()          # later segments read the activations the earlier ones wrote
            items.append(('graph', state['g']))

        def cut(lo, hi):
            if count() != state['n0']:                      # (consecutive buckets share one cut)
                end()
                items.append(('reduce', lo, hi))
                begin()
            else:
                items.append(('reduce', lo, hi))

        # Python's cyclic collector must not run while a capture is open: collecting an old step's CUDA graph there frees its memory
        # pool (cudaFree) and invalidates the capture ("operation failed due to a previous error during capture")
        gc_was = gc.isenabled()
        gc.collect()
        gc.disable()
        n0 = count()
        try:
            begin()
            self._forward_backward(is_mv, on_bucket=cut if self.bucketed else None)
            if self.world_size > 1 and not self.bucketed:
                cut(0, self.net._flat_grad.numel())
            self._optimizer()
            end()
        finally:
            if gc_was:
                gc.enable()
        self.launches_per_step = count() - n0     # kernels of this library inside one step
        return items

    def _replay(self, items):
        main = torch.cuda.current_stream(self.dev)
        last = max(i for i, it in enumerate(items) if it[0] == 'graph')
        reduced = False
        for i, it in enumerate(items):
            if it[0] == 'graph':
                if i == last and reduced:
                    main.wait_stream(self._comm)            # Adam reads the reduced gradient
                it[1].replay()
            else:
                self._reduce_bucket(it[1], it[2])
                reduced = True

    def step(self, is_mv=True):
        """One optimisation step on the buffers filled by load_batch / draw_randoms.  Returns the device tensor of the
        9 weighted loss terms (TERM_NAMES); reading it synchronises."""
        if not self.use_graph:
            self._eager_step(is_mv)
            return self.terms
        key = bool(is_mv)
        if key not in self._graphs:
            self._graphs[key] = self._capture(is_mv)
        self._replay(self._graphs[key])
        return self.terms

    # ------------------------------------------------------------------ checkpoints (reference format)
    def _flat_views(self, flat):
        net = self.net
        return [flat[off:off + n].view(p.shape) for p in net.parameters() for off, n in (net._offsets[id(p)],)]

    def checkpoint(self, epoch):
        """The dict `Engine.save_model` writes (network/engine.py:437-443): {'epoch', 'network_state_dict', 'optimizer_state_dict'}.
        `network_state_dict` has the keys of the reference's HeatmapEstimationNetwork (`hg.*` + the soft-argmax grids);
        `optimizer_state_dict` is produced by a real torch.optim.Adam holding views of the flat moment buffers, so it is in
        the installed torch's own format and `torch.optim.Adam.load_state_dict` of the reference takes it unchanged."""
        net = self.net
        sd = {'hg.' + k: v.detach().clone() for k, v in net.state_dict().items()}
        hm = self.hm
        v_grid, u_grid = torch.meshgrid(torch.arange(hm, dtype=torch.float32), torch.arange(hm, dtype=torch.float32), indexing='ij')
        sd['xyz_recover.u_grid'] = u_grid.reshape(1, 1, hm, hm).to(self.dev)
        sd['xyz_recover.v_grid'] = v_grid.reshape(1, 1, hm, hm).to(self.dev)
        params = list(net.parameters())
        opt = torch.optim.Adam(params, lr=self.lr, betas=self.betas, eps=self.eps_adam, weight_decay=self.weight_decay)
        step = float(self.step_dev.item())
        if step > 0:
            for p, m, v in zip(params, self._flat_views(self.adam_m), self._flat_views(self.adam_v)):
                opt.state[p] = {'step': torch.tensor(step), 'exp_avg': m.detach().clone(), 'exp_avg_sq': v.detach().clone()}
        return {'epoch': epoch, 'network_state_dict': sd, 'optimizer_state_dict': opt.state_dict()}

    def load_checkpoint(self, check_point, load_optimizer=True):
        """`Engine.load_model` (network/engine.py:445-459) for the fused step: weights into the flat buffer, and (for a numbered
        checkpoint) Adam's moments, step count and learning rate into the device-side optimiser state."""
        net = self.net
        sd = {k[3:]: v for k, v in check_point['network_state_dict'].items() if k.startswith('hg.')}
        net.load_state_dict(sd)
        net.flatten_parameters()
        if not load_optimizer:
            return
        params = list(net.parameters())
        opt = torch.optim.Adam(params, lr=self.lr, betas=self.betas, eps=self.eps_adam, weight_decay=self.weight_decay)
        opt.load_state_dict(check_point['optimizer_state_dict'])
        self.adam_m.zero_(); self.adam_v.zero_()
        steps = set()
        for p, m, v in zip(params, self._flat_views(self.adam_m), self._flat_views(self.adam_v)):
            st = opt.state.get(p)
            if st:
                m.copy_(st['exp_avg']); v.copy_(st['exp_avg_sq'])
                steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise ValueError('checkpoint holds different Adam step counts per parameter; the fused optimiser keeps one')
        self.step_dev.fill_(steps.pop() if steps else 0)
        self.set_lr(opt.param_groups[0]['lr'])

    def loss_dict(self, reduce=True):
        """The 9 weighted loss terms as floats.  Under data parallelism each rank's terms are already scaled by the reduction rule
        of the term (batch-MEAN terms by 1/world_size, batch-SUM terms not), so their SUM over ranks (reduce=True: one tiny
        all-reduce, a collective every rank must enter) is the loss of the global batch; reduce=False gives this rank's share."""
        t = self.terms.clone()
        if reduce and self.world_size > 1:
            parallel.allreduce_terms(t, self.world_size, self.pg)
        return dict(zip(TERM_NAMES, t.tolist()))

    def step_lr(self, step_size, gamma=0.1):
        """torch.optim.lr_scheduler.StepLR(optimizer, step_size, gamma) of the reference (engine.py:98-99: step_size = epochs // 3,
        gamma = 0.1) for the fused optimiser: `.step()` once per epoch moves the device-side learning rate, no re-capture."""
        return StepLR(self, step_size, gamma)


class StepLR:
    """lr = base_lr * gamma ** (epoch // step_size), epoch counted by `.step()` calls (torch.optim.lr_scheduler.StepLR)."""

    def __init__(self, train_step, step_size, gamma=0.1, last_epoch=0):
        if step_size < 1:
            raise ValueError('step_size must be >= 1 (the reference divides the epoch count by 3: needs >= 3 epochs)')
        self.train_step, self.step_size, self.gamma = train_step, int(step_size), float(gamma)
        self.base_lr = train_step.lr
        self.last_epoch = int(last_epoch)
        if last_epoch:
            train_step.set_lr(self.get_lr())

    def get_lr(self):
        return self.base_lr * self.gamma ** (self.last_epoch // self.step_size)

    def step(self):
        self.last_epoch += 1
        self.train_step.set_lr(self.get_lr())
        return self.train_step.lr
