"""Stacked-hourglass network on the B200 kernels — drop-in for /root/reference/network/hourglass.py.

`create_hourglass_network(num_outputs, num_stacks)` returns a module with the reference's parameter names and
shapes (`conv1.weight (64,1,5,5)`, `layer1.0.bn1.weight`, `hg.0.hg.1.2.0.conv2.weight`, `res.0.0...`, `fc.0.0/.1`,
`score.0`, `fc_.0`, `score_.0`; hourglass.py:89-120), so `pretrained/*.pth` load unchanged, and the reference's
`forward(x) -> (list[score], list[latent])` (hourglass.py:147-173).  The nn.Conv2d / nn.GroupNorm children are
parameter holders only: the arithmetic runs through the C ABI (tcgen05 implicit-GEMM convolutions, fused
GroupNorm/ReLU/pool/up-sample kernels) on NHWC bf16 activations with fp32 accumulation, fp32 master weights and
fp32 GroupNorm statistics.  Precision contract: bf16 operands => heat-maps within 2e-2 of the fp32 reference
(tests/test_gpu_hourglass.py); everything downstream of the heat-maps is fp32.

There is no CPU path: calling forward on CPU tensors raises.
"""
import os

import torch
import torch.nn as nn

from .. import ops

BF16 = torch.bfloat16
# GroupNorm + ReLU in front of a bottleneck's 1x1 layers is applied inside the convolution / weight-gradient kernels
# (sh_conv_fwd_gn / sh_conv_wgrad_gn) instead of by a separate sh_gn_relu_fwd pass.  SH_FUSE_GN=0 keeps the separate pass.
FUSE_GN = os.environ.get('SH_FUSE_GN', '1') != '0'


def _ceil(x, m):
    return (x + m - 1) // m * m


class Bottleneck(nn.Module):
    """Parameter holder for the pre-activation bottleneck (hourglass.py:7-41)."""
    expansion = 2

    def __init__(self, inplanes, planes, downsample=None):
        super().__init__()
        self.bn1 = nn.GroupNorm(16, inplanes)
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=True)
        self.bn2 = nn.GroupNorm(16, planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1, bias=True)
        self.bn3 = nn.GroupNorm(16, planes)
        self.conv3 = nn.Conv2d(planes, planes * 2, kernel_size=1, bias=True)
        self.downsample = downsample


class Hourglass(nn.Module):
    """Parameter holder for one depth-`depth` hourglass (hourglass.py:44-66)."""

    def __init__(self, planes, depth):
        super().__init__()
        self.depth = depth
        levels = []
        for i in range(depth):
            n_res = 4 if i == 0 else 3
            levels.append(nn.ModuleList([nn.Sequential(Bottleneck(planes * 2, planes)) for _ in range(n_res)]))
        self.hg = nn.ModuleList(levels)


class _Tensor:
    """NHWC bf16 activation + the fp32 [N,16,2] statistics the consuming GroupNorm needs."""
    __slots__ = ('buf', 'stats', 'H', 'W', 'C', 'bias_grads')

    def __init__(self, buf, stats, H, W, C):
        self.buf, self.stats, self.H, self.W, self.C = buf, stats, H, W, C
        # fp32 gradient views that equal the per-channel column sum of dL/d(this tensor): the biases of the convolutions
        # whose outputs sum to it.  The kernel that finalises dL/d(tensor) fuses that column sum (`_Run.cs`).
        self.bias_grads = []


class HourglassNet(nn.Module):
    def __init__(self, num_stacks, num_outputs):
        super().__init__()
        self.num_stacks = num_stacks
        self.num_outputs = num_outputs
        self.num_feats = 128
        ch = 256
        # creation order follows hourglass.py:95-120 so that a seeded default init matches the reference's
        self.conv1 = nn.Conv2d(1, 64, kernel_size=5, padding=2, stride=2, bias=True)
        self.bn1 = nn.GroupNorm(4, 64)
        self.layer1 = nn.Sequential(Bottleneck(64, 64, nn.Sequential(nn.Conv2d(64, 128, kernel_size=1, bias=True))))
        self.layer2 = nn.Sequential(Bottleneck(128, 128, nn.Sequential(nn.Conv2d(128, 256, kernel_size=1, bias=True))))
        self.layer3 = nn.Sequential(Bottleneck(256, 128))
        hg, res, fc, score, fc_, score_ = [], [], [], [], [], []
        for i in range(num_stacks):
            hg.append(Hourglass(self.num_feats, 2))
            res.append(nn.Sequential(Bottleneck(ch, self.num_feats)))
            fc.append(nn.Sequential(nn.Conv2d(ch, ch, kernel_size=1, bias=True), nn.GroupNorm(16, ch)))   # (:138-145)
            score.append(nn.Conv2d(ch, num_outputs, kernel_size=1, bias=True))
            if i < num_stacks - 1:
                fc_.append(nn.Conv2d(ch, ch, kernel_size=1, bias=True))
                score_.append(nn.Conv2d(num_outputs, ch, kernel_size=1, bias=True))
        self.hg = nn.ModuleList(hg)
        self.res = nn.ModuleList(res)
        self.fc = nn.ModuleList(fc)
        self.score = nn.ModuleList(score)
        self.fc_ = nn.ModuleList(fc_)
        self.score_ = nn.ModuleList(score_)
        self._flat = None
        self._flat_grad = None
        self._wcache = {}

    # ------------------------------------------------------------------ flat parameter storage
    def flatten_parameters(self):
        """Re-home every parameter as a view of one flat fp32 buffer (one Adam launch, one gradient all-reduce)."""
        params = list(self.parameters())
        dev = params[0].device
        if self._flat is not None and self._flat.device == dev and all(p.data_ptr() == q for p, q in zip(params, self._ptrs)):
            return self._flat
        total = sum(_ceil(p.numel(), 4) for p in params)
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        grad = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        self._offsets = {}
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1).float())
            p.data = flat[off:off + n].view(p.shape)
            self._offsets[id(p)] = (off, n)
            off += _ceil(n, 4)
        self._flat, self._flat_grad = flat, grad
        self._ptrs = [p.data_ptr() for p in params]
        self._wcache = {}
        return flat

    def _conv_geometry(self, conv):
        Cout, Cin, k, _ = conv.weight.shape
        taps = k * k
        cout_pad = _ceil(Cout, 128) if Cout > 64 else 64
        cin_pad = _ceil(Cin, 64) if Cin % 64 else Cin
        if Cin == self.num_outputs:
            cin_pad = 128
        # data-gradient GEMM: rows = Cin (padded to its N tile), cols = Cout (padded to a K multiple of 64)
        b_rows = _ceil(Cin, 128) if Cin > 64 else 64
        b_cols = _ceil(Cout, 64)
        if Cout == self.num_outputs:
            b_cols = 128
        return Cout, Cin, taps, cout_pad, cin_pad, b_rows, b_cols

    def pack_all_weights(self):
        """fp32 masters -> bf16 tensor-core layouts of every stride-1 convolution, one launch (sh_pack_weights_batch)."""
        self.flatten_parameters()
        if not self._wcache:
            convs = [m for m in self.modules() if isinstance(m, nn.Conv2d) and m is not self.conv1]
            rows, off = [], 0
            for c in convs:
                Cout, Cin, taps, cout_pad, cin_pad, b_rows, b_cols = self._conv_geometry(c)
                nf, nb = taps * cout_pad * cin_pad, taps * b_rows * b_cols
                rows.append([self._offsets[id(c.weight)][0], Cout, Cin, taps, cout_pad, cin_pad, b_rows, b_cols, off, off + nf])
                off += nf + nb
            dev = self._flat.device
            self._warena = torch.empty(off, device=dev, dtype=BF16)
            self._wtable = torch.tensor(rows, dtype=torch.int32, device=dev)
            for c, r in zip(convs, rows):
                _, Cout, Cin, taps, cout_pad, cin_pad, b_rows, b_cols, o_f, o_b = r
                wf = self._warena[o_f:o_b].view(taps, cout_pad, cin_pad)
                wb = self._warena[o_b:o_b + taps * b_rows * b_cols].view(taps, b_rows, b_cols)
                self._wcache[id(c)] = (wf, wb, cout_pad, cin_pad, b_rows, b_cols)
            # 3x3 layers: weight gradients accumulate in a [9][Cout][Cin] scratch (vector reductions), unpacked in one launch
            rows3, off3, self._wg3 = [], 0, {}
            for c in convs:
                Cout, Cin, k, _ = c.weight.shape
                if k == 3:
                    rows3.append([off3, self._offsets[id(c.weight)][0], Cout, Cin])
                    self._wg3[id(c)] = (off3, 9 * Cout * Cin)
                    off3 += 9 * Cout * Cin
            self._wg3_scratch = torch.zeros(max(off3, 4), device=dev, dtype=torch.float32)
            self._wg3_table = torch.tensor(rows3 or [[0, 0, 0, 0]], dtype=torch.int32, device=dev)[:len(rows3)].contiguous()
        ops.pack_weights_batch(self._flat, self._wtable, self._warena)

    def grad_view(self, p):
        off, n = self._offsets[id(p)]
        return self._flat_grad[off:off + n].view(p.shape)

    # ------------------------------------------------------------------ execution
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('spherehand_b200 HourglassNet runs on CUDA (sm_100a) only; there is no CPU fallback')
        if x.dim() == 3:
            x = x.unsqueeze(1)
        res = _HourglassFn.apply(self, x, *list(self.parameters()))
        return list(res[:self.num_stacks]), list(res[self.num_stacks:])

    def run_forward(self, x, on_score=None, want_latents=True):
        """x fp32 [N,S,S] or [N,1,S,S] -> (scores fp32 NCHW list, latents fp32 NCHW list); records the backward tape.
        on_score(i, score): called as soon as the score convolution of stack i has been launched (the fused step starts that
        stack's loss heads on another stream while the next stack's forward continues).  want_latents=False skips the fp32 NCHW
        copies of the bottleneck features (the reference returns them; its domain loss has weight 0)."""
        self.flatten_parameters()
        if x.dim() == 4:
            x = x[:, 0]
        x = x.contiguous().float()
        N, S = x.shape[0], x.shape[-1]
        self.pack_all_weights()
        if getattr(self, '_last_run', None) is not None:
            self._last_run.release()       # an abandoned tape: free its activations now, not at the next cyclic-GC pass
        ctx = _Run(self, N, x.device)
        scores, latents = ctx.forward(x, S, on_score, want_latents)
        self._last_run = ctx
        return scores, latents

    def grad_buckets(self):
        """Contiguous slices of the flat gradient in the order the backward pass FINISHES them:
        [('stack k', lo, hi, rows3) for k = last .. 0] + [('rest', ...)]: `hg.k` is final when stack k's backward ends (its
        3x3 weight gradients after their rows `rows3` = (first, count) of the unpack table have been unpacked); everything else
        (the trunk and the small per-stack heads that sit behind the hourglasses in the buffer) when the whole pass ends."""
        self.pack_all_weights()
        out = []
        convs3 = [c for c in self.modules() if isinstance(c, nn.Conv2d) and c is not self.conv1 and c.weight.shape[-1] == 3]
        for k in reversed(range(self.num_stacks)):
            ps = list(self.hg[k].parameters())
            lo = self._offsets[id(ps[0])][0]
            hi = self._offsets[id(ps[-1])][0] + _ceil(ps[-1].numel(), 4)
            mine = {id(c) for c in self.hg[k].modules() if isinstance(c, nn.Conv2d)}
            rows = [i for i, c in enumerate(convs3) if id(c) in mine]
            assert rows == list(range(rows[0], rows[0] + len(rows)))
            out.append(('stack %d' % k, lo, hi, (rows[0], len(rows))))
        return out

    def run_backward(self, grad_scores, run=None, on_bucket=None):
        """grad_scores: list of fp32 NCHW [N,num_outputs,h,w] (or bf16 NHWC [N,h,w,128], or None) -> gradients accumulated into the
        flat grad buffer.
        run: the tape of the forward pass to differentiate (default: the most recent run_forward).
        on_bucket(lo, hi): called as soon as flat_grad[lo:hi] is final (see grad_buckets; the slices together cover the buffer),
        so a data-parallel caller can all-reduce that bucket underneath the rest of the pass."""
        run = run if run is not None else self._last_run
        if run is None:
            raise RuntimeError('run_backward: no recorded forward pass (the last forward ran with gradients disabled)')
        self._flat_grad.zero_()
        self._wg3_scratch.zero_()
        if on_bucket is None:
            run.backward(grad_scores)
            if self._wg3_table.shape[0]:
                ops.unpack_wgrad_batch(self._wg3_table, self._wg3_scratch, self._flat_grad)
            return self._flat_grad
        buckets = {int(name.split()[1]): (lo, hi, rows) for name, lo, hi, rows in self.grad_buckets()}
        done = []

        def stack_done(k):
            lo, hi, (r0, nr) = buckets[k]
            ops.unpack_wgrad_batch(self._wg3_table[r0:r0 + nr], self._wg3_scratch, self._flat_grad)
            done.append((r0, nr))
            on_bucket(lo, hi)
        run.backward(grad_scores, stack_done)
        # the remaining 3x3 layers (trunk, res.k), then the two slices around the hourglass block
        taken = sorted(done)
        n3, r = self._wg3_table.shape[0], 0
        for r0, nr in taken + [(n3, 0)]:
            if r0 > r:
                ops.unpack_wgrad_batch(self._wg3_table[r:r0], self._wg3_scratch, self._flat_grad)
            r = r0 + nr
        lo_all = min(b[0] for b in buckets.values())
        hi_all = max(b[1] for b in buckets.values())
        if lo_all > 0:
            on_bucket(0, lo_all)
        if hi_all < self._flat_grad.numel():
            on_bucket(hi_all, self._flat_grad.numel())
        return self._flat_grad


class _HourglassFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, *params):
        scores, latents = net.run_forward(x)
        ctx.net = net
        # the tape of THIS forward: a later forward (an eval pass, the second branch of a two-forward step, gradient
        # accumulation) must not be the one this node back-propagates through
        ctx.run = net._last_run
        net._last_run = None                 # the autograd node owns the tape; nothing is pinned once the graph is freed
        if not any(ctx.needs_input_grad):
            ctx.run.release()                # inference: no backward will come
            ctx.run = None
        ctx.n_scores = len(scores)
        ctx.mark_non_differentiable(*latents)
        return (*scores, *latents)

    @staticmethod
    def backward(ctx, *grads):
        net = ctx.net
        gs = [g.contiguous() if g is not None else None for g in grads[:ctx.n_scores]]
        if ctx.run is None:
            raise RuntimeError('HourglassNet: backward through the same forward pass twice (activations are freed after the first)')
        net.run_backward(gs, run=ctx.run)
        ctx.run.release()
        ctx.run = None
        out = [net.grad_view(p).clone() for p in net.parameters()]
        return (None, None, *out)


class _Run:
    """One forward pass: allocates activations, runs the kernels, and records closures for the backward."""

    def __init__(self, net, N, device):
        self.net, self.N, self.dev = net, N, device
        self._gn_words = 0                 # scratch words the GroupNorm backward kernels of this pass need
        self.tape = []
        self.stats_pool = torch.zeros((160, N, 16, 2), device=device, dtype=torch.float32)
        self.stats_used = 0

    # -------------------------------------------------------------- helpers
    def new_stats(self):
        s = self.stats_pool[self.stats_used]
        self.stats_used += 1
        return s

    def cs(self, t):
        """Column-sum target for the kernel that finalises dL/dt (None if no bias hangs off t)."""
        return t.bias_grads[0] if t.bias_grads else None

    def cs_done(self, t):
        for extra in t.bias_grads[1:]:
            extra.copy_(t.bias_grads[0])

    def cs_fallback(self, t, dbuf):
        """dL/dt was produced by a kernel without a fused column sum (a convolution data-gradient)."""
        if t.bias_grads:
            ops.colsum(dbuf, self.N, t.H * t.W, t.C, t.bias_grads[0])
            self.cs_done(t)

    def act(self, H, W, C, stats=True):
        buf = torch.empty((self.N, H, W, C), device=self.dev, dtype=BF16)
        return _Tensor(buf, self.new_stats() if stats else None, H, W, C)

    def packed(self, conv):
        """bf16 forward / data-gradient weight layouts of a conv (packed from the fp32 masters at the start of this forward)."""
        return self.net._wcache[id(conv)]

    # -------------------------------------------------------------- layers
    def conv(self, conv, a, a_C, out, residual=None, y_nchw=None, want_stats=True, gn=None):
        """out = conv(a) + bias (+ residual).  a: bf16 buffer [N,H,W,a_C]; out: _Tensor (or None with y_nchw).
        gn = (x, GroupNorm module): `a` is x.buf, the RAW GroupNorm input, and relu(groupnorm(.)) is applied to the operand tiles
        inside the convolution and inside its weight gradient (no normalised copy of x is ever written)."""
        net, N = self.net, self.N
        Cout, Cin, k, _ = conv.weight.shape
        taps = k * k
        wf, wb, cout_pad, cin_pad, b_rows, b_cols = self.packed(conv)
        H, W = (out.H, out.W) if out is not None else y_nchw.shape[-2:]
        assert cin_pad == a_C, (cin_pad, a_C)
        gn_op = (gn[0].stats, gn[1].weight.data, gn[1].bias.data, 16, 1e-5) if gn is not None else None
        ops.conv_fwd(a, wf, conv.bias.data, N, H, W, a_C, Cout, cout_pad, taps,
                     y=out.buf if out is not None else None, y_ld=out.C if out is not None else 0, y_nchw=y_nchw,
                     residual=residual, stats=out.stats if (out is not None and want_stats) else None, groups=16, gn=gn_op)

        def bwd(dout, dout_C, need_dx=True, dx_addend=None):
            """dout: bf16 [N,H,W,dout_C] gradient of the conv output -> returns bf16 gradient w.r.t. `a`."""
            if taps == 9:
                off3, n3 = net._wg3[id(conv)]
                ops.conv_wgrad3x3(dout, a, N, H, W, a_C, Cin, dout_C, Cout, net._wg3_scratch[off3:off3 + n3])
            else:
                ops.conv_wgrad(dout, a, N, H, W, a_C, Cin, dout_C, Cout, taps, net.grad_view(conv.weight), gn=gn_op)
            if not need_dx:
                return None
            dx = torch.empty((N, H, W, a_C), device=self.dev, dtype=BF16)
            # data gradient = the same implicit GEMM on flipped/transposed weights: "Cin" := dout_C, "Cout" := a_C
            assert b_cols == dout_C, (b_cols, dout_C)
            ops.conv_fwd(dout, wb, None, N, H, W, dout_C, a_C, b_rows, taps, y=dx, y_ld=a_C,
                         residual=dx_addend)
            return dx
        return bwd

    def gn_relu(self, x, gn, G=16, out_stats=False, G_out=16, fused=False):
        """a = relu(groupnorm(x)); returns (a buffer, [stats of a], backward closure).  fused: the consumer applies the
        normalisation itself (conv(..., gn=)); only the backward closure is built and `a` is x.buf."""
        N = self.N
        st = self.new_stats() if out_stats else None
        if fused:
            a = x.buf
        else:
            a = torch.empty_like(x.buf)
            ops.gn_relu_fwd(x.buf, x.stats, gn.weight.data, gn.bias.data, N, x.H * x.W, x.C, G, a, st, G_out)
        net = self.net

        # scratch of this layer's backward: a slice of ONE arena zeroed once per backward pass (one fill instead of a memset
        # node per GroupNorm layer in the captured graph)
        words = (ops.gn_relu_bwd_scratch_words(N, G) + 3) // 4 * 4
        slot = self._gn_words
        self._gn_words += words

        def bwd(da, addend=None, colsum=None):
            red = self._gn_arena[slot:slot + words]
            dx = torch.empty_like(x.buf)
            ops.gn_relu_bwd(da, x.buf, x.stats, gn.weight.data, gn.bias.data, N, x.H * x.W, x.C, G, red,
                            net.grad_view(gn.weight), net.grad_view(gn.bias), dx, addend, colsum, prezeroed=True)
            return dx
        return a, st, bwd

    def bottleneck(self, blk, x, x_final=True):
        """Pre-activation bottleneck (hourglass.py:23-41).  x: _Tensor with stats -> _Tensor with stats.
        x_final: the gradient this block returns is the whole dL/dx (no other consumer of x adds to it afterwards)."""
        net, N = self.net, self.N
        planes = blk.conv1.weight.shape[0]
        H, W = x.H, x.W
        # GroupNorm + ReLU in front of the two 1x1 layers runs inside them (images of >= 64 pixels)
        fuse = FUSE_GN and H * W >= 64
        fuse3 = fuse and planes > 64        # (64 -> 128 with the residual at 64x64 measures slower fused: tools/bench_fused_gn.py)
        a1, _, gn1_b = self.gn_relu(x, blk.bn1, fused=fuse)
        t1 = self.act(H, W, planes)
        c1_b = self.conv(blk.conv1, a1, x.C, t1, gn=(x, blk.bn1) if fuse else None)
        a2, _, gn2_b = self.gn_relu(t1, blk.bn2)
        t2 = self.act(H, W, planes)
        c2_b = self.conv(blk.conv2, a2, planes, t2)
        a3, _, gn3_b = self.gn_relu(t2, blk.bn3, fused=fuse3)
        out = self.act(H, W, planes * 2)
        if blk.downsample is not None:
            res = self.act(H, W, planes * 2, stats=False)
            d_b = self.conv(blk.downsample[0], x.buf, x.C, res, want_stats=False)
            res_buf = res.buf
        else:
            d_b, res_buf = None, x.buf
        c3_b = self.conv(blk.conv3, a3, planes, out, residual=res_buf, gn=(t2, blk.bn3) if fuse3 else None)
        # bias gradients of conv3 (and of the downsample conv, which sees the same output gradient) = colsum(dL/dout)
        out.bias_grads = [net.grad_view(blk.conv3.bias)] + ([net.grad_view(blk.downsample[0].bias)] if d_b is not None else [])

        def bwd(dout):
            """dout = dL/dout, its column sum already delivered to out.bias_grads by whoever produced it."""
            da3 = c3_b(dout, planes * 2)
            dt2 = gn3_b(da3, colsum=net.grad_view(blk.conv2.bias))
            da2 = c2_b(dt2, planes)
            dt1 = gn2_b(da2, colsum=net.grad_view(blk.conv1.bias))
            da1 = c1_b(dt1, planes)
            dres = d_b(dout, planes * 2) if d_b is not None else dout
            dx = gn1_b(da1, addend=dres, colsum=self.cs(x) if x_final else None)
            if x_final:
                self.cs_done(x)
            return dx
        return out, bwd

    def hourglass(self, hgm, n, x, defer_cs=False):
        """hourglass.py:68-82.  Returns (out, latent, backward).  defer_cs: the caller adds another gradient to dL/dx
        and takes its column sum there."""
        N = self.N
        lvl = hgm.hg[n - 1]
        up1, up1_b = self.bottleneck(lvl[0][0], x, x_final=False)     # dL/dx is finalised by the max-pool backward
        pooled = self.act(x.H // 2, x.W // 2, x.C)
        ops.maxpool_fwd(x.buf, N, pooled.H, pooled.W, x.C, pooled.buf, pooled.stats, 16)
        low1, low1_b = self.bottleneck(lvl[1][0], pooled)
        if n > 1:
            low2, latent, low2_b = self.hourglass(hgm, n - 1, low1)
        else:
            low2, low2_b = self.bottleneck(lvl[3][0], low1)
            latent = low2
        low3, low3_b = self.bottleneck(lvl[2][0], low2)
        out = self.act(x.H, x.W, x.C)
        ops.upsample_add_fwd(up1.buf, low3.buf, N, low3.H, low3.W, x.C, out.buf, out.stats, 16)
        out.bias_grads = up1.bias_grads                 # dL/dup1 == dL/dout

        def bwd(dout, dlatent=None):
            dlow3 = torch.empty_like(low3.buf)
            ops.upsample_bwd(dout, N, low3.H, low3.W, x.C, dlow3, colsum=self.cs(low3))
            self.cs_done(low3)
            dlow2 = low3_b(dlow3)
            dlow1 = low2_b(dlow2)
            dpooled = low1_b(dlow1)
            dx_a = up1_b(dout)
            dx = torch.empty_like(x.buf)
            ops.maxpool_bwd(dpooled, x.buf, N, pooled.H, pooled.W, x.C, dx, addend=dx_a,
                            colsum=None if defer_cs else self.cs(x))
            if not defer_cs:
                self.cs_done(x)
            return dx
        return out, latent, bwd

    # -------------------------------------------------------------- whole network
    def forward(self, img, S, on_score=None, want_latents=True):
        net, N = self.net, self.N
        K = net.num_outputs
        self.img = img
        H = S // 2
        # stem: conv1 -> GroupNorm(4) -> ReLU (hourglass.py:153-155)
        c1 = _Tensor(torch.empty((N, H, H, 64), device=self.dev, dtype=BF16),
                     torch.zeros((N, 4, 2), device=self.dev, dtype=torch.float32), H, H, 64)
        ops.stem_conv_fwd(img, net.conv1.weight.data, net.conv1.bias.data, N, S, c1.buf, c1.stats, 4)
        x0_buf, x0_stats, stem_gn_b = self.gn_relu(c1, net.bn1, G=4, out_stats=True, G_out=16)
        x0 = _Tensor(x0_buf, x0_stats, H, H, 64)
        l1, l1_b = self.bottleneck(net.layer1[0], x0)
        H2 = H // 2
        p1 = self.act(H2, H2, 128)
        ops.maxpool_fwd(l1.buf, N, H2, H2, 128, p1.buf, p1.stats, 16)
        l2, l2_b = self.bottleneck(net.layer2[0], p1)
        x, l3_b = self.bottleneck(net.layer3[0], l2)
        h = H2
        scores, latents, stack_b = [], [], []
        for i in range(net.num_stacks):
            last = i == net.num_stacks - 1
            # (not last: dL/dx also receives dL/dx_next through the residual chain; the add below takes the column sum)
            y, latent, hg_b = self.hourglass(net.hg[i], 2, x, defer_cs=not last)
            y, res_b = self.bottleneck(net.res[i][0], y)
            f = self.act(h, h, 256)
            fc_b = self.conv(net.fc[i][0], y.buf, 256, f)
            yf_buf, _, fcgn_b = self.gn_relu(f, net.fc[i][1])
            score = torch.empty((N, K, h, h), device=self.dev, dtype=torch.float32)
            score_pad = None if last else _Tensor(torch.zeros((N, h, h, 128), device=self.dev, dtype=BF16), None, h, h, 128)
            sc_b = self.conv(net.score[i], yf_buf, 256, score_pad, y_nchw=score, want_stats=False)
            scores.append(score)
            if on_score is not None:
                on_score(i, score)
            if want_latents:
                lat = torch.empty((N, 256, latent.H, latent.W), device=self.dev, dtype=torch.float32)
                ops.nhwc_to_nchw(latent.buf, N, 256, latent.H * latent.W, lat)
                latents.append(lat)
            if not last:
                # x <- x + fc_(y) + score_(score)   (hourglass.py:169-172), both adds in the conv epilogues
                t = self.act(h, h, 256, stats=False)
                fcu_b = self.conv(net.fc_[i], yf_buf, 256, t, residual=x.buf, want_stats=False)
                xn = self.act(h, h, 256)
                scu_b = self.conv(net.score_[i], score_pad.buf, 128, xn, residual=t.buf)
                xn.bias_grads = [net.grad_view(net.fc_[i].bias), net.grad_view(net.score_[i].bias)]   # both see dL/dxn
            else:
                fcu_b = scu_b = xn = None
            stack_b.append((hg_b, res_b, fc_b, fcgn_b, sc_b, fcu_b, scu_b, x, y, h))
            if not last:
                x = xn

        def backward(grad_scores, stack_done=None):
            dx_next = None            # gradient w.r.t. the input of the following stack
            for i in reversed(range(net.num_stacks)):
                hg_b, res_b, fc_b, fcgn_b, sc_b, fcu_b, scu_b, xin, yres, hh = stack_b[i]
                g = grad_scores[i]
                if g is not None and g.dtype == BF16:
                    # already in the backward pass's own layout (bf16 NHWC, 128 padded channels): the fused step's soft-argmax backward
                    if tuple(g.shape) != (N, hh, hh, 128) or not g.is_contiguous():
                        raise RuntimeError('bf16 score gradient must be a contiguous [N,h,w,128] NHWC tensor')
                    dscore = g
                else:
                    dscore = torch.zeros((N, hh, hh, 128), device=self.dev, dtype=BF16)
                    if g is not None:
                        ops.nchw_to_nhwc(g.contiguous().float(), N, K, hh * hh, 128, dscore)
                dyf = None
                if dx_next is not None:
                    dsp = scu_b(dx_next, 256)                       # [N,h,h,128] gradient of the padded score copy
                    tmp = torch.empty_like(dscore)
                    ops.add(dscore, dsp, N, hh * hh, 128, tmp)
                    dscore = tmp
                    dyf = fcu_b(dx_next, 256)
                self._tmp128().zero_()
                ops.colsum(dscore, N, hh * hh, 128, self._tmp128())
                net.grad_view(net.score[i].bias).copy_(self._tmp128()[:K])
                dyf = sc_b(dscore, 128, dx_addend=dyf)
                df = fcgn_b(dyf, colsum=net.grad_view(net.fc[i][0].bias))
                dy = fc_b(df, 256)
                self.cs_fallback(yres, dy)                          # a conv data-gradient: no fused column sum
                dy = res_b(dy)
                dxin = hg_b(dy)
                if dx_next is not None:
                    tmp = torch.empty_like(dxin)
                    ops.add(dxin, dx_next, N, hh * hh, 256, tmp, colsum=self.cs(xin))
                    self.cs_done(xin)
                    dxin = tmp
                dx_next = dxin
                if stack_done is not None:
                    stack_done(i)         # every parameter gradient of hg[i] is final
            d = l3_b(dx_next)
            d = l2_b(d)
            dl1 = torch.empty_like(l1.buf)
            ops.maxpool_bwd(d, l1.buf, N, H2, H2, 128, dl1, colsum=self.cs(l1))
            self.cs_done(l1)
            dx0 = l1_b(dl1)
            dc1 = stem_gn_b(dx0)
            ops.stem_conv_wgrad(self.img, dc1, N, S, net.grad_view(net.conv1.weight), net.grad_view(net.conv1.bias))
        self._backward = backward
        return scores, latents

    def _tmp128(self):
        if not hasattr(self, '_t128'):
            self._t128 = torch.zeros(128, device=self.dev, dtype=torch.float32)
        return self._t128

    def release(self):
        """Drop the recorded closures (and with them every saved activation).  The tape references this object and this object the
        tape: without this the ~12 GB of a BASELINE-size pass would wait for Python's cyclic garbage collector, and the next
        pass would have to cudaMalloc its activations instead of reusing the cached blocks (measured on the module-by-module
        path: 19.5 ms steps with 30-370 ms spikes every few steps)."""
        self._backward = None
        self.tape = []
        self.img = None
        self.stats_pool = None
        self._gn_arena = None

    def backward(self, grad_scores, stack_done=None):
        if self._backward is None:
            raise RuntimeError('this forward pass has been released (already back-propagated, or run without gradients)')
        if hasattr(self, '_t128'):
            self._t128.zero_()
        self._gn_arena = torch.zeros(max(self._gn_words, 1), device=self.dev, dtype=torch.float32)
        self._backward(grad_scores, stack_done)


def create_hourglass_network(num_outputs, num_stacks=1):
    """Same signature as the reference factory (hourglass.py:175-176)."""
    return HourglassNet(num_stacks=num_stacks, num_outputs=num_outputs)
