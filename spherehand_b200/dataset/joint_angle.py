"""Drop-in mirror of /root/reference/dataset/joint_angle.py (JointAngleDataset, :7-233): the synthetic pose sampler of the
self-supervised step.  The reference builds ONE pose with ~35 `torch.rand(1)` calls and as many tiny
tensor ops in a DataLoader worker; here a batch of poses is one `torch.rand` call, one integer walk over the mode draws and
ONE kernel launch (`sh_sample_poses`), and, fed the same generator state, it returns the reference's poses bit for bit:
`torch.manual_seed(s); [ref[i] for i in range(n)]` == `torch.manual_seed(s); ours.sample_batch(n)`.

`sample_batch` (what the fused train step calls) produces the poses on the CUDA device: they feed the FK kernel directly and
there is no CPU path for it.  `__getitem__` is the reference's DataLoader protocol — one pose, a CPU tensor, computed inside
a forked worker process that must not touch CUDA (engine.py:331-333: `DataLoader(self.synt_dataset, batch_size=48,
num_workers=1)`) — so it is host code here as it is in the reference: `_host_pose`, the same fp32 operation sequence in numpy.
"""
import numpy as np
import torch
import torch.utils.data as data

from .. import ops

MAX_UNIFORMS = 44          # palm 6 + abduction 5 + thumb 4 + flexion mode 1 + 4 fingers x (1 + 6)


def sequential_offsets(u):
    """Where every pose starts in ONE sequential uniform stream (what consecutive `__getitem__` calls of the reference
    consume).  u: 1-D float32 numpy array -> (int32 first index [m], int64 one-past-last index [m]) for as many poses as fit.  The count per
    pose depends on its own mode draws only (`int(rand * k)` in fp32, joint_angle.py:130-214)."""
    u = np.asarray(u, np.float32)
    fixed = {0: 3, 1: 3, 2: 6, 3: 6, 4: 6}                     # draws of a finger shape: straight, open, half open, pinching, closed
    rules = {5: (-1, -2, -2, -2), 6: (-2, -2, -2, -1), 7: (-1, -1, -2, -2), 8: (-2, -1, -1, -1), 9: (-3, -3, -3, -3)}
    offs, ends, p, n = [], [], 0, u.shape[0]
    f32 = np.float32
    while p + MAX_UNIFORMS <= n:
        offs.append(p)
        q = p + 15                                             # palm 6, abduction 5, thumb 4
        mode = min(int(u[q] * f32(10)), 9)
        q += 1
        for rule in (rules[mode] if mode >= 5 else (mode,) * 4):
            if rule == -1:
                shape = min(int(u[q] * f32(3)), 2); q += 1
            elif rule == -2:
                shape = 3 + min(int(u[q] * f32(2)), 1); q += 1
            elif rule == -3:
                shape = min(int(u[q] * f32(5)), 4); q += 1
            else:
                shape = rule
            q += fixed[shape]
        p = q
        ends.append(p)
    return np.asarray(offs, np.int32), np.asarray(ends, np.int64)


_F = np.float32
_PI = _F(np.pi)


def _host_pose(u):
    """One pose [26] (float32) from the uniform stream u, joint_angle.py:21-233 statement by statement in numpy float32
    (every product / sum rounded to fp32 like the reference's 1-element tensors).  -> (pose, uniforms consumed)."""
    it = iter(np.asarray(u, np.float32))
    n = [0]

    def rand():
        n[0] += 1
        return _F(next(it))

    def deg(a):
        return _F(_F(a * _PI) / _F(180))

    def curled(b1, b2, b3):                                   # closed / pinching / half-open finger (:42-103)
        def curr(base):
            a = deg(_F(_F(rand() * _F(30)) + _F(base)))
            return _F(a + deg(_F(_F(rand() * _F(20)) - _F(10))))
        f1, f2, f3 = _F(-0.2), _F(-0.4), _F(-0.34)
        c = curr(b1); f1 = _F(f1 + c); f2 = _F(f2 + _F(_F(0.2) * c))
        c = curr(b2); f1 = _F(f1 + _F(_F(0.2) * c)); f2 = _F(f2 + c); f3 = _F(f3 + _F(_F(0.7) * c))
        c = curr(b3); f2 = _F(f2 + _F(_F(0.2) * c)); f3 = _F(f3 + c)
        return [f1, f2, f3]

    def finger(shape):
        if shape == 0:                                        # straight (:111-115)
            return [_F(_F(rand() * _F(0.25)) - _F(0.25)), _F(_F(rand() * _F(0.4)) - _F(0.4)), _F(_F(rand() * _F(0.34)) - _F(0.34))]
        if shape == 1:                                        # open (:105-109)
            return [_F(_F(rand() * _F(0.25)) - _F(0.1)), _F(_F(rand() * _F(0.4)) - _F(0.1)), _F(_F(rand() * _F(0.34)) - _F(0.1))]
        return curled(*((0, 60, 60), (60, 5, 5), (60, 60, 60))[shape - 2])

    p = np.zeros(26, np.float32)
    p[0] = _F(_F(rand() * _F(6.28)) - _F(3.14))
    p[1] = _F(-rand() * _F(3.14))
    p[2] = _F(_F(rand() * _F(6.28)) - _F(3.14))
    p[3] = _F(_F(rand() * _F(30)) - _F(15))
    p[4] = _F(_F(rand() * _F(30)) - _F(15))
    p[5] = _F(_F(rand() * _F(50)) - _F(35))
    spread = _F(_F(rand() - _F(0.35)) / _F(1.55))
    for f, k in enumerate((1.55, 0.75, -0.75, -2.2)):
        p[6 + 4 * f] = _F(_F(k) * _F(spread + deg(_F(_F(rand() * _F(10)) - _F(5)))))
    sel, v = rand(), rand()
    flex = _F(_F(v * _F(0.35)) - _F(0.25)) if sel < _F(0.5) else _F(_F(v * _F(0.6)) + _F(0.1))
    f3 = _F(_F(rand() * _F(2)) - _F(1.7))
    p[22] = _F(rand() - _F(0.5))
    p[23], p[24], p[25] = flex, _F(_F(0.25) * flex), f3
    mode = int(_F(rand() * _F(10)))
    rules = (mode,) * 4 if mode <= 4 else {5: (-1, -2, -2, -2), 6: (-2, -2, -2, -1), 7: (-1, -1, -2, -2), 8: (-2, -1, -1, -1)}.get(mode, (-3,) * 4)
    for f, rule in enumerate(rules):
        shape = rule if rule >= 0 else {-1: lambda: int(_F(rand() * _F(3))), -2: lambda: 3 + int(_F(rand() * _F(2))),
                                        -3: lambda: int(_F(rand() * _F(5)))}[rule]()
        p[7 + 4 * f: 10 + 4 * f] = finger(shape)
    return p, n[0]


class JointAngleDataset(data.Dataset):
    INDEX, MIDDLE, RING, PINKY, THUMB = 6, 10, 14, 18, 22
    ABDUCT, FLEX_1, FLEX_2, FLEX_3 = 0, 1, 2, 3

    def __init__(self, device='cuda'):
        super().__init__()
        self.num_parameter = 26
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('JointAngleDataset.sample_batch samples on a CUDA device; there is no CPU path for it')

    def sample_batch(self, n, generator=None, sequential=True):
        """n poses [n,26] on the device.  sequential=True consumes the (CPU) generator exactly like n consecutive
        `__getitem__` calls of the reference (one stream, 30-44 uniforms per pose); sequential=False gives every pose its own
        slice of MAX_UNIFORMS draws (no host walk: for large n)."""
        if n == 0:
            return torch.empty((0, self.num_parameter), device=self.device)
        if sequential:
            gen = generator if generator is not None else torch.default_generator
            state = gen.get_state()
            u = torch.rand(n * MAX_UNIFORMS, generator=generator)
            offs, ends = sequential_offsets(u.numpy())
            # leave the generator where the reference would: rewind and consume exactly the uniforms the n poses used
            gen.set_state(state)
            torch.rand(int(ends[n - 1]), generator=generator)
            offsets = torch.from_numpy(offs[:n].copy())
        else:
            u = torch.rand(n * MAX_UNIFORMS, generator=generator)
            offsets = torch.arange(n, dtype=torch.int32) * MAX_UNIFORMS
        return ops.sample_poses(u.to(self.device), offsets.to(self.device))

    def __getitem__(self, index):
        """One pose as a CPU tensor [26], consuming the default CPU generator exactly like the reference's `__getitem__` (the
        DataLoader-worker protocol: no CUDA in here)."""
        gen = torch.default_generator
        state = gen.get_state()
        pose, used = _host_pose(torch.rand(MAX_UNIFORMS).numpy())
        gen.set_state(state)
        torch.rand(used)
        return torch.from_numpy(pose)

    def __len__(self):
        return 400000
