"""Drop-in mirror of /root/reference/dataset/joint_angle.py (JointAngleDataset, :7-233): the synthetic pose sampler of the
self-supervised step.  The reference builds ONE pose with ~35 `torch.rand(1)` calls and as many tiny
tensor ops in a DataLoader worker; here a batch of poses is one `torch.rand` call, one integer walk over the mode draws and
ONE kernel launch (`sh_sample_poses`), and, fed the same generator state, it returns the reference's poses bit for bit:
`torch.manual_seed(s); [ref[i] for i in range(n)]` == `torch.manual_seed(s); ours.sample_batch(n)`.

No CPU fallback: the poses are produced on the CUDA device (they feed the FK kernel directly).
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


class JointAngleDataset(data.Dataset):
    INDEX, MIDDLE, RING, PINKY, THUMB = 6, 10, 14, 18, 22
    ABDUCT, FLEX_1, FLEX_2, FLEX_3 = 0, 1, 2, 3

    def __init__(self, device='cuda'):
        super().__init__()
        self.num_parameter = 26
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('JointAngleDataset samples on a CUDA device; there is no CPU path')

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
        return self.sample_batch(1)[0]

    def __len__(self):
        return 400000
