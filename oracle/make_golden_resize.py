"""Generate tests/golden/resize_crop.npz by running the UNMODIFIED reference ResizeCropImage on CPU (build container only).
TEST INFRASTRUCTURE (see oracle/make_golden.py).  Usage: python oracle/make_golden_resize.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import _refshim  # noqa: E402

_refshim.install()
import torch  # noqa: E402
from network.util_modules import ResizeCropImage  # noqa: E402  (the reference's)


def main():
    g = torch.Generator().manual_seed(77)
    out = {}
    for S in (64, 128):
        n = 8 if S == 64 else 5
        dm = torch.floor(torch.rand(n, S, S, generator=g) * 64) / 64            # 6-bit values: the fixture compresses
        rnd = torch.rand(n, generator=g) * 0.2 + 0.75                     # create_network_and_criterion.py:99-101
        u = rnd + torch.rand(n, generator=g) * 0.1 - 0.05
        v = rnd + torch.rand(n, generator=g) * 0.1 - 0.05
        u[0], v[0] = 1.0, 1.0
        u[1], v[1] = 1.1, 0.9                                             # u > 1 branch (crop), v <= 1
        u[2], v[2] = 0.9, 1.2                                             # v > 1: the reference writes nothing (stays 1.0)
        u[3], v[3] = 0.7031, 0.9961                                       # near the rounding edges of int(S*s + 0.5)
        res = ResizeCropImage()(dm, u, v)
        out.update({'dm%d' % S: dm.numpy(), 'u%d' % S: u.numpy(), 'v%d' % S: v.numpy(), 'out%d' % S: res.numpy()})
    path = os.path.join(ROOT, 'tests', 'golden', 'resize_crop.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
