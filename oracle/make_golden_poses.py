"""Generate tests/golden/joint_angle.npz by running the UNMODIFIED reference JointAngleDataset on CPU (build container only).
TEST INFRASTRUCTURE.  Usage: python oracle/make_golden_poses.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, '/root/reference')
from dataset.joint_angle import JointAngleDataset  # noqa: E402  (the reference's)


def main():
    n, seed = 256, 2024
    ds = JointAngleDataset()
    torch.manual_seed(seed)
    poses = torch.stack([ds[i] for i in range(n)]).numpy()
    after = torch.rand(4).numpy()                       # where the reference leaves the generator
    torch.manual_seed(seed)
    u = torch.rand(n * 44).numpy()                      # the same stream, drawn in one call
    path = os.path.join(ROOT, 'tests', 'golden', 'joint_angle.npz')
    np.savez_compressed(path, seed=np.int64(seed), poses=poses, u=u, after=after)
    print(path, os.path.getsize(path) // 1024, 'KB', poses.shape)


if __name__ == '__main__':
    main()
