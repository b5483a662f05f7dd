"""Generate tests/golden/shard/ (two tiny `mv_data_<k>` shards in the reference's on-disk format, written statement by statement
like dataset/nyu_generator.py:100-118) and tests/golden/shard_items.npz (what the UNMODIFIED reference reader returns for
them).  TEST INFRASTRUCTURE, build container only.  Usage: python oracle/make_golden_shard.py"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, '/root/reference')
from dataset.nyu_dataset import create_nyu_dataset  # noqa: E402  (the reference's)


def rot(axis, ang):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def main():
    rng = np.random.RandomState(11)
    out_dir = os.path.join(ROOT, 'tests', 'golden', 'shard')
    os.makedirs(out_dir, exist_ok=True)
    V, S, J = 3, 16, 36
    for k, n in enumerate((5, 3)):
        dms = np.full((n, V, S, S), 100.0)
        fg = rng.rand(n, V, S, S) < 0.3
        dms[fg] = np.round(rng.rand(int(fg.sum())) * 60 - 30, 2)
        joint_poses = rng.randn(n, V, J, 3) * 40
        camera_poses = np.tile(np.eye(4), (n, V, 1, 1))
        for i in range(n):
            for v in range(1, V):
                camera_poses[i, v, :3, :3] = rot(rng.randn(3), rng.rand() * 0.5)
                camera_poses[i, v, :3, 3] = rng.randn(3) * 5
        dms, joint_poses, camera_poses = dms.astype(np.float32), joint_poses.astype(np.float32), camera_poses.astype(np.float32)
        name = 'mv_data_%d' % k
        shape_info = {'dms': dms.shape, 'joint_poses': joint_poses.shape, 'camera_poses': camera_poses.shape}
        with open(os.path.join(out_dir, name + '_shape.pkl'), 'wb') as f:
            pickle.dump(shape_info, f, protocol=pickle.HIGHEST_PROTOCOL)
        fp = np.memmap(os.path.join(out_dir, name + '_dms.bat'), dtype='float32', mode='w+', shape=dms.shape)
        fp[:] = dms[:]
        fp.flush()
        del fp
        np.save(os.path.join(out_dir, name + '_joint_poses.npy'), joint_poses)
        np.save(os.path.join(out_dir, name + '_camera_poses.npy'), camera_poses)
    ds = create_nyu_dataset(out_dir)
    items = {}
    for i in range(len(ds)):
        dm, jp, cp, icp = ds[i]
        items.update({'dm%d' % i: np.array(dm), 'jp%d' % i: jp, 'cp%d' % i: cp, 'icp%d' % i: icp})
    items['n'] = np.int64(len(ds))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'shard_items.npz'), **items)
    print('wrote', out_dir, len(ds), 'items')


if __name__ == '__main__':
    main()
