"""Drop-in mirror of /root/reference/dataset/nyu_dataset.py (NpyDataset :9-31, create_nyu_dataset :34-51) and of the shard
format its writer produces (dataset/nyu_generator.py:89-119): `<prefix>_shape.pkl` (dict of shapes), `<prefix>_dms.bat`
(raw float32 memmap [n,V,S,S]), `<prefix>_joint_poses.npy`, `<prefix>_camera_poses.npy` [n,V,4,4].

`NpyDataset` / `create_nyu_dataset` read exactly those files and return exactly the reference's items, so the reference's
DataLoader code keeps working.  `ShardBatchLoader` is the B200-side feeder for `SelfSupTrainStep.load_batch`: a worker thread
gathers the next batch from the memmaps into PINNED staging buffers while the current one trains, and the host->device copies
run on their own stream; it yields device tensors (depth maps, joint poses, camera poses, inverse camera poses).
"""
import os
import pickle
import threading

import numpy as np
import torch
import torch.utils.data as data


class NpyDataset(data.Dataset):
    def __init__(self, file_path, transform=None):
        super().__init__()
        with open(file_path + '_shape.pkl', 'rb') as f:
            shape_info = pickle.load(f)
        self.dms = np.memmap(file_path + '_dms.bat', dtype='float32', mode='r', shape=tuple(shape_info['dms']))
        self.joint_poses = np.load(file_path + '_joint_poses.npy')
        self.camera_poses = np.load(file_path + '_camera_poses.npy')
        # one LAPACK inverse per 4x4 like the reference (:20-21): bit-identical inverse poses
        inv = [np.linalg.inv(m).reshape(1, 4, 4) for m in self.camera_poses.reshape(-1, 4, 4)]
        self.inv_camera_poses = np.concatenate(inv, axis=0).reshape(self.camera_poses.shape)
        self.transform = transform

    def __getitem__(self, index):
        item = (np.asarray(self.dms[index]), self.joint_poses[index], self.camera_poses[index], self.inv_camera_poses[index])
        return item if self.transform is None else self.transform(*item)

    def __len__(self):
        return self.joint_poses.shape[0]


def create_nyu_dataset(file_dir):
    """All shards `mv_data_0`, `mv_data_1`, ... of every directory, concatenated (:34-51)."""
    if type(file_dir) is not list:
        file_dir = [file_dir]
    datasets = []
    for d in file_dir:
        idx = 0
        curr_path = os.path.join(d, 'mv_data_%d' % idx)
        while os.path.exists(curr_path + '_shape.pkl'):
            datasets.append(NpyDataset(curr_path))
            idx += 1
            curr_path = os.path.join(d, 'mv_data_%d' % idx)
            if os.name == 'nt' and idx > 5:
                break
    return data.ConcatDataset(datasets)


def write_npy_shard(npy_dir, file_name, dms, joint_poses, camera_poses):
    """The file-writing half of NyuGenerator.create_npy_dataset_from_indices (dataset/nyu_generator.py:100-118)."""
    dms = np.asarray(dms).astype(np.float32)
    joint_poses = np.asarray(joint_poses).astype(np.float32)
    camera_poses = np.asarray(camera_poses).astype(np.float32)
    shape_info = {'dms': dms.shape, 'joint_poses': joint_poses.shape, 'camera_poses': camera_poses.shape}
    with open(os.path.join(npy_dir, file_name + '_shape.pkl'), 'wb') as f:
        pickle.dump(shape_info, f, protocol=pickle.HIGHEST_PROTOCOL)
    fp = np.memmap(os.path.join(npy_dir, file_name + '_dms.bat'), dtype='float32', mode='w+', shape=dms.shape)
    fp[:] = dms[:]
    fp.flush()
    del fp
    np.save(os.path.join(npy_dir, file_name + '_joint_poses.npy'), joint_poses)
    np.save(os.path.join(npy_dir, file_name + '_camera_poses.npy'), camera_poses)


class ShardBatchLoader:
    """Batches of a (Concat)Dataset of NpyDataset shards as DEVICE tensors, with pinned double-buffered staging.

    for dms, joints, cams, inv_cams in ShardBatchLoader(ds, batch_size, device): step.load_batch(dms, cams, inv_cams, poses)

    Order: a seeded permutation per epoch when shuffle=True (torch.randperm on `generator`), else sequential; the last
    partial batch is dropped (the train step's buffers are static).  The tensors of a batch are valid until the next-but-one
    batch is requested (two staging slots)."""

    def __init__(self, dataset, batch_size, device='cuda', shuffle=True, generator=None):
        self.ds, self.bs, self.shuffle, self.gen = dataset, int(batch_size), shuffle, generator
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('ShardBatchLoader stages through pinned memory to a CUDA device; there is no CPU path')
        if len(dataset) < self.bs:
            raise ValueError('dataset smaller than one batch')
        first = dataset[0]
        self._host = [[torch.empty((self.bs,) + np.shape(a), dtype=torch.float32).pin_memory() for a in first] for _ in range(2)]
        self._dev = [[torch.empty(h.shape, dtype=torch.float32, device=self.device) for h in slot] for slot in self._host]
        self._stream = torch.cuda.Stream(device=self.device)
        self._ready = [torch.cuda.Event() for _ in range(2)]           # H2D copy of the slot has finished
        self._consumed = [torch.cuda.Event() for _ in range(2)]        # the consumer's work on the slot has been enqueued

    def __len__(self):
        return len(self.ds) // self.bs

    def _gather(self, slot, indices):
        views = [h.numpy() for h in self._host[slot]]      # numpy views of the pinned buffers: memmap -> pinned memory, one copy
        for k, i in enumerate(indices):
            for dst, src in zip(views, self.ds[int(i)]):
                dst[k] = src

    def _issue_copy(self, slot):
        with torch.cuda.stream(self._stream):
            for h, d in zip(self._host[slot], self._dev[slot]):
                d.copy_(h, non_blocking=True)
            self._ready[slot].record(self._stream)

    def __iter__(self):
        n = len(self)
        order = torch.randperm(len(self.ds), generator=self.gen) if self.shuffle else torch.arange(len(self.ds))
        batches = [order[b * self.bs:(b + 1) * self.bs].tolist() for b in range(n)]
        # slot state carried over from the previous epoch: the host runs ahead of the stream (graph replays), so the last batches'
        # device buffers may still be waiting to be read and their H2D copies may still be in flight
        for s in range(2):
            self._ready[s].synchronize()                   # pinned host buffers: their last copy has left
            self._stream.wait_event(self._consumed[s])     # device buffers: the last consumer's reads come first (no-op if never recorded)
        self._gather(0, batches[0])                        # the first batch has nothing to overlap with
        self._issue_copy(0)
        worker = None
        try:
            for b in range(n):
                s = b & 1
                if b + 1 < n:
                    if b >= 1:
                        self._ready[s ^ 1].synchronize()   # pinned buffers of the other slot: their copy (batch b-1) is done
                    worker = threading.Thread(target=self._gather, args=(s ^ 1, batches[b + 1]))
                    worker.start()                         # gathers batch b+1 from the memmaps while batch b trains
                torch.cuda.current_stream(self.device).wait_event(self._ready[s])
                yield tuple(self._dev[s])
                # the consumer is back: what it enqueued so far is everything that reads the device buffers of this slot
                self._consumed[s].record(torch.cuda.current_stream(self.device))
                if b + 1 < n:
                    worker.join()
                    worker = None
                    if b >= 1:
                        self._stream.wait_event(self._consumed[s ^ 1])     # batch b-1's consumer is done with those device buffers
                    self._issue_copy(s ^ 1)                # H2D of batch b+1 overlaps the training step of batch b
        finally:
            if worker is not None:
                worker.join()
