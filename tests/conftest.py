import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(autouse=True)
def _strict_fp32():
    """torch-on-GPU references in the parity tests are fp32: cuDNN convolutions default to TF32 (torch.backends.cudnn.allow_tf32 is
    True out of the box), which alone moves trained-weight gradients by ~3 % (measured, gpurun_out/g1_tests.log)."""
    import torch
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def golden(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


@pytest.fixture(scope='session')
def hand_model():
    return golden('hand_model')


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def elem_err(a, b, floor=1e-2):
    """Per-element relative error with an absolute floor: max_i |a_i - b_i| / (|b_i| + floor * max|b|).  Unlike rel_err (a
    max-normalised error) a small component may not be arbitrarily wrong: it is held to `tol * (its own size + floor * max)`,
    i.e. with tol = 1e-4 and the default floor an absolute error of 1e-6 of the largest component (sums of fp32 terms with
    cancellation cannot be asked for more)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if b.size == 0:
        return 0.0
    return float((np.abs(a - b) / (np.abs(b) + floor * max(np.abs(b).max(), 1e-30))).max())


def mesh_dict(a):
    """The reference's `mesh` dict (mesh/preprocess.py output: vertices, faces, bones[{offset_matrix, weight_vertexid,
    weight_coeff, keypoint}]) rebuilt from the arrays of tests/golden/hand_model.npz."""
    bones = []
    for b in range(a['offset_mats'].shape[0]):
        sel = a['weight_bone'] == b
        bone = {'offset_matrix': np.asarray(a['offset_mats'][b], np.float64),
                'weight_vertexid': a['weight_vertexid'][sel].tolist(), 'weight_coeff': a['weight_coeff'][sel].tolist()}
        kp = [(np.asarray(a['keypoints'][k], np.float64), float(a['keypoint_radius'][k]))
              for k in range(len(a['keypoint_bone'])) if a['keypoint_bone'][k] == b]
        if kp:
            bone['keypoint'] = kp
        bones.append(bone)
    return {'vertices': np.asarray(a['vertices']), 'faces': np.array(a['faces'], copy=True), 'bones': bones}
