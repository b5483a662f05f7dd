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


def l2(a, b):
    import torch
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_grads(ours, ref, emul, what, cos_min=0.98, agg=1.25):
    """ours / ref / emul: dict name -> gradient.  The yardstick is the bf16 emulation's own error against the fp32 reference.  Two runs
    of the SAME step already differ by 6-12 % per tensor (tools/diag_noise.py: the order of the fp32 statistics atomics flips single
    bf16 roundings and the deep network amplifies them), so single tensors get a loose bound (2 x emulated + 5e-2) and the stable
    aggregates the tight ones: whole-gradient l2 and the median per-tensor l2 within 1.25 x the emulation's + 1e-2, cosine >= 0.98."""
    import torch
    rows = []
    for k in ref:
        e_o, e_e = l2(ours[k], ref[k]), l2(emul[k], ref[k])
        rows.append((e_o - 2.0 * e_e, k, e_o, e_e))
    rows.sort(reverse=True)
    fo = torch.cat([torch.as_tensor(ours[k]).double().cpu().reshape(-1) for k in ref])
    fr = torch.cat([torch.as_tensor(ref[k]).double().cpu().reshape(-1) for k in ref])
    fe = torch.cat([torch.as_tensor(emul[k]).double().cpu().reshape(-1) for k in ref])
    cos = float(torch.dot(fo, fr) / (fo.norm() * fr.norm()))
    cos_e = float(torch.dot(fe, fr) / (fe.norm() * fr.norm()))
    print('%s: whole-gradient cosine ours %.4f (emulated bf16 %.4f), l2 ours %.4f (emulated %.4f); median per-tensor l2 ours %.4f '
          '(emulated %.4f); worst vs yardstick: %s' % (what, cos, cos_e, l2(fo, fr), l2(fe, fr), float(np.median([r[2] for r in rows])),
                                                       float(np.median([r[3] for r in rows])), [(k, '%.3f' % a, '%.3f' % b) for _, k, a, b in rows[:3]]))
    assert rows[0][0] < 5e-2, rows[0]
    assert l2(fo, fr) <= agg * l2(fe, fr) + 1e-2
    assert float(np.median([r[2] for r in rows])) <= agg * float(np.median([r[3] for r in rows])) + 1e-2
    assert cos > cos_min and cos > cos_e - 0.01


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
