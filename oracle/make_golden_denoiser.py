"""Generate tests/golden/pose_denoiser.npz by running the UNMODIFIED reference PoseDenoiser (eval mode) on CPU with seeded
random weights (build container only).  TEST INFRASTRUCTURE.  Usage: python oracle/make_golden_denoiser.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import _refshim  # noqa: E402

_refshim.install()
import torch  # noqa: E402
from network.pose_denoiser import PoseDenoiser, input_indices, output_indices  # noqa: E402  (the reference's)


def main():
    torch.manual_seed(31)
    net = PoseDenoiser().eval()
    with torch.no_grad():
        for name, p in net.named_parameters():            # non-trivial GroupNorm affine parameters, biases
            if name.endswith('.1.weight') or name.endswith('.4.weight'):
                p.copy_(torch.rand_like(p) + 0.5)
            elif 'bias' in name:
                p.copy_(torch.randn_like(p) * 0.1)
        joints = torch.randn(7, 41, 3) * 40                # mm
        out3 = net(joints)
        out2 = net(joints.reshape(7, -1)[:3])
    out = {'sd.' + k: v.numpy() for k, v in net.state_dict().items()}
    out.update(joints=joints.numpy(), out3=out3.numpy(), out2=out2.numpy(), input_indices=np.asarray(input_indices),
               output_indices=np.asarray(output_indices), loss=np.float32(net.loss(joints, out3)))
    path = os.path.join(ROOT, 'tests', 'golden', 'pose_denoiser.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
