"""One eager (un-graphed) train step at the BASELINE config for ncu: every C-ABI call sits in an NVTX range named after its
entry (convolutions: `sh_conv_fwd_1x1` / `sh_conv_fwd_3x3` / `sh_conv_wgrad_1x1` / ...), profiling is limited to the third
step by cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --profile-from-start off --nvtx --nvtx-include "sh_conv_fwd_1x1/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ...
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spherehand_b200 import _lib, data, ops                     # noqa: E402
from spherehand_b200.engine import SelfSupTrainStep             # noqa: E402
from spherehand_b200.model import HandModel                     # noqa: E402
from spherehand_b200.network.hourglass import create_hourglass_network   # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
B, V, NS, S, STACKS, J = 64, 3, 64, 128, 2, 41


def tag(name, a):
    if name == 'sh_conv_fwd':
        return 'sh_conv_fwd_%s' % ('3x3' if a[10] == 9 else '1x1')
    if name == 'sh_conv_wgrad':
        return 'sh_conv_wgrad_%s' % ('3x3' if a[9] == 9 else '1x1')
    if name == 'sh_conv_fwd_gn':            # the 1x1 layer with the GroupNorm in front of it folded in: same family
        return 'sh_conv_fwd_1x1'
    if name == 'sh_conv_wgrad_gn':
        return 'sh_conv_wgrad_1x1'
    return name


def main():
    dev = torch.device('cuda', 0)
    hand = HandModel.from_arrays(dict(np.load(os.path.join(GOLD, 'hand_model.npz'))), dev)
    blob = ops.vae_blob_from_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, 'pose_vae.npz')).items()}, dev)
    torch.manual_seed(0)
    net = create_hourglass_network(2 * J, STACKS).to(dev)
    step = SelfSupTrainStep(net, hand, blob, B, V, NS, S, lr=1e-4, use_graph=False)
    gen = torch.Generator().manual_seed(1234)
    real, cams, inv = data.synthetic_real_batch(hand, B, V, S, gen)
    step.load_batch(real, cams, inv, data.random_poses(NS, gen))
    _lib.NVTX = tag
    for i in range(3):
        step.draw_randoms()
        torch.cuda.synchronize()
        if i == 2:
            torch.cuda.profiler.start()
        step.step(is_mv=True)
        torch.cuda.synchronize()
        if i == 2:
            torch.cuda.profiler.stop()
    print('terms', step.terms.tolist())


if __name__ == '__main__':
    main()
