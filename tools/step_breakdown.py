"""Per-shape breakdown of one eager train step at the BASELINE config (GPU box): CUDA events around every C-ABI call
(spherehand_b200._lib.PROFILE), aggregated by (entry, shape), with achieved GB/s / TFLOP/s from bench.call_work.
    python tools/step_breakdown.py > gpurun_out/step_breakdown.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                       # noqa: E402
from spherehand_b200 import _lib, data, ops                        # noqa: E402
from spherehand_b200.engine import SelfSupTrainStep                # noqa: E402
from spherehand_b200.model import HandModel                        # noqa: E402
from spherehand_b200.network.hourglass import create_hourglass_network   # noqa: E402

SHAPE_ARGS = {'sh_conv_fwd': (4, 5, 6, 7, 8, 10), 'sh_conv_fwd_gn': (9, 10, 11, 12, 13), 'sh_conv_wgrad_gn': (7, 8, 9, 11, 13), 'sh_conv_wgrad': (2, 3, 4, 6, 8, 9), 'sh_conv_wgrad3x3': (2, 3, 4, 6, 8),
              'sh_gn_relu_fwd': (4, 5, 6), 'sh_gn_relu_bwd_prezeroed': (6, 7, 8), 'sh_maxpool_fwd': (1, 2, 3, 4), 'sh_maxpool_bwd': (3, 4, 5, 6),
              'sh_upsample_add_fwd': (2, 3, 4, 5), 'sh_upsample_bwd': (1, 2, 3, 4), 'sh_add': (3, 4, 5), 'sh_colsum': (1, 2, 3)}


def main():
    dev = torch.device('cuda', 0)
    B, V, NS, S, STACKS, J = bench.B, bench.V, bench.NS, bench.S, bench.STACKS, bench.J
    hand = HandModel.from_arrays(dict(np.load(os.path.join(bench.GOLD, 'hand_model.npz'))), dev)
    blob = ops.vae_blob_from_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(bench.GOLD, 'pose_vae.npz')).items()}, dev)
    torch.manual_seed(0)
    net = create_hourglass_network(2 * J, STACKS).to(dev)
    step = SelfSupTrainStep(net, hand, blob, B, V, NS, S, lr=1e-4, use_graph=False)
    gen = torch.Generator().manual_seed(1234)
    real, cams, inv = data.synthetic_real_batch(hand, B, V, S, gen)
    step.load_batch(real, cams, inv, data.random_poses(NS, gen))
    for _ in range(3):
        step.draw_randoms(); step.step(is_mv=True)
    torch.cuda.synchronize()
    reps = 3
    _lib.PROFILE = []
    for _ in range(reps):
        step.draw_randoms(); step.step(is_mv=True)
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    agg = {}
    for name, a, e0, e1 in prof:
        key = (name,) + tuple(a[i] for i in SHAPE_ARGS.get(name, ())) + ((bool(a[3]),) if name == 'sh_conv_fwd' else ()) + ((bool(a[8]),) if name == 'sh_conv_fwd_gn' else ()) + \
              ((bool(a[5]),) if name == 'sh_gn_relu_bwd_prezeroed' else ())
        fl, by, fam = bench.call_work(name, a)
        d = agg.setdefault(key, [0, 0.0, 0.0, 0.0, fam])
        d[0] += 1; d[1] += e0.elapsed_time(e1) * 1e3; d[2] += fl; d[3] += by
    total = sum(d[1] for d in agg.values())
    print('# eager step, CUDA events per C-ABI call, %d reps; total %.1f us per step' % (reps, total / reps))
    print('%-70s %5s %9s %7s %8s %8s' % ('call (shape...)', 'n/step', 'us/call', 'share', 'GB/s', 'TFLOP/s'))
    for key, (n, us, fl, by, fam) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-70s %5d %9.1f %6.2f%% %8.0f %8.1f' % (' '.join(str(k) for k in key)[:70], n // reps, us / n, 100 * us / total,
                                                     by / us * 1e-3 if by else 0, fl / us * 1e-6 if fl else 0))


if __name__ == '__main__':
    main()
