#!/bin/bash
mkdir -p gpurun_out
for band in 32 16 8; do for tpb in 16 8 4 2; do
SH_R2_BAND=$band SH_R2_TPB=$tpb timeout 300 python - 2>&1 <<'PY' | tail -4 | sed "s/^/band=$band tpb=$tpb  /"
import os, json, numpy as np, torch, bench
from spherehand_b200.model import HandModel
dev = torch.device('cuda', 0)
hand = HandModel.from_arrays(dict(np.load(os.path.join(bench.GOLD, 'hand_model.npz'))), dev)
for r in bench.renderer_rooflines(hand, bench.peaks(), dev)[:4]:
    print(r['kernel'][:45], 'us=%.1f' % r['avg_launch_us'], 'frac=%.3f' % r.get('frac', 0))
PY
done; done > gpurun_out/g5_r2sweep.log
cat gpurun_out/g5_r2sweep.log
