#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py::test_two_shards_sum_to_the_global_step tests/test_gpu_nn.py::test_hourglass_backward_uses_its_own_forward_tape tests/test_gpu_modules.py -m gpu -q -s > gpurun_out/g4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g4_tests.log
grep -E "passed|failed|FAILED|^E  |two shards|own-tape|per-element" gpurun_out/g4_tests.log | head -40
timeout 300 python tools/diag_noise.py > gpurun_out/g4_noise.log 2>&1; cat gpurun_out/g4_noise.log | tail -12
timeout 300 python - > gpurun_out/g4_r2.log 2>&1 <<'PY'
import os, json, numpy as np, torch, bench
from spherehand_b200.model import HandModel
dev = torch.device('cuda', 0)
hand = HandModel.from_arrays(dict(np.load(os.path.join(bench.GOLD, 'hand_model.npz'))), dev)
for r in bench.renderer_rooflines(hand, bench.peaks(), dev):
    print(r['kernel'][:70], 'us=%.1f' % r['avg_launch_us'], 'frac=%.3f' % r.get('frac', 0))
PY
cat gpurun_out/g4_r2.log | tail -8
