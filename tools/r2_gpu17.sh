#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn.py tests/test_gpu_engine.py tests/test_gpu_trained.py -m gpu -q -x > gpurun_out/g17_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g17_tests.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/g17_tests.log | head -20
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --legs 0 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('six-block build, addend from global: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'gn_bwd', [ (round(r['frac'],3), r['step_share']) for r in [d['roofline']]+d['roofline_other'] if r['kernel'].startswith('gn_relu_bwd')])"
done
