#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_nn.py tests/test_gpu_engine.py tests/test_gpu_modules.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/g8_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g8_tests.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/g8_tests.log | head -30
timeout 600 python tools/step_breakdown.py > gpurun_out/g8_step_breakdown.txt 2>&1; head -60 gpurun_out/g8_step_breakdown.txt
timeout 600 python oracle/ref_gpu.py --mode dropin --S 128 --stacks 2 --B 64 --Ns 64 --steps 12 --warmup 3 --real_aug 1 2>&1 | tail -1 > gpurun_out/g8_dropin.json
python -c "
import json; d=json.load(open('gpurun_out/g8_dropin.json')); print('dropin', round(d['ms_per_step'],1), d['per_step_ms'], 'mem', round(d['peak_mem_gb'],1))"
