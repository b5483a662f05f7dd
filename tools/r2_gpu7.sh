#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_gpu_trained.py -m gpu -q -x > gpurun_out/g7_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g7_tests.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/g7_tests.log | head -30
timeout 600 python bench.py --steps 20 --warmup 5 --legs 0 > gpurun_out/g7_bench.json 2> gpurun_out/g7_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/g7_bench.json'))
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches']/20, d['roofline']['shares'])
PY
for aug in 1 0; do
timeout 600 python oracle/ref_gpu.py --mode dropin --S 128 --stacks 2 --B 64 --Ns 64 --steps 12 --warmup 3 --real_aug $aug 2>&1 | tail -1 > gpurun_out/g7_dropin_aug$aug.json
python -c "
import json; d=json.load(open('gpurun_out/g7_dropin_aug$aug.json')); print('dropin aug=$aug', round(d['ms_per_step'],1), d['per_step_ms'], 'mem', round(d['peak_mem_gb'],1))"
done
