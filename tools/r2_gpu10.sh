#!/bin/bash
# 2 GPUs: the data-parallel path (bucketed all-reduce inside the graph) + full GPU suite + bench at 1 and 2 GPUs
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/g10_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g10_tests.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/g10_tests.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/g10_dp_check.log 2>&1
echo "dp_check rc=$?"; grep -E "step|DP CHECK|Error|error" gpurun_out/g10_dp_check.log | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --legs 0 > gpurun_out/g10_bench1.json 2> gpurun_out/g10_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/g10_bench2.json 2> gpurun_out/g10_bench2.err
echo "bench2 rc=$?"; tail -3 gpurun_out/g10_bench2.err
python - <<PY
import json
for n in (1, 2):
    try:
        d=json.load(open('gpurun_out/g10_bench%d.json' % n))
        print(n, 'GPU value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches/step', d['gpu_launches']/20)
    except Exception as e:
        print(n, 'failed', e)
PY
