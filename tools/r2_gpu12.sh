#!/bin/bash
# 2 GPUs, tight timeouts: the data-parallel path
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/g12_dp_check.log 2>&1
echo "dp_check rc=$?"; grep -E "step|DP CHECK|barrier|Error|error" gpurun_out/g12_dp_check.log | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/g12_bench2.json 2> gpurun_out/g12_bench2.err
echo "bench2 rc=$?"; tail -3 gpurun_out/g12_bench2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/g12_bench2.json'))
    print('2 GPU value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), [round(x,3) for x in d['repeats']['ms_per_step']])
except Exception as e:
    print('failed', e)
PY
