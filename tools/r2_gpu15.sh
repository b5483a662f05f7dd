#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py > gpurun_out/g15_dp.log 2>&1
echo "dp rc=$?"; grep -E "DP CHECK|barrier|step 2|Error|error|Traceback" -A3 gpurun_out/g15_dp.log | head -40
for ch in 0 16 32; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$((ch/16)) bench.py --gpus 2 --steps 20 --warmup 5 --nccl_channels $ch > gpurun_out/g15_b$ch.json 2> gpurun_out/g15_b$ch.err
echo "bench ch=$ch rc=$?"; grep -E "Error|error" gpurun_out/g15_b$ch.err | head -5
grep '^{' gpurun_out/g15_b$ch.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nccl_channels=$ch 2 GPU value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), [round(x,3) for x in d['repeats']['ms_per_step']])"
done
