#!/bin/bash
mkdir -p gpurun_out
for b in 1 0 1 0; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus 2 --steps 20 --warmup 5 --bucketed $b 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bucketed=$b 2 GPU value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), [round(x,3) for x in d['repeats']['ms_per_step']])"
done
