#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/g6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g6_tests.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/g6_tests.log | head -30
for pdl in 1 0; do
  SH_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --legs 0 > gpurun_out/g6_bench_pdl$pdl.json 2> gpurun_out/g6_bench_pdl$pdl.err
  python - <<PY
import json
d=json.load(open('gpurun_out/g6_bench_pdl$pdl.json'))
print('PDL=$pdl value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), [round(x,3) for x in d['repeats']['ms_per_step']])
PY
done
timeout 600 python oracle/ref_gpu.py --mode dropin --S 128 --stacks 2 --B 64 --Ns 64 --steps 5 --warmup 2 --sections 1 2>&1 | tail -1 > gpurun_out/g6_dropin_sections.json
python -c "
import json; d=json.load(open('gpurun_out/g6_dropin_sections.json')); print('dropin', d['ms_per_step'], d['sections_ms'])"
