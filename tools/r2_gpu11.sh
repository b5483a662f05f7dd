#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/g11_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g11_tests.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/g11_tests.log | head -20
timeout 300 python bench.py --steps 20 --warmup 5 --legs 0 > gpurun_out/g11_bench1.json 2> gpurun_out/g11_bench1.err
python - <<PY
import json
d=json.load(open('gpurun_out/g11_bench1.json'))
print('1 GPU value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches/step', d['gpu_launches']/20)
for r in [d['roofline']]+d['roofline_other']:
    print('  %-50s frac=%.3f share=%s' % (r['kernel'][:50], r.get('frac',0), r.get('step_share')))
PY
SH_GN_BWD_MB=4 timeout 300 python bench.py --steps 20 --warmup 5 --legs 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('gn_bwd four-block build everywhere: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3))"
