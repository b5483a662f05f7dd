#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/g16_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g16_tests.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/g16_tests.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g16_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/g16_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
for v in "--real_aug 0" "--sample_poses 1"; do
timeout 300 python bench.py --steps 20 --warmup 5 --legs 0 $v 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))" | tee -a gpurun_out/g16_variants.txt
done
timeout 600 python tools/step_breakdown.py > gpurun_out/r2_step_breakdown.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_fwd_kernel|wgrad1x1_kernel|wgrad3x3_kernel|gn_relu_(fwd|bwd)_kernel|sphere_render_(fwd|bwd)_kernel|tri_raster_kernel|mvproj_main_kernel" -o gpurun_out/r2_full_shapes -f python tools/ncu_shapes.py > gpurun_out/r2_prof_shapes.log 2>&1
tail -3 gpurun_out/r2_prof_shapes.log
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench.json') if l.startswith('{')][-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config']['real_aug'][:20])
print('gpu_reference', d['gpu_reference'].get('value'), d['vs_reference_gpu'], 'dropin', d['e2e_dropin'].get('value'))
PY
