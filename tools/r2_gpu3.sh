#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/g3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g3_tests.log
grep -E "passed|failed|FAILED|^E  |joints|cosine|terms|heat-map|per-element|engine:|two shards|torch fp32" gpurun_out/g3_tests.log | head -120
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/g3_bench.json 2> gpurun_out/g3_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/g3_bench.err; cat gpurun_out/g3_bench.json | head -c 6000
