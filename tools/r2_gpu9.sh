#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn.py -m gpu -q -x > gpurun_out/g9_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g9_tests.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/g9_tests.log | head -20
for cfg in "0 1" "16 1" "8 1" "4 1" "16 2" "16 4" "16 8"; do set -- $cfg
SH_WGRAD_KSPLIT=$1 SH_W1_MINK=$2 timeout 300 python tools/step_breakdown.py 2>&1 | grep -E "total|sh_conv_wgrad3x3 256 8 8|sh_conv_wgrad3x3 256 16|sh_conv_wgrad 256 (8|16|32|64) " | sed "s/^/ks=$1 mink=$2 /"
done > gpurun_out/g9_wgrad_sweep.txt 2>&1
cat gpurun_out/g9_wgrad_sweep.txt
