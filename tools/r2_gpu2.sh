#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_trained.py tests/test_gpu_dropin.py "tests/test_gpu_kernels.py" "tests/test_gpu_nn.py::test_hourglass_backward_uses_its_own_forward_tape" -m gpu -q -s > gpurun_out/g2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g2_tests.log
grep -E "passed|failed|FAILED|Error|error|joints|cosine|terms|heat-map|per-element|engine:" gpurun_out/g2_tests.log | head -80
timeout 600 python oracle/ref_gpu.py --mode stock --S 128 --stacks 2 --B 64 --Ns 64 --tf32 default --steps 5 --warmup 2 2>&1 | tail -1 > gpurun_out/g2_ref.log
cat gpurun_out/g2_ref.log
