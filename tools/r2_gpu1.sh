#!/bin/bash
# round-2 GPU check 1: new parity tests + the reference's own GPU path timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/g1_smi.txt
timeout 1500 python -m pytest tests/test_gpu_trained.py tests/test_gpu_dropin.py "tests/test_gpu_kernels.py" "tests/test_gpu_nn.py::test_hourglass_backward_uses_its_own_forward_tape" -m gpu -q -s > gpurun_out/g1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g1_tests.log
tail -30 gpurun_out/g1_tests.log
for cfg in "--S 64 --stacks 1 --B 64 --Ns 64" "--S 128 --stacks 2 --B 64 --Ns 64" "--S 128 --stacks 2 --B 64 --Ns 64 --sync 1"; do
  timeout 600 python oracle/ref_gpu.py --mode stock $cfg --steps 5 --warmup 2 2>&1 | tail -2 >> gpurun_out/g1_ref.log
done
timeout 600 python oracle/ref_gpu.py --mode dropin --S 128 --stacks 2 --B 64 --Ns 64 --steps 10 --warmup 3 2>&1 | tail -2 >> gpurun_out/g1_ref.log
timeout 600 python oracle/ref_gpu.py --mode dropin --S 64 --stacks 1 --B 64 --Ns 64 --steps 10 --warmup 3 2>&1 | tail -2 >> gpurun_out/g1_ref.log
cat gpurun_out/g1_ref.log
