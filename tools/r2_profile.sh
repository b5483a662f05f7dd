#!/bin/bash
# Round-2 ncu evidence (one GPU).  Numbers printed under ncu are never bench values; what is kept are launch durations, DRAM bytes
# and the --set full metrics of the kernels north_star names.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
# (1) launch list of one eager train step, kernels named by the C-ABI call that launched them (NVTX range)
timeout 900 ncu --profile-from-start off --nvtx --print-nvtx-rename kernel --metrics $M --clock-control none --csv \
    --log-file gpurun_out/r2_launches_by_call.csv python tools/profile_step.py > gpurun_out/r2_prof1.log 2>&1
# (2) the same by kernel name
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python tools/profile_step.py > gpurun_out/r2_prof2.log 2>&1
# (3) --set full of the named kernels inside the step: one launch each of the big shapes
for k in "gn_relu_bwd_kernel:20" "wgrad3x3_kernel:2" "wgrad1x1_kernel:10" "mvproj_main_kernel:0" "conv_fwd_kernel:30" "gn_relu_fwd_kernel:20" "softargmax_bwd_nhwc_kernel:0"; do
  name=${k%%:*}; skip=${k##*:}
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$name -s $skip -c 2 \
      -o gpurun_out/r2_full_$name -f python tools/profile_step.py > gpurun_out/r2_prof_$name.log 2>&1
done
# (4) the renderers stand-alone (BASELINE config 2 and the pybind boundary)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sphere_render_(fwd|bwd)_kernel|tri_raster_kernel" -s 6 -c 4 \
    -o gpurun_out/r2_full_renderers -f python tools/bench_layers.py raster > gpurun_out/r2_prof_renderers.log 2>&1
ls -la gpurun_out/*.ncu-rep | head -20
