#!/bin/bash
mkdir -p gpurun_out
# one full capture each of the HBM-bound kernels inside a 16-frame bench step (second step: skip the first launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'unproject_kernel|voxelize_kernel|tail_mlp_kernel|softargmax_partial|feature_conv1x1|maxpool2' -s 12 -c 8 -o gpurun_out/prof_hbm \
   python bench.py --steps 1 --warmup 3 --frames-per-gpu 16 --no-cpu-baseline > gpurun_out/ncu_hbm.log 2>&1
tail -n 2 gpurun_out/ncu_hbm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel<4, 1, 1, 0, 1>' -s 2 -c 1 -o gpurun_out/prof_deconv \
   python tools/run_v2v_only.py 16 2 > gpurun_out/ncu_deconv.log 2>&1
tail -n 2 gpurun_out/ncu_deconv.log
