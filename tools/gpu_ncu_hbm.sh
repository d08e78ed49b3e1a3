#!/bin/bash
mkdir -p gpurun_out
# one full capture each of the memory-/latency-bound kernels inside a 64-frame bench step (skip the warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'unproject_kernel|voxelize_kernel|tail_tc_kernel|softargmax_partial|softmax_write|feature_conv1x1|upsample_pad' -s 21 -c 7 -o gpurun_out/prof_hbm \
   python bench.py --steps 1 --warmup 3 --frames-per-gpu 64 --no-cpu-baseline > gpurun_out/ncu_hbm.log 2>&1
tail -n 2 gpurun_out/ncu_hbm.log
