#!/bin/bash
# ncu --set full of the memory-/latency-bound kernels in their final round-2 form, at the bench's launch size (64 frames)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'unproject_kernel|voxelize_kernel|occ_expand_zwin_kernel|tail_tc_kernel|feature_conv1x1_tc_kernel' -c 5 \
    -o gpurun_out/r02_prof_mem_final -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_ncu_mem_final.log 2>&1
tail -3 gpurun_out/r02_ncu_mem_final.log
ls -la gpurun_out/r02_prof_mem_final.ncu-rep
