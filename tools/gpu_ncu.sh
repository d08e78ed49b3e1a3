#!/bin/bash
mkdir -p gpurun_out
# full captures, second pass of the V2V program at 16 frames per launch: marching convs, stem, fused tail
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_march|stem_s2d_tc|tail_tc' -s 12 -c 6 -o gpurun_out/prof_v2v_main \
   python tools/run_v2v_only.py 16 2 > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_full.log
