#!/bin/bash
mkdir -p gpurun_out
# second pass of the program: skip the 47 conv launches of the first pass, capture stem + first convs
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 47 -c 5 -o gpurun_out/prof_conv_tc \
   python tools/run_v2v_only.py 16 2 > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_full.log
