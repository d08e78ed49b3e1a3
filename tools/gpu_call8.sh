#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_march2_kernel' -c 3 \
    -o gpurun_out/r02_prof_march2_b64 -f python tools/run_v2v_only.py 64 1 > gpurun_out/r02_ncu_march2.log 2>&1
tail -2 gpurun_out/r02_ncu_march2.log
ls -la gpurun_out/r02_prof_march2_b64.ncu-rep
