#!/bin/bash
mkdir -p gpurun_out
# full capture of the marching conv: second pass of the program (skip the 10 launches of the first pass)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_march -s 10 -c 4 -o gpurun_out/prof_conv_march \
   python tools/run_v2v_only.py 16 2 > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_full.log
# launch list of the whole bench step (cold-cache, serialised): kernel shares of the step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
