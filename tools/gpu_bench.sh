#!/bin/bash
# tests + smoke + bench + ncu evidence, one box lease
mkdir -p gpurun_out
./tools/gpu_tests.sh > gpurun_out/tests_summary.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --profile-ops gpurun_out/v2v_ops.json > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 3 -o gpurun_out/prof_conv_tc \
   python bench.py --steps 1 --warmup 1 --frames-per-gpu 16 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/tests_summary.txt | tail -n 20
cat gpurun_out/smoke.log | tail -n 3
cat gpurun_out/bench.json | cut -c 1-1500
cat gpurun_out/bench.err | tail -n 5
