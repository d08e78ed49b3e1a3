#!/bin/bash
# BASELINE configs[2]: frame-batch sweep on one GPU (per-GPU batch; the 2/4/8-GPU runs shard the same per-GPU work)
mkdir -p gpurun_out; : > gpurun_out/sweep.txt
for b in 8 16 32 64 128; do
  timeout 600 python bench.py --steps 10 --warmup 3 --frames-per-gpu $b --no-cpu-baseline > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_b$b.json')); print('frames_per_gpu $b value %.1f frames/s  e2e %.1f  ms/step %.2f  sm_mhz %s' % (b['value'], b['e2e']['value'], b['ms_per_step'], b['clocks']['sm_mhz']))" | tee -a gpurun_out/sweep.txt
done
