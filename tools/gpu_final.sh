#!/bin/bash
# full GPU test suite (separate processes), smoke, bench (our arm + reference arm), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
for f in test_gpu_geometry test_gpu_softargmax test_gpu_eval test_gpu_v2v test_gpu_stage; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -n 6 > gpurun_out/$f.log
  echo "== $f: $(tail -n 1 gpurun_out/$f.log)"
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --profile-ops gpurun_out/v2v_ops.json > gpurun_out/bench.json 2> gpurun_out/bench.err
python tools/show_ops.py 2>/dev/null | head -9
tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cut -c 1-300 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
