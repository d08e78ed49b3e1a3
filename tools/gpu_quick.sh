#!/bin/bash
# quick iteration: v2v + stage parity, then bench with the per-op table
mkdir -p gpurun_out
python -m sceneego_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -n 5 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_v2v.py tests/test_gpu_stage.py tests/test_gpu_geometry.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -n 15 > gpurun_out/quick_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-ops gpurun_out/v2v_ops.json > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 5 gpurun_out/quick_tests.log
cut -c 1-400 gpurun_out/bench.json
tail -n 5 gpurun_out/bench.err
