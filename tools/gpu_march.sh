#!/bin/bash
# marching conv: parity tests first (own process), then whole-network tests, then the bench with the per-op table
mkdir -p gpurun_out
python -m sceneego_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -n 5 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_v2v.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "marching" 2>&1 | tail -n 40 > gpurun_out/march_tests.log
tail -n 25 gpurun_out/march_tests.log
timeout 900 python -m pytest tests/test_gpu_v2v.py tests/test_gpu_stage.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "not marching" 2>&1 | tail -n 15 > gpurun_out/quick_tests.log
tail -n 6 gpurun_out/quick_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-ops gpurun_out/v2v_ops.json > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c 1-300 gpurun_out/bench.json
tail -n 5 gpurun_out/bench.err
python tools/show_ops.py gpurun_out/v2v_ops.json 2>/dev/null | head -70
