#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_f16.py tests/test_gpu_v2v.py -m gpu -x -q -s > gpurun_out/r02_f16_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_f16_tests.log
grep -E "passed|failed|rc=|FAILED" gpurun_out/r02_f16_tests.log | tail -8
