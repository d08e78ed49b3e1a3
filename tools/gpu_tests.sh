#!/bin/bash
# Run the GPU test files in separate processes (a sticky CUDA error in one must not poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
for f in test_gpu_geometry test_gpu_softargmax; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/$f.log
done
timeout 600 python -m pytest tests/test_gpu_v2v.py -m gpu -q --timeout 300 -p no:cacheprovider -k "simt or maxpool" 2>&1 | tail -40 > gpurun_out/test_gpu_v2v_simt.log
timeout 900 python -m pytest tests/test_gpu_v2v.py -m gpu -q --timeout 300 -p no:cacheprovider -k "not simt and not maxpool" 2>&1 | tail -60 > gpurun_out/test_gpu_v2v_tc.log
timeout 900 python -m pytest tests/test_gpu_stage.py -m gpu -q --timeout 600 -p no:cacheprovider -s 2>&1 | tail -60 > gpurun_out/test_gpu_stage.log
for f in gpurun_out/test_*.log; do echo "== $f"; tail -n 3 $f; done
