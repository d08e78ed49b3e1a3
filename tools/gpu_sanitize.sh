#!/bin/bash
# compute-sanitizer over the round-2 kernels (marching stem, hand-off, z-window voxelisation, dataset voxelisation)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_v2v.py -q -x -k "test_stem_march and (16-2 or 32-3)" > gpurun_out/r02_sanitize_stem.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitize_stem.log
tail -6 gpurun_out/r02_sanitize_stem.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_handoff.py -q -x -k "3-32-32 or 5-16-24" > gpurun_out/r02_sanitize_handoff.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitize_handoff.log
tail -6 gpurun_out/r02_sanitize_handoff.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_geometry.py -q -x -k "dataset or demo_frames" > gpurun_out/r02_sanitize_geometry.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitize_geometry.log
tail -6 gpurun_out/r02_sanitize_geometry.log
# final round-2 kernels: fused tail with the hidden tile in tensor memory, split-operand feature conv, z-window expand (inside the demo-frame tests above)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_v2v.py -q -x -k "tail" > gpurun_out/r02_sanitize_tail.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitize_tail.log
tail -4 gpurun_out/r02_sanitize_tail.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_geometry.py -q -x -k "feature_conv" > gpurun_out/r02_sanitize_fc.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitize_fc.log
tail -4 gpurun_out/r02_sanitize_fc.log
