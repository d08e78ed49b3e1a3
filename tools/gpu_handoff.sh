#!/bin/bash
# hand-off measurement: timings vs the torch ops it replaces, forward() with/without it, one ncu capture
mkdir -p gpurun_out
timeout 600 python tools/bench_handoff.py --batch 64 --iters 10 ${HANDOFF_FORWARD:+--forward} --out gpurun_out/r02_handoff_bench.json 2> gpurun_out/r02_handoff_bench.err | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'handoff_tc_kernel|handoff_pack_kernel' -c 2 -s 4 \
    -o gpurun_out/r02_prof_handoff -f python tools/bench_handoff.py --batch 64 --kernel-only > gpurun_out/r02_ncu_handoff.log 2>&1
tail -3 gpurun_out/r02_ncu_handoff.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_handoff.py -x -q -k "kernel_vs_torch" > gpurun_out/r02_sanitizer_handoff.log 2>&1
tail -4 gpurun_out/r02_sanitizer_handoff.log
