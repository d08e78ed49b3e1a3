#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_v2v.py -m gpu -x -q -k "march or stem" > gpurun_out/r02_march_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_march_tests.log
tail -4 gpurun_out/r02_march_tests.log
for b in 1 2 4 8 16; do
  timeout 200 python bench.py --frames-per-gpu $b --steps 50 --graph --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_latency2_graph.jsonl > /dev/null 2>> gpurun_out/r02_latency2.err
done
timeout 200 python bench.py --frames-per-gpu 1 --steps 50 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_latency2_eager.jsonl > /dev/null 2>> gpurun_out/r02_latency2.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_bench5.json > gpurun_out/r02_bench5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel<4, 1, 1, 0, 1>|tail_tc_kernel|unproject_kernel|voxelize_kernel' -c 6 \
    -o gpurun_out/r02_prof_mem_b64 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_ncu_mem.log 2>&1
python - <<'PY'
import json
for p in ('gpurun_out/r02_latency2_graph.jsonl','gpurun_out/r02_latency2_eager.jsonl','gpurun_out/r02_bench5.json'):
    for l in open(p):
        d=json.loads(l); print(p.split('/')[-1], 'B', d['config']['frames_per_gpu'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms %.3f'%d['ms_per_step'])
PY
ls -la gpurun_out/r02_prof_mem_b64.ncu-rep
