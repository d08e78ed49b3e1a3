#!/bin/bash
# round-2 GPU call 1: full GPU test-suite, smoke, bench (both feature-buffer policies), batch-1 latency, reference-gpu
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpu_info.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests.log
tail -5 gpurun_out/r02_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 --out gpurun_out/r02_bench1.json --profile-ops gpurun_out/r02_v2v_ops.json > gpurun_out/r02_bench1.log 2>&1; tail -c 600 gpurun_out/r02_bench1.log
timeout 300 python bench.py --steps 10 --persistent-features --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_bench1_persist.json > /dev/null 2> gpurun_out/r02_bench1_persist.err
timeout 300 python bench.py --frames-per-gpu 1 --steps 50 --graph --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_lat_b1_graph.json > /dev/null 2> gpurun_out/r02_lat_b1_graph.err
timeout 300 python bench.py --frames-per-gpu 1 --steps 50 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_lat_b1_eager.json > /dev/null 2> gpurun_out/r02_lat_b1_eager.err
timeout 600 python bench.py --impl reference-gpu --steps 1 --out gpurun_out/r02_refgpu.json > /dev/null 2> gpurun_out/r02_refgpu.err
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r02_*.json')):
    if 'ops' in p: continue
    for l in open(p):
        d=json.loads(l)
        print(p, d.get('impl','ours'), 'value', d.get('value'), 'e2e', d.get('e2e',{}).get('value'), 'ms', d.get('ms_per_step'), d.get('unavailable'))
PY
