#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_lowres_start.jsonl > /dev/null 2> gpurun_out/r02_lowres.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table --no-features --out gpurun_out/r02_lowres_start.jsonl > /dev/null 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_lowres_start.jsonl > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_stage.py -m gpu -x -q > gpurun_out/r02_stage_tests2.log 2>&1; tail -2 gpurun_out/r02_stage_tests2.log
python - <<'PY'
import json
for l in open('gpurun_out/r02_lowres_start.jsonl'):
    d=json.loads(l); print('value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms %.3f'%d['ms_per_step'], d['steps'], d['config']['outputs'][:40])
PY
tail -3 gpurun_out/r02_lowres.err
