#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02_overlap.jsonl
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_overlap.jsonl > /dev/null 2>&1
SCENEEGO_FEATURES_SERIAL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_overlap.jsonl > /dev/null 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table --no-features --out gpurun_out/r02_overlap.jsonl > /dev/null 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r02_overlap.jsonl'):
    d=json.loads(l); print('value %.0f'%d['value'], 'ms %.3f'%d['ms_per_step'], d['config']['outputs'][:40])
PY
