#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_f16.py -m gpu -x -q -s > gpurun_out/r02_f16_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_f16_tests.log
grep -E "F16CHECK|dtype|passed|failed|rc=|Error" gpurun_out/r02_f16_tests.log | tail -8
timeout 400 python bench.py --act-dtype f16 --steps 10 --warmup 3 --no-cpu-baseline --out gpurun_out/r02_bench_f16.json > gpurun_out/r02_bench_f16.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_bench_bf16_same_box.json > /dev/null 2>&1
python - <<'PY'
import json
for p in ('gpurun_out/r02_bench_f16.json','gpurun_out/r02_bench_bf16_same_box.json'):
    try:
        d=json.loads(open(p).read().splitlines()[-1]); print(p, d['dtype'], 'value', d['value'], 'e2e', d['e2e']['value'])
    except Exception as e: print(p, 'failed', e)
PY
tail -3 gpurun_out/r02_bench_f16.log | cut -c1-300
