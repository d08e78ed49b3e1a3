#!/bin/bash
# round-2 GPU call 2: marching stem bring-up
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_v2v.py -m gpu -x -q -s -k "stem" > gpurun_out/r02_stem_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_stem_tests.log
tail -15 gpurun_out/r02_stem_tests.log
timeout 600 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_stage.py -m gpu -x -q -s > gpurun_out/r02_stage_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_stage_tests.log
tail -8 gpurun_out/r02_stage_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --out gpurun_out/r02_bench2.json --profile-ops gpurun_out/r02_v2v_ops2.json > gpurun_out/r02_bench2.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench2.json').read().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
    print('stem', d['roofline']['stem'])
except Exception as e: print('bench failed', e); print(open('gpurun_out/r02_bench2.log').read()[-2000:])
PY
