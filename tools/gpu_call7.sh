#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v2v.py -m gpu -x -q -k "march" > gpurun_out/r02_m4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_m4_tests.log
tail -6 gpurun_out/r02_m4_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --out gpurun_out/r02_bench7.json --profile-ops gpurun_out/r02_v2v_ops7.json > gpurun_out/r02_bench7.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench7.json').read().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
    r=d['roofline']; print('march', r['achieved'], r['frac'], r['launch_ms'])
    for e in json.load(open('gpurun_out/r02_v2v_ops7.json'))[:10]: print(e['op'], e['cin'], e['cout'], round(e['ms_per_frame']*1000,2), round(e.get('tflops',0)))
except Exception as ex:
    print('bench failed', ex); print(open('gpurun_out/r02_bench7.log').read()[-3000:])
PY
