#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v2v.py -m gpu -x -q -k "deconv or v2v_v32 or simple" > gpurun_out/r02_deconv_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_deconv_tests.log
tail -5 gpurun_out/r02_deconv_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --out gpurun_out/r02_bench14.json --profile-ops gpurun_out/r02_v2v_ops14.json > gpurun_out/r02_bench14.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench14.json').read().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
    print(d['kernels']['v2v']['families']['deconv'])
    for e in json.load(open('gpurun_out/r02_v2v_ops14.json')):
        if e['kind']=='deconv': print(e['op'], e['cin'], e['cout'], e['side'], round(e['ms_per_frame']*1000,2), round(e.get('GB_per_s',0)), round(e.get('frac',0),2))
except Exception as ex:
    print('bench failed', ex); print(open('gpurun_out/r02_bench14.log').read()[-2000:])
PY
