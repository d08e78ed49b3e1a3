#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --frames-per-gpu 1 --steps 30 --graph --no-cpu-baseline --out gpurun_out/r02_b1_profile.json --profile-ops gpurun_out/r02_v2v_ops_b1.json > /dev/null 2> gpurun_out/r02_b1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_b1_profile.json').read().splitlines()[-1])
print('B=1 ms', d['ms_per_step'], 'value', d['value'])
for k,v in d['kernels'].items():
    if k!='v2v': print(k, round(v['ms_per_frame']*1000,1),'us')
tot=0
for e in json.load(open('gpurun_out/r02_v2v_ops_b1.json')):
    tot+=e['ms_per_frame']; print(e['op'], e['kind'], e['cin'], e['cout'], e['side'], round(e['ms_per_frame']*1000,1))
print('v2v sum us', tot*1000)
PY
