#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_stage.py -m gpu -x -q -s > gpurun_out/r02_geom_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_geom_tests.log
grep -E "passed|failed|rc=|V=96" gpurun_out/r02_geom_tests.log | tail -5
timeout 300 python tools/microbench_geometry.py > gpurun_out/r02_microbench_geometry.txt 2>&1; tail -12 gpurun_out/r02_microbench_geometry.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --out gpurun_out/r02_bench16.json > gpurun_out/r02_bench16.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench16.json').read().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
k=d['kernels']['voxelize']; print('voxelize', round(k['ms_per_frame']*1000,2), 'us', round(k['GB_per_s']), round(k['frac'],3))
PY
