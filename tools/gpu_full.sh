#!/bin/bash
# full verification: GPU test-suite, smoke, bench (default line + reference arm)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_tests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_full.log
tail -4 gpurun_out/r02_tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 --out gpurun_out/r02_bench_full.json --profile-ops gpurun_out/r02_v2v_ops_full.json > gpurun_out/r02_bench_full.log 2>&1
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 --out gpurun_out/r02_bench_reference.json > /dev/null 2>&1
python - <<'PY'
import json
for p in ('gpurun_out/r02_bench_full.json','gpurun_out/r02_bench_reference.json'):
    d=json.loads(open(p).read().splitlines()[-1]); print(p, d.get('impl','ours'), 'value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
PY
