#!/bin/bash
# round-2 GPU call 3: ncu captures at the bench's launch size (64 frames) + side-stream cost experiment
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_march_kernel|stem_march_tc_kernel' -c 4 \
    -o gpurun_out/r02_prof_march_b64 -f python tools/run_v2v_only.py 64 1 > gpurun_out/r02_ncu_march.log 2>&1
tail -3 gpurun_out/r02_ncu_march.log
ls -la gpurun_out/*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_ncu_bench.log 2>&1
timeout 300 python bench.py --steps 10 --no-features --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_bench_nofeat.json > /dev/null 2> gpurun_out/r02_bench_nofeat.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_bench_20steps.json > /dev/null 2> gpurun_out/r02_bench_20.err
python - <<'PY'
import json,glob
for p in ('gpurun_out/r02_bench_nofeat.json','gpurun_out/r02_bench_20steps.json'):
    for l in open(p):
        d=json.loads(l); print(p, 'value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks'])
PY
