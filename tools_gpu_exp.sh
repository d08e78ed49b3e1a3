mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v2v.py tests/test_gpu_stage.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -n 12 > gpurun_out/quick_tests.log
tail -n 6 gpurun_out/quick_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-ops gpurun_out/v2v_ops.json > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c 1-200 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
ops=json.load(open('gpurun_out/v2v_ops.json'))
for o in ops:
    if o['kind'] in ('deconv','pool','tail') or o['op']<5: print(o['op'],o['kind'],o['cin'],o['cout'],o['k'],o['side'],round(o['ms_per_frame']*1000,1),round(o.get('tflops') or 0))
print(sum(o['ms_per_frame'] for o in ops)*1000)
b=json.load(open('gpurun_out/bench.json')); print(b['value'], b['e2e']['value'], b['clocks'])
PY
