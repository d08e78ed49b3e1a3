#!/bin/bash
# round-2 sweeps (BASELINE configs[2] / [3]): bash tools/gpu_sweep_r2.sh N
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --sweep 8,16,32,64,128,256,512,1024 --steps 10 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_sweep_1gpu.jsonl > /dev/null 2> gpurun_out/r02_sweep_1gpu.err
  timeout 600 python bench.py --volume-size 128 --frames-per-gpu 8 --chunk 8 --steps 10 --no-cpu-baseline --out gpurun_out/r02_bench_v128.json > /dev/null 2> gpurun_out/r02_bench_v128.err
  timeout 300 python bench.py --steps 10 --raw-depth --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_bench_rawdepth_1gpu.json > /dev/null 2> gpurun_out/r02_bench_rawdepth_1gpu.err
  for b in 1 2 4 8 16; do
    timeout 200 python bench.py --frames-per-gpu $b --steps 50 --graph --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_latency_graph.jsonl > /dev/null 2>> gpurun_out/r02_latency.err
    timeout 200 python bench.py --frames-per-gpu $b --steps 50 --no-cpu-baseline --no-kernel-table --out gpurun_out/r02_latency_eager.jsonl > /dev/null 2>> gpurun_out/r02_latency.err
  done
else
  P=$((29500 + N))
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P"
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-kernel-table --out gpurun_out/r02_bench_${N}gpu.json > /dev/null 2> gpurun_out/r02_bench_${N}gpu.err
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-kernel-table --raw-depth --out gpurun_out/r02_bench_rawdepth_${N}gpu.json > /dev/null 2> gpurun_out/r02_bench_rawdepth_${N}gpu.err
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-kernel-table --no-numa-bind --out gpurun_out/r02_bench_nonuma_${N}gpu.json > /dev/null 2> gpurun_out/r02_bench_nonuma_${N}gpu.err
  timeout 900 $TR bench.py --gpus $N --sweep 8,64,256,512,1024 --steps 10 --no-kernel-table --out gpurun_out/r02_sweep_${N}gpu.jsonl > /dev/null 2> gpurun_out/r02_sweep_${N}gpu.err
fi
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/r02_sweep_*gpu.jsonl')+glob.glob('gpurun_out/r02_bench_*gpu.json')+glob.glob('gpurun_out/r02_bench_v128.json')+glob.glob('gpurun_out/r02_latency_*.jsonl')):
    for l in open(p):
        d=json.loads(l); c=d['config']
        print(p.split('/')[-1], 'gpus', d['n_gpus'], 'B/gpu', c['frames_per_gpu'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'ms %.3f' % d['ms_per_step'], 'h2d %.1f GB/s' % d['e2e']['h2d_alone_GB_per_s'], 'numa', d['e2e'].get('numa_node'), d['clocks']['reasons'] if d.get('clocks') else '')
PY
tail -3 gpurun_out/*.err | tail -30
