mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v2v.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "stem_s2d" 2>&1 | tail -n 15 > gpurun_out/stem_tests.log
tail -n 15 gpurun_out/stem_tests.log
timeout 900 python -m pytest tests/test_gpu_v2v.py tests/test_gpu_stage.py tests/test_gpu_geometry.py -m gpu -q --timeout 600 -p no:cacheprovider -k "not stem_s2d and not simt" 2>&1 | tail -n 12 > gpurun_out/quick_tests.log
tail -n 12 gpurun_out/quick_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-ops gpurun_out/v2v_ops.json > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c 1-300 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
