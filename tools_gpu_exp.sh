mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_v2v.py tests/test_gpu_stage.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "not simt and not maxpool" 2>&1 | tail -n 8 > gpurun_out/quick_tests.log
tail -n 4 gpurun_out/quick_tests.log
./tools/mma_replay > gpurun_out/mma_replay.txt 2>&1; cat gpurun_out/mma_replay.txt
for d in 1 7; do SCENEEGO_DEBUG=$d timeout 120 python tools/debug_conv.py one 2>&1 | head -3; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 2 -o gpurun_out/prof_conv_tc2 python tools/run_v2v_only.py 16 2 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
