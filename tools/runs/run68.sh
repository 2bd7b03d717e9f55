timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | grep -E "rel|passed|failed|rror" | tail -7
timeout 300 python tools/step_time.py 32 red1 2>&1 | sed -n 2,5p | grep -E "graph| gemm"
grep -E "gemm M131072 N320 K320|gemm M32768 N640 K640|gemm M8192 N1280 K1280 " gpurun_out/step_time_red1.txt
SDOD_TMA_REDUCE=0 timeout 300 python tools/step_time.py 32 red0 2>&1 | sed -n 2,5p | grep -E "graph| gemm"
grep -E "gemm M131072 N320 K320|gemm M32768 N640 K640|gemm M8192 N1280 K1280 " gpurun_out/step_time_red0.txt
timeout 300 python tools/step_time.py 2 red1b2 2>&1 | sed -n 2,2p
