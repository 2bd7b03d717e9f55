timeout 300 python tools/step_time.py 32 pres 2>&1 | sed -n 2,5p | grep -E "graph| gemm"
grep -E "gemm M131072 N320 K320|gemm M32768 N640 K640" gpurun_out/step_time_pres.txt
SDOD_PERSIST_RES=1 timeout 300 python tools/step_time.py 32 pres1 2>&1 | sed -n 2,5p | grep -E "graph| gemm"
grep -E "gemm M131072 N320 K320|gemm M32768 N640 K640" gpurun_out/step_time_pres1.txt
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -k "unet" 2>&1 | tail -1
