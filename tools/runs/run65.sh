SDOD_GEMM_PERSIST=0 timeout 300 python tools/step_time.py 32 np 2>&1 | sed -n 2,5p | grep -E "graph| gemm"
grep -E "gemm M131072 N320 K320|gemm M32768 N640 K640|gemm M8192 N1280 K1280 |gemm M131072 N2560|gemm M131072 N960|gemm M131072 N320 K1280" gpurun_out/step_time_np.txt
