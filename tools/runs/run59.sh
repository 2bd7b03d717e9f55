timeout 300 python tools/step_time.py 32 gld 2>&1 | sed -n 2,5p | grep -E "graph| gemm"
grep -E "gemm M131072 N2560 K320|gemm M32768 N5120 K640|gemm M8192 N10240" gpurun_out/step_time_gld.txt
timeout 300 python tools/step_time.py 2 gldb2 2>&1 | sed -n 2,2p
