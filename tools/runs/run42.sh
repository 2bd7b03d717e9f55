timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "persist or geglu or epilogue or gemm" 2>&1 | tail -2
SDOD_TMEM_ACCS=2 timeout 300 python tools/step_time.py 32 acc2 2>&1 | sed -n 2,6p | grep -E "graph| gemm"
timeout 300 python tools/step_time.py 32 acc4 2>&1 | sed -n 2,6p | grep -E "graph| gemm"
for f in acc2 acc4; do echo $f; grep -E "gemm M131072 N320 K320|gemm M131072 N2560 K320|gemm M131072 N960 K320|gemm M32768 N640 K640|gemm M131072 N320 K1280|gemm M32768 N5120" gpurun_out/step_time_$f.txt | head -8; done
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | grep -E "rel|passed|failed" | tail -6
