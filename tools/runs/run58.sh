timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "geglu or persistent" 2>&1 | tail -1
timeout 120 python tools/hot_kernels.py geglu 8 2>&1 | tail -1
timeout 120 python tools/hot_kernels.py geglu 32 2>&1 | tail -1
