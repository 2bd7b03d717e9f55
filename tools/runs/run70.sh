timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_libsdod.py -x -q 2>&1 | tail -2
timeout 300 python tools/step_time.py 32 ip32 2>&1 | sed -n 2,2p
timeout 300 python tools/step_time.py 2 ip2 2>&1 | sed -n 2,2p
