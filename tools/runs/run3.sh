timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gemm or conv or split or stream or qkv" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | tail -6
timeout 200 python tools/step_time.py 2 sk 2>&1 | sed -n 1,14p
SDOD_STREAMK=0 timeout 200 python tools/step_time.py 2 nosk 2>&1 | sed -n 1,3p
SDOD_SPLITK_CLUSTER=0 timeout 200 python tools/step_time.py 2 sk_nocl 2>&1 | sed -n 1,3p
timeout 200 python tools/gemm_timeline.py sk 2>&1 | tail -18
