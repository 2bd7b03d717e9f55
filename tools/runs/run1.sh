timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "gn or split or second or skip or conv" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s 2>&1 | tail -15
timeout 200 python tools/step_time.py 2 new 2>&1 | tail -18
SDOD_W_PREFETCH=0 timeout 200 python tools/step_time.py 2 nopf 2>&1 | sed -n 2,3p
SDOD_SPLITK_CLUSTER=0 timeout 200 python tools/step_time.py 2 nocluster 2>&1 | sed -n 1,3p
SDOD_GN_FUSED=0 timeout 200 python tools/step_time.py 2 nognf 2>&1 | sed -n 1,3p
