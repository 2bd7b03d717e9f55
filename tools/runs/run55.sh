SDOD_SPLITK_CLUSTER=2 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "split_k or second_operand or fused_skip or conv3x3 or gemm_plain" 2>&1 | tail -3
SDOD_SPLITK_CLUSTER=2 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | grep -E "rel|passed|failed|rror" | tail -6
SDOD_SPLITK_CLUSTER=2 timeout 300 python tools/step_time.py 2 skc2 2>&1 | sed -n 2,4p
SDOD_SPLITK_CLUSTER=1 timeout 300 python tools/step_time.py 2 skc1 2>&1 | sed -n 2,4p
timeout 300 python tools/step_time.py 2 skc0 2>&1 | sed -n 2,4p
