timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn or group_norm or GroupNorm or efficient or groupnorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -2
