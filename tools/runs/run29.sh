timeout 900 python -m pytest tests/test_gpu_model.py -x -q -k "vae" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "geglu" 2>&1 | tail -2
for v in 1 0; do echo "SDOD_GEGLU_TANH=$v"; SDOD_GEGLU_TANH=$v timeout 120 python tools/hot_kernels.py geglu 8 2>&1 | tail -1; SDOD_GEGLU_TANH=$v timeout 120 python tools/hot_kernels.py geglu 32 2>&1 | tail -1; SDOD_GEGLU_TANH=$v timeout 300 python tools/step_time.py 32 gt$v 2>&1 | sed -n 2,2p; SDOD_GEGLU_TANH=$v timeout 300 python tools/step_time.py 2 gtb2$v 2>&1 | sed -n 2,2p; done
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | grep -E "rel|passed|failed" | tail -6
