timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention or qkv" 2>&1 | tail -4
for pt in 1 0; do echo "SDOD_ATTN_PTMEM=$pt"; SDOD_ATTN_PTMEM=$pt timeout 120 python tools/hot_kernels.py attn 8 2>&1 | tail -1; SDOD_ATTN_PTMEM=$pt timeout 120 python tools/hot_kernels.py attn 2 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "unet" 2>&1 | tail -3
timeout 200 python tools/step_time.py 2 r2c 2>&1 | sed -n 1,3p
