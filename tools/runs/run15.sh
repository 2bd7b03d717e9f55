timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention or qkv" 2>&1 | tail -3
for m in 3 2 1 0; do echo "SDOD_ATTN_MODE=$m"; SDOD_ATTN_MODE=$m timeout 120 python tools/hot_kernels.py attn 8 2>&1 | tail -1; done
SDOD_ATTN_MODE=3 timeout 120 python tools/hot_kernels.py attn 32 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "unet" 2>&1 | tail -3
