timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" 2>&1 | grep -E "passed|failed|Error|error" | tail -5
for h in 0 1; do echo "SDOD_ATTN_HOIST=$h"; SDOD_ATTN_HOIST=$h timeout 120 python tools/hot_kernels.py attn 8 2>&1 | tail -1; SDOD_ATTN_HOIST=$h timeout 120 python tools/hot_kernels.py attn 2 2>&1 | tail -1; done
