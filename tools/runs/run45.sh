timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" 2>&1 | grep -E "passed|failed|Error|error" | tail -5
for c in 1 2 4 8 0; do echo "SDOD_ATTN_QPC=$c"; SDOD_ATTN_QPC=$c timeout 120 python tools/hot_kernels.py xattn 32 2>&1 | grep xattn; done
timeout 120 python tools/hot_kernels.py attn 8 2>&1 | tail -1
