timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_text.py -x -q -k "attention or clip or text or encoder" 2>&1 | grep -E "passed|failed|Error|error" | tail -5
timeout 120 python tools/hot_kernels.py attn 8 2>&1 | tail -1
timeout 300 python tools/step_time.py 32 qpc 2>&1 | sed -n 2,6p | grep -E "graph| attn"
timeout 300 python tools/step_time.py 2 qpcb2 2>&1 | sed -n 2,2p
grep " attn " gpurun_out/step_time_qpc.txt | head -6
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | grep -E "rel|passed|failed" | tail -6
