timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -s -k unet 2>&1 | grep -E "rel|passed|failed" | tail -6
timeout 300 python tools/step_time.py 32 sb1 2>&1 | sed -n 2,6p | grep -E "graph| attn"
SDOD_ATTN_SB=0 timeout 300 python tools/step_time.py 32 sb0 2>&1 | sed -n 2,6p | grep -E "graph| attn"
timeout 300 python tools/step_time.py 2 sb1b2 2>&1 | sed -n 2,2p
grep "d80" gpurun_out/step_time_sb1.txt gpurun_out/step_time_sb0.txt gpurun_out/step_time_sb1b2.txt
