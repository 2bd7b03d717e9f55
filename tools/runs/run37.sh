timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "layer_norm" 2>&1 | tail -2
SDOD_LN_TMA=1 timeout 300 python tools/step_time.py 32 lnt1 2>&1 | sed -n 2,14p | grep -E "graph| ln"
grep -E " ln rows" gpurun_out/step_time_lnt1.txt
