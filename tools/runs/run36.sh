timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "layer_norm" 2>&1 | tail -3
for v in 1 0; do echo "SDOD_LN_TMA=$v"; SDOD_LN_TMA=$v timeout 300 python tools/step_time.py 32 lnt$v 2>&1 | sed -n 2,14p | grep -E "graph| ln"; done
grep -E " ln rows" gpurun_out/step_time_lnt1.txt gpurun_out/step_time_lnt0.txt
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | grep -E "rel|passed|failed" | tail -4
