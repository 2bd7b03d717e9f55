timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "layer_norm" 2>&1 | tail -2
for v in 1 0; do echo "SDOD_LN_VARIANT=$v"; SDOD_LN_VARIANT=$v timeout 300 python tools/step_time.py 32 ln$v 2>&1 | sed -n 2,14p | grep -E "graph| ln"; done
grep -E " ln rows" gpurun_out/step_time_ln1.txt gpurun_out/step_time_ln0.txt
