timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "layer_norm or gemm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -s -k unet 2>&1 | grep -E "rel|passed|failed" | tail -6
timeout 300 python tools/step_time.py 2 lnpush 2>&1 | sed -n 2,6p
grep -E "\+ln" gpurun_out/step_time_lnpush.txt | head -5
