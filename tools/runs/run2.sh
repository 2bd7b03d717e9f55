timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "layer_norm or split or second or skip" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | tail -8
timeout 200 python tools/step_time.py 2 ln 2>&1 | sed -n 1,14p
SDOD_LN_FUSE=0 timeout 200 python tools/step_time.py 2 noln 2>&1 | sed -n 1,3p
timeout 200 python tools/gemm_timeline.py 2>&1 | tail -16
timeout 120 python tools/unet_step.py 2 2 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tcgen05|gn_|layer_norm|attention|splitk|cast_|concat|im2col|upsample|silu" -c 900 --csv --log-file gpurun_out/r02_step_b2_launches_a.csv python tools/unet_step.py 2 2 > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
