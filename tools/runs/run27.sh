timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "stride2 or upsample or conv" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -s -k unet 2>&1 | grep -E "passed|failed|rel" | tail -8
for v in 1 0; do echo "SDOD_CONV_S2=$v"; SDOD_CONV_S2=$v timeout 300 python tools/step_time.py 32 s2_$v 2>&1 | sed -n 1,2p; SDOD_CONV_S2=$v timeout 300 python tools/step_time.py 2 s2b2_$v 2>&1 | sed -n 1,2p; done
grep -E "conv3s2|im2col|cast" gpurun_out/step_time_s2_1.txt gpurun_out/step_time_s2_0.txt | head -14
