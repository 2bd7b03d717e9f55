timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn or group_norm or epilogue" 2>&1 | tail -2
timeout 300 python tools/step_time.py 32 gnf 2>&1 | sed -n 2,14p | grep -E "graph| gn1"
timeout 300 python tools/step_time.py 2 gnfb2 2>&1 | sed -n 2,2p
grep -E " gn1 f32 HW4096 C320$| gn1 bf16 HW4096 C320$| gn1 f32 HW1024 C640$" gpurun_out/step_time_gnf.txt
timeout 200 python tools/gn_c1.py 2>&1 | tail -1 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s 2>&1 | grep -E "rel|passed|failed|dB" | tail -10
