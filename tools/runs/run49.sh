timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_efficient_gn.py -x -q -k "gn or group_norm or GroupNorm or efficient" 2>&1 | grep -E "passed|failed|Error|error" | tail -3
timeout 200 python tools/gn_c1.py 2>&1 | tail -1 | cut -c1-330
timeout 300 python tools/step_time.py 32 gnpush 2>&1 | sed -n 2,14p | grep -E "graph| gn1"
timeout 300 python tools/step_time.py 2 gnpushb2 2>&1 | sed -n 2,14p | grep -E "graph| gn1"
grep -E " gn1 (f32|bf16) HW4096 C320$| gn1 f32 HW1024 C640$" gpurun_out/step_time_gnpush.txt gpurun_out/step_time_gnpushb2.txt
