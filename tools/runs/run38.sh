SDOD_GN_ASYNC=1 timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn or group_norm" 2>&1 | tail -2
for v in 1 0; do echo "SDOD_GN_ASYNC=$v"; SDOD_GN_ASYNC=$v timeout 300 python tools/step_time.py 32 gna$v 2>&1 | sed -n 2,14p | grep -E "graph| gn1"; SDOD_GN_ASYNC=$v timeout 300 python tools/step_time.py 2 gnab2$v 2>&1 | sed -n 2,2p; done
grep -E " gn1 f32 HW4096 C320$| gn1 bf16 HW4096 C320$| gn1 f32 HW1024 C640$" gpurun_out/step_time_gna1.txt gpurun_out/step_time_gna0.txt
