export SDOD_STREAMK=0 SDOD_SPLITK_CLUSTER=0
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn or group_norm" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | tail -5
timeout 200 python tools/step_time.py 2 coop 2>&1 | sed -n 1,3p
SDOD_GN_COOP=0 timeout 200 python tools/step_time.py 2 nocoop 2>&1 | sed -n 1,3p
timeout 300 python tools/graph_trace.py 2 64 b2_hw64_coop 2>&1 | sed -n 3,12p
