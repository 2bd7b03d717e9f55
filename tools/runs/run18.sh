timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn or group_norm" 2>&1 | tail -3
for cs in 0 1 2 4 8; do echo "SDOD_GN_CS=$cs"; SDOD_GN_CS=$cs timeout 200 python tools/gn_c1.py 2>&1 | tail -1 | cut -c1-330; done
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
