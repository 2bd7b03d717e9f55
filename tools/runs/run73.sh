SDOD_GN_PIPE=1 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "gn" 2>&1 | tail -2
SDOD_GN_PIPE=1 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "unet" 2>&1 | grep -E "rel|passed|failed|rror" | tail -6
SDOD_GN_PIPE=1 timeout 300 python tools/step_time.py 32 gp1 2>&1 | sed -n 2,14p | grep -E "graph| gn1"
timeout 300 python tools/step_time.py 32 gp0 2>&1 | sed -n 2,14p | grep -E "graph| gn1"
for f in gp1 gp0; do echo $f; grep " gn1 " gpurun_out/step_time_$f.txt | sed -n 2,9p; done
