timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "conv or pair" 2>&1 | tail -2
SDOD_PAIR_RELAXED=0 timeout 300 python tools/step_time.py 32 pr0 2>&1 | sed -n 2,6p | grep -E "graph| conv3"
timeout 300 python tools/step_time.py 32 pr1 2>&1 | sed -n 2,6p | grep -E "graph| conv3"
for f in pr0 pr1; do echo $f; grep -E " conv3" gpurun_out/step_time_$f.txt | sed -n 5,12p; done
timeout 300 python tools/vae_step.py 8 2>&1 | tail -3
SDOD_PAIR_RELAXED=0 timeout 300 python tools/vae_step.py 8 2>&1 | tail -3
