timeout 900 python -m pytest tests/test_gpu_pair.py -x -q -s 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn or group_norm" 2>&1 | tail -2
timeout 200 python tools/gn_c1.py 2>&1 | tail -1 | cut -c1-420
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_2gpu_b.json 2> gpurun_out/bench_r2_2gpu_b.err; tail -c 3500 gpurun_out/bench_r2_2gpu_b.json; tail -5 gpurun_out/bench_r2_2gpu_b.err
