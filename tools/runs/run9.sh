timeout 900 python -m pytest tests/test_gpu_pair.py -x -q -s 2>&1 | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/pair_check.py --S 64 --images 1 --perf 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_2gpu_a.json 2> gpurun_out/bench_r2_2gpu_a.err; tail -c 2500 gpurun_out/bench_r2_2gpu_a.json; tail -5 gpurun_out/bench_r2_2gpu_a.err
