timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "gn or group_norm or split or stream or optional" 2>&1 | tail -5
timeout 200 python tools/step_time.py 2 pol1 2>&1 | sed -n 1,3p
SDOD_SPLIT_POLICY=0 timeout 200 python tools/step_time.py 2 pol0 2>&1 | sed -n 1,3p
timeout 300 python tools/graph_trace.py 2 64 b2_hw64_pol1 2>&1 | sed -n 3,12p
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; tail -c 3000 gpurun_out/bench_r2_a.json; tail -5 gpurun_out/bench_r2_a.err
