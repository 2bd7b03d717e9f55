timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 200 python tools/step_time.py 2 r2a 2>&1 | sed -n 1,20p
timeout 300 python tools/graph_trace.py 2 64 r2a 2>&1 | sed -n 1,24p
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; tail -c 4000 gpurun_out/bench_r2_a.json; tail -5 gpurun_out/bench_r2_a.err
