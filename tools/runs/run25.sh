for kb in 200 100; do echo "SDOD_GN_GROUP_SLABKB=$kb"; SDOD_GN_GROUP_SLABKB=$kb timeout 300 python tools/step_time.py 32 slab$kb 2>&1 | sed -n 2,12p | grep -E "graph|gn"; done
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err; tail -c 3800 gpurun_out/bench_r2_c.json; tail -3 gpurun_out/bench_r2_c.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 1500 gpurun_out/bench_r2_ref.json
