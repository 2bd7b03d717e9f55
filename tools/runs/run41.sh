timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err; echo rc=$?; tail -c 600 gpurun_out/bench_r2_final2.json; tail -3 gpurun_out/bench_r2_final2.err | cut -c1-300
timeout 200 python tools/gn_c1.py > gpurun_out/gn_c1_final2.json 2>&1; tail -c 400 gpurun_out/gn_c1_final2.json
timeout 300 python tools/step_time.py 32 final2_b32 2>&1 | sed -n 2,14p
timeout 300 python tools/step_time.py 2 final2_b2 2>&1 | sed -n 2,12p
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
