timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_d.json 2> gpurun_out/bench_r2_d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_d.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ("value","ms_per_step","unet_step_p50_ms","iteration_ms_in_loop","gpu_launches")}, "e2e", d["e2e"]["value"], "roofline", round(d["roofline"]["frac"],3), "step", round(d["step_roofline"]["frac"],3), d["parity"], d["clocks"])
PY
tail -3 gpurun_out/bench_r2_d.err
timeout 300 python tools/step_time.py 2 final_b2 2>&1 | sed -n 1,12p
timeout 300 python tools/step_time.py 32 final_b32 2>&1 | sed -n 1,14p
timeout 300 python tools/graph_trace.py 2 64 final 2>&1 | sed -n 1,12p
timeout 300 python tools/config_sweep.py > gpurun_out/r02_config_sweep_b.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/r02_config_sweep_b.json')); print({k:(round(v.get('ms',v.get('images_per_s')),2)) for k,v in d.items()})"
