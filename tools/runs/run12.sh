timeout 300 python tools/graph_trace.py 2 64 r2b 2>&1 | sed -n 1,95p
timeout 300 python tools/step_time.py 32 r2_b32 2>&1 | sed -n 1,60p
