export SDOD_STREAMK=0 SDOD_SPLITK_CLUSTER=0
timeout 300 python tools/graph_trace.py 2 64 b2_hw64 2>&1 | tail -30
timeout 300 python tools/graph_trace.py 2 8 b2_hw8 2>&1 | tail -30
