for kb in 100 50 25; do echo "SDOD_GN_GROUP_SLABKB=$kb"; SDOD_GN_GROUP_SLABKB=$kb timeout 300 python tools/step_time.py 32 slab$kb 2>&1 | sed -n 2,12p | grep -E "graph|gn"; done
SDOD_GN_GROUP_SLABKB=50 timeout 200 python tools/step_time.py 2 slab50b2 2>&1 | sed -n 2,2p
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
