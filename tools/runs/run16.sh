timeout 900 python -m pytest tests/test_gpu_text.py -x -q -s 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_libsdod.py -x -q -s -k "text or prompt or app_flow" 2>&1 | tail -8
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
