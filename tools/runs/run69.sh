timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "in_place or gemm_plain or epilogues" 2>&1 | tail -3
