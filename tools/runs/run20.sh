timeout 900 python -m pytest tests/test_gpu_pair.py -x -q -s 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/pair_check.py --S 64 --images 1 --perf 2>&1 | tail -3
