timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention or qkv" 2>&1 | tail -3
for m in 2 1; do echo "SDOD_ATTN_MODE=$m"; SDOD_ATTN_MODE=$m timeout 120 python tools/hot_kernels.py attn 8 2>&1 | tail -1; done
HOT_ONCE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_kernel -c 1 -o gpurun_out/r02_attn40_m2 -f python tools/hot_kernels.py attn 8 > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
ls -la gpurun_out/*.ncu-rep
