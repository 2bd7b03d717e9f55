mkdir -p /tmp/rep
HOT_ONCE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05_kernel -c 1 -o /tmp/rep/geglu -f python tools/hot_kernels.py geglu 8 > /tmp/rep/log 2>&1
python tools/ncu_full_summary.py /tmp/rep/geglu.ncu-rep > gpurun_out/r02_geglu_source_top.txt
python tools/ncu_source_top.py /tmp/rep/geglu.ncu-rep 30 >> gpurun_out/r02_geglu_source_top.txt
wc -l gpurun_out/r02_geglu_source_top.txt
