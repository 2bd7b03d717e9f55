mkdir -p /tmp/rep
HOT_ONCE=1 timeout 120 python tools/hot_kernels.py geglu 32 > /dev/null 2>&1
HOT_ONCE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -c 1 -o /tmp/rep/geglu -f python tools/hot_kernels.py geglu 32 > /tmp/rep/log 2>&1
python tools/ncu_full_summary.py /tmp/rep/geglu.ncu-rep > gpurun_out/r02_geglu_b32_source_top_final.txt
python tools/ncu_source_top.py /tmp/rep/geglu.ncu-rep 45 >> gpurun_out/r02_geglu_b32_source_top_final.txt
head -56 gpurun_out/r02_geglu_b32_source_top_final.txt | cut -c1-200
