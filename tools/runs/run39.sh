mkdir -p /tmp/rep
timeout 120 python tools/unet_step.py 1 32 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gn_nhwc_group_kernel -s 1 -c 1 -o /tmp/rep/gn -f python tools/unet_step.py 1 32 > /tmp/rep/log 2>&1
python tools/ncu_full_summary.py /tmp/rep/gn.ncu-rep > gpurun_out/r02_gn_group_b32_source_top.txt
python tools/ncu_source_top.py /tmp/rep/gn.ncu-rep 40 >> gpurun_out/r02_gn_group_b32_source_top.txt
head -50 gpurun_out/r02_gn_group_b32_source_top.txt | cut -c1-230
